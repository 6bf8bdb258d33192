// kv_kernels.cuh -- paged INT4 KV cache: append (decode / prefill) and batch decode attention with fused RoPE.
//
// Replace AppendPagedKVCacheDecodeKernel / AppendPagedKVCachePrefillKernel
//         (/root/reference/kernels/include/flashinfer/page.cuh:119-216) and
//         BatchDecodeWithPagedKVCacheKernel (/root/reference/kernels/include/flashinfer/decode.cuh:480-689).
// Cache layout is the reference's (utils/kvcache.py:17-24):
//   data  u8  [pages][L][2][H][P][64]   two INT4 per byte, element 2j in the low nibble
//   param f16 [pages][L][2][H][P][2]    (scale, zero);  x = nibble * scale - zero   (quantization.cuh:76)
// K is stored pre-RoPE; RoPE(theta = 1e4) is applied to q at position len-1 and to k at its index.
#pragma once
#include "ptx_sm100.cuh"

namespace atom {

struct KvArgs {
  uint8_t* data;
  __half2* param;
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* last_page_offset;
  int L, layer, H, P, B;
};

__device__ __forceinline__ size_t kv_row(const KvArgs& kv, int page, int which, int head, int entry) {
  return ((((size_t)page * kv.L + kv.layer) * 2 + which) * kv.H + head) * kv.P + entry;
}

// ---------------------------------------------------------------- K7 / K8: append
// One 16-thread group moves one (token, head): 64 B of K, 64 B of V (4 B per lane each) and the two params.
// append_indptr == nullptr: decode append (one token per sequence, at position seq_len-1).
__global__ void __launch_bounds__(256)
append_kv_kernel(KvArgs kv, const uint8_t* __restrict__ k, const uint8_t* __restrict__ v, const __half2* __restrict__ kp,
                 const __half2* __restrict__ vp, const int32_t* __restrict__ append_indptr, int total_tokens) {
  const long long unit = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // (token, head)
  const int sub = threadIdx.x & 15;
  if (unit >= (long long)total_tokens * kv.H) return;
  const int tok = (int)(unit / kv.H), head = (int)(unit % kv.H);
  int b, pos;
  if (append_indptr == nullptr) {
    b = tok;
    pos = (kv.indptr[b + 1] - kv.indptr[b] - 1) * kv.P + kv.last_page_offset[b] - 1;
  } else {
    int lo = 0, hi = kv.B;                          // largest b with append_indptr[b] <= tok
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (append_indptr[mid] <= tok) lo = mid; else hi = mid; }
    b = lo;
    const int seq_len = (kv.indptr[b + 1] - kv.indptr[b] - 1) * kv.P + kv.last_page_offset[b];
    const int app = append_indptr[b + 1] - append_indptr[b];
    pos = seq_len - app + (tok - append_indptr[b]);
  }
  const int page = kv.indices[kv.indptr[b] + pos / kv.P], entry = pos % kv.P;
  const size_t rk = kv_row(kv, page, 0, head, entry), rv = kv_row(kv, page, 1, head, entry);
  const size_t src = ((size_t)tok * kv.H + head);
  reinterpret_cast<uint32_t*>(kv.data + rk * 64)[sub] = reinterpret_cast<const uint32_t*>(k + src * 64)[sub];
  reinterpret_cast<uint32_t*>(kv.data + rv * 64)[sub] = reinterpret_cast<const uint32_t*>(v + src * 64)[sub];
  if (sub == 0) { kv.param[rk] = kp[src]; kv.param[rv] = vp[src]; }
}

// ---------------------------------------------------------------- K6: batch decode
// grid (B, H), 128 threads.  Warp w owns pages w, w+4, ...; inside a page a warp handles 8 tokens per step:
// lane = (token slot ts = lane/4, quarter c = lane%4).  For QK the lane holds RoPE pairs i = 16c .. 16c+15
// (elements i and i+64: two 8-B loads); for PV it holds V elements 32c .. 32c+31 (one 16-B load).  A warp step reads
// 512 contiguous bytes of K and of V.
// RoPE: with z = x_i + j x_{i+64}, rope(x, p) = z e^{j p theta_i} and q.k = Re(zq conj(zk)), so
//   score(t) = Re( [zq e^{j(len-1)theta} e^{-j pagebase theta}] * conj( zk e^{j t_lo theta} ) ):
// the bracket is advanced once per page by a constant rotation, e^{j t_lo theta} comes from a P x 64 smem table.
constexpr int DEC_THREADS = 128;

__device__ __forceinline__ float nib_f(uint32_t w, int e) { return (float)((w >> (4 * e)) & 0xFu); }

__global__ void __launch_bounds__(DEC_THREADS)
batch_decode_kernel(__half* __restrict__ o, const __half* __restrict__ q, KvArgs kv) {
  extern __shared__ __align__(16) uint8_t smem_d[];
  float2* tab = reinterpret_cast<float2*>(smem_d);                 // [P][64]  (cos, sin)(t_lo * theta_i)
  float2* stepr = tab + kv.P * 64;                                 // [64]     e^{-j 4P theta_i}
  float* merge = reinterpret_cast<float*>(stepr + 64);             // [4 warps][4 quarters][34]

  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ts = lane >> 2, c = lane & 3;
  const int page_begin = kv.indptr[b], npages = kv.indptr[b + 1] - page_begin;
  const int seq_len = (npages - 1) * kv.P + kv.last_page_offset[b];
  constexpr float kLog2Theta = 13.287712379549449f;                // log2(1e4)
  constexpr float kSmScale = 0.08838834764831845f * 1.4426950408889634f;   // 1/sqrt(128) * log2(e)

  for (int i = tid; i < kv.P * 64; i += DEC_THREADS) {
    const float f = exp2f(-(float)(i & 63) * (kLog2Theta / 64.f));
    float sn, cs; sincosf((float)(i >> 6) * f, &sn, &cs);
    tab[i] = make_float2(cs, sn);
  }
  if (tid < 64) {
    const float f = exp2f(-(float)tid * (kLog2Theta / 64.f));
    float sn, cs; sincosf((float)(4 * kv.P) * f, &sn, &cs);
    stepr[tid] = make_float2(cs, -sn);
  }

  // rotated query for this lane's 16 pairs, pre-multiplied by e^{-j (warp*P) theta}
  float qre[16], qim[16];
  {
    const __half* qh = q + ((size_t)b * kv.H + h) * 128;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = 16 * c + j;
      const float f = exp2f(-(float)i * (kLog2Theta / 64.f));
      const float xr = __half2float(qh[i]), xi = __half2float(qh[i + 64]);
      float sn, cs; sincosf((float)(seq_len - 1 - warp * kv.P) * f, &sn, &cs);
      qre[j] = xr * cs - xi * sn;
      qim[j] = xi * cs + xr * sn;
    }
  }
  __syncthreads();

  float m = -5e4f, d = 0.f, acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;

  for (int pg = warp; pg < npages; pg += 4) {
    const int page = kv.indices[page_begin + pg];
    const int valid = (pg == npages - 1) ? kv.last_page_offset[b] : kv.P;
    const size_t rk = kv_row(kv, page, 0, h, 0), rv = kv_row(kv, page, 1, h, 0);
    for (int s0 = 0; s0 < valid; s0 += 8) {
      const int tl = s0 + ts;
      const bool active = tl < valid;
      float x = 0.f;
      uint4 vw = make_uint4(0, 0, 0, 0);
      float2 vpar = make_float2(0.f, 0.f);
      if (active) {
        const uint8_t* kr = kv.data + (rk + tl) * 64;
        const uint2 k_lo = *reinterpret_cast<const uint2*>(kr + c * 8);        // elements 16c .. 16c+15
        const uint2 k_hi = *reinterpret_cast<const uint2*>(kr + 32 + c * 8);   // elements 64+16c ..
        vw = *reinterpret_cast<const uint4*>(kv.data + (rv + tl) * 64 + c * 16);
        const float2 kpar = __half22float2(kv.param[rk + tl]);
        vpar = __half22float2(kv.param[rv + tl]);
        const float2* trow = tab + tl * 64 + 16 * c;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t wl = (j < 8) ? k_lo.x : k_lo.y, wh = (j < 8) ? k_hi.x : k_hi.y;
          const float kre = fmaf(nib_f(wl, j & 7), kpar.x, -kpar.y);
          const float kim = fmaf(nib_f(wh, j & 7), kpar.x, -kpar.y);
          const float2 t = trow[j];
          const float rr = kre * t.x - kim * t.y;     // Re(zk e^{j t_lo theta})
          const float ri = kim * t.x + kre * t.y;
          x = fmaf(qre[j], rr, x);
          x = fmaf(qim[j], ri, x);
        }
      }
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      if (active) {
        x *= kSmScale;
        const float m_new = fmaxf(m, x);
        const float sc = exp2f(m - m_new), p = exp2f(x - m_new);
        d = fmaf(d, sc, p);
        m = m_new;
        const uint32_t w4[4] = {vw.x, vw.y, vw.z, vw.w};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float vv = fmaf(nib_f(w4[i >> 3], i & 7), vpar.x, -vpar.y);
          acc[i] = fmaf(acc[i], sc, p * vv);
        }
      }
    }
    // advance the query rotation by 4 pages
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 st = stepr[16 * c + j];
      const float nr = qre[j] * st.x - qim[j] * st.y;
      qim[j] = qim[j] * st.x + qre[j] * st.y;
      qre[j] = nr;
    }
  }

  // merge the 8 token-slot states of the warp (same quarter c: lanes differ in bits 2..4)
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, off), d_o = __shfl_xor_sync(0xffffffffu, d, off);
    const float m_new = fmaxf(m, m_o);
    const float s_a = exp2f(m - m_new), s_b = exp2f(m_o - m_new);
    d = d * s_a + d_o * s_b;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float a_o = __shfl_xor_sync(0xffffffffu, acc[i], off);
      acc[i] = acc[i] * s_a + a_o * s_b;
    }
    m = m_new;
  }
  if (ts == 0) {
    float* dst = merge + (warp * 4 + c) * 34;
    dst[0] = m; dst[1] = d;
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[2 + i] = acc[i];
  }
  __syncthreads();
  if (warp == 0 && ts == 0) {
    float mm = -5e4f;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, merge[(w * 4 + c) * 34]);
    float dd = 0.f, out[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float* src = merge + (w * 4 + c) * 34;
      const float s = exp2f(src[0] - mm);
      dd = fmaf(src[1], s, dd);
#pragma unroll
      for (int i = 0; i < 32; ++i) out[i] = fmaf(src[2 + i], s, out[i]);
    }
    const float inv = 1.f / dd;
    __half* dst = o + ((size_t)b * kv.H + h) * 128 + 32 * c;
#pragma unroll
    for (int i = 0; i < 32; i += 2)
      *reinterpret_cast<__half2*>(dst + i) = __floats2half2_rn(out[i] * inv, out[i + 1] * inv);
  }
}

}  // namespace atom
