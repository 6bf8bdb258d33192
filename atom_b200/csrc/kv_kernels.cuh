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
// grid (B, H), 160 threads: warp 4 streams whole pages (K block, V block and their params are each contiguous) into an
// 8-stage smem ring with cp.async.bulk + mbarrier complete_tx -- enough bytes in flight per SM to cover HBM latency
// (a register-load version of this kernel was latency bound at 0.8 TB/s).  Warp w < 4 consumes pages w, w+4, ...
// Inside a page a lane = (token slot ts = lane/4, quarter c = lane%4) handles tokens ts, ts+8, ...: for QK it holds the
// RoPE pairs i = 16c..16c+15 (elements i and i+64), for PV the V elements 32c..32c+31.
// RoPE: with z = x_i + j x_{i+64}, rope(x, p) = z e^{j p theta_i} and q.k = Re(zq conj(zk)), so
//   score(t) = Re( [zq e^{j(len-1)theta} e^{-j pagebase theta}] * conj( zk e^{j t_lo theta} ) ):
// the bracket is advanced once per page by a constant rotation, e^{j t_lo theta} comes from a P x 64 smem table -- no
// transcendental per token (the reference evaluates __sincosf per element per token, decode.cuh:39-71).
// Softmax is blocked per page: one rescale of the accumulator per page instead of per token.
// V dequant is folded: sum_t p_t (n s_t - z_t) = sum_t (p_t s_t) n - sum_t p_t z_t (the last sum is a scalar).
constexpr int DEC_CONSUMERS = 4;
constexpr int DEC_THREADS = 32 * (DEC_CONSUMERS + 1);
constexpr int DEC_STAGES = 8;
constexpr int DEC_MAX_TPL = 8;   // tokens per lane per page = P / 8  (P <= 64)

__device__ __forceinline__ float nib_f(uint32_t w, int e) { return (float)((w >> (4 * e)) & 0xFu); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DEC_THREADS, 4)
batch_decode_kernel(__half* __restrict__ o, const __half* __restrict__ q, KvArgs kv) {
  extern __shared__ __align__(128) uint8_t smem_d[];
  const int P = kv.P;
  const int stage_bytes = 2 * 64 * P + 2 * 4 * P;                   // K | V | K params | V params
  uint8_t* ring = smem_d;
  float2* tab = reinterpret_cast<float2*>(smem_d + DEC_STAGES * stage_bytes);   // [P][64]  (cos, sin)(t_lo * theta_i)
  float2* stepr = tab + P * 64;                                     // [64]     e^{-j 4P theta_i}
  float* merge = reinterpret_cast<float*>(stepr + 64);              // [4 warps][4 quarters][34]
  uint64_t* full = reinterpret_cast<uint64_t*>(merge + 4 * 4 * 34);
  uint64_t* empty = full + DEC_STAGES;

  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ts = lane >> 2, c = lane & 3;
  const int page_begin = kv.indptr[b], npages = kv.indptr[b + 1] - page_begin;
  const int last_valid = kv.last_page_offset[b];
  const int seq_len = (npages - 1) * P + last_valid;
  constexpr float kLog2Theta = 13.287712379549449f;                // log2(1e4)
  constexpr float kSmScale = 0.08838834764831845f * 1.4426950408889634f;   // 1/sqrt(128) * log2(e)

  if (tid == 0) {
    for (int i = 0; i < DEC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  // table layout [j/2][t_lo][c][j%2] (pair = 16c + j): for a fixed j the 32 lanes (t_lo = ts + 8i, c) of a warp read 32
  // consecutive 16-B words -- conflict free.  (A [t_lo][pair] layout put all 32 lanes on the same bank: 82 % of the
  // kernel's shared-memory wavefronts were conflict replays.)
  for (int i = tid; i < P * 64; i += DEC_THREADS) {
    const int tl = i >> 6, pair = i & 63, cc = pair >> 4, j = pair & 15;
    const float f = exp2f(-(float)pair * (kLog2Theta / 64.f));
    float sn, cs; sincosf((float)tl * f, &sn, &cs);
    tab[(((j >> 1) * P + tl) * 4 + cc) * 2 + (j & 1)] = make_float2(cs, sn);
  }
  if (tid < 64) {
    const float f = exp2f(-(float)tid * (kLog2Theta / 64.f));
    float sn, cs; sincosf((float)(DEC_CONSUMERS * P) * f, &sn, &cs);
    stepr[tid] = make_float2(cs, -sn);
  }
  __syncthreads();

  if (warp == DEC_CONSUMERS) {
    // ------------------------------------------------------------ producer: one elected lane streams the pages
    if (lane == 0) {
      for (int pg = 0; pg < npages; ++pg) {
        const int s = pg % DEC_STAGES;
        mbar_wait(&empty[s], ((pg / DEC_STAGES) & 1) ^ 1);
        const int page = kv.indices[page_begin + pg];
        const size_t rk = kv_row(kv, page, 0, h, 0), rv = kv_row(kv, page, 1, h, 0);
        uint8_t* st = ring + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        bulk_g2s(st, kv.data + rk * 64, 64 * P, &full[s]);
        bulk_g2s(st + 64 * P, kv.data + rv * 64, 64 * P, &full[s]);
        bulk_g2s(st + 128 * P, kv.param + rk, 4 * P, &full[s]);
        bulk_g2s(st + 132 * P, kv.param + rv, 4 * P, &full[s]);
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  // rotated query for this lane's 16 pairs, pre-multiplied by e^{-j (warp*P) theta} (this warp's first page)
  float qre[16], qim[16];
  {
    const __half* qh = q + ((size_t)b * kv.H + h) * 128;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = 16 * c + j;
      const float f = exp2f(-(float)i * (kLog2Theta / 64.f));
      const float xr = __half2float(qh[i]), xi = __half2float(qh[i + 64]);
      float sn, cs; sincosf((float)(seq_len - 1 - warp * P) * f, &sn, &cs);
      qre[j] = xr * cs - xi * sn;
      qim[j] = xi * cs + xr * sn;
    }
  }
  float m = -5e4f, d = 0.f, zsum = 0.f, acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int tpl = P >> 3;                                          // tokens per lane per page

  for (int pg = warp; pg < npages; pg += DEC_CONSUMERS) {
    const int s = pg % DEC_STAGES;
    const int valid = (pg == npages - 1) ? last_valid : P;
    mbar_wait(&full[s], (pg / DEC_STAGES) & 1);
    const uint8_t* st = ring + s * stage_bytes;
    const uint8_t* kblk = st;
    const uint8_t* vblk = st + 64 * P;
    const __half2* kpar = reinterpret_cast<const __half2*>(st + 128 * P);
    const __half2* vpar = reinterpret_cast<const __half2*>(st + 132 * P);

    // ---- scores of this lane's tokens
    float x[DEC_MAX_TPL];
    float xmax = -5e4f;
#pragma unroll
    for (int i = 0; i < DEC_MAX_TPL; ++i) {
      x[i] = 0.f;
      if (i < tpl) {
        const int tl = ts + 8 * i;
        const uint8_t* kr = kblk + tl * 64;
        // the two 8-B halves are read in opposite order by alternate token pairs so that one instruction touches both
        // 32-B halves of the 64-B rows (2-way instead of 4-way bank conflict on the linear page layout)
        const bool swp = (ts & 2) != 0;
        const uint2 k_a = *reinterpret_cast<const uint2*>(kr + (swp ? 32 : 0) + c * 8);
        const uint2 k_b = *reinterpret_cast<const uint2*>(kr + (swp ? 0 : 32) + c * 8);
        const uint2 k_lo = swp ? k_b : k_a;                                    // elements 16c .. 16c+15
        const uint2 k_hi = swp ? k_a : k_b;                                    // elements 64+16c ..
        const float2 kp = __half22float2(kpar[tl]);
        const float4* trow = reinterpret_cast<const float4*>(tab) + tl * 4 + c;
        float xa = 0.f, xb = 0.f;      // two chains: the 32 fma of a token are no longer one dependent sequence
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 t2 = trow[(j >> 1) * P * 4];                          // (cos, sin) of pairs j and j+1
          const uint32_t wl = (j < 8) ? k_lo.x : k_lo.y, wh = (j < 8) ? k_hi.x : k_hi.y;
          {
            const float kre = fmaf(nib_f(wl, j & 7), kp.x, -kp.y), kim = fmaf(nib_f(wh, j & 7), kp.x, -kp.y);
            const float rr = kre * t2.x - kim * t2.y, ri = kim * t2.x + kre * t2.y;     // zk e^{j t_lo theta}
            xa = fmaf(qre[j], rr, xa);
            xa = fmaf(qim[j], ri, xa);
          }
          {
            const float kre = fmaf(nib_f(wl, (j + 1) & 7), kp.x, -kp.y), kim = fmaf(nib_f(wh, (j + 1) & 7), kp.x, -kp.y);
            const float rr = kre * t2.z - kim * t2.w, ri = kim * t2.z + kre * t2.w;
            xb = fmaf(qre[j + 1], rr, xb);
            xb = fmaf(qim[j + 1], ri, xb);
          }
        }
        xa += xb;
        xa += __shfl_xor_sync(0xffffffffu, xa, 1);
        xa += __shfl_xor_sync(0xffffffffu, xa, 2);
        x[i] = xa * kSmScale;
        if (tl < valid) xmax = fmaxf(xmax, x[i]);
      }
    }
    // ---- one rescale per page, then p * v with the dequant folded
    const float m_new = fmaxf(m, xmax);
    const float sc = exp2f(m - m_new);
    m = m_new;
    d *= sc; zsum *= sc;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] *= sc;
#pragma unroll
    for (int i = 0; i < DEC_MAX_TPL; ++i) {
      if (i < tpl) {
        const int tl = ts + 8 * i;
        if (tl < valid) {
          const uint4 vw = *reinterpret_cast<const uint4*>(vblk + tl * 64 + c * 16);
          const float2 vp = __half22float2(vpar[tl]);
          const float p = exp2f(x[i] - m_new);
          d += p;
          zsum = fmaf(p, vp.y, zsum);
          const float ps = p * vp.x;
          const uint32_t w4[4] = {vw.x, vw.y, vw.z, vw.w};
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = fmaf(ps, nib_f(w4[e >> 3], e & 7), acc[e]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    // advance the query rotation to this warp's next page
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 st2 = stepr[16 * c + j];
      const float nr = qre[j] * st2.x - qim[j] * st2.y;
      qim[j] = qim[j] * st2.x + qre[j] * st2.y;
      qre[j] = nr;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] -= zsum;      // sum_t p_t z_t is common to all elements of the head

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
  asm volatile("bar.sync 1, 128;" ::: "memory");     // the 4 consumer warps (the producer warp has left)
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
      const float sx = exp2f(src[0] - mm);
      dd = fmaf(src[1], sx, dd);
#pragma unroll
      for (int i = 0; i < 32; ++i) out[i] = fmaf(src[2 + i], sx, out[i]);
    }
    const float inv = 1.f / dd;
    __half* dst = o + ((size_t)b * kv.H + h) * 128 + 32 * c;
#pragma unroll
    for (int i = 0; i < 32; i += 2)
      *reinterpret_cast<__half2*>(dst + i) = __floats2half2_rn(out[i] * inv, out[i + 1] * inv);
  }
}

}  // namespace atom
