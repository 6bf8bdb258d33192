// gemm_i4_sm100.cuh -- W4A4 group-quantised GEMM with INT8 keeper for B200 (sm_100a).
//
// Replaces the reference's compute_gemm_imma / DenseLayerGEMM_i4[_o4]_kernel
// (/root/reference/kernels/include/GEMM/Dense_layer_gemm_i4_o16.cuh:436-710,
//  /root/reference/e2e/punica-atom/punica/ops/csrc/GEMM/DenseLayerGEMM_i4_o4.cu:440-787).
// Same operands, same layouts, same arithmetic (exact INT32 group sums, one FP16 multiply of the
// two scales, FP32 fma accumulation in group order, keeper last, RN cast to half) -- different machine:
//
//   TMA (packed INT4 tiles, 64 B rows)  ->  smem "packed ring"
//   converter warpgroup: nibble -> INT8 (value*16, no sign-extension work) into the canonical
//       K-major SWIZZLE_128B operand layout ("expanded ring"); the INT8 keeper group is TMA'd straight
//       into that layout.  Blackwell has no INT4 MMA: kind::i8 is the integer tensor path.
//   one thread: tcgen05.mma.kind::i8  (128 x BN x 32) x 4 per 128-wide quantisation group, INT32 in TMEM,
//       a fresh TMEM buffer per group (ring of NB buffers)
//   epilogue warpgroups: tcgen05.ld -> acc += float(c) * half(sA*sB) in registers, overlapped with the
//       MMAs of the following groups; FP16 (o16) or asymmetric INT4 (o4) output.
//
// kSwap=false ("tall"):   MMA-M = 128 tokens (A), MMA-N = BN output channels (B).
// kSwap=true  ("skinny"): MMA-M = 128 output channels (B), MMA-N = BN tokens (A) -- decode shapes;
//       optionally split along K over a thread-block cluster, partial sums reduced through DSMEM.
#pragma once
#include "ptx_sm100.cuh"

namespace atom {

struct GemmArgs {
  const __half* a_scale;         // [G][S(M)]  ldmatrix-replicated layout (Reorder.cuh:39-50)
  const __half* b_scale;         // [G][N]
  const __half* a_keeper_scale;  // [S(M)]
  const __half* b_keeper_scale;  // [N]
  __half* d;                     // o16: [M][N]
  uint8_t* d4;                   // o4 : [M][N/2]
  __half2* d_scale;              // o4 : [M][N/128] (scale, zero)
  int M, N, G;                   // G = number of INT4 groups = K/128 - 1
  int lda_scale;                 // S(M)
};

__host__ __device__ __forceinline__ int scale_index(int row) { return (row / 16) * 64 + (row % 8) * 8 + ((row / 8) % 2); }
__host__ __device__ __forceinline__ int scale_size(int m) { return m / 16 * 64 + 64 - (1 - (m % 16) / 8) * (8 - (m % 8)) * 8; }

template <bool kSwap, int BN, int kPack, int kExp, int kSplit, bool kO4>
struct GemmCfg {
  static constexpr int BM = 128;                                  // MMA M (TMEM lanes)
  static constexpr int NB = (BN <= 128) ? 4 : 2;                  // TMEM accumulator buffers
  static constexpr int TMEM_COLS = (NB * BN < 32) ? 32 : NB * BN; // power of two for BN in {16..256}
  static constexpr int SCALE_SLOTS = kExp + NB + 1;               // see converter/epilogue lifetime argument
  static constexpr int EPI_WGS = (BN >= 64) ? 2 : 1;              // epilogue warpgroups
  static constexpr int CPT = BN / EPI_WGS;                        // accumulator columns per epilogue thread
  static constexpr int THREADS = 256 + 128 * EPI_WGS;
  static constexpr int PACK_P = BM * 64, PACK_Q = BN * 64;        // bytes per packed stage
  static constexpr int EXP_P = BM * 128, EXP_Q = BN * 128;        // bytes per expanded stage
  static constexpr int OFF_EXP_P = 0;
  static constexpr int OFF_EXP_Q = OFF_EXP_P + kExp * EXP_P;
  static constexpr int OFF_PACK_P = OFF_EXP_Q + kExp * ((EXP_Q + 1023) / 1024 * 1024);
  static constexpr int OFF_PACK_Q = OFF_PACK_P + kPack * PACK_P;
  static constexpr int OFF_SM = OFF_PACK_Q + kPack * PACK_Q;      // half2 per MMA-M row per slot
  static constexpr int OFF_SN = OFF_SM + SCALE_SLOTS * BM * 4;    // half per MMA-N column per slot
  static constexpr int OFF_BAR = (OFF_SN + SCALE_SLOTS * BN * 2 + 15) / 16 * 16;
  static constexpr int NUM_BARS = 2 * kPack + 2 * kExp + 2 * NB + SCALE_SLOTS;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;     // + slack for the 1024-B alignment fix-up
  static constexpr int RED_BYTES = BM * BN * 4;                   // split-K partials, aliased on the expanded ring
  static_assert(!kO4 || (BN == 128 || kSwap), "o4 (tall) quantises one 128-column head per CTA");
  static_assert(kSplit == 1 || RED_BYTES <= OFF_SM, "split-K reduction buffer must fit in the tile rings");
  static_assert(NB * BN <= 512, "TMEM has 512 columns");
};

// nibble -> int8(value * 16): element 2b of the word lands in byte b of `lo`, element 2b+1 in byte b of `hi`.
// The same permutation of K inside a 32-element chunk is applied to both operands, so dot products are unchanged;
// the uniform factor 16*16 = 256 is carried by the accumulator and removed exactly at the output cast.
__device__ __forceinline__ void expand_chunk(const uint4& w, uint4& lo, uint4& hi) {
  lo.x = (w.x << 4) & 0xF0F0F0F0u; hi.x = w.x & 0xF0F0F0F0u;
  lo.y = (w.y << 4) & 0xF0F0F0F0u; hi.y = w.y & 0xF0F0F0F0u;
  lo.z = (w.z << 4) & 0xF0F0F0F0u; hi.z = w.z & 0xF0F0F0F0u;
  lo.w = (w.w << 4) & 0xF0F0F0F0u; hi.w = w.w & 0xF0F0F0F0u;
}

// convert `rows` packed rows (64 B each, dense) into the K-major SWIZZLE_128B layout (128 B rows, 8-row atoms)
template <int kRows>
__device__ __forceinline__ void convert_tile(const uint8_t* __restrict__ packed, uint8_t* __restrict__ expanded, int t) {
#pragma unroll
  for (int c = t; c < kRows * 4; c += 128) {
    const int r = c >> 2, j = c & 3;
    const uint4 w = *reinterpret_cast<const uint4*>(packed + c * 16);
    uint4 lo, hi;
    expand_chunk(w, lo, hi);
    uint8_t* row = expanded + (r >> 3) * 1024 + (r & 7) * 128;
    *reinterpret_cast<uint4*>(row + (((2 * j) ^ (r & 7)) << 4)) = lo;
    *reinterpret_cast<uint4*>(row + (((2 * j + 1) ^ (r & 7)) << 4)) = hi;
  }
}

template <bool kSwap, int BN, int kPack, int kExp, int kSplit, bool kO4>
__global__ void __launch_bounds__(GemmCfg<kSwap, BN, kPack, kExp, kSplit, kO4>::THREADS, 1)
gemm_i4_kernel(const __grid_constant__ CUtensorMap tm_p4,   // packed INT4, MMA-M operand  (box 64 B x 128 rows)
               const __grid_constant__ CUtensorMap tm_q4,   // packed INT4, MMA-N operand  (box 64 B x BN rows)
               const __grid_constant__ CUtensorMap tm_p8,   // INT8 keeper, MMA-M operand  (box 128 B x 128 rows, SW128)
               const __grid_constant__ CUtensorMap tm_q8,   // INT8 keeper, MMA-N operand  (box 128 B x BN rows, SW128)
               const GemmArgs args) {
  using C = GemmCfg<kSwap, BN, kPack, kExp, kSplit, kO4>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;
  uint64_t* pack_empty = pack_full + kPack;
  uint64_t* exp_full = pack_empty + kPack;
  uint64_t* exp_empty = exp_full + kExp;
  uint64_t* tmem_full = exp_empty + kExp;
  uint64_t* tmem_empty = tmem_full + C::NB;
  uint64_t* scale_full = tmem_empty + C::NB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates: blockIdx.x walks the MMA-N operand's tiles fastest? No: x = output-channel tile so that
  // CTAs launched together share the token tile (L2-resident) and stream disjoint weights.
  const int tile_ch = blockIdx.x, tile_tok = blockIdx.y;
  const int p0 = (kSwap ? tile_ch : tile_tok) * C::BM;   // first row of the MMA-M operand
  const int q0 = (kSwap ? tile_tok : tile_ch) * BN;      // first row of the MMA-N operand
  const int m0 = kSwap ? q0 : p0, n0 = kSwap ? p0 : q0;  // token / channel origin of the tile

  // K split over the cluster: groups [g_begin, g_end) of the G+1 groups (index G = INT8 keeper)
  const int total_groups = args.G + 1;
  int g_begin = 0, g_end = total_groups;
  uint32_t krank = 0;
  if constexpr (kSplit > 1) {
    krank = cluster_ctarank();
    const int per = (total_groups + kSplit - 1) / kSplit;
    g_begin = min((int)krank * per, total_groups);
    g_end = min(g_begin + per, total_groups);
  }
  const int iters = g_end - g_begin;

  // ---------------------------------------------------------------- one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_p4); tma_prefetch_desc(&tm_q4); tma_prefetch_desc(&tm_p8); tma_prefetch_desc(&tm_q8);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPack; ++i) { mbar_init(&pack_full[i], 1); mbar_init(&pack_empty[i], 4); }
    for (int i = 0; i < kExp; ++i) { mbar_init(&exp_full[i], 4); mbar_init(&exp_empty[i], 1); }
    for (int i = 0; i < C::NB; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 * C::EPI_WGS); }
    for (int i = 0; i < C::SCALE_SLOTS; ++i) mbar_init(&scale_full[i], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============================================================ TMA producer (INT4 groups only)
    if (lane == 0) {
      int itp = 0;
      for (int it = 0; it < iters; ++it) {
        const int g = g_begin + it;
        if (g >= args.G) break;  // keeper is loaded by the converter straight into the expanded ring
        const int s = itp % kPack;
        mbar_wait(&pack_empty[s], ((itp / kPack) & 1) ^ 1);
        mbar_arrive_expect_tx(&pack_full[s], C::PACK_P + C::PACK_Q);
        tma_load_2d(smem + C::OFF_PACK_P + s * C::PACK_P, &tm_p4, &pack_full[s], g * 64, p0);
        tma_load_2d(smem + C::OFF_PACK_Q + s * C::PACK_Q, &tm_q4, &pack_full[s], g * 64, q0);
        ++itp;
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_i8(C::BM, BN);
      for (int it = 0; it < iters; ++it) {
        const int e = it % kExp, b = it % C::NB;
        mbar_wait(&tmem_empty[b], ((it / C::NB) & 1) ^ 1);
        mbar_wait(&exp_full[e], (it / kExp) & 1);
        tc_fence_after();
        const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_P + e * C::EXP_P));
        const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_Q + e * ((C::EXP_Q + 1023) / 1024 * 1024)));
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x K=32 per 128-wide group; +32 B inside the swizzle atom per step
          umma_i8(tmem_base + b * BN, dp + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc, k > 0);
        umma_commit(&exp_empty[e]);   // smem stage may be refilled once these MMAs have read it
        umma_commit(&tmem_full[b]);   // accumulator of this group is complete
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================================================ converter warpgroup
    const int t = threadIdx.x - 128;
    int itp = 0;
    for (int it = 0; it < iters; ++it) {
      const int g = g_begin + it;
      const bool keeper = (g == args.G);
      // scales of this group, fetched early so that their latency hides behind the barrier waits
      const __half* as_row = keeper ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
      const __half* bs_row = keeper ? args.b_keeper_scale : args.b_scale + (size_t)g * args.N;
      __half2 sm_val = __half2half2(__ushort_as_half(0));
      __half sn_val = __ushort_as_half(0);
      if constexpr (!kSwap) {
        if (m0 + t < args.M) sm_val = __half2half2(as_row[scale_index(m0 + t)]);
        if (t < BN && n0 + t < args.N) sn_val = bs_row[n0 + t];
      } else {
        const int n = n0 + t;   // reference column pairing: rows m%16<8 use sB[n&~1], others sB[n|1]
        if (n < args.N) sm_val = __halves2half2(bs_row[n & ~1], bs_row[min(n | 1, args.N - 1)]);
        if (t < BN && m0 + t < args.M) sn_val = as_row[scale_index(m0 + t)];
      }
      const int e = it % kExp;
      uint8_t* exp_p = smem + C::OFF_EXP_P + e * C::EXP_P;
      uint8_t* exp_q = smem + C::OFF_EXP_Q + e * ((C::EXP_Q + 1023) / 1024 * 1024);
      mbar_wait(&exp_empty[e], ((it / kExp) & 1) ^ 1);
      const int slot = it % C::SCALE_SLOTS;
      reinterpret_cast<__half2*>(smem + C::OFF_SM)[slot * C::BM + t] = sm_val;
      if (t < BN) reinterpret_cast<__half*>(smem + C::OFF_SN)[slot * BN + t] = sn_val;
      if (!keeper) {
        const int s = itp % kPack;
        mbar_wait(&pack_full[s], (itp / kPack) & 1);
        convert_tile<C::BM>(smem + C::OFF_PACK_P + s * C::PACK_P, exp_p, t);
        convert_tile<BN>(smem + C::OFF_PACK_Q + s * C::PACK_Q, exp_q, t);
        fence_proxy_async_smem();     // generic-proxy stores -> visible to tcgen05.mma operand fetch
        __syncwarp();
        if (lane == 0) { mbar_arrive(&pack_empty[s]); mbar_arrive(&scale_full[slot]); mbar_arrive(&exp_full[e]); }
        ++itp;
      } else {
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&scale_full[slot]);
          if (t == 0) {
            mbar_arrive_expect_tx(&exp_full[e], C::EXP_P + C::EXP_Q);
            tma_load_2d(exp_p, &tm_p8, &exp_full[e], 0, p0);
            tma_load_2d(exp_q, &tm_q8, &exp_full[e], 0, q0);
          } else {
            mbar_arrive(&exp_full[e]);
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ============================================================ epilogue warpgroup(s)
    const int wq = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;                // MMA-M row == TMEM lane
    const int colbase = ((warp - 8) >> 2) * C::CPT;
    float acc[C::CPT];
#pragma unroll
    for (int i = 0; i < C::CPT; ++i) acc[i] = 0.f;
    const bool upper = kSwap ? false : (((m0 + row) & 15) >= 8);

    for (int it = 0; it < iters; ++it) {
      const int g = g_begin + it;
      const bool keeper = (g == args.G);
      const int slot = it % C::SCALE_SLOTS, b = it % C::NB;
      mbar_wait(&scale_full[slot], (it / C::SCALE_SLOTS) & 1);
      const __half2 sm2 = reinterpret_cast<const __half2*>(smem + C::OFF_SM)[slot * C::BM + row];
      const __half* sn = reinterpret_cast<const __half*>(smem + C::OFF_SN) + slot * BN + colbase;
      mbar_wait(&tmem_full[b], (it / C::NB) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * BN + colbase);
      constexpr int CH = (C::CPT >= 32) ? 32 : 16;
#pragma unroll
      for (int c0 = 0; c0 < C::CPT; c0 += CH) {
        uint32_t r[CH];
        if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, r); else tmem_ld_32x32b_x16(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + CH == C::CPT) {      // whole accumulator buffer is in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[b]);
        }
        // INT4 groups carry a factor 256 (both operands are value*16); lift the keeper to the same domain
        if (keeper) {
#pragma unroll
          for (int j = 0; j < CH; ++j) r[j] = (uint32_t)((int32_t)r[j] << 8);
        }
        if constexpr (!kSwap) {
          // thread = token row; columns = channels.  rs is shared by the column pair (2p, 2p+1).
#pragma unroll
          for (int p = 0; p < CH / 2; p += 2) {
            const __half2 n01 = *reinterpret_cast<const __half2*>(sn + c0 + 2 * p);
            const __half2 n23 = *reinterpret_cast<const __half2*>(sn + c0 + 2 * p + 2);
            const __half2 sel = upper ? __halves2half2(__high2half(n01), __high2half(n23))
                                      : __halves2half2(__low2half(n01), __low2half(n23));
            const float2 rs = __half22float2(__hmul2(sm2, sel));
            acc[c0 + 2 * p + 0] = fmaf((float)(int32_t)r[2 * p + 0], rs.x, acc[c0 + 2 * p + 0]);
            acc[c0 + 2 * p + 1] = fmaf((float)(int32_t)r[2 * p + 1], rs.x, acc[c0 + 2 * p + 1]);
            acc[c0 + 2 * p + 2] = fmaf((float)(int32_t)r[2 * p + 2], rs.y, acc[c0 + 2 * p + 2]);
            acc[c0 + 2 * p + 3] = fmaf((float)(int32_t)r[2 * p + 3], rs.y, acc[c0 + 2 * p + 3]);
          }
        } else {
          // thread = channel row (sm2 = {sB[n&~1], sB[n|1]}); columns = tokens, 16-aligned tile origin:
          // token j with j%16<8 pairs with sm2.x, j%16>=8 with sm2.y
#pragma unroll
          for (int j = 0; j < CH; j += 16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 a2 = __halves2half2(sn[c0 + j + i], sn[c0 + j + i + 8]);
              const float2 rs = __half22float2(__hmul2(a2, sm2));
              acc[c0 + j + i] = fmaf((float)(int32_t)r[j + i], rs.x, acc[c0 + j + i]);
              acc[c0 + j + i + 8] = fmaf((float)(int32_t)r[j + i + 8], rs.y, acc[c0 + j + i + 8]);
            }
          }
        }
      }
    }

    // ------------------------------------------------------------ split-K: publish partials in smem
    if constexpr (kSplit > 1) {
      float* red = reinterpret_cast<float*>(smem);   // [col][row], aliases the (now idle) expanded ring
#pragma unroll
      for (int i = 0; i < C::CPT; ++i) red[(colbase + i) * C::BM + row] = acc[i];
    }
    if constexpr (kSplit > 1) {
      cluster_arrive(); cluster_wait();
      if (krank == 0) {
        const uint32_t red_local = smem_u32(smem);
#pragma unroll 1
        for (uint32_t rk = 1; rk < (uint32_t)kSplit; ++rk) {
          const uint32_t remote = mapa_shared(red_local, rk);
#pragma unroll
          for (int i = 0; i < C::CPT; ++i) acc[i] += ld_dsmem_f32(remote + ((colbase + i) * C::BM + row) * 4);
        }
      }
    }

    // ------------------------------------------------------------ output
    if (kSplit == 1 || krank == 0) {
      constexpr float kInv = 1.0f / 256.0f;   // exact: removes the 16*16 operand factor
      if constexpr (!kO4) {
        if constexpr (!kSwap) {
          const int m = m0 + row;
          if (m < args.M) {
            __half* drow = args.d + (size_t)m * args.N + n0 + colbase;
#pragma unroll
            for (int i = 0; i < C::CPT; i += 8) {
              if (n0 + colbase + i < args.N) {   // N is a multiple of 8 (16-B rows)
                uint4 v;
                __half2 h0 = __floats2half2_rn(acc[i + 0] * kInv, acc[i + 1] * kInv);
                __half2 h1 = __floats2half2_rn(acc[i + 2] * kInv, acc[i + 3] * kInv);
                __half2 h2 = __floats2half2_rn(acc[i + 4] * kInv, acc[i + 5] * kInv);
                __half2 h3 = __floats2half2_rn(acc[i + 6] * kInv, acc[i + 7] * kInv);
                v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
                v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(drow + i) = v;
              }
            }
          }
        } else {
          const int n = n0 + row;
          if (n < args.N) {
#pragma unroll
            for (int i = 0; i < C::CPT; ++i) {
              const int m = m0 + colbase + i;
              if (m < args.M) args.d[(size_t)m * args.N + n] = __float2half_rn(acc[i] * kInv);   // warp writes 64 B runs
            }
          }
        }
      } else {
        // o4 epilogue (DenseLayerGEMM_i4_o4.cu:705-787): per (token, 128-channel head) asymmetric INT4 with the
        // reference's |v| min/max.  tall: two epilogue WGs hold 64 columns each of the same row.
        float* xch = reinterpret_cast<float*>(smem + C::OFF_PACK_P);   // packed ring is idle by now
        if constexpr (!kSwap) {
          float mx = -INFINITY, mn = INFINITY;
#pragma unroll
          for (int i = 0; i < C::CPT; ++i) { acc[i] *= kInv; const float a = fabsf(acc[i]); mx = fmaxf(mx, a); mn = fminf(mn, a); }
          const int half_id = (warp - 8) >> 2;
          xch[(half_id * 2 + 0) * C::BM + row] = mx;
          xch[(half_id * 2 + 1) * C::BM + row] = mn;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(mx, xch[((half_id ^ 1) * 2 + 0) * C::BM + row]);
          mn = fminf(mn, xch[((half_id ^ 1) * 2 + 1) * C::BM + row]);
          const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
          const int m = m0 + row;
          if (m < args.M) {
            if (half_id == 0) args.d_scale[(size_t)m * (args.N / 128) + tile_ch] = __floats2half2_rn(scale, zero);
            uint32_t pk[C::CPT / 8];
#pragma unroll
            for (int i = 0; i < C::CPT; i += 8) {
              uint32_t w = 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) w |= ((uint32_t)((int)roundf((acc[i + j] + zero) * r_scale) & 0xF)) << (4 * j);
              pk[i / 8] = w;
            }
            uint4* dst = reinterpret_cast<uint4*>(args.d4 + (size_t)m * (args.N / 2) + (n0 + colbase) / 2);
#pragma unroll
            for (int i = 0; i < C::CPT / 32; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        } else {
          // skinny: thread = channel; reduce |v| min/max over the 128 channels of the head for every token column
          float* xmx = xch;                 // [4 warps][BN]
          float* xmn = xch + 4 * BN;
          const int ewarp = warp - 8;       // EPI_WGS == 1 or 2; with 2 WGs each owns CPT token columns
#pragma unroll
          for (int i = 0; i < C::CPT; ++i) {
            acc[i] *= kInv;
            float a = fabsf(acc[i]), mx = a, mn = a;
            if (n0 + row >= args.N) { mx = -INFINITY; mn = INFINITY; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
            if (lane == 0) { xmx[(ewarp & 3) * BN + colbase + i] = mx; xmn[(ewarp & 3) * BN + colbase + i] = mn; }
          }
          if constexpr (C::EPI_WGS == 2) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 1, 128;" ::: "memory");
          const int n = n0 + row;
#pragma unroll
          for (int i = 0; i < C::CPT; ++i) {
            const int c = colbase + i, m = m0 + c;
            const float mx = fmaxf(fmaxf(xmx[c], xmx[BN + c]), fmaxf(xmx[2 * BN + c], xmx[3 * BN + c]));
            const float mn = fminf(fminf(xmn[c], xmn[BN + c]), fminf(xmn[2 * BN + c], xmn[3 * BN + c]));
            const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
            int q = (int)roundf((acc[i] + zero) * r_scale) & 0xF;
            const int qn = __shfl_down_sync(0xffffffffu, q, 1);     // channel n+1 lives in the next lane
            if (m < args.M && n < args.N) {
              if ((lane & 1) == 0) args.d4[(size_t)m * (args.N / 2) + n / 2] = (uint8_t)(q | (qn << 4));
              if (row == 0) args.d_scale[(size_t)m * (args.N / 128) + tile_ch] = __floats2half2_rn(scale, zero);
            }
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  if constexpr (kSplit > 1) {
    if (warp < 8) { cluster_arrive(); cluster_wait(); }   // pairs with the epilogue's first cluster barrier
    cluster_arrive(); cluster_wait();                     // nobody exits while its smem may still be read remotely
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

}  // namespace atom
