// gemm_i4_skinny_sm100.cuh -- decode-shape (M <= 64 tokens) W4A4 GEMM for B200: weights are the MMA-M operand and are
// expanded from packed INT4 straight into TENSOR MEMORY; two CTAs share an SM so that consecutive launches overlap.
//
// Same contract as gemm_i4_sm100.cuh (replaces compute_gemm_imma / DenseLayerGEMM_i4[_o4]_kernel,
// /root/reference/kernels/include/GEMM/Dense_layer_gemm_i4_o16.cuh:436-710 and
// /root/reference/e2e/punica-atom/punica/ops/csrc/GEMM/DenseLayerGEMM_i4_o4.cu:440-787): exact INT32 group sums, one
// FP16 multiply of the two scales, FP32 fma accumulation in group order, keeper last, RN cast to half.
//
// A decode GEMM is a weight stream: 9 MB at M=16, N=K=4096 against 0.5 GOP.  What bounded the first design
// (profiles/r01_gemm_pipeline_trace_final.jsonl) was not bandwidth but the latency chain
//   TMA -> converter (LDS, expand, STS, proxy fence) -> MMA -> commit -> converter may reuse the 16 KB operand slot
// with only two operand slots per CTA and one CTA per SM (150 KB of shared memory, most of it expanded INT8 weights).
// Here:
//   * TMA lands packed INT4 weight tiles (64 B rows, SWIZZLE_64B so that a warp reading one 16-B chunk per ROW is
//     conflict free) in a 6-deep ring;
//   * converter thread r owns weight row r of the tile: 4 x LDS.128 -> 48 ALU ops -> ONE tcgen05.st (32 columns) puts
//     the row's 128 INT8 values (value * 16) into a tensor-memory operand slot -- no shared-memory store, no swizzle
//     arithmetic, no generic->async proxy fence for the big operand, and half the shared-memory traffic;
//   * tcgen05.mma.kind::i8 reads A from tensor memory (TS form) and the (tiny) token operand from shared memory;
//   * shared memory per CTA drops to ~90 KB and tensor memory to 256 columns, so TWO CTAs are resident per SM: with
//     programmatic dependent launch the next GEMM's CTAs start streaming their weights while this one drains its
//     epilogue (griddepcontrol: weights do not depend on the preceding kernel, activations and the output do);
//   * K may be split over a cluster of 2 or 4 CTAs; every rank reduces and stores its own slice of the token columns
//     (partials pushed through DSMEM, summed in rank order: deterministic).
#pragma once
#include "gemm_i4_sm100.cuh"

namespace atom {

enum { EPI_O16 = 0, EPI_O4 = 1 };

template <int BN, int kSplit, int kEpi>
struct SkinnyCfg {
  static constexpr int BM = 128;                       // weight rows per tile = TMEM lanes
  static constexpr int A_RING = 4;                     // tensor-memory operand slots, 32 columns (one group) each
  static constexpr int ACC = 128 / BN;                 // accumulator slots of BN columns: 8 / 4 / 2
  static constexpr int NB = A_RING > ACC ? A_RING : ACC;   // "group's MMAs completed" barriers
  static constexpr int A_COL0 = 0, ACC_COL0 = A_RING * 32;
  static constexpr int TMEM_COLS = 256;
  static constexpr int PACK = BN == 16 ? 6 : (BN == 32 ? 5 : 4);   // packed ring depth (groups)
  static constexpr int SCALE_SLOTS = 8;
  static constexpr int THREADS = 384;                  // 4 service warps, 4 converter warps, 4 epilogue warps
  static constexpr int PACK_P = BM * 64, PACK_Q = BN * 64;
  static constexpr int EXP_Q = (BN * 128 + 1023) / 1024 * 1024;
  static constexpr int CPR = BN / kSplit;              // token columns reduced + stored by one split-K rank
  static constexpr int OFF_PACK_P = 0;
  static constexpr int OFF_EXP_Q = OFF_PACK_P + PACK * PACK_P;
  static constexpr int OFF_KEEP_P = OFF_EXP_Q + A_RING * EXP_Q;
  static constexpr int OFF_KEEP_Q = OFF_KEEP_P + BM * 128;
  static constexpr int OFF_PACK_Q = OFF_KEEP_Q + EXP_Q;
  static constexpr int OFF_SM = OFF_PACK_Q + PACK * PACK_Q;
  static constexpr int RED_BYTES = BM * CPR * 4;       // one source rank's partial for this rank's columns
  static constexpr int OFF_RED = OFF_SM + SCALE_SLOTS * 512;
  static constexpr int OFF_XCH = OFF_RED + (kSplit > 1 ? (kSplit - 1) * RED_BYTES : 0);   // o4: per-warp |v| min/max
  static constexpr int OFF_BAR = OFF_XCH + (kEpi == EPI_O4 ? 8 * BN * 4 : 0);
  static constexpr int NUM_BARS = 2 * PACK + A_RING + NB + ACC + 2 * SCALE_SLOTS + 1;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
  static constexpr int CTAS_PER_SM = SMEM_BYTES <= 112 * 1024 ? 2 : 1;
  static_assert(BN == 16 || BN == 32 || BN == 64, "token tile");
  static_assert(BN % kSplit == 0 && CPR >= 4, "every split-K rank owns at least 4 token columns");
  static_assert(kEpi == EPI_O16 || kSplit == 1, "the INT4-output epilogue quantises un-split FP32 sums");
  static_assert(A_RING * 32 + ACC * BN <= TMEM_COLS, "tensor memory budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// tcgen05.mma, A operand in tensor memory, B through a shared-memory descriptor
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_dsmem_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int kSplit, int kEpi>
__global__ void __launch_bounds__(SkinnyCfg<BN, kSplit, kEpi>::THREADS, SkinnyCfg<BN, kSplit, kEpi>::CTAS_PER_SM)
gemm_i4_skinny_kernel(const __grid_constant__ CUtensorMap tm_p4,   // packed INT4 weights   (box 64 B x 128 rows, SWIZZLE_64B)
                      const __grid_constant__ CUtensorMap tm_q4,   // packed INT4 tokens    (box 64 B x BN rows)
                      const __grid_constant__ CUtensorMap tm_p8,   // INT8 keeper weights   (box 128 B x 128 rows, SWIZZLE_128B)
                      const __grid_constant__ CUtensorMap tm_q8,   // INT8 keeper tokens    (box 128 B x BN rows, SWIZZLE_128B)
                      const GemmArgs args) {
  using C = SkinnyCfg<BN, kSplit, kEpi>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;                           // TMA landed a group's packed tiles                 (1 + tx)
  uint64_t* pack_empty = pack_full + C::PACK;           // the converter warps have read them                (4)
  uint64_t* a_full = pack_empty + C::PACK;              // operand slot written: TMEM weights + smem tokens  (4)
  uint64_t* mma_done = a_full + C::A_RING;              // the group's MMAs completed (tcgen05.commit)       (1)
  uint64_t* acc_empty = mma_done + C::NB;               // epilogue has read the accumulator slot            (4)
  uint64_t* scale_full = acc_empty + C::ACC;            // the group's scales landed (cp.async, no-inc)      (32)
  uint64_t* scale_empty = scale_full + C::SCALE_SLOTS;  //                                                   (4)
  uint64_t* keep_full = scale_empty + C::SCALE_SLOTS;   // INT8 keeper operands landed                       (1 + tx)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * C::BM;        // first weight row (output channel) of the tile
  const int q0 = blockIdx.y * BN;           // first token
  const int m0 = q0, n0 = p0;

  const int total_groups = args.G + 1;      // index G = INT8 keeper
  int g_begin = 0, g_end = total_groups;
  uint32_t krank = 0;
  if constexpr (kSplit > 1) {
    krank = cluster_ctarank();
    const int per = (total_groups + kSplit - 1) / kSplit;
    g_begin = min((int)krank * per, total_groups);
    g_end = min(g_begin + per, total_groups);
  }
  const int iters = g_end - g_begin;                                  // groups of this CTA, the keeper (if any) last
  const int n4 = max(0, min(g_end, args.G) - g_begin);                // ... of which INT4
  if (threadIdx.x == 0) { griddep_launch_dependents(); trace_stamp(args, 0); }

  // ---------------------------------------------------------------- setup
  // part 1 = weight tile (independent of the preceding kernel), part 2 = token tile
  auto issue_group = [&](int i, int part) {
    const int ps = i % C::PACK, g = g_begin + i;
    if (part & 1) {
      mbar_arrive_expect_tx(&pack_full[ps], C::PACK_P + C::PACK_Q);
      tma_load_2d(smem + C::OFF_PACK_P + ps * C::PACK_P, &tm_p4, &pack_full[ps], g * 64, p0);
    }
    if (part & 2) tma_load_2d(smem + C::OFF_PACK_Q + ps * C::PACK_Q, &tm_q4, &pack_full[ps], g * 64, q0);
    if (i < 16 && (part & 2)) trace_stamp(args, 8 + i);
  };
  const int first = min(C::PACK, n4);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_p4); tma_prefetch_desc(&tm_q4);
    for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_full[i], 1);
    mbar_init(keep_full, 1);
    fence_barrier_init();
    // weights first: they do not depend on the preceding kernel, so under programmatic dependent launch they stream
    // (and are converted) while that kernel is still running
    for (int i = 0; i < first; ++i) issue_group(i, 1);
    if (n4 < iters) {                                                  // this rank owns the keeper group
      tma_prefetch_desc(&tm_p8); tma_prefetch_desc(&tm_q8);
      mbar_arrive_expect_tx(keep_full, C::BM * 128 + BN * 128);
      tma_load_2d(smem + C::OFF_KEEP_P, &tm_p8, keep_full, 0, p0);
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_empty[i], 4);
    for (int i = 0; i < C::A_RING; ++i) mbar_init(&a_full[i], 4);
    for (int i = 0; i < C::NB; ++i) mbar_init(&mma_done[i], 1);
    for (int i = 0; i < C::ACC; ++i) mbar_init(&acc_empty[i], 4);
    for (int i = 0; i < C::SCALE_SLOTS; ++i) { mbar_init(&scale_full[i], 32); mbar_init(&scale_empty[i], 4); }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (kSplit > 1) cluster_arrive_relaxed();                  // see the reduction: remote smem needs a running CTA
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) trace_stamp(args, 1);

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      griddep_wait();                                                  // activations are the preceding kernel's output
      for (int i = 0; i < first; ++i) issue_group(i, 2);
      if (n4 < iters) tma_load_2d(smem + C::OFF_KEEP_Q, &tm_q8, keep_full, 0, q0);
      for (int i = first; i < n4; ++i) {
        mbar_wait(&pack_empty[i % C::PACK], ((i / C::PACK) & 1) ^ 1);
        issue_group(i, 3);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_i8(C::BM, BN);
      for (int i = 0; i < iters; ++i) {
        const int as = i % C::ACC;
        if (i >= C::ACC) mbar_wait(&acc_empty[as], ((i / C::ACC) - 1) & 1);
        const uint32_t d_tmem = tmem_base + C::ACC_COL0 + as * BN;
        if (i < n4) {
          const int ar = i % C::A_RING;
          mbar_wait(&a_full[ar], (i / C::A_RING) & 1);
          tc_fence_after();
          if (i < 16) trace_stamp(args, 88 + i);
          const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_Q + ar * C::EXP_Q));
#pragma unroll
          for (int k = 0; k < 4; ++k)        // 4 x K=32: 8 tensor-memory columns of A, 32 B of each token row
            umma_i8_ts(d_tmem, tmem_base + C::A_COL0 + ar * 32 + k * 8, dq + (uint64_t)(k * 2), idesc, k > 0);
        } else {
          mbar_wait(keep_full, 0);
          tc_fence_after();
          if (i < 16) trace_stamp(args, 88 + i);
          const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + C::OFF_KEEP_P));
          const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_KEEP_Q));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_i8(d_tmem, dp + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc, k > 0);
        }
        umma_commit(&mma_done[i % C::NB]);   // accumulators ready AND operand slot reusable
      }
    }
  } else if (warp == 2) {
    // ============================================================ scale loader (cp.async into an 8-group ring)
    // slot: [0,256) the tile's 128 weight-scale halves (thread n reads the pair word n/2); [256, ...) the (lower, upper)
    // activation-scale words of the token rows, word = (r/16)*8 + r%8 (raw copy of the reference layout, Reorder.cuh:39-50)
    griddep_wait();
    for (int i = 0; i < iters; ++i) {
      const int ss = i % C::SCALE_SLOTS, g = g_begin + i;
      if (i >= C::SCALE_SLOTS) mbar_wait(&scale_empty[ss], ((i / C::SCALE_SLOTS) - 1) & 1);
      const bool keeper = (g == args.G);
      const __half* as_row = keeper ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
      const __half* bs_row = keeper ? args.b_keeper_scale : args.b_scale + (size_t)g * args.N;
      uint8_t* slot = smem + C::OFF_SM + ss * 512;
#pragma unroll
      for (int w = lane; w < BN / 2; w += 32) {
        const int blk = w >> 3, r = w & 7;
        if (m0 + 16 * blk + r < args.M) cp_async_4(slot + 256 + w * 4, as_row + 64 * (m0 / 16 + blk) + 8 * r);
      }
      if (lane < 16 && n0 + 8 * lane < args.N) cp_async_16(slot + lane * 16, bs_row + n0 + 8 * lane);
      cp_async_mbar_arrive_noinc(&scale_full[ss]);
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================================================ converters: thread = weight row
    const int wq = warp & 3, row = wq * 32 + lane, t = (warp - 4) * 32 + lane;
    const int xr = (row >> 1) & 3;                       // SWIZZLE_64B: 16-B chunk index ^= address bits [7,9)
    for (int i = 0; i < n4; ++i) {
      const int ps = i % C::PACK, ar = i % C::A_RING;
      if (i >= C::A_RING) mbar_wait(&mma_done[(i - C::A_RING) % C::NB], ((i - C::A_RING) / C::NB) & 1);
      if (t == 0 && i < 16) trace_stamp(args, 24 + i);
      mbar_wait(&pack_full[ps], (i / C::PACK) & 1);
      if (t == 0 && i < 16) trace_stamp(args, 40 + i);
      const uint8_t* prow = smem + C::OFF_PACK_P + ps * C::PACK_P + row * 64;
      uint4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const uint4*>(prow + ((j ^ xr) << 4));
      // token tile -> canonical K-major SWIZZLE_128B INT8 operand in shared memory (same K permutation as below)
      convert_tile<BN, 128>(smem + C::OFF_PACK_Q + ps * C::PACK_Q, smem + C::OFF_EXP_Q + ar * C::EXP_Q, t);
      // chunk j (32 consecutive K) -> columns 8j..8j+7: four "even element" words, then four "odd element" words
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 lo, hi;
        expand_chunk(w[j], lo, hi);
        r[8 * j + 0] = lo.x; r[8 * j + 1] = lo.y; r[8 * j + 2] = lo.z; r[8 * j + 3] = lo.w;
        r[8 * j + 4] = hi.x; r[8 * j + 5] = hi.y; r[8 * j + 6] = hi.z; r[8 * j + 7] = hi.w;
      }
      tmem_st_32x32b_x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::A_COL0 + ar * 32), r);
      tmem_st_wait();
      if (t == 0 && i < 16) trace_stamp(args, 56 + i);
      fence_proxy_async_smem();          // the token tile's generic-proxy stores -> visible to the MMA's operand fetch
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&pack_empty[ps]); mbar_arrive(&a_full[ar]); }
      if (t == 0 && i < 16) trace_stamp(args, 72 + i);
    }
  } else if (warp >= 8) {
    // ============================================================ epilogue: thread = output channel (TMEM lane)
    const int wq = warp & 3, row = wq * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int c = 0; c < BN; ++c) acc[c] = 0.f;

    for (int i = 0; i < iters; ++i) {
      const int as = i % C::ACC, ss = i % C::SCALE_SLOTS;
      const bool keeper = (g_begin + i == args.G);
      mbar_wait(&scale_full[ss], (i / C::SCALE_SLOTS) & 1);
      mbar_wait(&mma_done[i % C::NB], (i / C::NB) & 1);
      tc_fence_after();
      if (warp == 8 && lane == 0 && i < 16) trace_stamp(args, 104 + i);
      const uint8_t* slot = smem + C::OFF_SM + ss * 512;
      const __half2 sm2 = reinterpret_cast<const __half2*>(slot)[row >> 1];          // {sB[n & ~1], sB[n | 1]}
      const __half2* snw = reinterpret_cast<const __half2*>(slot + 256);             // (sA[tok], sA[tok + 8]) words
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::ACC_COL0 + as * BN);
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t r[CH];
        if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, r); else tmem_ld_32x32b_x16(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + CH == BN) {               // the group's accumulators are in registers: hand the slot back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (keeper) {                       // INT4 groups carry 16 * 16 = 256; lift the keeper to the same domain
#pragma unroll
          for (int e = 0; e < CH; ++e) r[e] = (uint32_t)((int32_t)r[e] << 8);
        }
        // token c with c%16 < 8 pairs with sB[n & ~1], c%16 >= 8 with sB[n | 1] (the reference's column pairing,
        // Dense_layer_gemm_i4_o16.cuh:417-431)
#pragma unroll
        for (int q = 0; q < CH; q += 16) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const __half2 a2 = snw[((c0 + q) >> 4) * 8 + e];
            const float2 rs = __half22float2(__hmul2(a2, sm2));
            acc[c0 + q + e] = fmaf((float)(int32_t)r[q + e], rs.x, acc[c0 + q + e]);
            acc[c0 + q + e + 8] = fmaf((float)(int32_t)r[q + e + 8], rs.y, acc[c0 + q + e + 8]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&scale_empty[ss]);
      if (warp == 8 && lane == 0 && i < 8) trace_stamp(args, 120 + i);
    }
    if (warp == 8 && lane == 0) trace_stamp(args, 2);
    constexpr float kInv = 1.0f / 256.0f;   // exact: removes the 16 * 16 operand factor
    griddep_wait();                          // the output buffer may still be read by the preceding kernel

    if constexpr (kEpi == EPI_O16) {
      // ---------------------------------------------------------- split-K: rank d reduces + stores token columns
      // [d * CPR, (d + 1) * CPR).  Partials are PUSHED (no DSMEM read latency on the critical path), published by one
      // cluster barrier and summed in rank order 0..kSplit-1, so the result does not depend on arrival order.
      if constexpr (kSplit > 1) {
        cluster_wait();                      // pairs with the setup arrive: every CTA of the cluster is running
#pragma unroll
        for (int d = 0; d < kSplit; ++d) {
          if (d != (int)krank) {
            const int slot_in_dst = (int)krank < d ? (int)krank : (int)krank - 1;
            const uint32_t remote = mapa_shared(smem_u32(smem + C::OFF_RED), d) + slot_in_dst * C::RED_BYTES;
#pragma unroll
            for (int c = 0; c < C::CPR; ++c) st_dsmem_f32(remote + (c * C::BM + row) * 4, acc[d * C::CPR + c]);
          }
        }
        cluster_arrive(); cluster_wait();
        const float* red = reinterpret_cast<const float*>(smem + C::OFF_RED);
#pragma unroll
        for (int d = 0; d < kSplit; ++d) {
          if (d == (int)krank) {
#pragma unroll
            for (int c = 0; c < C::CPR; ++c) {
              float s = 0.f;
#pragma unroll
              for (int rk = 0; rk < kSplit; ++rk) {
                const float v = rk == d ? acc[d * C::CPR + c] : red[(rk < d ? rk : rk - 1) * (C::RED_BYTES / 4) + c * C::BM + row];
                s = rk == 0 ? v : s + v;
              }
              acc[d * C::CPR + c] = s;
            }
          }
        }
      }
      if (warp == 8 && lane == 0) trace_stamp(args, 3);
      const int n = n0 + row;
      if (n < args.N) {
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          const int m = m0 + c;
          if ((kSplit == 1 || c / C::CPR == (int)krank) && m < args.M)
            args.d[(size_t)m * args.N + n] = __float2half_rn(acc[c] * kInv);      // a warp writes 64-B runs
        }
      }
    } else {
      // ---------------------------------------------------------- o4 (DenseLayerGEMM_i4_o4.cu:705-787): per (token,
      // 128-channel head) asymmetric INT4 with the reference's |v| min/max; thread = channel, reduce over the 128 rows
      float* xmx = reinterpret_cast<float*>(smem + C::OFF_XCH);   // [4 warps][BN]
      float* xmn = xmx + 4 * BN;
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        acc[c] *= kInv;
        // |v| >= 0: IEEE bit patterns order like unsigned integers, so one REDUX each replaces 5 shuffle rounds
        uint32_t ua = __float_as_uint(fabsf(acc[c])), umx = ua, umn = ua;
        if (n0 + row >= args.N) { umx = 0u; umn = 0x7f800000u; }
        umx = __reduce_max_sync(0xffffffffu, umx);
        umn = __reduce_min_sync(0xffffffffu, umn);
        if (lane == 0) { xmx[wq * BN + c] = __uint_as_float(umx); xmn[wq * BN + c] = __uint_as_float(umn); }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int n = n0 + row;
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        const int m = m0 + c;
        const float mx = fmaxf(fmaxf(xmx[c], xmx[BN + c]), fmaxf(xmx[2 * BN + c], xmx[3 * BN + c]));
        const float mn = fminf(fminf(xmn[c], xmn[BN + c]), fminf(xmn[2 * BN + c], xmn[3 * BN + c]));
        const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
        const int q = (int)roundf((acc[c] + zero) * r_scale) & 0xF;
        const int qn = __shfl_down_sync(0xffffffffu, q, 1);     // channel n+1 lives in the next lane
        if (m < args.M && n < args.N) {
          if ((lane & 1) == 0) args.d4[(size_t)m * (args.N / 2) + n / 2] = (uint8_t)(q | (qn << 4));
          if (row == 0) args.d_scale[(size_t)m * (args.N / 128) + blockIdx.x] = __floats2half2_rn(scale, zero);
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  if constexpr (kSplit > 1) {
    if (warp < 8) { cluster_wait(); cluster_arrive(); cluster_wait(); }   // setup phase, then the epilogue's publish phase
    // no CTA may exit while a peer can still push into its shared memory: the publish barrier above is that guarantee
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_stamp(args, 4);
}

}  // namespace atom
