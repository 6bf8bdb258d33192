// gemm_i4_skinny_sm100.cuh -- decode-shape (M <= 64 tokens) W4A4 GEMM for B200: weights are the MMA-M operand and are
// expanded from packed INT4 straight into TENSOR MEMORY; two CTAs share an SM so that consecutive launches overlap.
//
// Same contract as gemm_i4_sm100.cuh (replaces compute_gemm_imma / DenseLayerGEMM_i4[_o4]_kernel,
// /root/reference/kernels/include/GEMM/Dense_layer_gemm_i4_o16.cuh:436-710 and
// /root/reference/e2e/punica-atom/punica/ops/csrc/GEMM/DenseLayerGEMM_i4_o4.cu:440-787): exact INT32 group sums, one
// FP16 multiply of the two scales, FP32 fma accumulation in group order, keeper last, RN cast to half.
//
// A decode GEMM is a weight stream (9 MB at M=16, N=K=4096 against 0.5 GOP) followed by a short dependent tail.
// Measured (profiles/r01_gemm_pipeline_trace_final.jsonl, r02_gemm_skinny_v2_pipeline_trace.jsonl): the stream itself is
// cheap; what costs is (1) everything that is serialised behind the arrival of the ACTIVATIONS -- they are the output of
// the kernel in front, so under programmatic dependent launch that part cannot overlap it -- and (2) fixed per-CTA
// latencies.  Hence two decoupled pipelines:
//
//   WEIGHTS (no dependency on the preceding kernel; start at CTA launch)
//     TMA: packed INT4 tiles (64 B rows, SWIZZLE_64B => a warp reading one 16-B chunk per ROW is conflict free), 6-deep ring
//     converter thread r owns weight row r: 4 x LDS.128 -> 48 ALU ops -> ONE tcgen05.st (32 columns) puts the row's 128
//       INT8 values (value * 16) into a tensor-memory operand slot: no shared-memory store, no swizzle arithmetic, no
//       proxy fence, half the shared-memory traffic.  Up to 6 groups wait converted in tensor memory, 6 more packed in smem.
//   ACTIVATIONS (after griddepcontrol.wait)
//     two warps read the packed token tile straight from global memory (L2: it was just written), expand it into the
//       canonical K-major SWIZZLE_128B INT8 operand in shared memory, four groups per hand-off;
//     the epilogue warps stage the scale rows themselves (no loader warp, no scale ring);
//     tcgen05.mma.kind::i8, A from tensor memory (TS form), B = tokens from shared memory; INT32 accumulators in TMEM;
//     epilogue: tcgen05.ld -> acc = fmaf(float(c), float(hmul(sA, sB)), acc); split-K partials go to the rank that owns
//       the token columns with st.async (data + mbarrier complete_tx in one message: no cluster barrier on the path),
//       are summed in rank order (deterministic) and stored by that rank.
//   Shared memory ~100 KB and 256 tensor-memory columns per CTA => TWO CTAs per SM: the next GEMM's CTAs stream and
//   convert their weights while this one waits for its activations or drains its tail.
#pragma once
#include "gemm_i4_sm100.cuh"

namespace atom {

// EPI_QKV: fused q/k/v projection (o16 for the q tiles, o4 for the k and v tiles).
// EPI_GATEUP: fused gate/up projection: a cluster of two CTAs (rank 0 = the gate tile, rank 1 = the up tile of the same 128
// channels, each streaming the whole K range); rank 1 hands its FP32 sums to rank 0 through DSMEM, rank 0 applies
// SiLU(gate) * up and the dynamic per-(token, 128-channel group) quantisation of activate_fp16_i4 and writes the INT4 / INT8
// operands of the down projection directly: no FP16 round trip through HBM, no activation kernel.
// EPI_PUSH: EPI_O16 whose result goes to every rank's all-reduce receive buffer (GemmArgs::ar; tp.py row-parallel projections).
enum { EPI_O16 = 0, EPI_O4 = 1, EPI_QKV = 2, EPI_GATEUP = 3, EPI_PUSH = 4 };

template <int BN, int kSplit, int kEpi>
struct SkinnyCfg {
  static constexpr int BM = 128;                       // weight rows per tile = TMEM lanes
  // One CTA per SM: 64 accumulators per thread need > 80 registers; and the fused q/k/v grid (3 H / 128 tiles: 96 at Llama-7B)
  // never fills 148 SMs twice, cannot split K (o4 tiles) and is bound by the bytes it keeps in flight: a 16-deep weight ring
  // instead of 6 (Little: 96 CTAs x 48 KB / ~1.5 us = 3 TB/s at best; measured 1.35 TB/s, 18.6 us for 25 MB).
  static constexpr bool ONE_PER_SM = BN == 64 || kEpi == EPI_QKV;
  static constexpr int TMEM_COLS = ONE_PER_SM ? 512 : 256;
  // The unit of every hand-off (converter -> MMA -> epilogue) is a PAIR of quantisation groups: a wait on an mbarrier
  // costs ~100 cycles even when the phase has completed and a tcgen05.commit ~200 cycles of the single issuing thread,
  // against 32 tensor cycles for one group at 16 tokens (r02_gemm_skinny_v3_pipeline_trace.jsonl: 760 cycles per group
  // in the MMA thread when every hand-off covered one group).
  static constexpr int A_PAIRS = ONE_PER_SM ? 4 : (BN == 16 ? 3 : 2);   // tensor-memory operand slots (2 groups = 64 columns each)
  static constexpr int ACC_PAIRS = 2;                  // accumulator slots (2 groups = 2 * BN columns each)
  static constexpr int A_COL0 = 0, ACC_COL0 = A_PAIRS * 64;
  static constexpr int PACK = (kEpi == EPI_QKV && BN <= 32) ? 16 : (ONE_PER_SM ? 8 : (BN == 16 ? 6 : (kEpi == EPI_GATEUP ? 4 : 5)));   // packed weight ring depth (groups)
  static constexpr int QB = 4, QS = 2 * QB;            // token tiles: groups per hand-off, expanded slots
  static constexpr int SC = BN == 16 ? 32 : 16;        // groups per staged scale chunk
  // 4 service warps, 4 converter warps, 4 epilogue warps; the one-CTA-per-SM instances add a second converter warpgroup
  // (warps 12-15, the unit's second group): a 33-group chain without a K split is paced by the converters (~650 cycles per
  // group with 4 warps, tools/microbench), and with one CTA per SM the registers for 16 warps are there
  static constexpr int CONV_WGS = ONE_PER_SM ? 2 : 1;
  static constexpr int THREADS = 384 + 128 * (CONV_WGS - 1);
  static constexpr int PACK_P = BM * 64;
  static constexpr int EXP_Q = (BN * 128 + 1023) / 1024 * 1024;
  static constexpr int CPR = kEpi == EPI_GATEUP ? BN : BN / kSplit;   // token columns reduced + stored by one split-K rank (gate/up: all of them go to rank 0)
  static constexpr int OFF_PACK_P = 0;
  static constexpr int OFF_EXP_Q = OFF_PACK_P + PACK * PACK_P;
  static constexpr int OFF_KEEP_P = OFF_EXP_Q + QS * EXP_Q;
  static constexpr int OFF_KEEP_Q = OFF_KEEP_P + BM * 128;
  static constexpr int OFF_SB = OFF_KEEP_Q + EXP_Q;                    // [SC][128] weight-scale halves
  static constexpr int OFF_SA = OFF_SB + SC * 256;                     // [SC][BN/2] (lower, upper) activation-scale words
  static constexpr int RED_BYTES = BM * CPR * 4;                       // one source rank's partial for this rank's columns
  static constexpr int OFF_RED = OFF_SA + SC * BN * 2;
  static constexpr int OFF_XCH = OFF_RED + (kSplit > 1 ? (kSplit - 1) * RED_BYTES : 0);   // o4: per-warp |v| min/max
  static constexpr int OFF_BAR = OFF_XCH + (kEpi != EPI_O16 && kEpi != EPI_PUSH ? 8 * BN * 4 : 0);     // gate/up uses the first 4 * BN floats
  static constexpr int NUM_BARS = 2 * PACK + 2 * A_PAIRS + 4 + ACC_PAIRS + 3;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
  static constexpr int CTAS_PER_SM = ONE_PER_SM ? 1 : 2;
  static_assert(BN == 16 || BN == 32 || BN == 64, "token tile");
  static_assert(BN % kSplit == 0 && CPR >= 2, "every split-K rank owns at least 2 token columns (one 8-byte st.async)");
  static_assert(kEpi == EPI_O16 || kEpi == EPI_PUSH || (kEpi == EPI_GATEUP && kSplit == 2) || kSplit == 1, "the quantising epilogues work on un-split FP32 sums");
  static_assert(A_PAIRS * 64 + ACC_PAIRS * 2 * BN <= TMEM_COLS, "tensor memory budget");
  static_assert(A_PAIRS >= ACC_PAIRS, "mma_done is indexed by operand slot");
  static_assert(ONE_PER_SM ? SMEM_BYTES <= 227 * 1024 : SMEM_BYTES <= 113 * 1024, "shared memory budget (two CTAs per SM)");
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
// 16 bytes into a peer CTA's shared memory; the same message completes 16 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar) : "memory");
}

__device__ __forceinline__ void st_async_v2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar) : "memory");
}

// 16 token columns of one group: acc[c] = fmaf(float(c_int), float(hmul(sA[c], sB)), acc[c]).  Token c with c%16 < 8 pairs
// with sB[n & ~1], c%16 >= 8 with sB[n | 1] (the reference's column pairing, Dense_layer_gemm_i4_o16.cuh:417-431).
__device__ __forceinline__ void dequant16(float* acc, const uint32_t* r, const __half2* snw, __half2 sm2, bool keeper) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 rs = __half22float2(__hmul2(snw[e], sm2));
    int32_t lo = (int32_t)r[e], hi = (int32_t)r[e + 8];
    if (keeper) { lo <<= 8; hi <<= 8; }        // INT4 groups carry 16 * 16 = 256; lift the keeper to the same domain
    acc[e] = fmaf((float)lo, rs.x, acc[e]);
    acc[e + 8] = fmaf((float)hi, rs.y, acc[e + 8]);
  }
}

template <int BN, int kSplit, int kEpi>
__global__ void __launch_bounds__(SkinnyCfg<BN, kSplit, kEpi>::THREADS, SkinnyCfg<BN, kSplit, kEpi>::CTAS_PER_SM)
gemm_i4_skinny_kernel(const __grid_constant__ CUtensorMap tm_p4,   // packed INT4 weights   (box 64 B x 128 rows, SWIZZLE_64B)
                      const __grid_constant__ CUtensorMap tm_p8,   // INT8 keeper weights   (box 128 B x 128 rows, SWIZZLE_128B)
                      const GemmArgs args) {
  using C = SkinnyCfg<BN, kSplit, kEpi>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;                           // TMA landed a group's packed weight tile            (1 + tx)
  uint64_t* pack_empty = pack_full + C::PACK;           // the converter warps have read it                   (4)
  uint64_t* a_full = pack_empty + C::PACK;              // tensor-memory operand slot (a PAIR of groups) written (4)
  uint64_t* mma_done = a_full + C::A_PAIRS;             // the pair's MMAs completed (tcgen05.commit)         (1)
  uint64_t* qx_full = mma_done + C::A_PAIRS;            // a batch of QB expanded token tiles is in place     (2)
  uint64_t* q_empty = qx_full + 2;                      // ... and has been consumed (tcgen05.commit)         (1)
  uint64_t* acc_empty = q_empty + 2;                    // epilogue has read the accumulator pair             (4)
  uint64_t* keep_full = acc_empty + C::ACC_PAIRS;       // INT8 keeper weights landed                         (1 + tx)
  uint64_t* kq_full = keep_full + 1;                    // INT8 keeper tokens copied                          (2)
  uint64_t* red_full = kq_full + 1;                     // split-K: every peer's partial has arrived          (1 + tx)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * C::BM;        // first weight row (output channel) of the tile
  const int m0 = blockIdx.y * BN;           // first token

  const int total_groups = args.G + 1;      // index G = INT8 keeper
  int g_begin = 0, g_end = total_groups;
  uint32_t krank = 0;
  if constexpr (kSplit > 1) {
    krank = cluster_ctarank();
    if constexpr (kEpi != EPI_GATEUP) {
      // balanced K ranges (33 groups over 8 ranks: 4,4,4,4,4,4,4,5 -- the cheap keeper group goes to the last rank)
      g_begin = (int)krank * total_groups / kSplit;
      g_end = ((int)krank + 1) * total_groups / kSplit;
    }
  }
  // gate/up: the two ranks are not K slices but the gate (rank 0) and up (rank 1) rows of the same channel tile
  const int wrow0 = n0 + (kEpi == EPI_GATEUP ? (int)krank * args.gu_rows : 0);    // first row of this CTA's weight tile
  const int iters = g_end - g_begin;                                  // groups of this CTA, the keeper (if any) last
  const int n4 = max(0, min(g_end, args.G) - g_begin);                // ... of which INT4
  const bool has_keeper = n4 < iters;
  // A unit = two consecutive groups of this CTA (the last unit may hold one); the keeper, always the CTA's last group, shares
  // a unit with the last INT4 group when their count is odd, so the rank that owns the keeper does not run an extra unit.
  const int np4 = (n4 + 1) >> 1;                                      // units that contain INT4 groups
  const int nu = (iters + 1) >> 1;
  if (threadIdx.x == 0) { griddep_launch_dependents(); trace_stamp(args, 0); }

  // ---------------------------------------------------------------- setup
  if (warp == 0) {
    // the weight stream starts before anything else exists: it depends on nothing.  One ELECTED lane of the converged
    // warp issues (uniform operands -> no per-lane broadcast loop around every TMA instruction)
    const int first = min(C::PACK, n4);
    if (elect_one_sync()) {
      tma_prefetch_desc(&tm_p4);
      for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_full[i], 1);
      mbar_init(keep_full, 1);
      fence_barrier_init();
      for (int i = 0; i < first; ++i) {
        mbar_arrive_expect_tx(&pack_full[i], C::PACK_P);
        tma_load_2d(smem + C::OFF_PACK_P + i * C::PACK_P, &tm_p4, &pack_full[i], (g_begin + i) * 64, wrow0);
        if (i < 8) trace_stamp(args, 8 + i);
      }
      if (has_keeper) {
        tma_prefetch_desc(&tm_p8);
        mbar_arrive_expect_tx(keep_full, C::BM * 128);
        tma_load_2d(smem + C::OFF_KEEP_P, &tm_p8, keep_full, 0, wrow0);
      }
    }
    __syncwarp();
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_empty[i], 4);
    for (int i = 0; i < C::A_PAIRS; ++i) { mbar_init(&a_full[i], 4 * C::CONV_WGS); mbar_init(&mma_done[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&qx_full[i], 2); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < C::ACC_PAIRS; ++i) mbar_init(&acc_empty[i], 4);
    mbar_init(kq_full, 2);
    mbar_init(red_full, 1);
    fence_barrier_init();
    if constexpr (kSplit > 1) mbar_arrive_expect_tx(red_full, (kSplit - 1) * C::RED_BYTES);   // (gate/up: only rank 0 ever waits on it)
  } else if (warp == 2) {
    tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (kSplit > 1) cluster_arrive_relaxed();                  // peers may write our smem once we are running
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) trace_stamp(args, 1);

  if (warp == 0) {
    // ============================================================ weight TMA producer (whole warp loops, one lane issues)
    for (int i = C::PACK; i < n4; ++i) {
      const int ps = i % C::PACK;
      mbar_wait(&pack_empty[ps], ((i / C::PACK) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&pack_full[ps], C::PACK_P);
        tma_load_2d(smem + C::OFF_PACK_P + ps * C::PACK_P, &tm_p4, &pack_full[ps], (g_begin + i) * 64, wrow0);
        if (i < 8) trace_stamp(args, 8 + i);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: the whole warp runs the loop (waits are
    // uniform), ONE elected lane issues the tcgen05 instructions; one commit per pair
    constexpr uint32_t idesc = umma_idesc_i8(C::BM, BN);
    for (int u = 0; u < nu; ++u) {
      const int as = u % C::ACC_PAIRS, ar = u % C::A_PAIRS, b = u >> 1;           // QB = 4 groups = 2 units per token batch
      const int ng = min(2, iters - 2 * u), ng4 = max(0, min(2, n4 - 2 * u));    // groups of the unit, of which INT4
      if (lane == 0 && u < 8) trace_stamp(args, 96 + u);
      if (u >= C::ACC_PAIRS) mbar_wait(&acc_empty[as], ((u / C::ACC_PAIRS) - 1) & 1);
      const uint32_t d0 = tmem_base + C::ACC_COL0 + as * 2 * BN;
      if (ng4 > 0) {
        if ((u & 1) == 0) mbar_wait(&qx_full[b & 1], (b >> 1) & 1);
        mbar_wait(&a_full[ar], (u / C::A_PAIRS) & 1);
      }
      if (ng > ng4) { mbar_wait(keep_full, 0); mbar_wait(kq_full, 0); }
      tc_fence_after();
      if (elect_one_sync()) {
        if (u < 8) trace_stamp(args, 88 + u);
        for (int j = 0; j < ng4; ++j) {
          const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_Q + ((2 * u + j) % C::QS) * C::EXP_Q));
#pragma unroll
          for (int k = 0; k < 4; ++k)      // 4 x K=32: 8 tensor-memory columns of A, 32 B of each token row
            umma_i8_ts(d0 + j * BN, tmem_base + C::A_COL0 + ar * 64 + j * 32 + k * 8, dq + (uint64_t)(k * 2), idesc, k > 0);
        }
        if (ng > ng4) {                    // the INT8 keeper group: both operands from shared memory
          const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + C::OFF_KEEP_P));
          const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_KEEP_Q));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_i8(d0 + ng4 * BN, dp + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc, k > 0);
        }
        if (u < 8) trace_stamp(args, 16 + u);
        if (ng4 > 0 && ((u & 1) == 1 || u == np4 - 1)) umma_commit(&q_empty[b & 1]);    // the batch's token tiles are consumed
        umma_commit(&mma_done[ar]);          // accumulators ready AND operand slot reusable
        if (u < 8) trace_stamp(args, 32 + u);
      }
      __syncwarp();
    }
  } else if (warp == 2 || warp == 3) {
    // ============================================================ token tiles: global -> INT8 operand in smem
    // chunk c of a batch: group c / (4 BN), token row (c % (4 BN)) / 4, 16-byte piece c % 4 of the row's 64 packed bytes
    const int tq = (warp - 2) * 32 + lane;
    const size_t kp = (size_t)args.G * 64;                // packed bytes per token row
    griddep_wait();                                       // the activations are the preceding kernel's output
    if (tq == 0) trace_stamp(args, 6);
    const int nbatch = (n4 + C::QB - 1) / C::QB;
    // One pass = up to 256 chunks (4 per thread).  The loads of the NEXT pass are issued before the current one is expanded
    // and stored, so the global-memory latency of consecutive batches overlaps (it is the longest item on the dependent path).
    auto load_pass = [&](int b, int c0, uint4 (&w)[4]) {
      const int ng = min(C::QB, n4 - b * C::QB), chunks = ng * BN * 4;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int c = c0 + x * 64 + tq;
        w[x] = make_uint4(0, 0, 0, 0);
        if (c < chunks) {
          const int grp = c / (BN * 4), r = (c % (BN * 4)) >> 2, j = c & 3;
          if (m0 + r < args.M)
            w[x] = ld_cg_v4(args.a4 + (size_t)(m0 + r) * kp + (size_t)(g_begin + b * C::QB + grp) * 64 + j * 16);
        }
      }
    };
    auto store_pass = [&](int b, int c0, const uint4 (&w)[4]) {
      const int ng = min(C::QB, n4 - b * C::QB), chunks = ng * BN * 4;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int c = c0 + x * 64 + tq;
        if (c < chunks) {
          const int grp = c / (BN * 4), r = (c % (BN * 4)) >> 2, j = c & 3;
          uint4 lo, hi;
          expand_chunk(w[x], lo, hi);
          uint8_t* row = smem + C::OFF_EXP_Q + ((b * C::QB + grp) % C::QS) * C::EXP_Q + (r >> 3) * 1024 + (r & 7) * 128;
          *reinterpret_cast<uint4*>(row + (((2 * j) ^ (r & 7)) << 4)) = lo;
          *reinterpret_cast<uint4*>(row + (((2 * j + 1) ^ (r & 7)) << 4)) = hi;
        }
      }
    };
    constexpr int PASSES = BN * C::QB * 4 / 256;           // passes per full batch: 1 / 2 / 4
    uint4 wa[4], wb[4];
    if (nbatch > 0) load_pass(0, 0, wa);
    for (int b = 0; b < nbatch; ++b) {
      if (b >= 2) mbar_wait(&q_empty[b & 1], ((b >> 1) - 1) & 1);         // the slots this batch overwrites are free
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        // next pass: same batch, or the first pass of the next batch (its loads do not touch shared memory, so they may be
        // issued before that batch's slots are known to be free)
        const bool more = p + 1 < PASSES || b + 1 < nbatch;
        const int nb = p + 1 < PASSES ? b : b + 1, nc0 = p + 1 < PASSES ? (p + 1) * 256 : 0;
        if ((p & 1) == 0) { if (more) load_pass(nb, nc0, wb); store_pass(b, p * 256, wa); }
        else              { if (more) load_pass(nb, nc0, wa); store_pass(b, p * 256, wb); }
      }
      if constexpr (PASSES & 1) {                           // odd number of passes per batch: the buffers swap roles per batch
#pragma unroll
        for (int x = 0; x < 4; ++x) { const uint4 t4 = wa[x]; wa[x] = wb[x]; wb[x] = t4; }
      }
      fence_proxy_async_smem();            // generic-proxy stores -> visible to the MMA's operand fetch
      __syncwarp();
      if (lane == 0) mbar_arrive(&qx_full[b & 1]);
      if (tq == 0 && b == 0) trace_stamp(args, 5);
    }
    if (has_keeper) {                                      // keeper tokens: plain copy into the SWIZZLE_128B layout
      for (int c = tq; c < BN * 8; c += 64) {
        const int r = c >> 3, j = c & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (m0 + r < args.M) v = ld_cg_v4(args.a8 + (size_t)(m0 + r) * 128 + j * 16);
        *reinterpret_cast<uint4*>(smem + C::OFF_KEEP_Q + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(kq_full);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ============================================================ weight converters: thread = weight row; with two converter
    // warpgroups, warps 4-7 take the first group of every unit and warps 12-15 the second
    const int wq = warp & 3, row = wq * 32 + lane, t = (warp - 4) * 32 + lane;
    const int jsel = warp >= 12 ? 1 : 0;
    const int xr = (row >> 1) & 3;                       // SWIZZLE_64B: 16-B chunk index ^= address bits [7,9)
    for (int u = 0; u < np4; ++u) {
      const int ar = u % C::A_PAIRS, ng = min(2, n4 - 2 * u);
      if (u >= C::A_PAIRS) mbar_wait(&mma_done[ar], ((u / C::A_PAIRS) - 1) & 1);
      if (t == 0 && u < 8) trace_stamp(args, 24 + u);
      for (int j = (C::CONV_WGS == 2 ? jsel : 0); j < (C::CONV_WGS == 2 ? min(ng, jsel + 1) : ng); ++j) {
        const int i = 2 * u + j, ps = i % C::PACK;
        mbar_wait(&pack_full[ps], (i / C::PACK) & 1);
        if (t == 0 && u < 8 && j == 0) trace_stamp(args, 40 + u);
        const uint8_t* prow = smem + C::OFF_PACK_P + ps * C::PACK_P + row * 64;
        uint4 w[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) w[x] = *reinterpret_cast<const uint4*>(prow + ((x ^ xr) << 4));
        // chunk x (32 consecutive K) -> columns 8x..8x+7: four "even element" words, then four "odd element" words
        // (the token tiles use the same K permutation, so dot products are unchanged)
        uint32_t r[32];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 lo, hi;
          expand_chunk(w[x], lo, hi);
          r[8 * x + 0] = lo.x; r[8 * x + 1] = lo.y; r[8 * x + 2] = lo.z; r[8 * x + 3] = lo.w;
          r[8 * x + 4] = hi.x; r[8 * x + 5] = hi.y; r[8 * x + 6] = hi.z; r[8 * x + 7] = hi.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pack_empty[ps]);     // the packed tile is in registers: the slot can be refilled
        tmem_st_32x32b_x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::A_COL0 + ar * 64 + j * 32), r);
      }
      tmem_st_wait();
      if (t == 0 && u < 8) trace_stamp(args, 56 + u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[ar]);
      if (t == 0 && u < 8) trace_stamp(args, 72 + u);
    }
  } else if (warp >= 8 && warp < 12) {
    // ============================================================ epilogue: thread = output channel (TMEM lane)
    const int wq = warp & 3, row = wq * 32 + lane, te = (warp - 8) * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int c = 0; c < BN; ++c) acc[c] = 0.f;
    __half* sb_s = reinterpret_cast<__half*>(smem + C::OFF_SB);
    uint32_t* sa_s = reinterpret_cast<uint32_t*>(smem + C::OFF_SA);
    int staged_from = 0, staged_to = 0;      // groups [staged_from, staged_to) have their scale rows in shared memory

    for (int u = 0; u < nu; ++u) {
      const int i0 = 2 * u, ng = min(2, iters - 2 * u);
      if (i0 + ng > staged_to) {
        // ---- stage the scale rows of groups [i0, i0 + SC): weight scales first (no dependency), then activation scales
        const int cnt = min(C::SC, iters - i0);
        if (u > 0) asm volatile("bar.sync 2, 128;" ::: "memory");        // everyone is done with the previous chunk
        for (int c = te; c < cnt * 16; c += 128) {
          const int gi = c >> 4, part = c & 15, g = g_begin + i0 + gi;
          const __half* bs_row = (g == args.G) ? args.b_keeper_scale : args.b_scale + (size_t)g * args.ldb_scale;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (n0 + 8 * part < (kEpi == EPI_GATEUP ? args.gu_rows : args.N)) v = ld_cg_v4(bs_row + wrow0 + 8 * part);
          reinterpret_cast<uint4*>(sb_s)[c] = v;
        }
        if (u == 0) griddep_wait();
        for (int c = te; c < cnt * (BN / 2); c += 128) {
          const int gi = c / (BN / 2), w = c % (BN / 2), blk = w >> 3, r = w & 7, g = g_begin + i0 + gi;
          const __half* as_row = (g == args.G) ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
          uint32_t v = 0;
          if (m0 + 16 * blk + r < args.M) v = ld_cg_u32(as_row + 64 * (m0 / 16 + blk) + 8 * r);
          sa_s[c] = v;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        staged_from = i0; staged_to = i0 + cnt;
        if (te == 0 && u == 0) trace_stamp(args, 7);
      }
      const int as = u % C::ACC_PAIRS;
      mbar_wait(&mma_done[u % C::A_PAIRS], (u / C::A_PAIRS) & 1);
      tc_fence_after();
      if (warp == 8 && lane == 0 && u < 8) trace_stamp(args, 104 + u);
      const uint32_t tpair = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::ACC_COL0 + as * 2 * BN);
      if constexpr (BN == 16) {
        // both groups of the pair with one tcgen05.ld (a single group: the upper half is stale and ignored)
        uint32_t r[32];
        tmem_ld_32x32b_x32(tpair, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);       // the pair's accumulators are in registers: hand the slot back
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < ng) {
            const int si = i0 + j - staged_from;
            const __half2 sm2 = reinterpret_cast<const __half2*>(sb_s + si * 128)[row >> 1];   // {sB[n & ~1], sB[n | 1]}
            const __half2* snw = reinterpret_cast<const __half2*>(sa_s + si * (BN / 2));       // (sA[tok], sA[tok + 8]) words
            dequant16(acc, r + 16 * j, snw, sm2, g_begin + i0 + j == args.G);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < ng) {
            const int si = i0 + j - staged_from;
            const bool keeper = (g_begin + i0 + j == args.G);
            const __half2 sm2 = reinterpret_cast<const __half2*>(sb_s + si * 128)[row >> 1];
            const __half2* snw = reinterpret_cast<const __half2*>(sa_s + si * (BN / 2));
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 16) {
              uint32_t r[16];
              tmem_ld_32x32b_x16(tpair + j * BN + c0, r);
              tmem_ld_wait();
              if (c0 + 16 == BN && j == ng - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
              }
              dequant16(acc + c0, r, snw + (c0 >> 4) * 8, sm2, keeper);
            }
          }
        }
      }
      if (warp == 8 && lane == 0 && u < 8) trace_stamp(args, 120 + u);
    }
    if (nu == 0) griddep_wait();             // (an empty split-K rank still orders its stores behind the preceding kernel)
    if (warp == 8 && lane == 0) trace_stamp(args, 2);
    constexpr float kInv = 1.0f / 256.0f;   // exact: removes the 16 * 16 operand factor

    if constexpr (kEpi == EPI_GATEUP) {
      cluster_wait();                        // pairs with the setup arrive: both CTAs of the pair are running
      if (krank == 1) {
        // ---------------------------------------------------------- up tile: hand the FP32 sums to the gate CTA
        const uint32_t remote = mapa_shared(smem_u32(smem + C::OFF_RED), 0) + row * (BN * 4);
        const uint32_t rbar = mapa_shared(smem_u32(red_full), 0);
#pragma unroll
        for (int c = 0; c < BN; c += 4) st_async_v4(remote + c * 4, acc[c], acc[c + 1], acc[c + 2], acc[c + 3], rbar);
      } else {
        // ---------------------------------------------------------- gate tile: SiLU(gate) * up, then the dynamic
        // quantisation of activate_fp16_i4 (Activate.cuh:102-166) for this 128-channel group of every token.  Both
        // projections are rounded to FP16 first, as they are when the reference stores them between the kernels.
        mbar_wait(red_full, 0);
        const float* red = reinterpret_cast<const float*>(smem + C::OFF_RED);
        float* xmx = reinterpret_cast<float*>(smem + C::OFF_XCH);   // [4 warps][BN]
        const bool last = ((int)blockIdx.x == args.gu_rows / 128 - 1);          // the INT8 keeper group of the down projection
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          const float g = __half2float(__float2half_rn(acc[c] * kInv));
          const float up = __half2float(__float2half_rn(red[row * BN + c] * kInv));
          const float t = silu_ref(g) * up;
          acc[c] = t;
          uint32_t ua = __float_as_uint(fabsf(t));
          if (n0 + row >= args.gu_rows) ua = 0u;
          ua = __reduce_max_sync(0xffffffffu, ua);        // |t| >= 0: IEEE bit patterns order like unsigned integers
          if (lane == 0) xmx[wq * BN + c] = __uint_as_float(ua);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const size_t q4_pitch = (size_t)(args.gu_rows - 128) / 2;
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          const int m = m0 + c;
          float maxv = fmaxf(fmaxf(xmx[c], xmx[BN + c]), fmaxf(xmx[2 * BN + c], xmx[3 * BN + c]));
          maxv = maxv / (last ? 127.f : 7.f);                      // IEEE division, as `maxv /= 7`
          const float r_scale = 1.f / maxv;
          const int tq = (int)roundf(acc[c] * r_scale);            // round half away from zero (CUDA round())
          const int q = last ? max(-128, min(127, tq)) : max(-8, min(7, tq));
          const int qn = __shfl_down_sync(0xffffffffu, q, 1);      // channel n+1 lives in the next lane
          if (m < args.M && n0 + row < args.gu_rows) {
            if (last) args.q8_out[(size_t)m * 128 + row] = (int8_t)q;
            else if ((lane & 1) == 0) args.q4_out[(size_t)m * q4_pitch + (size_t)blockIdx.x * 64 + (row >> 1)] = (uint8_t)((q & 0xF) | ((qn & 0xF) << 4));
            if (row == 0) {
              const __half hs = __float2half_rn(maxv);
              __half* dst = last ? args.q8_scale : args.q4_scale + (size_t)blockIdx.x * args.lda_scale;
              const int si = scale_index(m);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[si + 2 * j] = hs;     // replicated x4 (ldmatrix layout of the reference GEMM)
            }
          }
        }
      }
    } else {
    // fused q/k/v: the channel tile decides the epilogue (uniform per CTA)
    const int seg = kEpi == EPI_QKV ? (int)blockIdx.x / args.seg_tiles : 0;
    const int tile = kEpi == EPI_QKV ? (int)blockIdx.x % args.seg_tiles : (int)blockIdx.x;
    const int n_out_dim = kEpi == EPI_QKV ? args.seg_tiles * 128 : args.N;       // row length of the output this tile writes
    if (kEpi == EPI_O16 || kEpi == EPI_PUSH || (kEpi == EPI_QKV && seg == 0)) {
      // ---------------------------------------------------------- split-K: rank d reduces + stores token columns
      // [d * CPR, (d + 1) * CPR).  Partials travel as st.async messages that also complete transaction bytes on the
      // owner's mbarrier: no cluster barrier, and the owner sums in rank order 0..kSplit-1 (deterministic).
      if constexpr (kSplit > 1) {
        cluster_wait();                      // pairs with the setup arrive: every CTA of the cluster is running
        const uint32_t red_local = smem_u32(smem + C::OFF_RED), bar_local = smem_u32(red_full);
#pragma unroll
        for (int d = 0; d < kSplit; ++d) {
          if (d != (int)krank) {
            const int slot_in_dst = (int)krank < d ? (int)krank : (int)krank - 1;
            const uint32_t remote = mapa_shared(red_local, d) + slot_in_dst * C::RED_BYTES + row * (C::CPR * 4);
            const uint32_t rbar = mapa_shared(bar_local, d);
            if constexpr (C::CPR == 2) {
              st_async_v2(remote, acc[d * 2], acc[d * 2 + 1], rbar);
            } else {
#pragma unroll
              for (int c = 0; c < C::CPR; c += 4)
                st_async_v4(remote + c * 4, acc[d * C::CPR + c], acc[d * C::CPR + c + 1], acc[d * C::CPR + c + 2],
                            acc[d * C::CPR + c + 3], rbar);
            }
          }
        }
        mbar_wait(red_full, 0);
        const float* red = reinterpret_cast<const float*>(smem + C::OFF_RED);
#pragma unroll
        for (int d = 0; d < kSplit; ++d) {
          if (d == (int)krank) {
#pragma unroll
            for (int c = 0; c < C::CPR; ++c) {
              float s = 0.f;
#pragma unroll
              for (int rk = 0; rk < kSplit; ++rk) {
                const float v = rk == d ? acc[d * C::CPR + c] : red[(rk < d ? rk : rk - 1) * (C::RED_BYTES / 4) + row * C::CPR + c];
                s = rk == 0 ? v : s + v;
              }
              acc[d * C::CPR + c] = s;
            }
          }
        }
      }
      if (warp == 8 && lane == 0) trace_stamp(args, 3);
      const int n = tile * C::BM + row;
      if constexpr (kEpi == EPI_PUSH) {
        // fused all-reduce, push half: this rank's CPR token rows of the tile go to slot [call % 3][rank] of EVERY rank's receive
        // buffer (-0.0, the buffers' "not yet arrived" pattern, travels as +0.0); the following add+RMSNorm kernel polls and
        // sums the slots.  The tile is transposed through the (now idle) packed-weight ring so that the peer stores are 16 bytes
        // of 8 consecutive channels: 2-byte stores over NVLink cost an order of magnitude more per byte.
        __half* stg = reinterpret_cast<__half*>(smem + C::OFF_PACK_P);         // [CPR tokens][128 channels]
#pragma unroll
        for (int d = 0; d < kSplit; ++d) {
          if (kSplit == 1 || d == (int)krank) {            // uniform per CTA
#pragma unroll
            for (int j = 0; j < C::CPR; ++j) {
              unsigned short hv = __half_as_ushort(__float2half_rn(acc[d * C::CPR + j] * kInv));
              if (hv == 0x8000u) hv = 0;
              stg[j * C::BM + row] = __ushort_as_half(hv);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t cur = (ar_ld_state(args.ar.state) + 1) % 3;
        const size_t base = ((size_t)cur * args.ar.world + args.ar.rank) * (size_t)args.ar.slot;
        const int pt = (warp - 8) * 32 + lane;
        for (int q = pt; q < C::CPR * 16; q += 128) {
          const int j = q >> 4, piece = q & 15;
          const int m = m0 + (int)krank * C::CPR + j, nn = n0 + piece * 8;
          if (m < args.M && nn < args.N) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + j * C::BM + piece * 8);
            const size_t off = base + (size_t)m * n_out_dim + nn;
#pragma unroll 1
            for (int r = 0; r < args.ar.world; ++r) {
              int peer = args.ar.rank + r;
              if (peer >= args.ar.world) peer -= args.ar.world;
              ar_st_v4(reinterpret_cast<__half*>(args.ar.bufs[peer]) + off, v);
            }
          }
        }
      } else if (n < n_out_dim && n0 + row < args.N) {
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          const int m = m0 + c;
          if ((kSplit == 1 || c / C::CPR == (int)krank) && m < args.M)
            args.d[(size_t)m * n_out_dim + n] = __float2half_rn(acc[c] * kInv);      // a warp writes 64-B runs
        }
      }
    } else {
      // ---------------------------------------------------------- o4 (DenseLayerGEMM_i4_o4.cu:705-787): per (token,
      // 128-channel head) asymmetric INT4 with the reference's |v| min/max; thread = channel, reduce over the 128 rows
      uint8_t* d4 = (kEpi == EPI_QKV && seg == 2) ? args.d4_v : args.d4;
      __half2* d_scale = (kEpi == EPI_QKV && seg == 2) ? args.d_scale_v : args.d_scale;
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
      const int n = tile * C::BM + row;
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        const int m = m0 + c;
        const float mx = fmaxf(fmaxf(xmx[c], xmx[BN + c]), fmaxf(xmx[2 * BN + c], xmx[3 * BN + c]));
        const float mn = fminf(fminf(xmn[c], xmn[BN + c]), fminf(xmn[2 * BN + c], xmn[3 * BN + c]));
        const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
        const int q = (int)roundf((acc[c] + zero) * r_scale) & 0xF;
        const int qn = __shfl_down_sync(0xffffffffu, q, 1);     // channel n+1 lives in the next lane
        if (m < args.M && n < n_out_dim && n0 + row < args.N) {
          if ((lane & 1) == 0) d4[(size_t)m * (n_out_dim / 2) + n / 2] = (uint8_t)(q | (qn << 4));
          if (row == 0) d_scale[(size_t)m * (n_out_dim / 128) + tile] = __floats2half2_rn(scale, zero);
        }
      }
    }
    }   // !EPI_GATEUP
  }

  // ---------------------------------------------------------------- teardown
  // split-K: this CTA's shared memory is a target only until red_full completed (the epilogue waited for it), so no
  // closing cluster barrier is needed; every thread still pairs its setup arrive with one wait.
  if constexpr (kSplit > 1) { if (warp < 8 || warp >= 12) cluster_wait(); }
  tc_fence_before();
  __syncthreads();
  if constexpr (kEpi == EPI_GATEUP) { if (args.dbg & 1) { cluster_arrive(); cluster_wait(); } }   // the up CTA outlives the hand-off
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_stamp(args, 4);
}

}  // namespace atom
