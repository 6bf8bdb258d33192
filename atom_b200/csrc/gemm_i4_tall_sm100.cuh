// gemm_i4_tall_sm100.cuh -- prefill-shape (M > 64 tokens) W4A4 GEMM for B200: tokens on the MMA-M axis, 128 x 128 tiles.
//
// Same contract as the reference's compute_gemm_imma / DenseLayerGEMM_i4[_o4]_kernel
// (/root/reference/kernels/include/GEMM/Dense_layer_gemm_i4_o16.cuh:436-710,
//  /root/reference/e2e/punica-atom/punica/ops/csrc/GEMM/DenseLayerGEMM_i4_o4.cu:440-787): exact INT32 group sums, one
// FP16 multiply of the two scales, FP32 fma accumulation in group order, keeper last, RN cast to half -- bit-identical output.
//
// Pipeline (one CTA per SM, 16 warps):
//   warp 0     TMA producer: packed INT4 tiles (64 B rows) of both operands, GS = 2 quantisation groups per stage
//   warps 4-7  converters: nibble -> int8(value * 16) into the canonical K-major SWIZZLE_128B operand layout
//   warp 1     one thread issues tcgen05.mma.kind::i8 (128 x 128 x 32) x 4 per group, one commit per stage
//   warp 3     scale loader
//   warps 8-15 two epilogue warpgroups, 64 accumulator columns each
//
// What bounds this kernel is not the tensor pipe but the per-group FP32 dequantisation the reference's arithmetic
// demands: at kind::i8 rate a 128x128x128 group is 256 tensor cycles, i.e. 1024 issue slots per SM for 16384
// accumulator elements.  Round 1 spent 3.5 SASS instructions per element there (I2F, FFMA, half of an F2F, a quarter of a
// PRMT / HMUL2 / two LDS) -- 1.7x the budget before the converter issues anything.  This version spends 1.8:
//   * the INT32 accumulators live in tensor memory BIASED by 0x4B400000 (the bit pattern of 12582912.0f = 1.5 * 2^23):
//     the epilogue re-arms a slot with tcgen05.st after reading it and every MMA accumulates.  The word read back,
//     reinterpreted as FP32, IS 12582912 + c exactly (|c| < 2^22: a group sum is at most 2^21 after the 16 * 16 operand
//     factor), so one packed FADD2 (-12582912, exact) replaces two I2F on the quarter-rate conversion pipe;
//   * the fma is the packed FFMA2 (same IEEE single rounding per lane as fmaf);
//   * the weight-tile rows are permuted BY THE TMA (4-D tensor map, rows 4q+{0,2,1,3}), so that two neighbouring
//     accumulator columns belong to two different channel pairs: the (rs_p, rs_p+1) couple that one HMUL2 + two F2F
//     produce is directly the 64-bit multiplier operand of two FFMA2 -- no duplication moves;
//   * the scale loader de-interleaves the weight scales (even / odd channel of every pair, the reference's column
//     pairing, Dense_layer_gemm_i4_o16.cuh:417-431) so a thread fetches 8 pair scales with one LDS.128 and no PRMT;
//   * the keeper's 2^8 (its operands are not pre-multiplied by 16) is folded into the scale product (exact power of two).
#pragma once
#include "gemm_i4_sm100.cuh"

namespace atom {

template <bool kO4>
struct TallCfg {
  static constexpr int BM = 128, BN = 128;
  static constexpr int GS = 2;                                    // quantisation groups per pipeline stage
  static constexpr int RING = 2;                                  // operand-slot / accumulator stages
  static constexpr int PACK = 2;                                  // packed-tile stages
  static constexpr int TMEM_COLS = RING * GS * BN;                // 512
  static constexpr int SCALE_STAGES = 4;
  static constexpr int EPI_WGS = 2, CPT = BN / EPI_WGS;           // 64 accumulator columns per epilogue thread
  static constexpr int CONV_WARPS = 4;                            // 8 measured slower (782 vs 829 TOP/s at 4096^3): the SM is issue-bound, not converter-latency-bound
  static constexpr int EPI_WARP0 = 4 + CONV_WARPS;                // first epilogue warp
  static constexpr int THREADS = 32 * (EPI_WARP0 + 4 * EPI_WGS);  // 512
  static constexpr int PACK_T = 128 * 64, EXP_T = 128 * 128;      // bytes per packed / expanded group tile
  static constexpr int OFF_EXP_P = 0;
  static constexpr int OFF_EXP_Q = OFF_EXP_P + RING * GS * EXP_T;
  static constexpr int OFF_PACK_P = OFF_EXP_Q + RING * GS * EXP_T;
  static constexpr int OFF_PACK_Q = OFF_PACK_P + PACK * GS * PACK_T;
  static constexpr int OFF_SM = OFF_PACK_Q + PACK * GS * PACK_T;  // scale ring: [stage][group][512 B]
  static constexpr int OFF_BAR = OFF_SM + SCALE_STAGES * GS * 512;
  static constexpr int NUM_BARS = 2 * PACK + 3 * RING + 2 * SCALE_STAGES;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

constexpr uint32_t kAccBias = 0x4B400000u;   // bit pattern of 12582912.0f

template <int kRegs> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// 16 registers that hold the bias pattern for the whole kernel: tcgen05.st needs 16 consecutive source registers, and a
// plain constant (or anything ptxas can prove uniform) makes the compiler rebuild all 16 with MOVs before every store --
// one MOV per accumulator element.  The vector is therefore read back from TMEM once (tcgen05.ld results are opaque).
struct BiasRegs { uint32_t r[16]; };
__device__ __forceinline__ BiasRegs make_bias_regs(uint32_t armed_taddr) {   // 16 columns at armed_taddr already hold kAccBias
  BiasRegs b;
  tmem_ld_32x32b_x16(armed_taddr, b.r);
  tmem_ld_wait();
  return b;
}
__device__ __forceinline__ void tmem_st_const_x16(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const BiasRegs& b) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(b.r[0]), "r"(b.r[1]), "r"(b.r[2]), "r"(b.r[3]), "r"(b.r[4]), "r"(b.r[5]), "r"(b.r[6]), "r"(b.r[7]),
        "r"(b.r[8]), "r"(b.r[9]), "r"(b.r[10]), "r"(b.r[11]), "r"(b.r[12]), "r"(b.r[13]), "r"(b.r[14]), "r"(b.r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// {as_float(a), as_float(b)} * mul - 12582912 * mul: the biased accumulator words as exact FP32 integers, times the group's
// power-of-two factor (1 for INT4 groups, 256 for the keeper) -- one FFMA2, exact (24 significant bits, no rounding)
__device__ __forceinline__ float2 unbias2(uint32_t a, uint32_t b, float mul, float nbias) {
  float2 r;
  asm("{\n\t.reg .b64 t, u, v;\n\t"
      "mov.b64 t, {%2, %3};\n\t"
      "mov.b64 u, {%4, %4};\n\t"
      "mov.b64 v, {%5, %5};\n\t"
      "fma.rn.f32x2 t, t, u, v;\n\t"
      "mov.b64 {%0, %1}, t;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "r"(a), "r"(b), "f"(mul), "f"(nbias));
  return r;
}
// acc = c * rs + acc per lane (FFMA2: two IEEE fmaf)
__device__ __forceinline__ void ffma2(float2& acc, const float2 c, const float2 rs) {
  asm("{\n\t.reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "mov.b64 d, {%0, %1};\n\t"
      "fma.rn.f32x2 d, a, b, d;\n\t"
      "mov.b64 {%0, %1}, d;\n\t}"
      : "+f"(acc.x), "+f"(acc.y) : "f"(c.x), "f"(c.y), "f"(rs.x), "f"(rs.y));
}

template <bool kO4>
__global__ void __launch_bounds__(TallCfg<kO4>::THREADS, 1)
gemm_i4_tall_kernel(const __grid_constant__ CUtensorMap tm_p4,   // packed INT4 tokens   (2-D, box 64 B x 128 rows)
                    const __grid_constant__ CUtensorMap tm_q4,   // packed INT4 weights  (4-D view: rows 4q+{0,2,1,3}, box 64 B x 2 x 2 x 32)
                    const __grid_constant__ CUtensorMap tm_p8,   // INT8 keeper tokens   (2-D, box 128 B x 128 rows, SWIZZLE_128B)
                    const __grid_constant__ CUtensorMap tm_q8,   // INT8 keeper weights  (4-D view, box 128 B x 2 x 2 x 32, SWIZZLE_128B)
                    const GemmArgs args) {
  using C = TallCfg<kO4>;
  constexpr int BN = C::BN, GS = C::GS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;                          // TMA landed the stage's packed tiles          (1 + tx)
  uint64_t* pack_empty = pack_full + C::PACK;          // converters have read them                    (4)
  uint64_t* exp_full = pack_empty + C::PACK;           // expanded operands of the stage are in place  (4 [+ tx])
  uint64_t* mma_done = exp_full + C::RING;             // the stage's MMAs completed                   (1)
  uint64_t* tmem_empty = mma_done + C::RING;           // epilogue has read AND re-armed the stage's accumulators (8)
  uint64_t* scale_full = tmem_empty + C::RING;         // scales of the stage staged                   (32)
  uint64_t* scale_empty = scale_full + C::SCALE_STAGES;//                                              (8)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = output-channel tile: CTAs launched together share the token tile (L2) and stream disjoint weights
  const int m0 = blockIdx.y * C::BM, n0 = blockIdx.x * BN;
  const int total_groups = args.G + 1;                 // index G = INT8 keeper
  const int nstages = (total_groups + GS - 1) / GS;
  auto stage_groups = [&](int s) { return min(GS, total_groups - s * GS); };
  auto stage_int4 = [&](int s) { return max(0, min(s * GS + stage_groups(s), args.G) - s * GS); };
  if (threadIdx.x == 0) { griddep_launch_dependents(); trace_stamp(args, 0); }

  // part: 1 = weight tiles (independent of the preceding kernel), 2 = token tiles, 3 = both
  auto issue_stage = [&](int s, int ps, int part) {
    const int n4 = stage_int4(s);
    if (part & 1) mbar_arrive_expect_tx(&pack_full[ps], n4 * 2 * C::PACK_T);
    for (int j = 0; j < n4; ++j) {
      const int g = s * GS + j;
      if (part & 2) tma_load_2d(smem + C::OFF_PACK_P + (ps * GS + j) * C::PACK_T, &tm_p4, &pack_full[ps], g * 64, m0);
      if (part & 1) tma_load_4d(smem + C::OFF_PACK_Q + (ps * GS + j) * C::PACK_T, &tm_q4, &pack_full[ps], g * 64, 0, 0, n0 / 4);
    }
    if (s < 16 && (part & 2)) trace_stamp(args, 8 + s);
  };
  int s_w = 0;
  for (; s_w < C::PACK && s_w < nstages && stage_int4(s_w) > 0; ++s_w) {}     // stages whose weight tiles are issued up front
  if (warp == 0 && elect_one_sync()) {       // an elected lane of the converged warp: uniform TMA operands
    tma_prefetch_desc(&tm_p4); tma_prefetch_desc(&tm_q4);
    for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_full[i], 1);
    fence_barrier_init();
    for (int s = 0; s < s_w; ++s) issue_stage(s, s, 1);   // weights first
    tma_prefetch_desc(&tm_p8); tma_prefetch_desc(&tm_q8);
    for (int i = 0; i < C::PACK; ++i) mbar_init(&pack_empty[i], C::CONV_WARPS);
    for (int i = 0; i < C::RING; ++i) { mbar_init(&exp_full[i], C::CONV_WARPS); mbar_init(&mma_done[i], 1); mbar_init(&tmem_empty[i], 4 * C::EPI_WGS); }
    for (int i = 0; i < C::SCALE_STAGES; ++i) { mbar_init(&scale_full[i], 32); mbar_init(&scale_empty[i], 4 * C::EPI_WGS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) trace_stamp(args, 1);

  // 512 threads => 128 registers each at launch.  setmaxnreg.inc can only claim what warps of the SAME CTA released (CTA
  // pool; a claim beyond it spins forever): service 128 -> 40 and the converter warpgroup 128 -> 80 release 88 + 48 = 136
  // slices of 128 registers, the two epilogue warpgroups claim 2 x (160 - 128) = 64
  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ============================================================ TMA producer (warp loops, one elected lane issues)
      griddep_wait();                              // the token tiles are the preceding kernel's output
      if (elect_one_sync()) { for (int s = 0; s < s_w; ++s) issue_stage(s, s, 2); }
      __syncwarp();
      for (int s = s_w; s < nstages; ++s) {
        if (stage_int4(s) == 0) break;             // a trailing keeper-only stage has nothing in the packed ring
        const int ps = s % C::PACK;
        mbar_wait(&pack_empty[ps], ((s / C::PACK) & 1) ^ 1);
        if (elect_one_sync()) issue_stage(s, ps, 3);
        __syncwarp();
      }
    } else if (warp == 1) {
      // ============================================================ MMA issuer (warp loops, one elected lane issues)
      constexpr uint32_t idesc = umma_idesc_i8(C::BM, BN);
      for (int s = 0; s < nstages; ++s) {
        const int es = s % C::RING;
        mbar_wait(&tmem_empty[es], (s / C::RING) & 1);        // completion #0 is the initial arming of the slot
        mbar_wait(&exp_full[es], (s / C::RING) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          if (s < 16) trace_stamp(args, 88 + s);
          const int ng = stage_groups(s);
          for (int j = 0; j < ng; ++j) {
            const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_P + (es * GS + j) * C::EXP_T));
            const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_Q + (es * GS + j) * C::EXP_T));
#pragma unroll
            for (int k = 0; k < 4; ++k)     // always accumulating: the slot holds the bias pattern, not zero
              umma_i8(tmem_base + (es * GS + j) * BN, dp + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc, 1u);
          }
          umma_commit(&mma_done[es]);       // one commit per stage: epilogue may read, converter may refill
        }
        __syncwarp();
      }
    } else if (warp == 3) {
      // ============================================================ scale loader.  Group slot (512 B):
      //   [0,256)   (lower, upper) activation-scale words of the tile's token rows (word = (r/16)*8 + r%8, raw copy)
      //   [256,384) weight scale of the EVEN channel of each of the tile's 64 channel pairs; [384,512) of the ODD one
      griddep_wait();
      for (int s = 0; s < nstages; ++s) {
        const int ss = s % C::SCALE_STAGES, ng = stage_groups(s);
        uint32_t aw[GS][2];
        uint4 bw[GS];
#pragma unroll
        for (int j = 0; j < GS; ++j) {
          aw[j][0] = aw[j][1] = 0u; bw[j] = make_uint4(0, 0, 0, 0);
          if (j < ng) {
            const int g = s * GS + j;
            const bool keeper = (g == args.G);
            const __half* as_row = keeper ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
            const __half* bs_row = keeper ? args.b_keeper_scale : args.b_scale + (size_t)g * args.ldb_scale;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int w = lane + 32 * h, blk = w >> 3, i = w & 7;
              if (m0 + 16 * blk + i < args.M) aw[j][h] = ld_cg_u32(as_row + 64 * (m0 / 16 + blk) + 8 * i);
            }
            if (lane < 16 && n0 + 8 * lane < args.N) bw[j] = ld_cg_v4(bs_row + n0 + 8 * lane);
          }
        }
        if (s >= C::SCALE_STAGES) mbar_wait(&scale_empty[ss], ((s / C::SCALE_STAGES) - 1) & 1);
#pragma unroll
        for (int j = 0; j < GS; ++j) {
          if (j < ng) {
            uint8_t* slot = smem + C::OFF_SM + (ss * GS + j) * 512;
            reinterpret_cast<uint32_t*>(slot)[lane] = aw[j][0];
            reinterpret_cast<uint32_t*>(slot)[lane + 32] = aw[j][1];
            if (lane < 16) {
              const uint2 ev = make_uint2(__byte_perm(bw[j].x, bw[j].y, 0x5410), __byte_perm(bw[j].z, bw[j].w, 0x5410));
              const uint2 od = make_uint2(__byte_perm(bw[j].x, bw[j].y, 0x7632), __byte_perm(bw[j].z, bw[j].w, 0x7632));
              reinterpret_cast<uint2*>(slot + 256)[lane] = ev;
              reinterpret_cast<uint2*>(slot + 384)[lane] = od;
            }
          }
        }
        mbar_arrive(&scale_full[ss]);
      }
    }
  } else if (warp < C::EPI_WARP0) {
    // ============================================================ converter warps
    reg_dealloc<80>();
    const int t = (warp - 4) * 32 + lane;
    for (int s = 0; s < nstages; ++s) {
      const int es = s % C::RING, ps = s % C::PACK;
      const int ng = stage_groups(s), n4 = stage_int4(s);
      if (s >= C::RING) mbar_wait(&mma_done[es], ((s / C::RING) - 1) & 1);   // MMAs that read this slot have completed
      if (t == 0 && s < 16) trace_stamp(args, 24 + s);
      if (n4 > 0) {
        mbar_wait(&pack_full[ps], (s / C::PACK) & 1);
        if (t == 0 && s < 16) trace_stamp(args, 40 + s);
#pragma unroll
        for (int j = 0; j < GS; ++j) {
          if (j < n4) {
            convert_tile<128, 32 * C::CONV_WARPS>(smem + C::OFF_PACK_P + (ps * GS + j) * C::PACK_T, smem + C::OFF_EXP_P + (es * GS + j) * C::EXP_T, t);
            convert_tile<128, 32 * C::CONV_WARPS>(smem + C::OFF_PACK_Q + (ps * GS + j) * C::PACK_T, smem + C::OFF_EXP_Q + (es * GS + j) * C::EXP_T, t);
          }
        }
        if (t == 0 && s < 16) trace_stamp(args, 56 + s);
        fence_proxy_async_smem();     // generic-proxy stores -> visible to tcgen05.mma operand fetch
      }
      __syncwarp();
      if (lane == 0) {
        if (n4 > 0) mbar_arrive(&pack_empty[ps]);
        if (n4 < ng && t == 0) {        // the keeper is this stage's last group: TMA it into its operand slot
          griddep_wait();
          mbar_arrive_expect_tx(&exp_full[es], 2 * C::EXP_T);
          tma_load_2d(smem + C::OFF_EXP_P + (es * GS + n4) * C::EXP_T, &tm_p8, &exp_full[es], 0, m0);
          tma_load_4d(smem + C::OFF_EXP_Q + (es * GS + n4) * C::EXP_T, &tm_q8, &exp_full[es], 0, 0, 0, n0 / 4);
        } else {
          mbar_arrive(&exp_full[es]);
        }
      }
      if (t == 0 && s < 16) trace_stamp(args, 72 + s);
    }
  } else {
    // ============================================================ epilogue warpgroups
    reg_alloc<160>();
    const int wq = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;                // token row == TMEM lane
    const int colbase = ((warp - C::EPI_WARP0) >> 2) * C::CPT;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    // arm every accumulator slot with the bias pattern (completion #0 of tmem_empty)
#pragma unroll
    for (int slot = 0; slot < C::RING * GS; ++slot)
#pragma unroll
      for (int c0 = 0; c0 < C::CPT; c0 += 16) tmem_st_const_x16(lane_addr + slot * BN + colbase + c0, kAccBias);
    tmem_st_wait_();
    const BiasRegs kBias = make_bias_regs(lane_addr + colbase);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { for (int es = 0; es < C::RING; ++es) mbar_arrive(&tmem_empty[es]); }

    // position k of this thread's 64 columns: TMEM column colbase + k = channel colbase + (k & ~3) + {0,2,1,3}[k & 3]
    float2 acc[C::CPT / 2];
#pragma unroll
    for (int i = 0; i < C::CPT / 2; ++i) acc[i] = make_float2(0.f, 0.f);
    const bool upper = (((m0 + row) & 15) >= 8);

    for (int s = 0; s < nstages; ++s) {
      const int es = s % C::RING, ss = s % C::SCALE_STAGES;
      const int ng = stage_groups(s);
      mbar_wait(&scale_full[ss], (s / C::SCALE_STAGES) & 1);
      mbar_wait(&mma_done[es], (s / C::RING) & 1);
      tc_fence_after();
      if (warp == C::EPI_WARP0 && lane == 0 && s < 16) trace_stamp(args, 104 + s);
      // The stage's 2 x 4 chunks of 16 columns are software-pipelined: the tcgen05.ld of chunk n+1 is in flight while chunk
      // n is dequantised (an exposed TMEM round trip per chunk is what bounded this loop before: ~16 cycles per element).
      uint32_t rbuf[2][16];
      tmem_ld_32x32b_x16(lane_addr + (uint32_t)((es * GS) * BN + colbase), rbuf[0]);
#pragma unroll
      for (int j = 0; j < GS; ++j) {
        if (j < ng) {
          // keeper operands carry no 16 * 16 factor: its exact sums are scaled by 2^8 while they are unbiased
          const float mul = (s * GS + j == args.G) ? 256.f : 1.f, nbias = -12582912.f * mul;
          const uint8_t* slot = smem + C::OFF_SM + (ss * GS + j) * 512;
          const __half2 pw = reinterpret_cast<const __half2*>(slot)[(row >> 4) * 8 + (row & 7)];
          const __half2 sm2 = __half2half2(upper ? __high2half(pw) : __low2half(pw));
          // pair scales of this row's half (even channel for rows 0-7 of each 16, odd for rows 8-15), 8 per LDS.128
          const uint4* sel = reinterpret_cast<const uint4*>(slot + 256 + (upper ? 128 : 0) + colbase);
          const uint32_t taddr = lane_addr + (uint32_t)((es * GS + j) * BN + colbase);
#pragma unroll
          for (int c = 0; c < 4; ++c) {                          // 16 columns = 8 channel pairs = one LDS.128
            const int n = j * 4 + c;                             // chunk number within the stage (compile-time after unrolling)
            const uint4 sv = sel[c];
            tmem_ld_wait();                                      // chunk n has landed in rbuf[n & 1]
            const bool last = (c == 3 && j == ng - 1);
            if (!last) {                                         // next chunk: same group, or the first of the stage's second group
              const uint32_t nxt = (c < 3) ? taddr + 16 * (c + 1) : lane_addr + (uint32_t)((es * GS + j + 1) * BN + colbase);
              tmem_ld_32x32b_x16(nxt, rbuf[(n + 1) & 1]);
            }
            tmem_st_32x32b_x16(taddr + 16 * c, kBias);           // re-arm the chunk just read
            if (last) {                                          // the stage's accumulators are read and re-armed
              tmem_st_wait_();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[es]);
            }
            const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {                         // 4 columns: pairs (2q', 2q'+1) of this 16-column run
              const float2 rs = __half22float2(__hmul2(sm2, *reinterpret_cast<const __half2*>(&sw[q])));
              const int k = 16 * c + 4 * q;
              ffma2(acc[(k >> 1) + 0], unbias2(rbuf[n & 1][4 * q + 0], rbuf[n & 1][4 * q + 1], mul, nbias), rs);
              ffma2(acc[(k >> 1) + 1], unbias2(rbuf[n & 1][4 * q + 2], rbuf[n & 1][4 * q + 3], mul, nbias), rs);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&scale_empty[ss]);       // this warp no longer reads the stage's scales
      if (warp == C::EPI_WARP0 && lane == 0 && s < 8) trace_stamp(args, 120 + s);
    }
    if (warp == C::EPI_WARP0 && lane == 0) trace_stamp(args, 2);
    griddep_wait();                                        // the output buffer may still be read by the preceding kernel

    // ------------------------------------------------------------ output.  acc[i] = positions (2i, 2i+1); channel of
    // position 4q+t is 4q + {0,2,1,3}[t]: channels (4q, 4q+1) = (acc[2q].x, acc[2q+1].x), (4q+2, 4q+3) = (acc[2q].y, acc[2q+1].y)
    constexpr float kInv = 1.0f / 256.0f;   // exact: removes the 16*16 operand factor
    const int m = m0 + row;
    if constexpr (!kO4) {
      if (m < args.M) {
        __half* drow = args.d + (size_t)m * args.N + n0 + colbase;
#pragma unroll
        for (int i = 0; i < C::CPT; i += 8) {
          if (n0 + colbase + i < args.N) {   // N is a multiple of 8 (16-B rows)
            const int q = i >> 2;
            uint4 v;
            const __half2 h0 = __floats2half2_rn(acc[2 * q].x * kInv, acc[2 * q + 1].x * kInv);
            const __half2 h1 = __floats2half2_rn(acc[2 * q].y * kInv, acc[2 * q + 1].y * kInv);
            const __half2 h2 = __floats2half2_rn(acc[2 * q + 2].x * kInv, acc[2 * q + 3].x * kInv);
            const __half2 h3 = __floats2half2_rn(acc[2 * q + 2].y * kInv, acc[2 * q + 3].y * kInv);
            v.x = *reinterpret_cast<const uint32_t*>(&h0); v.y = *reinterpret_cast<const uint32_t*>(&h1);
            v.z = *reinterpret_cast<const uint32_t*>(&h2); v.w = *reinterpret_cast<const uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(drow + i) = v;
          }
        }
      }
    } else {
      // o4 epilogue (DenseLayerGEMM_i4_o4.cu:705-787): per (token, 128-channel head) asymmetric INT4 with the
      // reference's |v| min/max.  The two epilogue warpgroups hold 64 columns each of the same row.
      float* xch = reinterpret_cast<float*>(smem + C::OFF_PACK_P);   // packed ring is idle by now
      float mx = -INFINITY, mn = INFINITY;
#pragma unroll
      for (int i = 0; i < C::CPT / 2; ++i) {
        acc[i].x *= kInv; acc[i].y *= kInv;
        const float a0 = fabsf(acc[i].x), a1 = fabsf(acc[i].y);
        mx = fmaxf(mx, fmaxf(a0, a1)); mn = fminf(mn, fminf(a0, a1));
      }
      const int part = (warp - C::EPI_WARP0) >> 2;
      xch[(part * 2 + 0) * C::BM + row] = mx;
      xch[(part * 2 + 1) * C::BM + row] = mn;
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int o = 0; o < C::EPI_WGS; ++o) {
        mx = fmaxf(mx, xch[(o * 2 + 0) * C::BM + row]);
        mn = fminf(mn, xch[(o * 2 + 1) * C::BM + row]);
      }
      const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
      if (m < args.M) {
        if (part == 0) args.d_scale[(size_t)m * (args.N / 128) + blockIdx.x] = __floats2half2_rn(scale, zero);
        uint32_t pk[C::CPT / 8];
#pragma unroll
        for (int i = 0; i < C::CPT; i += 8) {
          uint32_t w = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int ch = i + e, q = ch >> 2, tt = ch & 3;                 // channel -> position 4q + {0,2,1,3}[tt]
            const float v = (tt == 0) ? acc[2 * q].x : (tt == 1) ? acc[2 * q + 1].x : (tt == 2) ? acc[2 * q].y : acc[2 * q + 1].y;
            w |= ((uint32_t)((int)roundf((v + zero) * r_scale) & 0xF)) << (4 * e);
          }
          pk[i / 8] = w;
        }
        uint4* dst = reinterpret_cast<uint4*>(args.d4 + (size_t)m * (args.N / 2) + (n0 + colbase) / 2);
#pragma unroll
        for (int i = 0; i < C::CPT / 32; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_stamp(args, 4);
}

}  // namespace atom
