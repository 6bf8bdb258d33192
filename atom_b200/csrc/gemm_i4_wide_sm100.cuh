// gemm_i4_wide_sm100.cuh -- prefill-shape W4A4 GEMM, 128 x 256 tiles with the token operand in TENSOR MEMORY.
//
// Same contract and arithmetic as gemm_i4_tall_sm100.cuh (bit-identical output; reference:
// /root/reference/kernels/include/GEMM/Dense_layer_gemm_i4_o16.cuh:436-710).  What changes is where the bytes move.
//
// Measured on the 128 x 128 kernel (profiles/r02_gemm_v5_pipeline_trace.jsonl): a stage of two groups takes
// ~2500 cycles against 512 of tensor work, and the instruction-count cut of the packed epilogue bought only 7 % -- the
// kernel is bound by SHARED-MEMORY BANDWIDTH (128 B/clk/SM).  tcgen05.mma with both operands in shared memory reads
// 8 KB per 128x128x32 MMA = 128 B/clk by itself; add the TMA writes, the converter's reads and its expanded stores and a
// group moves 96 KB through a port that passes 32 KB in the group's 256 tensor cycles: a 33 % ceiling.
// Here, per 128 x 256 x 128 group (512 tensor cycles):
//   * the 128 token rows are expanded by one thread each (4 x LDS.128 -> 48 ALU -> one tcgen05.st) into a tensor-memory
//     operand slot and used by BOTH 128-column halves of the tile: no expanded store, no operand fetch through shared
//     memory for the M side;
//   * only the weight halves go through shared memory (TMA 16 KB in, converter 16 KB read + 32 KB write, MMA 32 KB read):
//     112 KB per 512 cycles = 219 B/clk -> a 58 % ceiling, up from 33 %.
// Tensor memory: 3 operand slots (96 columns) + one 128-column accumulator per half (biased INT32, see the tall kernel).
// Two MMA issuers (one per half) share the per-stage waits; epilogue threads keep 2 x 64 accumulators (setmaxnreg moves
// registers from the service and converter warpgroups to the two epilogue warpgroups).
#pragma once
#include "gemm_i4_tall_sm100.cuh"
#include "gemm_i4_skinny_sm100.cuh"   // umma_i8_ts, tmem_st_32x32b_x32

namespace atom {

template <bool kO4>
struct WideCfg {
  static constexpr int BM = 128, BH = 128, NH = 2;               // tokens, channels per half, halves
  static constexpr int R = 3;                                     // operand stages (one quantisation group each)
  static constexpr int A_COL0 = 0, D_COL0 = 128;                  // TMEM: [0, 96) token operand slots, [128, 384) accumulators
  static constexpr int TMEM_COLS = 512;
  static constexpr int SCALE_STAGES = 4, SCALE_BYTES = 768;       // per group: 256 B token words + 2 x (128 even + 128 odd halves)
  static constexpr int THREADS = 512;
  static constexpr int PACK_T = 128 * 64, EXP_T = 128 * 128;
  static constexpr int OFF_EXP_B = 0;                             // [R][NH] expanded weight halves (also the keeper's)
  static constexpr int OFF_KEEP_A = OFF_EXP_B + R * NH * EXP_T;   // INT8 keeper tokens (SWIZZLE_128B operand)
  static constexpr int OFF_PACK_A = OFF_KEEP_A + EXP_T;           // [R] packed token tiles (SWIZZLE_64B)
  static constexpr int OFF_PACK_B = OFF_PACK_A + R * PACK_T;      // [R][NH] packed weight halves
  static constexpr int OFF_SM = OFF_PACK_B + R * NH * PACK_T;
  static constexpr int OFF_BAR = OFF_SM + SCALE_STAGES * SCALE_BYTES;
  static constexpr int NUM_BARS = 4 * R + 2 * NH + 2 * SCALE_STAGES;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};


template <bool kO4>
__global__ void __launch_bounds__(WideCfg<kO4>::THREADS, 1)
gemm_i4_wide_kernel(const __grid_constant__ CUtensorMap tm_a4,   // packed INT4 tokens   (2-D, box 64 B x 128 rows, SWIZZLE_64B)
                    const __grid_constant__ CUtensorMap tm_b4,   // packed INT4 weights  (4-D quad-swapped view, box 64 B x 2 x 2 x 32)
                    const __grid_constant__ CUtensorMap tm_a8,   // INT8 keeper tokens   (2-D, box 128 B x 128 rows, SWIZZLE_128B)
                    const __grid_constant__ CUtensorMap tm_b8,   // INT8 keeper weights  (4-D quad-swapped view, SWIZZLE_128B)
                    const GemmArgs args) {
  using C = WideCfg<kO4>;
  constexpr int R = C::R;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;                        // TMA landed the group's packed tiles                 (1 + tx)
  uint64_t* pack_empty = pack_full + R;              // converters have read them                           (4)
  uint64_t* ab_full = pack_empty + R;                // the group's operands are in place (TMEM + smem)     (4 [+ tx for the keeper])
  uint64_t* ab_empty = ab_full + R;                  // both halves' MMAs on the stage completed            (2, tcgen05.commit)
  uint64_t* mma_done = ab_empty + R;                 // half h: accumulators of the current group ready     (1)
  uint64_t* tmem_empty = mma_done + C::NH;           // half h: accumulator read and re-armed               (8)
  uint64_t* scale_full = tmem_empty + C::NH;         //                                                     (32)
  uint64_t* scale_empty = scale_full + C::SCALE_STAGES;   //                                                (8)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * C::BM, n0 = blockIdx.x * (C::NH * C::BH);
  const int nh = (n0 + C::BH < args.N) ? 2 : 1;      // the last tile of an N that is not a multiple of 256 has one half
  const int groups = args.G + 1;                     // index G = INT8 keeper
  if (threadIdx.x == 0) { griddep_launch_dependents(); trace_stamp(args, 0); }

  // part: 1 = weight halves (independent of the preceding kernel), 2 = token tile
  auto issue_group = [&](int g, int part) {
    const int ps = g % R;
    if (part & 1) {
      mbar_arrive_expect_tx(&pack_full[ps], (1 + nh) * C::PACK_T);
      for (int h = 0; h < nh; ++h)
        tma_load_4d(smem + C::OFF_PACK_B + (ps * C::NH + h) * C::PACK_T, &tm_b4, &pack_full[ps], g * 64, 0, 0, (n0 + h * C::BH) / 4);
    }
    if (part & 2) {
      tma_load_2d(smem + C::OFF_PACK_A + ps * C::PACK_T, &tm_a4, &pack_full[ps], g * 64, m0);
      if (g < 8) trace_stamp(args, 8 + g);
    }
  };
  const int first = min(R, args.G);                  // INT4 groups whose weight tiles are issued before the dependency resolves
  if (warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&tm_a4); tma_prefetch_desc(&tm_b4);
    for (int i = 0; i < R; ++i) mbar_init(&pack_full[i], 1);
    fence_barrier_init();
    for (int g = 0; g < first; ++g) issue_group(g, 1);
    tma_prefetch_desc(&tm_a8); tma_prefetch_desc(&tm_b8);
    for (int i = 0; i < R; ++i) { mbar_init(&pack_empty[i], 4); mbar_init(&ab_full[i], 4); mbar_init(&ab_empty[i], 2); }
    for (int i = 0; i < C::NH; ++i) { mbar_init(&mma_done[i], 1); mbar_init(&tmem_empty[i], 8); }
    for (int i = 0; i < C::SCALE_STAGES; ++i) { mbar_init(&scale_full[i], 32); mbar_init(&scale_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) trace_stamp(args, 1);

  if (warp < 4) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ============================================================ TMA producer (INT4 groups; the keeper is loaded by a converter)
      griddep_wait();                                // token tiles are the preceding kernel's output
      if (elect_one_sync()) { for (int g = 0; g < first; ++g) issue_group(g, 2); }
      __syncwarp();
      for (int g = first; g < args.G; ++g) {
        mbar_wait(&pack_empty[g % R], ((g / R) & 1) ^ 1);
        if (elect_one_sync()) issue_group(g, 3);
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      // ============================================================ MMA issuers: warp 1 -> channels [0,128), warp 2 -> [128,256)
      const int h = warp - 1;
      constexpr uint32_t idesc = umma_idesc_i8(C::BM, C::BH);
      const uint32_t d_tmem = tmem_base + C::D_COL0 + h * C::BH;
      for (int g = 0; g < groups; ++g) {
        const int st = g % R;
        mbar_wait(&ab_full[st], (g / R) & 1);
        if (h < nh) mbar_wait(&tmem_empty[h], g & 1);          // completion #0 is the initial arming
        tc_fence_after();
        if (elect_one_sync()) {
          if (h == 0 && g < 8) trace_stamp(args, 88 + g);
          if (h < nh) {
            const uint64_t db = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_B + (st * C::NH + h) * C::EXP_T));
            if (g < args.G) {
#pragma unroll
              for (int k = 0; k < 4; ++k)                      // always accumulating onto the bias pattern
                umma_i8_ts(d_tmem, tmem_base + C::A_COL0 + st * 32 + k * 8, db + (uint64_t)(k * 2), idesc, 1u);
            } else {                                           // keeper: both operands from shared memory
              const uint64_t da = umma_desc_k_sw128(smem_u32(smem + C::OFF_KEEP_A));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_i8(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            }
            umma_commit(&mma_done[h]);
          }
          umma_commit(&ab_empty[st]);                          // (an idle second half still releases the stage)
        }
        __syncwarp();
      }
    } else {
      // ============================================================ scale loader.  Group slot (768 B):
      //   [0,256)  (lower, upper) activation-scale words of the 128 token rows (word = (r/16)*8 + r%8)
      //   [256 + 256 h, +128) weight scale of the EVEN channel of the 64 channel pairs of half h; [+128, +256) of the ODD one
      griddep_wait();
      for (int g = 0; g < groups; ++g) {
        const int ss = g % C::SCALE_STAGES;
          const bool keeper = (g == args.G);
        const __half* as_row = keeper ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
        const __half* bs_row = keeper ? args.b_keeper_scale : args.b_scale + (size_t)g * args.ldb_scale;
        uint32_t aw[2] = {0u, 0u};
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int w = lane + 32 * x, blk = w >> 3, i = w & 7;
          if (m0 + 16 * blk + i < args.M) aw[x] = ld_cg_u32(as_row + 64 * (m0 / 16 + blk) + 8 * i);
        }
        uint4 bw = make_uint4(0, 0, 0, 0);               // lane l: channels n0 + 8 l .. + 7 (32 lanes cover the 256 channels)
        if (n0 + 8 * lane < args.N) bw = ld_cg_v4(bs_row + n0 + 8 * lane);
        if (g >= C::SCALE_STAGES) mbar_wait(&scale_empty[ss], ((g / C::SCALE_STAGES) - 1) & 1);
        uint8_t* slot = smem + C::OFF_SM + ss * C::SCALE_BYTES;
        reinterpret_cast<uint32_t*>(slot)[lane] = aw[0];
        reinterpret_cast<uint32_t*>(slot)[lane + 32] = aw[1];
        const int hh = lane >> 4, l16 = lane & 15;
        reinterpret_cast<uint2*>(slot + 256 + 256 * hh)[l16] = make_uint2(__byte_perm(bw.x, bw.y, 0x5410), __byte_perm(bw.z, bw.w, 0x5410));
        reinterpret_cast<uint2*>(slot + 256 + 256 * hh + 128)[l16] = make_uint2(__byte_perm(bw.x, bw.y, 0x7632), __byte_perm(bw.z, bw.w, 0x7632));
        mbar_arrive(&scale_full[ss]);
      }
    }
  } else if (warp < 8) {
    // ============================================================ converters
    reg_dealloc<104>();
    const int wq = warp & 3, row = wq * 32 + lane, t = (warp - 4) * 32 + lane;
    const int xr = (row >> 1) & 3;                       // SWIZZLE_64B: 16-B chunk index ^= address bits [7,9)
    for (int g = 0; g < groups; ++g) {
      const int st = g % R;
      if (g >= R) mbar_wait(&ab_empty[st], ((g / R) - 1) & 1);           // both halves' MMAs on this stage completed
      if (t == 0 && g < 8) trace_stamp(args, 24 + g);
      if (g < args.G) {
        mbar_wait(&pack_full[st], (g / R) & 1);
        if (t == 0 && g < 8) trace_stamp(args, 40 + g);
        // token row -> tensor memory (chunk x = 32 consecutive K: four "even element" words, then four "odd element" words;
        // convert_tile below applies the same K permutation to the weight rows)
        const uint8_t* prow = smem + C::OFF_PACK_A + st * C::PACK_T + row * 64;
        uint4 w[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) w[x] = *reinterpret_cast<const uint4*>(prow + ((x ^ xr) << 4));
        uint32_t r[32];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 lo, hi;
          expand_chunk(w[x], lo, hi);
          r[8 * x + 0] = lo.x; r[8 * x + 1] = lo.y; r[8 * x + 2] = lo.z; r[8 * x + 3] = lo.w;
          r[8 * x + 4] = hi.x; r[8 * x + 5] = hi.y; r[8 * x + 6] = hi.z; r[8 * x + 7] = hi.w;
        }
        tmem_st_32x32b_x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(C::A_COL0 + st * 32), r);
        for (int h = 0; h < nh; ++h)
          convert_tile<128, 128>(smem + C::OFF_PACK_B + (st * C::NH + h) * C::PACK_T, smem + C::OFF_EXP_B + (st * C::NH + h) * C::EXP_T, t);
        tmem_st_wait();
        if (t == 0 && g < 8) trace_stamp(args, 56 + g);
        fence_proxy_async_smem();        // generic-proxy stores -> visible to tcgen05.mma operand fetch
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&pack_empty[st]); mbar_arrive(&ab_full[st]); }
      } else {
        // keeper: TMA straight into the operand layouts (token rows for the SS-form MMAs, weight halves into the stage)
        __syncwarp();
        if (lane == 0) {
          if (t == 0) {
            griddep_wait();
            mbar_arrive_expect_tx(&ab_full[st], (1 + nh) * C::EXP_T);
            tma_load_2d(smem + C::OFF_KEEP_A, &tm_a8, &ab_full[st], 0, m0);
            for (int h = 0; h < nh; ++h)
              tma_load_4d(smem + C::OFF_EXP_B + (st * C::NH + h) * C::EXP_T, &tm_b8, &ab_full[st], 0, 0, 0, (n0 + h * C::BH) / 4);
          } else {
            mbar_arrive(&ab_full[st]);
          }
        }
      }
      if (t == 0 && g < 8) trace_stamp(args, 72 + g);
    }
  } else {
    // ============================================================ epilogue warpgroups: thread = token row, 64 columns of each half
    reg_alloc<176>();
    const int wq = warp & 3, row = wq * 32 + lane;
    const int colbase = ((warp - 8) >> 2) * 64;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + C::D_COL0;
#pragma unroll
    for (int h = 0; h < C::NH; ++h)
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) tmem_st_const_x16(lane_addr + h * C::BH + colbase + c0, kAccBias);
    tmem_st_wait_();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(&tmem_empty[0]); mbar_arrive(&tmem_empty[1]); }

    // position k of a half's 64 columns: TMEM column colbase + k = channel colbase + (k & ~3) + {0,2,1,3}[k & 3]
    float2 acc[C::NH][32];
#pragma unroll
    for (int h = 0; h < C::NH; ++h)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[h][i] = make_float2(0.f, 0.f);
    const bool upper = (((m0 + row) & 15) >= 8);

    for (int g = 0; g < groups; ++g) {
      const int ss = g % C::SCALE_STAGES;
      const float mul = (g == args.G) ? 256.f : 1.f, nbias = -12582912.f * mul;   // keeper: exact sums x 2^8 (no 16 * 16 operand factor)
      mbar_wait(&scale_full[ss], (g / C::SCALE_STAGES) & 1);
      const uint8_t* slot = smem + C::OFF_SM + ss * C::SCALE_BYTES;
      const __half2 pw = reinterpret_cast<const __half2*>(slot)[(row >> 4) * 8 + (row & 7)];
      const __half2 sm2 = __half2half2(upper ? __high2half(pw) : __low2half(pw));
      // 2 halves x 4 chunks of 16 columns, software-pipelined: the tcgen05.ld of the next chunk is in flight while the
      // current one is dequantised (an exposed TMEM round trip per chunk costs ~16 cycles per element otherwise)
      uint32_t rbuf[2][16];
      mbar_wait(&mma_done[0], g & 1);
      tc_fence_after();
      if (warp == 8 && lane == 0 && g < 8) trace_stamp(args, 104 + g);
      tmem_ld_32x32b_x16(lane_addr + (uint32_t)colbase, rbuf[0]);
#pragma unroll
      for (int h = 0; h < C::NH; ++h) {
        if (h < nh) {
          const uint4* sel = reinterpret_cast<const uint4*>(slot + 256 + 256 * h + (upper ? 128 : 0) + colbase);
          const uint32_t taddr = lane_addr + (uint32_t)(h * C::BH + colbase);
#pragma unroll
          for (int c = 0; c < 4; ++c) {                          // 16 columns = 8 channel pairs = one LDS.128
            const int n = h * 4 + c;
            const uint4 sv = sel[c];
            tmem_ld_wait();                                      // chunk n has landed in rbuf[n & 1]
            if (c < 3) {
              tmem_ld_32x32b_x16(taddr + 16 * (c + 1), rbuf[(n + 1) & 1]);
            } else if (h + 1 < nh) {                             // first chunk of the other half: its MMAs must be complete
              mbar_wait(&mma_done[h + 1], g & 1);
              tc_fence_after();
              tmem_ld_32x32b_x16(lane_addr + (uint32_t)((h + 1) * C::BH + colbase), rbuf[(n + 1) & 1]);
            }
            tmem_st_const_x16(taddr + 16 * c, kAccBias);      // re-arm the chunk just read
            if (c == 3) {                                        // this half's accumulator is read and re-armed
              tmem_st_wait_();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[h]);
            }
            const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 rs = __half22float2(__hmul2(sm2, *reinterpret_cast<const __half2*>(&sw[q])));
              const int k = 16 * c + 4 * q;
              ffma2(acc[h][(k >> 1) + 0], unbias2(rbuf[n & 1][4 * q + 0], rbuf[n & 1][4 * q + 1], mul, nbias), rs);
              ffma2(acc[h][(k >> 1) + 1], unbias2(rbuf[n & 1][4 * q + 2], rbuf[n & 1][4 * q + 3], mul, nbias), rs);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&scale_empty[ss]);
      if (warp == 8 && lane == 0 && g < 8) trace_stamp(args, 120 + g);
    }
    if (warp == 8 && lane == 0) trace_stamp(args, 2);
    griddep_wait();                                        // the output buffer may still be read by the preceding kernel

    constexpr float kInv = 1.0f / 256.0f;   // exact: removes the 16*16 operand factor
    const int m = m0 + row;
#pragma unroll
    for (int h = 0; h < C::NH; ++h) {
      if (h >= nh) continue;
      const int nb = n0 + h * C::BH + colbase;               // first channel of this thread's 64 columns
      if constexpr (!kO4) {
        if (m < args.M) {
          __half* drow = args.d + (size_t)m * args.N + nb;
#pragma unroll
          for (int i = 0; i < 64; i += 8) {
            if (nb + i < args.N) {   // N is a multiple of 8 (16-B rows)
              const int q = i >> 2;
              uint4 v;
              const __half2 h0 = __floats2half2_rn(acc[h][2 * q].x * kInv, acc[h][2 * q + 1].x * kInv);
              const __half2 h1 = __floats2half2_rn(acc[h][2 * q].y * kInv, acc[h][2 * q + 1].y * kInv);
              const __half2 h2 = __floats2half2_rn(acc[h][2 * q + 2].x * kInv, acc[h][2 * q + 3].x * kInv);
              const __half2 h3 = __floats2half2_rn(acc[h][2 * q + 2].y * kInv, acc[h][2 * q + 3].y * kInv);
              v.x = *reinterpret_cast<const uint32_t*>(&h0); v.y = *reinterpret_cast<const uint32_t*>(&h1);
              v.z = *reinterpret_cast<const uint32_t*>(&h2); v.w = *reinterpret_cast<const uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(drow + i) = v;
            }
          }
        }
      } else {
        // o4 (DenseLayerGEMM_i4_o4.cu:705-787): one 128-channel head per half; the two warpgroups hold 64 columns each
        float* xch = reinterpret_cast<float*>(smem + C::OFF_PACK_A) + h * 4 * C::BM;   // packed ring is idle by now
        float mx = -INFINITY, mn = INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[h][i].x *= kInv; acc[h][i].y *= kInv;
          const float a0 = fabsf(acc[h][i].x), a1 = fabsf(acc[h][i].y);
          mx = fmaxf(mx, fmaxf(a0, a1)); mn = fminf(mn, fminf(a0, a1));
        }
        const int part = (warp - 8) >> 2;
        xch[(part * 2 + 0) * C::BM + row] = mx;
        xch[(part * 2 + 1) * C::BM + row] = mn;
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int o = 0; o < 2; ++o) { mx = fmaxf(mx, xch[(o * 2 + 0) * C::BM + row]); mn = fminf(mn, xch[(o * 2 + 1) * C::BM + row]); }
        const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
        if (m < args.M) {
          if (part == 0) args.d_scale[(size_t)m * (args.N / 128) + (n0 / 128 + h)] = __floats2half2_rn(scale, zero);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 64; i += 8) {
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ch = i + e, q = ch >> 2, tt = ch & 3;                 // channel -> position 4q + {0,2,1,3}[tt]
              const float v = (tt == 0) ? acc[h][2 * q].x : (tt == 1) ? acc[h][2 * q + 1].x : (tt == 2) ? acc[h][2 * q].y : acc[h][2 * q + 1].y;
              w |= ((uint32_t)((int)roundf((v + zero) * r_scale) & 0xF)) << (4 * e);
            }
            pk[i / 8] = w;
          }
          uint4* dst = reinterpret_cast<uint4*>(args.d4 + (size_t)m * (args.N / 2) + nb / 2);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
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
