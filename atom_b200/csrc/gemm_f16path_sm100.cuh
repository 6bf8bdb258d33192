// gemm_f16path_sm100.cuh -- EXPERIMENTAL (round-2 work item, compiled but only reachable with ATOM_GEMM_FP16_PATH;
// not yet run on hardware): the prefill-sized W4A4 GEMM through the FP16 tensor path.
//
// Why: with INT8 MMAs (gemm_i4_sm100.cuh) the per-group dequantisation has to happen on the INT32 accumulators -- about 3.3
// CUDA-core instructions per accumulator element per 128-wide group (I2FP, FFMA, HMUL2/cvt of the scale product), 420 issue
// cycles per 128x128x128 group against 256 tensor cycles: the SM's schedulers, not the tensor pipe, pace the kernel
// (measured 0.74 POP/s = 16 % of the INT8 peak, slower than the reference's RTX 4090 number).
// Here the scales are applied to the OPERANDS instead: the converter warps turn the packed INT4 (and the INT8 keeper)
// into  a' = fp16(a * sA[m,g])  and  b' = fp16(b * sB[n,g] * 2^8)  while they expand them into the K-major SWIZZLE_128B operand
// layout, tcgen05.mma.kind::f16 accumulates the WHOLE K range in FP32 in tensor memory, and the epilogue runs once per
// tile (x 2^-8, cast, store).  Each product a'*b' is exact in the FP32 accumulator; what is lost is the rounding of a' and b'
// to 11 bits (2^-12 relative each, zero-mean): ~3e-4 of the output scale at K = 4096, inside the 1e-3 the operator contract
// allows (include/atom_b200.h) but NOT bit-identical to the reference -- hence a flag, not the default.
// The FP16 tensor rate is half the INT8 rate, but nothing is left on the CUDA cores per group except the conversion
// (15 instructions per 8 elements, w4_f16_convert.cuh -- unit-tested on the host), which a 128 x 256 tile amortises.
//
// Pipeline (same roles and barrier idiom as the INT8 kernel, which is validated on the B200):
//   warp 0  TMA producer: packed "units" of 64 B per row (one INT4 group, or half of the INT8 keeper) -> packed ring
//   warp 3  scale loader: cp.async of the group's raw scale rows -> scale ring (mbarrier-counted)
//   warps 4..: converters: one operand-ring STAGE = 64 K-elements = one 128-B swizzled row per operand row
//           (an INT4 unit feeds two stages, a keeper unit one); fence.proxy.async; arrive exp_full
//   warp 1  MMA issuer: 4 x tcgen05.mma.kind::f16 (128 x BN x 16) per stage into ONE accumulator, commit -> slot_free
//   converters again: after the last commit, drain TMEM (tcgen05.ld), x 2^-8, FP16 store.
#pragma once
#include "gemm_i4_sm100.cuh"
#include "w4_f16_convert.cuh"

namespace atom {

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (format 0), both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kBDirect: the weights were expanded once to FP16 in HBM (expand_weights_f16_kernel below; same element order and 2^8 shift
// as the in-kernel conversion) and are TMA'd straight into the operand ring -- only the activations are converted.
template <int BN, int kPack, int kRing, int kConvWarps, bool kBDirect = false>
struct F16Cfg {
  static constexpr int BM = 128;
  static constexpr int EXP_A = BM * 128, EXP_B = BN * 128;        // one stage: 64 fp16 = 128 B per row
  static constexpr int PACK_A = BM * 64, PACK_B = kBDirect ? 0 : BN * 64;   // one packed unit: 64 B per row
  static constexpr int SCALE_STAGES = 4;                          // scale ring depth, in groups
  static constexpr int SCALE_SLOT = 256 + BN * 2;                 // 64 (lower, upper) A-scale words | BN B-scale halves
  static constexpr int CONV_THREADS = 32 * kConvWarps;
  static constexpr int THREADS = 128 + CONV_THREADS;
  static constexpr int EPI_SLICES = kConvWarps / 4;               // converter warps double as the epilogue
  static constexpr int CPT = BN / EPI_SLICES;                     // accumulator columns per epilogue thread
  static constexpr int TMEM_COLS = BN;
  static constexpr int OFF_EXP_A = 0;
  static constexpr int OFF_EXP_B = OFF_EXP_A + kRing * EXP_A;
  static constexpr int OFF_PACK_A = OFF_EXP_B + kRing * EXP_B;
  static constexpr int OFF_PACK_B = OFF_PACK_A + kPack * PACK_A;
  static constexpr int OFF_SM = OFF_PACK_B + kPack * PACK_B;
  static constexpr int OFF_BAR = OFF_SM + SCALE_STAGES * SCALE_SLOT;
  static constexpr int NUM_BARS = 2 * kPack + 2 * kRing + 2 * SCALE_STAGES + 1;
  static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;     // + slack for the 1024-B alignment fix-up
  static constexpr float OUT_SCALE = 1.0f / 256.0f;               // undoes the 2^8 carried by the B operand
  static_assert(BN == 128 || BN == 256, "TMEM allocation is a power of two; tcgen05.mma N <= 256");
  static_assert(kConvWarps % 4 == 0 && CPT % 32 == 0, "epilogue: each warp reads its lane quarter in 32-column steps");
  static_assert(SCALE_SLOT % 16 == 0 && OFF_SM % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(THREADS <= 1024, "block size");
};

template <int BN, int kPack, int kRing, int kConvWarps, bool kBDirect = false, bool kO4 = false>
__global__ void __launch_bounds__(F16Cfg<BN, kPack, kRing, kConvWarps, kBDirect>::THREADS, 1)
gemm_w4a4_f16path_kernel(const __grid_constant__ CUtensorMap tm_a4,   // packed INT4 activations  (box 64 B x 128 rows)
                         const __grid_constant__ CUtensorMap tm_b4,   // packed INT4 weights      (box 64 B x BN rows)
                                                                      //   kBDirect: expanded FP16 weights (box 128 B x BN rows, SW128)
                         const __grid_constant__ CUtensorMap tm_a8,   // INT8 keeper, activations (box 64 B x 128 rows, no swizzle)
                         const __grid_constant__ CUtensorMap tm_b8,   // INT8 keeper, weights     (box 64 B x BN rows, no swizzle; unused if kBDirect)
                         const GemmArgs args) {
  using C = F16Cfg<BN, kPack, kRing, kConvWarps, kBDirect>;
  static_assert(!kO4 || (BN == 128 && C::CPT == 32), "o4: one 128-channel head per tile, 4 column slices of 32");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* pack_full = bars;                          // TMA landed a packed unit                         (1 + tx)
  uint64_t* pack_empty = pack_full + kPack;            // converters have consumed it                      (kConvWarps)
  uint64_t* exp_full = pack_empty + kPack;             // a stage's FP16 operands are in place             (kConvWarps)
  uint64_t* slot_free = exp_full + kRing;              // the stage's MMAs completed: slot reusable        (1, tcgen05.commit)
  uint64_t* scale_full = slot_free + kRing;            // a group's scales landed                          (32, cp.async noinc)
  uint64_t* scale_empty = scale_full + C::SCALE_STAGES;   // converters are done with them                 (kConvWarps)
  uint64_t* acc_ready = scale_empty + C::SCALE_STAGES;    // every MMA of the tile completed               (1, tcgen05.commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * C::BM;
  const int G = args.G;                      // INT4 groups; units G, G+1 are the two halves of the INT8 keeper
  const int nunits = G + 2;
  const int nstages = 2 * G + 2;             // an INT4 unit feeds two 64-element stages, a keeper unit one
  auto unit_of = [&](int t) { return t < 2 * G ? (t >> 1) : G + (t - 2 * G); };

  // ---------------------------------------------------------------- setup (as in the INT8 kernel: first loads before the sync)
  auto issue_unit = [&](int u) {
    const int ps = u % kPack;
    mbar_arrive_expect_tx(&pack_full[ps], C::PACK_A + C::PACK_B);
    uint8_t* pa = smem + C::OFF_PACK_A + ps * C::PACK_A;
    uint8_t* pb = smem + C::OFF_PACK_B + ps * C::PACK_B;
    if (u < G) {
      tma_load_2d(pa, &tm_a4, &pack_full[ps], u * 64, m0);
      if constexpr (!kBDirect) tma_load_2d(pb, &tm_b4, &pack_full[ps], u * 64, n0);
    } else {
      tma_load_2d(pa, &tm_a8, &pack_full[ps], (u - G) * 64, m0);
      if constexpr (!kBDirect) tma_load_2d(pb, &tm_b8, &pack_full[ps], (u - G) * 64, n0);
    }
  };
  int u_issued = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a4); tma_prefetch_desc(&tm_b4);
    for (int i = 0; i < kPack; ++i) mbar_init(&pack_full[i], 1);
    fence_barrier_init();
    for (; u_issued < kPack && u_issued < nunits; ++u_issued) issue_unit(u_issued);
    tma_prefetch_desc(&tm_a8);
    if constexpr (!kBDirect) tma_prefetch_desc(&tm_b8);
    for (int i = 0; i < kPack; ++i) mbar_init(&pack_empty[i], kConvWarps);
    for (int i = 0; i < kRing; ++i) { mbar_init(&exp_full[i], kConvWarps); mbar_init(&slot_free[i], 1); }
    for (int i = 0; i < C::SCALE_STAGES; ++i) { mbar_init(&scale_full[i], 32); mbar_init(&scale_empty[i], kConvWarps); }
    mbar_init(acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      for (int u = u_issued; u < nunits; ++u) {
        mbar_wait(&pack_empty[u % kPack], ((u / kPack) & 1) ^ 1);
        issue_unit(u);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: one accumulator for the whole K range
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(C::BM, BN);
      for (int t = 0; t < nstages; ++t) {
        const int es = t % kRing;
        mbar_wait(&exp_full[es], (t / kRing) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_A + es * C::EXP_A));
        const uint64_t db = umma_desc_k_sw128(smem_u32(smem + C::OFF_EXP_B + es * C::EXP_B));
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 4 x K=16 halves = 4 x 32 B inside the 128-B swizzle atom
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (t > 0 || k > 0) ? 1u : 0u);
        umma_commit(&slot_free[es]);
      }
      umma_commit(acc_ready);
    }
  } else if (warp == 3) {
    // ============================================================ scale loader (groups 0..G-1, then the keeper's scales)
    for (int g = 0; g <= G; ++g) {
      const int ss = g % C::SCALE_STAGES;
      if (g >= C::SCALE_STAGES) mbar_wait(&scale_empty[ss], ((g / C::SCALE_STAGES) - 1) & 1);
      const bool keeper = (g == G);
      const __half* as_row = keeper ? args.a_keeper_scale : args.a_scale + (size_t)g * args.lda_scale;
      const __half* bs_row = keeper ? args.b_keeper_scale : args.b_scale + (size_t)g * args.N;
      uint8_t* slot = smem + C::OFF_SM + ss * C::SCALE_SLOT;
#pragma unroll
      for (int w = lane; w < C::BM / 2; w += 32) {            // A-scale words: rows (16 blk + i, 16 blk + i + 8)
        const int blk = w >> 3, i = w & 7;
        if (m0 + 16 * blk + i < args.M) cp_async_4(slot + w * 4, as_row + 64 * (m0 / 16 + blk) + 8 * i);
      }
      if constexpr (!kBDirect) {
#pragma unroll
        for (int c = lane; c < BN / 8; c += 32)                // B-scale: 8 channels per 16-B chunk
          if (n0 + 8 * c < args.N) cp_async_16(slot + 256 + c * 16, bs_row + n0 + 8 * c);
      }
      cp_async_mbar_arrive_noinc(&scale_full[ss]);
    }
  } else if (warp >= 4) {
    // ============================================================ converter warps (and, at the end, the epilogue)
    const int cw = warp - 4;
    const int t_id = cw * 32 + lane;
    const __half2 shift_b = __float2half2_rn(256.0f);
    for (int t = 0; t < nstages; ++t) {
      const int es = t % kRing;
      const int u = unit_of(t), ps = u % kPack;
      const bool is_i4 = t < 2 * G;
      const int half_sel = is_i4 ? (t & 1) : 0;
      const int sg = is_i4 ? u : G, ss = sg % C::SCALE_STAGES;
      const bool first_of_unit = !is_i4 || half_sel == 0, last_of_unit = !is_i4 || half_sel == 1;
      const bool first_of_group = is_i4 ? half_sel == 0 : (u == G), last_of_group = is_i4 ? half_sel == 1 : (u == G + 1);
      if (t >= kRing) mbar_wait(&slot_free[es], ((t / kRing) - 1) & 1);     // the MMAs that read this slot have completed
      if constexpr (kBDirect) {
        // stage t of the expanded weights = bytes [128 t, 128 t + 128) of every W' row (both use the converter's K order);
        // the load counts on exp_full next to the converter warps' arrivals and overlaps the A conversion below
        if (t_id == 0) {
          mbar_expect_tx(&exp_full[es], C::EXP_B);
          tma_load_2d(smem + C::OFF_EXP_B + es * C::EXP_B, &tm_b4, &exp_full[es], t * 128, n0);
        }
      }
      if (first_of_unit) mbar_wait(&pack_full[ps], (u / kPack) & 1);
      if (first_of_group) mbar_wait(&scale_full[ss], (sg / C::SCALE_STAGES) & 1);
      const uint8_t* slot = smem + C::OFF_SM + ss * C::SCALE_SLOT;
      const __half2* sa_words = reinterpret_cast<const __half2*>(slot);
      const __half* sb_halves = reinterpret_cast<const __half*>(slot + 256);
      const uint8_t* pa = smem + C::OFF_PACK_A + ps * C::PACK_A;
      const uint8_t* pb = smem + C::OFF_PACK_B + ps * C::PACK_B;
      uint8_t* ea = smem + C::OFF_EXP_A + es * C::EXP_A;
      uint8_t* eb = smem + C::OFF_EXP_B + es * C::EXP_B;
      if (is_i4) {
        // thread = (row r0 + RP*k, word j of the 32-B half row), j and r0 fixed per thread: 8 nibbles -> one 16-B chunk of
        // the 128-B expanded row.  Every address is a per-thread base plus a compile-time multiple of k.
        constexpr int RP = C::CONV_THREADS / 8;                      // rows per pass
        const int j = t_id & 7, r0 = t_id >> 3;
        const uint32_t src_off = (uint32_t)(r0 * 64 + half_sel * 32 + j * 4), dst_off = sw128_chunk_offset(r0, j);
        static_assert(RP % 16 == 0 && C::BM % RP == 0 && BN % RP == 0, "row passes keep r%16 and the swizzle phase fixed");
        {
          uint32_t w[C::BM / RP];
#pragma unroll
          for (int k = 0; k < C::BM / RP; ++k) w[k] = *reinterpret_cast<const uint32_t*>(pa + src_off + k * (RP * 64));
#pragma unroll
          for (int k = 0; k < C::BM / RP; ++k) {
            const __half2 pw = sa_words[((r0 >> 4) + k * (RP / 16)) * 8 + (r0 & 7)];
            const __half2 s2 = __half2half2((r0 & 8) ? __high2half(pw) : __low2half(pw));
            __half2 o[4];
            nib8_to_f16(w[k], s2, o);
            *reinterpret_cast<uint4*>(ea + dst_off + k * (RP * 128)) =
                make_uint4(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]),
                           *reinterpret_cast<uint32_t*>(&o[2]), *reinterpret_cast<uint32_t*>(&o[3]));
          }
        }
        if constexpr (!kBDirect) {
          uint32_t w[BN / RP];
#pragma unroll
          for (int k = 0; k < BN / RP; ++k) w[k] = *reinterpret_cast<const uint32_t*>(pb + src_off + k * (RP * 64));
#pragma unroll
          for (int k = 0; k < BN / RP; ++k) {
            const __half2 s2 = __hmul2(__half2half2(sb_halves[r0 + k * RP]), shift_b);
            __half2 o[4];
            nib8_to_f16(w[k], s2, o);
            *reinterpret_cast<uint4*>(eb + dst_off + k * (RP * 128)) =
                make_uint4(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]),
                           *reinterpret_cast<uint32_t*>(&o[2]), *reinterpret_cast<uint32_t*>(&o[3]));
          }
        }
      } else {
        // keeper: thread = (row r0 + RP*k, word j of the 64-B half row): 4 int8 -> 8 B of the expanded row
        constexpr int RP = C::CONV_THREADS / 16;
        const int j = t_id & 15, r0 = t_id >> 4;
        const uint32_t src_off = (uint32_t)(r0 * 64 + j * 4), dst_off = sw128_chunk_offset(r0, j >> 1) + (uint32_t)((j & 1) * 8);
        static_assert(RP % 16 == 0 && C::BM % RP == 0 && BN % RP == 0, "row passes keep r%16 and the swizzle phase fixed");
#pragma unroll
        for (int k = 0; k < C::BM / RP; ++k) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(pa + src_off + k * (RP * 64));
          const __half2 pw = sa_words[((r0 >> 4) + k * (RP / 16)) * 8 + (r0 & 7)];
          const __half2 s2 = __half2half2((r0 & 8) ? __high2half(pw) : __low2half(pw));
          __half2 o[2];
          i8x4_to_f16(w, s2, o);
          *reinterpret_cast<uint2*>(ea + dst_off + k * (RP * 128)) = make_uint2(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]));
        }
        if constexpr (!kBDirect) {
#pragma unroll
          for (int k = 0; k < BN / RP; ++k) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(pb + src_off + k * (RP * 64));
            const __half2 s2 = __hmul2(__half2half2(sb_halves[r0 + k * RP]), shift_b);
            __half2 o[2];
            i8x4_to_f16(w, s2, o);
            *reinterpret_cast<uint2*>(eb + dst_off + k * (RP * 128)) = make_uint2(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]));
          }
        }
      }
      fence_proxy_async_smem();          // generic-proxy stores -> visible to the tcgen05.mma operand fetch
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&exp_full[es]);
        if (last_of_unit) mbar_arrive(&pack_empty[ps]);
        if (last_of_group) mbar_arrive(&scale_empty[ss]);
      }
    }

    // ------------------------------------------------------------ epilogue: once per tile
    mbar_wait(acc_ready, 0);
    tc_fence_after();
    const int wq = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane, m = m0 + row;
    const int colbase = (cw >> 2) * C::CPT;
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)colbase;
    if constexpr (!kO4) {
#pragma unroll 1
      for (int c0 = 0; c0 < C::CPT; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        if (m < args.M) {
          __half* drow = args.d + (size_t)m * args.N + n0 + colbase + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            if (n0 + colbase + c0 + i < args.N) {       // N is a multiple of 8 (16-B rows)
              const __half2 h0 = __floats2half2_rn(__uint_as_float(r[i + 0]) * C::OUT_SCALE, __uint_as_float(r[i + 1]) * C::OUT_SCALE);
              const __half2 h1 = __floats2half2_rn(__uint_as_float(r[i + 2]) * C::OUT_SCALE, __uint_as_float(r[i + 3]) * C::OUT_SCALE);
              const __half2 h2 = __floats2half2_rn(__uint_as_float(r[i + 4]) * C::OUT_SCALE, __uint_as_float(r[i + 5]) * C::OUT_SCALE);
              const __half2 h3 = __floats2half2_rn(__uint_as_float(r[i + 6]) * C::OUT_SCALE, __uint_as_float(r[i + 7]) * C::OUT_SCALE);
              uint4 v;
              v.x = *reinterpret_cast<const uint32_t*>(&h0); v.y = *reinterpret_cast<const uint32_t*>(&h1);
              v.z = *reinterpret_cast<const uint32_t*>(&h2); v.w = *reinterpret_cast<const uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(drow + i) = v;
            }
          }
        }
      }
    } else {
      // o4 (DenseLayerGEMM_i4_o4.cu:705-787): per (token, 128-channel head) asymmetric INT4 with the reference's |v| min/max.
      // The tile is one head; a row's 128 columns sit in 4 warps (same lane quarter, slices of 32): exchange through the
      // packed ring, which is idle by now (every unit was consumed by these very warps).
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr, r);
      tmem_ld_wait();
      float v[32];
      float mx = -INFINITY, mn = INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r[i]) * C::OUT_SCALE;
        const float a = fabsf(v[i]);
        mx = fmaxf(mx, a); mn = fminf(mn, a);
      }
      float* xch = reinterpret_cast<float*>(smem + C::OFF_PACK_A);      // [slice][max | min][128 rows]
      const int slice = cw >> 2;
      xch[(slice * 2 + 0) * C::BM + row] = mx;
      xch[(slice * 2 + 1) * C::BM + row] = mn;
      asm volatile("bar.sync 1, %0;" ::"n"(C::CONV_THREADS) : "memory");
#pragma unroll
      for (int o = 0; o < C::EPI_SLICES; ++o) {
        mx = fmaxf(mx, xch[(o * 2 + 0) * C::BM + row]);
        mn = fminf(mn, xch[(o * 2 + 1) * C::BM + row]);
      }
      const float scale = (mx - mn) / 15.f, zero = -mn, r_scale = 1.f / scale;
      if (m < args.M) {
        if (slice == 0) args.d_scale[(size_t)m * (args.N / 128) + blockIdx.x] = __floats2half2_rn(scale, zero);
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint32_t w = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) w |= ((uint32_t)((int)roundf((v[i + e] + zero) * r_scale) & 0xF)) << (4 * e);
          pk[i / 8] = w;
        }
        *reinterpret_cast<uint4*>(args.d4 + (size_t)m * (args.N / 2) + (n0 + colbase) / 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// One-time weight expansion for kBDirect: W'[n][k'] = fp16(w[n][k] * fp16(sB[g][n] * 2^8)) with k' the converter's element
// order (nib8_to_f16 / i8x4_to_f16 per 32-bit word), row-major [N][K] halves, keeper in the last 128 columns.
// One thread per 32-bit source word; plain coalesced loads and 16-B / 8-B stores (runs once per weight, HBM-bound).
__global__ void __launch_bounds__(256)
expand_weights_f16_kernel(const uint8_t* __restrict__ b, const __half* __restrict__ b_scale, const int8_t* __restrict__ b_keeper,
                          const __half* __restrict__ b_keeper_scale, __half* __restrict__ out, int N, int K) {
  const int words4 = (K - 128) / 8, words = words4 + 32;             // INT4 words + keeper words per row
  const long long total = (long long)N * words;
  const __half2 shift_b = __float2half2_rn(256.0f);
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(id / words), wi = (int)(id % words);
    __half* orow = out + (size_t)n * K;
    if (wi < words4) {
      const uint32_t w = reinterpret_cast<const uint32_t*>(b + (size_t)n * ((K - 128) / 2))[wi];
      const __half2 s2 = __hmul2(__half2half2(b_scale[(size_t)(wi >> 4) * N + n]), shift_b);
      __half2 o[4];
      nib8_to_f16(w, s2, o);
      *reinterpret_cast<uint4*>(orow + wi * 8) = make_uint4(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]),
                                                            *reinterpret_cast<uint32_t*>(&o[2]), *reinterpret_cast<uint32_t*>(&o[3]));
    } else {
      const int j = wi - words4;
      const uint32_t w = reinterpret_cast<const uint32_t*>(b_keeper + (size_t)n * 128)[j];
      const __half2 s2 = __hmul2(__half2half2(b_keeper_scale[n]), shift_b);
      __half2 o[2];
      i8x4_to_f16(w, s2, o);
      *reinterpret_cast<uint2*>(orow + (K - 128) + j * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&o[0]), *reinterpret_cast<uint32_t*>(&o[1]));
    }
  }
}

}  // namespace atom
