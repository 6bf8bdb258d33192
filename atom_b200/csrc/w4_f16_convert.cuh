// w4_f16_convert.cuh -- packed INT4 / INT8 -> scaled FP16 operand conversion for the FP16-path GEMM
// (gemm_f16path_sm100.cuh).  Pure bit tricks + packed half arithmetic, __host__ __device__ so that the exact functions
// the kernel runs are unit-tested on the CPU (tools/host_check_f16conv.cu): no GPU is needed to pin this part.
//
// INT4: a 32-bit word holds 8 two's-complement nibbles e0..e7 (e_i in bits 4i..4i+3, Reorder.cuh:16-19).  For the
// couple (q, q+4):   x = ((w >> 4q) & 0x000F000F ^ 0x00080008) | 0x64006400   is the half2 (1024 + e_q + 8, 1024 + e_{q+4} + 8)
// -- the XOR turns two's complement into offset binary and is folded into one LOP3 with the AND/OR (the OR constant
// carries the XOR bit: f(w,A,B) = A ? (B ? ~w : w) : B, truth table 0x6A).  x - 1032 is exact, the product with the group
// scale is one correctly rounded HMUL2:  out = fp16(e * s).  15 instructions per 8 elements (3 SHF, 4 LOP3, 4 HSUB2, 4 HMUL2).
// Element order inside the 16-byte output chunk: [e0 e4 e1 e5 e2 e6 e3 e7]; both GEMM operands go through this function,
// so the permutation of K cancels in every dot product.
// INT8 (keeper): bytes b0..b3 -> (b ^ 0x80) placed under 0x64 by PRMT = 1024 + b + 128; minus 1152; times scale.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace atom {

template <int kLut>
__host__ __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(kLut));
  return d;
#else
  uint32_t d = 0;
  for (int i = 0; i < 32; ++i) {
    const uint32_t idx = (((a >> i) & 1u) << 2) | (((b >> i) & 1u) << 1) | ((c >> i) & 1u);
    d |= (((uint32_t)kLut >> idx) & 1u) << i;
  }
  return d;
#endif
}

__host__ __device__ __forceinline__ __half2 h2_from_bits(uint32_t u) {
  __half2 h;
  *reinterpret_cast<uint32_t*>(&h) = u;
  return h;
}

// 8 nibbles -> 4 half2 couples (q, q+4), each element = fp16(e * s) with s2 = (s, s)
__host__ __device__ __forceinline__ void nib8_to_f16(uint32_t w, __half2 s2, __half2 (&out)[4]) {
  const __half2 bias = h2_from_bits(0x64086408u);   // 1032 = 1024 + 8
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t x = lop3<0x6A>(w >> (4 * q), 0x000F000Fu, 0x64086408u);
    out[q] = __hmul2(__hsub2(h2_from_bits(x), bias), s2);
  }
}

// 4 int8 -> 2 half2 (b0, b1), (b2, b3), each element = fp16(b * s)
__host__ __device__ __forceinline__ void i8x4_to_f16(uint32_t w, __half2 s2, __half2 (&out)[2]) {
  const uint32_t v = w ^ 0x80808080u;                                 // offset binary: b + 128 in [0, 255]
  const __half2 bias = h2_from_bits(0x64806480u);                     // 1152 = 1024 + 128
#ifdef __CUDA_ARCH__
  const uint32_t lo = __byte_perm(v, 0x64646464u, 0x4140), hi = __byte_perm(v, 0x64646464u, 0x4342);
#else
  const uint32_t lo = (v & 0xFFu) | 0x6400u | ((v & 0xFF00u) << 8) | 0x64000000u;
  const uint32_t hi = ((v >> 16) & 0xFFu) | 0x6400u | ((v >> 24) << 16) | 0x64000000u;
#endif
  out[0] = __hmul2(__hsub2(h2_from_bits(lo), bias), s2);
  out[1] = __hmul2(__hsub2(h2_from_bits(hi), bias), s2);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SWIZZLE_128B operand tile (128-B rows, 8-row atoms of
// 1024 B): the layout tcgen05.mma reads through umma_desc_k_sw128 -- identical to the INT8 kernel's convert_tile addressing.
__host__ __device__ __forceinline__ uint32_t sw128_chunk_offset(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((c) ^ (r & 7)) << 4));
}

}  // namespace atom
