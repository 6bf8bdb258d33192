// microbench.cu -- B200 micro-benchmarks that size the GEMM design:
//   (1) UTCIMMA (tcgen05.mma.kind::i8) issue-rate peak per SM and chip  -> the "INT4 tensor peak" roofline denominator
//   (2) cost per accumulator element of candidate dequant epilogues (I2F+FFMA, +HMUL2/cvt, packed FFMA2, magic-number)
//   (3) nibble->int8 expansion throughput (smem -> smem)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../atom_b200/csrc/ptx_sm100.cuh"
using namespace atom;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------ (1) UTCIMMA peak
template <int BN>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 + BN) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 1 && lane == 0) {
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem)), db = umma_desc_k_sw128(smem_u32(smem + 128 * 128));
    constexpr uint32_t idesc = umma_idesc_i8(128, BN);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_i8(tbase + (it & 1) * BN, da + k * 2, db + k * 2, idesc, k > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ------------------------------------------------------------------ (2) epilogue op mixes
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) epi_kernel(int iters, float* out, unsigned long long* cycles, int seed) {
  int c[32]; float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { c[i] = (threadIdx.x * 37 + i * 101 + seed) << 8; acc[i] = 0.f; }
  __half2 sa = __floats2half2_rn(0.05f + seed, 0.05f + seed), sb = __floats2half2_rn(0.011f, 0.013f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (VARIANT == 0) {          // I2F + FFMA
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf((float)c[i], 1.0001f, acc[i]);
    } else if constexpr (VARIANT == 1) {   // faithful: HMUL2 per 4 elements, cvt per 2, I2F, FFMA
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float2 rs = __half22float2(__hmul2(sa, sb));
        acc[i] = fmaf((float)c[i], rs.x, acc[i]); acc[i + 1] = fmaf((float)c[i + 1], rs.x, acc[i + 1]);
        acc[i + 2] = fmaf((float)c[i + 2], rs.y, acc[i + 2]); acc[i + 3] = fmaf((float)c[i + 3], rs.y, acc[i + 3]);
        sb = __hadd2(sb, sa);
      }
    } else if constexpr (VARIANT == 2) {   // I2F + packed FFMA2
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float f0 = (float)c[i], f1 = (float)c[i + 1];
        asm volatile("{.reg .b64 a, b, d; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %4}; mov.b64 d, {%0, %1};\n\t"
                     "fma.rn.f32x2 d, a, b, d; mov.b64 {%0, %1}, d;}" : "+f"(acc[i]), "+f"(acc[i + 1]) : "f"(f0), "f"(f1), "f"(1.0001f));
      }
    } else if constexpr (VARIANT == 3) {   // magic number: (c + 0x4B400000) as float, minus 1.5*2^23, FFMA
#pragma unroll
      for (int i = 0; i < 32; ++i) { float f = __int_as_float(c[i] + 0x4B400000) - 12582912.f; acc[i] = fmaf(f, 1.0001f, acc[i]); }
    } else if constexpr (VARIANT == 4) {   // FFMA only (upper bound)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(__int_as_float(c[i]), 1.0001f, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] += it;   // 1 IADD per element keeps values live (counted in all variants)
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
}

// ------------------------------------------------------------------ (3) nibble expansion smem->smem
__global__ void __launch_bounds__(128, 1) expand_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* packed = raw; uint8_t* expd = raw + 16384;
  for (int i = threadIdx.x; i < 4096; i += 128) reinterpret_cast<uint32_t*>(packed)[i] = i * 2654435761u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = threadIdx.x; c < 1024; c += 128) {   // 256 rows x 4 chunks = one 128x128 + 128x128 group
      const int r = c >> 2, j = c & 3;
      uint4 w = *reinterpret_cast<const uint4*>(packed + c * 16), lo, hi;
      lo.x = (w.x << 4) & 0xF0F0F0F0u; hi.x = w.x & 0xF0F0F0F0u; lo.y = (w.y << 4) & 0xF0F0F0F0u; hi.y = w.y & 0xF0F0F0F0u;
      lo.z = (w.z << 4) & 0xF0F0F0F0u; hi.z = w.z & 0xF0F0F0F0u; lo.w = (w.w << 4) & 0xF0F0F0F0u; hi.w = w.w & 0xF0F0F0F0u;
      uint8_t* row = expd + (r >> 3) * 1024 + (r & 7) * 128;
      *reinterpret_cast<uint4*>(row + (((2 * j) ^ (r & 7)) << 4)) = lo;
      *reinterpret_cast<uint4*>(row + (((2 * j + 1) ^ (r & 7)) << 4)) = hi;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cycles = (unsigned long long)(t1 - t0); *sink = reinterpret_cast<uint32_t*>(expd)[iters & 1023]; }
}

int main() {
  unsigned long long* d_cyc; CK(cudaMalloc(&d_cyc, 8));
  float* d_out; CK(cudaMalloc(&d_out, 148 * 512 * 4));
  uint32_t* d_sink; CK(cudaMalloc(&d_sink, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned long long cyc; float ms;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);

  auto umma = [&](auto kern, int bn) {
    const int iters = 20000, smem = (128 + bn) * 128 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {1, 148}) {
      kern<<<grid, 128, smem>>>(200, d_cyc); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); kern<<<grid, 128, smem>>>(iters, d_cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      double ops = 2.0 * 128 * bn * 128 * iters * grid;
      printf("{\"bench\": \"umma_i8_peak\", \"M\": 128, \"N\": %d, \"grid\": %d, \"ms\": %.4f, \"TOPS\": %.1f, \"cycles_per_group_K128\": %.1f}\n",
             bn, grid, ms, ops / ms * 1e-9, (double)cyc / iters);
    }
  };
  umma(umma_peak_kernel<256>, 256); umma(umma_peak_kernel<128>, 128); umma(umma_peak_kernel<64>, 64); umma(umma_peak_kernel<16>, 16);

  auto epi = [&](auto kern, const char* name, int threads) {
    const int iters = 4000;
    kern<<<148, threads>>>(10, d_out, d_cyc, 1); CK(cudaDeviceSynchronize());
    kern<<<148, threads>>>(iters, d_out, d_cyc, 1); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    // elements processed per SMSP per iteration = (threads/128 warps per SMSP) * 32 elements per thread (per lane)
    double per_elem = (double)cyc / iters / (32.0 * (threads / 128));
    printf("{\"bench\": \"epilogue_mix\", \"variant\": \"%s\", \"threads\": %d, \"cycles_per_warp_element_per_SMSP\": %.3f}\n", name, threads, per_elem);
  };
  for (int th : {128, 256, 512}) {
    epi(epi_kernel<0>, "i2f+ffma(+iadd)", th); epi(epi_kernel<1>, "faithful hmul2/4+cvt/2+i2f+ffma(+iadd)", th);
    epi(epi_kernel<2>, "i2f+ffma2(+iadd)", th); epi(epi_kernel<3>, "magic iadd+fadd+ffma(+iadd)", th); epi(epi_kernel<4>, "ffma only(+iadd)", th);
  }
  {
    CK(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768));
    expand_kernel<<<148, 128, 16384 + 32768>>>(10, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    expand_kernel<<<148, 128, 16384 + 32768>>>(2000, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"expand_int4_to_int8\", \"warps\": 4, \"cycles_per_group_128x128_A_and_B\": %.1f}\n", (double)cyc / 2000);
  }
  return 0;
}
