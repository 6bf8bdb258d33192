// tmem_bench.cu -- TMEM -> register bandwidth (tcgen05.ld 32x32b.x32 / .x64) with 4..16 warps: bounds the per-group
// accumulator drain of the W4A4 GEMM (a 128x128 INT32 group = 64 KiB).
#include <cstdio>
#include <cstdlib>
#include "../atom_b200/csrc/ptx_sm100.cuh"
using namespace atom;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
        "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
        "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
        "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
      : "r"(taddr) : "memory");
}

template <int X>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, int warps, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < warps) {
    const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    for (int it = 0; it < iters; ++it) {
      if constexpr (X == 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + (it & 1) * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) acc ^= r[i];
      } else {
        uint32_t r[64];
        ld64(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; i += 8) acc ^= r[i];
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *out = (unsigned long long)(t1 - t0);
  if (acc == 0x12345) *sink = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 8));
  uint32_t* sink; CK(cudaMalloc(&sink, 4));
  for (int x : {32, 64}) for (int warps : {1, 4, 8, 16}) {
    const int iters = 4000;
    if (x == 32) { tmem_kernel<32><<<1, 512>>>(10, warps, d, sink); CK(cudaDeviceSynchronize()); tmem_kernel<32><<<1, 512>>>(iters, warps, d, sink); }
    else { tmem_kernel<64><<<1, 512>>>(10, warps, d, sink); CK(cudaDeviceSynchronize()); tmem_kernel<64><<<1, 512>>>(iters, warps, d, sink); }
    CK(cudaDeviceSynchronize());
    unsigned long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    double bytes = (double)iters * warps * 32 * x * 4;
    printf("{\"bench\": \"tmem_ld\", \"shape\": \"32x32b.x%d\", \"warps\": %d, \"cycles_per_ld\": %.1f, \"bytes_per_cycle_per_SM\": %.1f}\n", x, warps, (double)c / iters, bytes / c);
  }
  return 0;
}
