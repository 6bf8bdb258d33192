// sync_bench.cu -- cost of the hand-off primitives the GEMM pipeline is built from (cycles per iteration, one CTA).
#include <cstdio>
#include <cstdlib>
#include "../atom_b200/csrc/ptx_sm100.cuh"
using namespace atom;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) sync_kernel(int scenario, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 32); mbar_init(&bars[5], 4);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(&tptr);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  const uint64_t da = umma_desc_k_sw128(smem_u32(smem)), db = umma_desc_k_sw128(smem_u32(smem + 16384));
  constexpr uint32_t idesc = umma_idesc_i8(128, 16);
  long long t0 = 0, t1 = 0;
  if (scenario == 1 || scenario == 2 || scenario == 7) {       // one thread: 4 MMA (+1/2 commits, 7: commit + wait each iter)
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_i8(tb + (it & 7) * 16, da + k * 2, db + k * 2, idesc, k > 0);
        umma_commit(&bars[0]);
        if (scenario == 2) umma_commit(&bars[1]);
        if (scenario == 7) mbar_wait(&bars[0], it & 1);
      }
      if (scenario != 7) { umma_commit(&bars[2]); mbar_wait(&bars[2], 0); }
      t1 = clock64();
    }
  } else if (scenario == 3) {                                 // arrive + wait on the same thread
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) { mbar_arrive(&bars[0]); mbar_wait(&bars[0], it & 1); }
      t1 = clock64();
    }
  } else if (scenario == 4) {                                 // ping-pong between two warps
    if (warp < 2 && lane == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (warp == 0) { mbar_arrive(&bars[0]); mbar_wait(&bars[1], it & 1); }
        else { mbar_wait(&bars[0], it & 1); mbar_arrive(&bars[1]); }
      }
      t1 = clock64();
    }
  } else if (scenario == 5) {                                 // chain: warp0 arrive -> warp1 (wait, 4 MMA, commit) -> warp2 wait -> arrive back
    if (lane == 0 && warp < 3) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (warp == 0) { mbar_arrive(&bars[0]); mbar_wait(&bars[2], it & 1); }
        else if (warp == 1) {
          mbar_wait(&bars[0], it & 1); tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_i8(tb, da + k * 2, db + k * 2, idesc, k > 0);
          umma_commit(&bars[1]);
        } else { mbar_wait(&bars[1], it & 1); tc_fence_after(); mbar_arrive(&bars[2]); }
      }
      t1 = clock64();
    }
  } else if (scenario == 6) {                                 // 2 x STS.128 + fence.proxy.async + syncwarp + elected arrive (4 warps)
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint4 v = make_uint4(it, it, it, it);
      *reinterpret_cast<uint4*>(smem + threadIdx.x * 16) = v;
      *reinterpret_cast<uint4*>(smem + 4096 + threadIdx.x * 16) = v;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[5]);
    }
    t1 = clock64();
  } else if (scenario == 8) {                                 // same as 6 without the proxy fence
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint4 v = make_uint4(it, it, it, it);
      *reinterpret_cast<uint4*>(smem + threadIdx.x * 16) = v;
      *reinterpret_cast<uint4*>(smem + 4096 + threadIdx.x * 16) = v;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[5]);
    }
    t1 = clock64();
  } else if (scenario == 9) {                                 // wait on an already completed phase
    if (threadIdx.x == 0) {
      mbar_arrive(&bars[0]);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) mbar_wait(&bars[0], 0);
      t1 = clock64();
    }
  }
  if (threadIdx.x == 0 || (scenario == 4 && warp == 0 && lane == 0) || (scenario == 5 && warp == 0 && lane == 0)) *out = (unsigned long long)(t1 - t0);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tb);
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 8));
  CK(cudaFuncSetAttribute(sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  const char* names[] = {"", "4xMMA(N16)+1 commit, no wait", "4xMMA+2 commits, no wait", "arrive+wait same thread", "2-warp mbarrier ping-pong (round trip)",
                         "chain arrive->wait,4MMA,commit->wait->arrive (3 warps)", "2xSTS.128+fence.proxy.async+syncwarp+arrive", "4xMMA+commit+wait (serialised)",
                         "2xSTS.128+syncwarp+arrive (no fence)", "wait on completed phase"};
  for (int s = 1; s <= 9; ++s) {
    const int iters = 2000;
    sync_kernel<<<1, 128, 40960>>>(s, 10, d); CK(cudaDeviceSynchronize());
    sync_kernel<<<1, 128, 40960>>>(s, iters, d); CK(cudaDeviceSynchronize());
    unsigned long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"sync\", \"scenario\": \"%s\", \"cycles_per_iter\": %.1f}\n", names[s], (double)c / iters);
  }
  return 0;
}
