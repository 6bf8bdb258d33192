// tma_bench.cu -- how fast can ONE SM's TMA engine deliver narrow 2-D boxes?  (sizes the packed-operand staging)
// box = inner bytes x rows, always 16 KiB per load, source rows have pitch 1984 B (K=4096 packed INT4), L2 resident.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../atom_b200/csrc/ptx_sm100.cuh"
using namespace atom;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int iters, int inner, int rows, int stages,
                                                     int row_span, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = inner * rows;
    long long t0 = clock64();
    for (int it = 0; it < iters + stages; ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(&full[s], ((it / stages) - 1) & 1);
      if (it < iters) {
        mbar_arrive_expect_tx(&full[s], bytes);
        const int col = (it * inner) % 1920, row = ((blockIdx.x * 131 + it * 7) * rows) % row_span;
        tma_load_2d(smem + s * 16384, &tm, &full[s], col - col % inner, row);
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  const size_t pitch = 1984, nrows = 16384;   // 32 MB, L2 resident
  uint8_t* d; CK(cudaMalloc(&d, pitch * nrows)); CK(cudaMemset(d, 1, pitch * nrows));
  unsigned long long* d_cyc; CK(cudaMalloc(&d_cyc, 8));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024));
  for (int inner : {64, 128, 256, 512, 1024}) {
    const int rows = 16384 / inner;
    CUtensorMap tm;
    cuuint64_t dims[2], strides[1] = {pitch}; cuuint32_t box[2], es[2] = {1, 1};
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
    if (inner <= 256) { dims[0] = pitch; box[0] = inner; } else { dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; dims[0] = pitch / 4; box[0] = inner / 4; }
    dims[1] = nrows; box[1] = rows;
    CUresult r = enc(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"inner\": %d, \"encode_error\": %d}\n", inner, (int)r); continue; }
    for (int grid : {1, 148}) for (int stages : {4, 8}) {
      const int iters = 2000;
      tma_kernel<<<grid, 128, 8 * 16384 + 1024>>>(tm, 50, inner, rows, stages, (int)nrows - 512, d_cyc); CK(cudaDeviceSynchronize());
      tma_kernel<<<grid, 128, 8 * 16384 + 1024>>>(tm, iters, inner, rows, stages, (int)nrows - 512, d_cyc); CK(cudaDeviceSynchronize());
      unsigned long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("{\"bench\": \"tma_box\", \"inner_bytes\": %d, \"rows\": %d, \"grid\": %d, \"stages\": %d, \"cycles_per_16KiB\": %.1f, \"bytes_per_cycle_per_SM\": %.1f}\n",
             inner, rows, grid, stages, (double)cyc / iters, 16384.0 * iters / cyc);
    }
  }
  return 0;
}
