// ref_shim.cu -- extern "C" launchers around the REFERENCE's own CUDA kernels.
//
// TEST INFRASTRUCTURE ONLY.  The kernels themselves are compiled, unmodified, from
// /root/reference/e2e/punica-atom/punica/ops/csrc/**.cu (and kernels/include/GEMM for the
// nvbench o16 variant) by oracle/Makefile into oracle/_ref/libatom_ref.so.  This file adds
// no arithmetic: it only forwards raw pointers to the reference's launchers
// (punica_ops.cc:73-262 does the same from torch tensors).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

#include "GEMM/DenseLayerGEMM_i4.h"
#include "GEMM/DenseLayerGEMM_i4_o4.h"
#include "Reorder/Reorder.h"
#include "Norm/RMSNorm.h"
#include "Activate/Activate.h"
#include "flashinfer_adapter/flashinfer_config.h"

// explicit instantiations that exist in the reference .cu files
extern template void run_reorder_fp16_i4<128, 4096>(half*, int, int16_t*, int8_t*, int8_t*, half*, half*);
extern template void run_rmsnorm_fp16_i4<128, 4096>(nv_half*, nv_half*, float, int, int16_t*, int8_t*, int8_t*, nv_half*, nv_half*);
extern template void run_activate_fp16_i4<128, 11008>(nv_half*, nv_half*, int, int8_t*, int8_t*, nv_half*, nv_half*);

static int done() {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}

extern "C" {

// K incl. keeper, exactly like punica_ops.cc:240
int atom_ref_gemm_i4_o16(const void* A, const void* B, const void* As, const void* Bs, const void* Ak,
                         const void* Bk, const void* Aks, const void* Bks, void* D, size_t M, size_t N, size_t K,
                         int sync) {
  DenseLayerGEMM_i4<nv_half>((const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)As, (const uint8_t*)Bs,
                             (const uint8_t*)Ak, (const uint8_t*)Bk, (const uint8_t*)Aks, (const uint8_t*)Bks,
                             (nv_half*)D, M, N, K);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_gemm_i4_o4(const void* A, const void* B, const void* As, const void* Bs, const void* Ak,
                        const void* Bk, const void* Aks, const void* Bks, void* D, void* Dscale, size_t M,
                        size_t N, size_t K, int sync) {
  DenseLayerGEMM_i4_o4((const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)As, (const uint8_t*)Bs,
                       (const uint8_t*)Ak, (const uint8_t*)Bk, (const uint8_t*)Aks, (const uint8_t*)Bks,
                       (uint8_t*)D, M, N, K, (half2*)Dscale);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_reorder_fp16_i4(void* x, int seq_len, void* idx, void* o8, void* o4, void* s8, void* s4, int sync) {
  run_reorder_fp16_i4<128, 4096>((half*)x, seq_len, (int16_t*)idx, (int8_t*)o8, (int8_t*)o4, (half*)s8, (half*)s4);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_rmsnorm_fp16_i4(void* x, void* w, float eps, int seq_len, void* idx, void* o8, void* o4, void* s8,
                             void* s4, int sync) {
  run_rmsnorm_fp16_i4<128, 4096>((nv_half*)x, (nv_half*)w, eps, seq_len, (int16_t*)idx, (int8_t*)o8, (int8_t*)o4,
                                 (nv_half*)s8, (nv_half*)s4);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_activate_fp16_i4(void* a, void* b, int seq_len, void* o8, void* o4, void* s8, void* s4, int sync) {
  run_activate_fp16_i4<128, 11008>((nv_half*)a, (nv_half*)b, seq_len, (int8_t*)o8, (int8_t*)o4, (nv_half*)s8,
                                   (nv_half*)s4);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_batch_decode_i4(void* o, void* q, void* kv_data, void* kv_param, void* indptr, void* indices,
                             void* last_off, int L, int layer, int H, int P, int B, int sync) {
  FlashInferBatchDecodeKernel_i4<128>((nv_half*)o, (nv_half*)q, kv_data, (nv_half2*)kv_param, (int32_t*)indptr,
                                      (int32_t*)indices, (int32_t*)last_off, L, layer, H, P, B);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_append_kv_i4(void* kv_data, void* kv_param, void* indptr, void* indices, void* last_off, void* k,
                          void* v, void* kp, void* vp, int L, int layer, int H, int P, int B, int sync) {
  FlashInferAppendKvKernel_i4<128>(kv_data, (nv_half2*)kv_param, (int32_t*)indptr, (int32_t*)indices,
                                   (int32_t*)last_off, k, v, (nv_half2*)kp, (nv_half2*)vp, L, layer, H, P, B);
  return sync ? done() : (int)cudaGetLastError();
}

int atom_ref_init_kv_i4(void* kv_data, void* kv_param, void* indptr, void* indices, void* last_off, void* k,
                        void* v, void* kp, void* vp, void* seqlen_indptr, int L, int layer, int H, int P, int B,
                        int sync) {
  FlashInferInitKvKernel_i4<128>(kv_data, (nv_half2*)kv_param, (int32_t*)indptr, (int32_t*)indices,
                                 (int32_t*)last_off, k, v, (nv_half2*)kp, (nv_half2*)vp, (int32_t*)seqlen_indptr, L,
                                 layer, H, P, B);
  return sync ? done() : (int)cudaGetLastError();
}

}  // extern "C"
