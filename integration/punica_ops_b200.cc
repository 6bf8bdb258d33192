// punica_ops_b200.cc -- the reference-side binding: a drop-in replacement for
// /root/reference/e2e/punica-atom/punica/ops/csrc/punica_ops.cc (pybind module `punica.ops._kernels`, :270-279).
// Same eight Python-visible functions with the same argument lists; each forwards plain pointers + sizes to the C ABI of
// libatom_b200.so (include/atom_b200.h).  Nothing else of the reference changes: punica/ops/__init__.py keeps allocating
// the outputs and calling these functions.
//
// Build (what the maintainer's setup.py would do; checked here with `g++ -fsyntax-only`, see tests/test_host_cabi.py):
//   CppExtension("punica.ops._kernels", ["punica_ops_b200.cc"], include_dirs=[<repo>/include],
//                library_dirs=[<repo>/atom_b200], libraries=["atom_b200"], runtime_library_dirs=[<repo>/atom_b200])
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include "atom_b200.h"

namespace {

inline void* cur_stream() { return static_cast<void*>(c10::cuda::getCurrentCUDAStream().stream()); }

#define ATOM_CHECK(call)                                                   \
  do {                                                                     \
    const int atom_rc_ = (call);                                           \
    TORCH_CHECK(atom_rc_ == ATOM_OK, #call " failed: ", atom_last_error()); \
  } while (0)

#define CHECK_CUDA_CONTIG(x) TORCH_CHECK((x).is_cuda() && (x).is_contiguous(), #x " must be a contiguous CUDA tensor")

struct KvDims {
  int num_layers, num_heads, page_size;
};
inline KvDims kv_dims(const torch::Tensor& kv_data) {  // [pages, L, 2, H, P, 64]
  TORCH_CHECK(kv_data.dim() == 6 && kv_data.size(5) == 64, "kv_data must be [pages, L, 2, H, P, 64] (head_dim 128)");
  return {static_cast<int>(kv_data.size(1)), static_cast<int>(kv_data.size(3)), static_cast<int>(kv_data.size(4))};
}

}  // namespace

void activate_fp16_i4(torch::Tensor A, torch::Tensor B, int seq_len, torch::Tensor o_outliers, torch::Tensor o_norms,
                      torch::Tensor outlier_scales, torch::Tensor norm_scales) {
  CHECK_CUDA_CONTIG(A);
  CHECK_CUDA_CONTIG(B);
  ATOM_CHECK(atom_activate_fp16_i4(A.data_ptr(), B.data_ptr(), seq_len, static_cast<int>(A.size(1)), o_outliers.data_ptr(),
                                   o_norms.data_ptr(), outlier_scales.data_ptr(), norm_scales.data_ptr(), cur_stream()));
}

void rmsnorm_fp16_i4(torch::Tensor hidden_states, torch::Tensor weight, float eps, torch::Tensor reorder_index,
                     torch::Tensor o_outliers, torch::Tensor o_norms, torch::Tensor outlier_scales,
                     torch::Tensor norm_scales) {
  CHECK_CUDA_CONTIG(hidden_states);
  // the reference reinterprets whatever `weight` holds as half (punica_ops.cc:243); llama.py:237 keeps fp32 ones
  torch::Tensor w = weight.scalar_type() == at::ScalarType::Half ? weight : weight.to(at::ScalarType::Half);
  ATOM_CHECK(atom_rmsnorm_fp16_i4(hidden_states.data_ptr(), w.data_ptr(), eps, reorder_index.data_ptr(),
                                  static_cast<int>(hidden_states.size(0)), static_cast<int>(hidden_states.size(1)),
                                  o_outliers.data_ptr(), o_norms.data_ptr(), outlier_scales.data_ptr(),
                                  norm_scales.data_ptr(), cur_stream()));
}

void reorder_fp16_i4(torch::Tensor hidden_states, torch::Tensor reorder_index, torch::Tensor o_outliers,
                     torch::Tensor o_norms, torch::Tensor outlier_scales, torch::Tensor norm_scales) {
  CHECK_CUDA_CONTIG(hidden_states);
  ATOM_CHECK(atom_reorder_fp16_i4(hidden_states.data_ptr(), reorder_index.data_ptr(), static_cast<int>(hidden_states.size(0)),
                                  static_cast<int>(hidden_states.size(1)), o_outliers.data_ptr(), o_norms.data_ptr(),
                                  outlier_scales.data_ptr(), norm_scales.data_ptr(), cur_stream()));
}

void dense_layer_gemm_i4_fp16(torch::Tensor a, torch::Tensor b, torch::Tensor a_scale, torch::Tensor b_scale,
                              torch::Tensor a_keeper, torch::Tensor b_keeper, torch::Tensor a_keeper_scale,
                              torch::Tensor b_keeper_scale, torch::Tensor d) {
  CHECK_CUDA_CONTIG(a);
  CHECK_CUDA_CONTIG(b);
  ATOM_CHECK(atom_gemm_i4_o16(a.data_ptr(), b.data_ptr(), a_scale.data_ptr(), b_scale.data_ptr(), a_keeper.data_ptr(),
                              b_keeper.data_ptr(), a_keeper_scale.data_ptr(), b_keeper_scale.data_ptr(), d.data_ptr(),
                              a.size(0), b.size(0), a.size(1) * 2 + a_keeper.size(1), ATOM_GEMM_AUTO, cur_stream()));
}

void dense_layer_gemm_i4_o4(torch::Tensor a, torch::Tensor b, torch::Tensor a_scale, torch::Tensor b_scale,
                            torch::Tensor a_keeper, torch::Tensor b_keeper, torch::Tensor a_keeper_scale,
                            torch::Tensor b_keeper_scale, torch::Tensor d, torch::Tensor d_scale) {
  CHECK_CUDA_CONTIG(a);
  CHECK_CUDA_CONTIG(b);
  ATOM_CHECK(atom_gemm_i4_o4(a.data_ptr(), b.data_ptr(), a_scale.data_ptr(), b_scale.data_ptr(), a_keeper.data_ptr(),
                             b_keeper.data_ptr(), a_keeper_scale.data_ptr(), b_keeper_scale.data_ptr(), d.data_ptr(),
                             d_scale.data_ptr(), a.size(0), b.size(0), a.size(1) * 2 + a_keeper.size(1), ATOM_GEMM_AUTO,
                             cur_stream()));
}

void batch_decode_i4(torch::Tensor o, torch::Tensor q, torch::Tensor kv_data, torch::Tensor kv_param,
                     torch::Tensor kv_indptr, torch::Tensor kv_indicies, torch::Tensor last_page_offset, int layer_idx) {
  CHECK_CUDA_CONTIG(o);
  CHECK_CUDA_CONTIG(q);
  const KvDims k = kv_dims(kv_data);
  const int batch_size = static_cast<int>(o.size(0));
  TORCH_CHECK(kv_indptr.size(0) == batch_size + 1 && last_page_offset.size(0) == batch_size, "page table / batch mismatch");
  ATOM_CHECK(atom_batch_decode_i4(o.data_ptr(), q.data_ptr(), kv_data.data_ptr(), kv_param.data_ptr(), kv_indptr.data_ptr(),
                                  kv_indicies.data_ptr(), last_page_offset.data_ptr(), k.num_layers, layer_idx, k.num_heads,
                                  k.page_size, batch_size, cur_stream()));
}

void init_kv_i4(torch::Tensor kv_data, torch::Tensor kv_param, torch::Tensor kv_indptr, torch::Tensor kv_indicies,
                torch::Tensor last_page_offset, torch::Tensor k, torch::Tensor v, torch::Tensor k_param,
                torch::Tensor v_param, torch::Tensor seqlen_indptr, int layer_idx) {
  CHECK_CUDA_CONTIG(k);
  CHECK_CUDA_CONTIG(v);
  const KvDims d = kv_dims(kv_data);
  const int batch_size = static_cast<int>(last_page_offset.size(0));
  // total_tokens == seqlen_indptr[B] == k.size(0): known on the host, no device read needed
  ATOM_CHECK(atom_init_kv_i4(kv_data.data_ptr(), kv_param.data_ptr(), kv_indptr.data_ptr(), kv_indicies.data_ptr(),
                             last_page_offset.data_ptr(), k.data_ptr(), v.data_ptr(), k_param.data_ptr(), v_param.data_ptr(),
                             seqlen_indptr.data_ptr(), static_cast<int>(k.size(0)), d.num_layers, layer_idx, d.num_heads,
                             d.page_size, batch_size, cur_stream()));
}

void append_kv_i4(torch::Tensor kv_data, torch::Tensor kv_param, torch::Tensor kv_indptr, torch::Tensor kv_indicies,
                  torch::Tensor last_page_offset, torch::Tensor k, torch::Tensor v, torch::Tensor k_param,
                  torch::Tensor v_param, int layer_idx) {
  CHECK_CUDA_CONTIG(k);
  CHECK_CUDA_CONTIG(v);
  const KvDims d = kv_dims(kv_data);
  ATOM_CHECK(atom_append_kv_i4(kv_data.data_ptr(), kv_param.data_ptr(), kv_indptr.data_ptr(), kv_indicies.data_ptr(),
                               last_page_offset.data_ptr(), k.data_ptr(), v.data_ptr(), k_param.data_ptr(), v_param.data_ptr(),
                               d.num_layers, layer_idx, d.num_heads, d.page_size, static_cast<int>(k.size(0)), cur_stream()));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("activate_fp16_i4", &activate_fp16_i4, "");
  m.def("batch_decode_i4", &batch_decode_i4, "");
  m.def("init_kv_i4", &init_kv_i4, "");
  m.def("append_kv_i4", &append_kv_i4, "");
  m.def("dense_layer_gemm_i4_o4", &dense_layer_gemm_i4_o4, "");
  m.def("dense_layer_gemm_i4_fp16", &dense_layer_gemm_i4_fp16, "");
  m.def("rmsnorm_fp16_i4", &rmsnorm_fp16_i4, "");
  m.def("reorder_fp16_i4", &reorder_fp16_i4, "");
}
