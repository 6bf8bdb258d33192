// atom_b200.cu -- C ABI (include/atom_b200.h) and host-side launch logic of libatom_b200.so.
//
// No torch types, no allocation, no synchronisation: every entry point validates its arguments, builds the TMA
// descriptors it needs on the host (cached by (pointer, shape)) and enqueues kernels on the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>

#include "../../include/atom_b200.h"
#include "gemm_i4_sm100.cuh"
#include "gemm_i4_skinny_sm100.cuh"
#include "gemm_i4_tall_sm100.cuh"
#include "gemm_i4_wide_sm100.cuh"
#include "kv_kernels.cuh"
#include "prefill_kernels.cuh"
#include "comm_kernels.cuh"
#include "quant_kernels.cuh"

namespace {

thread_local std::string g_err;
unsigned long long* g_trace = nullptr;   // atom_gemm_set_trace

// Programmatic dependent launch (opt-in: ATOM_B200_PDL=1 in the environment, or atom_set_pdl): kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and told so (`pdl` argument), so that their prologue -- for the GEMM
// including the first weight tiles -- overlaps the tail of the preceding kernel.  Off: plain stream order, and the
// kernels never execute a griddepcontrol instruction.
int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ATOM_B200_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl == 1;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define ATOM_REQUIRE(cond, ...) \
  do { if (!(cond)) return fail(ATOM_E_INVALID, __VA_ARGS__); } while (0)

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return ATOM_OK;
}

// <<<>>> when PDL is off (the validated path, untouched); cudaLaunchKernelEx with the attribute when it is on
template <typename... KArgs, typename... Args>
int launch_k(const char* what, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  if (!pdl_enabled()) {
    kern<<<grid, block, smem, stream>>>(args..., 0);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args..., 1);
    if (e != cudaSuccess) return fail(ATOM_E_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return ATOM_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): raise it once for each pair, under a lock
// (the reference calls cudaFuncSetAttribute on every launch, GEMM.cuh:757).
template <typename F>
int ensure_dynamic_smem(F* func, int bytes, const char* what, bool max_carveout = false) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(ATOM_E_CUDA, "%s: no current CUDA device", what);
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(func), dev);
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(key)) return ATOM_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "%s: cudaFuncSetAttribute(smem=%d): %s", what, bytes, cudaGetErrorString(e));
  if (max_carveout) {   // two ~100 KB CTAs per SM only fit when the SM's L1/shared split is at its shared-memory maximum
    e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(ATOM_E_CUDA, "%s: cudaFuncSetAttribute(carveout): %s", what, cudaGetErrorString(e));
  }
  done.insert(key);
  return ATOM_OK;
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t inner, rows, pitch; uint32_t box_inner, box_rows, swizzle;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && rows == o.rows && pitch == o.pitch && box_inner == o.box_inner &&
           box_rows == o.box_rows && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.rows); mix(k.pitch); mix(k.box_inner); mix(k.box_rows); mix(k.swizzle);
    return h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D byte tensor [rows][inner] with row pitch `pitch`; box = box_rows x box_inner bytes; OOB rows read as zero.
// swizzle: 0 = none, 1 = SWIZZLE_128B, 2 = SWIZZLE_64B
int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_inner,
             uint32_t box_rows, int swizzle) {
  MapKey key{ptr, inner, rows, pitch, box_inner, box_rows, (uint32_t)swizzle};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return ATOM_OK; }
  }
  EncodeFn fn = encode_fn();
  if (!fn) return fail(ATOM_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ATOM_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return ATOM_OK;
}

// The same byte matrix viewed as [rows/4][2][2][inner] with the two middle dimensions swapped: a box of `box_rows` rows
// lands in shared memory in the order 4q+0, 4q+2, 4q+1, 4q+3 (gemm_i4_tall_sm100.cuh: neighbouring accumulator columns
// then belong to different channel pairs).  rows % 4 == 0; rows beyond the matrix read as zero.
int make_map_quadswap(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_inner,
                      uint32_t box_rows, int swizzle) {
  MapKey key{ptr, inner, rows, pitch, box_inner, box_rows, (uint32_t)swizzle | 0x100u};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return ATOM_OK; }
  }
  EncodeFn fn = encode_fn();
  if (!fn) return fail(ATOM_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {inner, 2, 2, (rows + 3) / 4};
  cuuint64_t strides[3] = {2 * pitch, pitch, 4 * pitch};
  cuuint32_t box[4] = {box_inner, 2, 2, box_rows / 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ATOM_E_CUDA, "cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return ATOM_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM launch
struct GemmOperands {
  const void *a, *b, *ak, *bk;
  int64_t M, N, K;
};

template <bool kSwap, int BN, int GS, int kPack, int kSplit, bool kO4, int kConvWarps, int kEpiWgs>
int launch_gemm(const GemmOperands& op, const atom::GemmArgs& args, cudaStream_t stream) {
  using C = atom::GemmCfg<kSwap, BN, GS, kPack, kSplit, kO4, kConvWarps, kEpiWgs>;
  auto kern = atom::gemm_i4_kernel<kSwap, BN, GS, kPack, kSplit, kO4, kConvWarps, kEpiWgs>;
  int rc = ensure_dynamic_smem(kern, C::SMEM_BYTES, "gemm_i4");
  if (rc) return rc;
  const uint64_t kp = (uint64_t)(op.K - 128) / 2;
  // MMA-M operand = kSwap ? weights : tokens
  const void* p4 = kSwap ? op.b : op.a;  const void* q4 = kSwap ? op.a : op.b;
  const void* p8 = kSwap ? op.bk : op.ak; const void* q8 = kSwap ? op.ak : op.bk;
  const uint64_t prow = kSwap ? op.N : op.M, qrow = kSwap ? op.M : op.N;
  CUtensorMap tp4, tq4, tp8, tq8;
  if ((rc = make_map(&tp4, p4, kp, prow, kp, 64, C::BM, 0))) return rc;
  if ((rc = make_map(&tq4, q4, kp, qrow, kp, 64, BN, 0))) return rc;
  if ((rc = make_map(&tp8, p8, 128, prow, 128, 128, C::BM, 1))) return rc;
  if ((rc = make_map(&tq8, q8, 128, qrow, 128, 128, BN, 1))) return rc;

  const int ch_tile = kSwap ? C::BM : BN, tok_tile = kSwap ? BN : C::BM;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((op.N + ch_tile - 1) / ch_tile), (unsigned)((op.M + tok_tile - 1) / tok_tile), kSplit);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = kSplit;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  atom::GemmArgs largs = args;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
    largs.pdl = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tp4, tq4, tp8, tq8, largs);
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "gemm_i4 launch: %s", cudaGetErrorString(e));
  return ATOM_OK;
}

// Decode-shape kernel (gemm_i4_skinny_sm100.cuh): weights expanded into tensor memory, two CTAs per SM, always launched
// with programmatic stream serialization (the kernel fetches only weights before griddepcontrol.wait).
int g_gemm_pdl = -1;
bool gemm_pdl_enabled() {
  if (g_gemm_pdl < 0) {
    const char* e = getenv("ATOM_B200_GEMM_PDL");
    g_gemm_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_gemm_pdl == 1;
}

template <int BN, int kSplit, int kEpi>
int launch_skinny(const GemmOperands& op, const atom::GemmArgs& args, cudaStream_t stream, int64_t out_channels = -1) {
  using C = atom::SkinnyCfg<BN, kSplit, kEpi>;
  auto kern = atom::gemm_i4_skinny_kernel<BN, kSplit, kEpi>;
  int rc = ensure_dynamic_smem(kern, C::SMEM_BYTES, "gemm_i4 (skinny)", true);
  if (rc) return rc;
  const uint64_t kp = (uint64_t)(op.K - 128) / 2;
  CUtensorMap tp4, tp8;       // weights only: the token operand is read with plain loads (no descriptor per activation buffer)
  if ((rc = make_map(&tp4, op.b, kp, op.N, kp, 64, C::BM, 2))) return rc;
  if ((rc = make_map(&tp8, op.bk, 128, op.N, 128, 128, C::BM, 1))) return rc;
  cudaLaunchConfig_t cfg{};
  const int64_t chan = out_channels > 0 ? out_channels : op.N;       // (gate/up: op.N = 2 I weight rows, I output channel tiles)
  cfg.gridDim = dim3((unsigned)((chan + C::BM - 1) / C::BM), (unsigned)((op.M + BN - 1) / BN), kSplit);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  if (kEpi == atom::EPI_GATEUP && (args.dbg & 2) && C::SMEM_BYTES < 120 * 1024) {     // experiment: one CTA per SM
    static bool raised = false;
    if (!raised) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024); raised = true; }
    cfg.dynamicSmemBytes = 120 * 1024;
  }
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = kSplit;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (gemm_pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tp4, tp8, args);
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "gemm_i4 (skinny) launch: %s", cudaGetErrorString(e));
  return ATOM_OK;
}

// Prefill-shape kernel (gemm_i4_tall_sm100.cuh): biased accumulators, packed FP32 epilogue, TMA-permuted weight rows.
template <bool kO4>
int launch_tall(const GemmOperands& op, const atom::GemmArgs& args, cudaStream_t stream) {
  using C = atom::TallCfg<kO4>;
  auto kern = atom::gemm_i4_tall_kernel<kO4>;
  int rc = ensure_dynamic_smem(kern, C::SMEM_BYTES, "gemm_i4 (tall)");
  if (rc) return rc;
  const uint64_t kp = (uint64_t)(op.K - 128) / 2;
  CUtensorMap tp4, tq4, tp8, tq8;
  if ((rc = make_map(&tp4, op.a, kp, op.M, kp, 64, C::BM, 0))) return rc;
  if ((rc = make_map_quadswap(&tq4, op.b, kp, op.N, kp, 64, C::BN, 0))) return rc;
  if ((rc = make_map(&tp8, op.ak, 128, op.M, 128, 128, C::BM, 1))) return rc;
  if ((rc = make_map_quadswap(&tq8, op.bk, 128, op.N, 128, 128, C::BN, 1))) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((op.N + C::BN - 1) / C::BN), (unsigned)((op.M + C::BM - 1) / C::BM), 1);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (gemm_pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tp4, tq4, tp8, tq8, args);
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "gemm_i4 (tall) launch: %s", cudaGetErrorString(e));
  return ATOM_OK;
}

// Prefill-shape kernel with the token operand in tensor memory, 128 x 256 tiles (gemm_i4_wide_sm100.cuh).
template <bool kO4>
int launch_wide(const GemmOperands& op, const atom::GemmArgs& args, cudaStream_t stream) {
  using C = atom::WideCfg<kO4>;
  auto kern = atom::gemm_i4_wide_kernel<kO4>;
  int rc = ensure_dynamic_smem(kern, C::SMEM_BYTES, "gemm_i4 (wide)");
  if (rc) return rc;
  const uint64_t kp = (uint64_t)(op.K - 128) / 2;
  CUtensorMap ta4, tb4, ta8, tb8;
  if ((rc = make_map(&ta4, op.a, kp, op.M, kp, 64, C::BM, 2))) return rc;
  if ((rc = make_map_quadswap(&tb4, op.b, kp, op.N, kp, 64, C::BH, 0))) return rc;
  if ((rc = make_map(&ta8, op.ak, 128, op.M, 128, 128, C::BM, 1))) return rc;
  if ((rc = make_map_quadswap(&tb8, op.bk, 128, op.N, 128, 128, C::BH, 1))) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((op.N + C::NH * C::BH - 1) / (C::NH * C::BH)), (unsigned)((op.M + C::BM - 1) / C::BM), 1);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (gemm_pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta4, tb4, ta8, tb8, args);
  if (e != cudaSuccess) return fail(ATOM_E_CUDA, "gemm_i4 (wide) launch: %s", cudaGetErrorString(e));
  return ATOM_OK;
}

template <bool kO4, bool kPush = false>
int skinny_dispatch(const GemmOperands& op, const atom::GemmArgs& args, uint32_t flags, cudaStream_t stream) {
  constexpr int kEpi = kO4 ? atom::EPI_O4 : (kPush ? atom::EPI_PUSH : atom::EPI_O16);
  const int64_t ch_tiles = (op.N + 127) / 128;
  const int groups = args.G + 1;
  const int bn = op.M <= 16 ? 16 : (op.M <= 32 ? 32 : 64);
  const int64_t tiles = ch_tiles * ((op.M + bn - 1) / bn);
  // K split over a cluster (1, 2 or 4 ranks) until the CTAs cover the machine: 148 SMs, two CTAs resident on each for
  // the 16- and 32-token tiles.  Every rank must keep at least ~4 groups, or the fixed cost per CTA dominates.
  int ksplit = 1;
  if (!kO4 && !(flags & ATOM_GEMM_NO_SPLITK)) {
    if (flags & ATOM_GEMM_SPLITK2) ksplit = 2;
    else if (flags & ATOM_GEMM_SPLITK4) ksplit = 4;
    else if ((flags & ATOM_GEMM_SPLITK8) && bn <= 32) ksplit = 8;
    else if (groups >= 8) {
      // The kernel is latency-, not bandwidth-bound: the largest split whose CTAs are all resident at once wins (148 SMs, two
      // CTAs each for the 16/32-token tiles), as long as every rank keeps >= 4 groups (8-way: from 32 groups on).
      // Measured (r02_gemm_skinny_v8_shape_table.jsonl): 16x4096x4096 5.9 us at 8 vs 6.3 at 4; 16x11008x4096 10.3 at 2 vs 12.6 at 1
      // and 12.9 at 4 (344 CTAs: a second wave); 32x13824x5120 19.1 at 2 vs 29.0 at 1.
      const int64_t slots = bn <= 32 ? 296 : 160;
      if (bn <= 32 && groups >= 32 && tiles * 8 <= slots) ksplit = 8;
      else ksplit = tiles * 4 <= slots ? 4 : (tiles * 2 <= slots ? 2 : 1);
    }
  }
  if constexpr (kO4) {
    if (bn == 16) return launch_skinny<16, 1, kEpi>(op, args, stream);
    if (bn == 32) return launch_skinny<32, 1, kEpi>(op, args, stream);
    return launch_skinny<64, 1, kEpi>(op, args, stream);
  } else {
#define ATOM_SK(BN_)                                                                 \
  return ksplit == 4   ? launch_skinny<BN_, 4, kEpi>(op, args, stream)               \
         : ksplit == 2 ? launch_skinny<BN_, 2, kEpi>(op, args, stream)               \
                       : launch_skinny<BN_, 1, kEpi>(op, args, stream)
    if (bn == 16) { if (ksplit == 8) return launch_skinny<16, 8, kEpi>(op, args, stream); ATOM_SK(16); }
    if (bn == 32) { if (ksplit == 8) return launch_skinny<32, 8, kEpi>(op, args, stream); ATOM_SK(32); }
    ATOM_SK(64);
#undef ATOM_SK
  }
}

template <bool kO4>
int gemm_dispatch(const GemmOperands& op, const atom::GemmArgs& args, uint32_t flags, cudaStream_t stream) {
  // up to 128 tokens the weight-stationary orientation wins (4096^2, M=128: 13.9 us as two 64-token tiles vs 21.3 us for one
  // wave of 32 128x128 tiles); from 3 token tiles on the 128x128 kernel is faster
  // (65..128 tokens are two 64-token tiles per channel tile: only while those CTAs still fit one wave of 148 SMs)
  const bool skinny = (flags & ATOM_GEMM_FORCE_SKINNY) ||
                      (!(flags & ATOM_GEMM_FORCE_TALL) && (op.M <= 64 || (op.M <= 128 && (op.N + 127) / 128 * 2 <= 148)));
  // <swap, BN, groups per pipeline stage, packed stages, K split, o4, converter warps, epilogue warpgroups>
  if (!skinny) {
    if (flags & ATOM_GEMM_LEGACY_TALL) return launch_gemm<false, 128, 2, 2, 1, kO4, 4, 2>(op, args, stream);
    // 128 x 128 tiles by default.  The 128 x 256 kernel (token operand in tensor memory) halves the token-side conversion
    // per MMA but measured slower at every size (4096^3: 704 vs 783-829 TOP/s, its 128-accumulator epilogue spills), so it
    // only runs on request.
    return (flags & ATOM_GEMM_FORCE_WIDE) ? launch_wide<kO4>(op, args, stream) : launch_tall<kO4>(op, args, stream);
  }
  // 33..64 tokens: the round-1 configuration (one CTA per SM, 8 converter warps) still measures faster than the new
  // kernel's BN=64 instance (4096^2: 13.0-13.3 vs 15.0 us), which then only runs on request and under the fused epilogues
  if (op.M <= 64 && !(flags & ATOM_GEMM_LEGACY_SKINNY) && (op.M <= 32 || (flags & ATOM_GEMM_FORCE_SKINNY))) {
    if constexpr (!kO4) { if (args.ar.bufs != nullptr) return skinny_dispatch<false, true>(op, args, flags, stream); }
    return skinny_dispatch<kO4>(op, args, flags, stream);
  }
  // decode shapes: weights on the MMA-M axis; K split 4-way over a cluster when one wave of CTAs would not
  // cover the machine (the kernel is HBM-bound on the weights: more CTAs = more bytes in flight)
  const int64_t ch_tiles = (op.N + 127) / 128;
  const int groups = args.G + 1;
  // K split over a cluster (1, 2 or 4 ranks): as many CTAs as it takes to give every SM about one, no more -- each extra
  // wave of CTAs costs a full pipeline ramp (first TMA tile ~2 us after launch).
  const int64_t tiles = ch_tiles * ((op.M + 63) / 64);
  int ksplit = 1;
  // o4 quantises from the FP32 sums: a K split would reorder them and move INT4 codes by one, so the KV projections
  // never split (bit-identical codes to the reference kernel on every path)
  if (!kO4 && !(flags & ATOM_GEMM_NO_SPLITK) && groups >= 8) {
    if (flags & ATOM_GEMM_SPLITK2) ksplit = 2;
    else if (flags & ATOM_GEMM_SPLITK4) ksplit = 4;
    else ksplit = tiles * 4 <= 148 ? 4 : (tiles * 2 <= 148 ? 2 : 1);   // the largest split that still fits one wave of 148 SMs
  }
#define ATOM_SKINNY(BN, GS4, GS2, GS1, KP4, KP2, KP1, EW)                                                      \
  return ksplit == 4   ? launch_gemm<true, BN, GS4, KP4, 4, kO4, 8, EW>(op, args, stream)                      \
         : ksplit == 2 ? launch_gemm<true, BN, GS2, KP2, 2, kO4, 8, EW>(op, args, stream)                      \
                       : launch_gemm<true, BN, GS1, KP1, 1, kO4, 8, EW>(op, args, stream)
  if (op.M <= 16) { ATOM_SKINNY(16, 2, 2, 2, 4, 4, 4, 1); }
  if (op.M <= 32) { ATOM_SKINNY(32, 2, 2, 2, 4, 4, 4, 1); }
  ATOM_SKINNY(64, 1, 2, 2, 4, 2, 3, 2);
#undef ATOM_SKINNY
}

int gemm_common(const void* a, const void* b, const void* a_scale, const void* b_scale, const void* a_keeper,
                const void* b_keeper, const void* a_keeper_scale, const void* b_keeper_scale, void* d, void* d_scale,
                int64_t M, int64_t N, int64_t K, uint32_t flags, void* stream, bool o4, const atom::ArArgs* ar = nullptr) {
  ATOM_REQUIRE(a && b && a_scale && b_scale && a_keeper && b_keeper && a_keeper_scale && b_keeper_scale && (d || ar),
               "gemm_i4: null pointer argument");
  ATOM_REQUIRE(M > 0 && N > 0, "gemm_i4: M=%lld N=%lld must be positive", (long long)M, (long long)N);
  ATOM_REQUIRE(K >= 256 && K % 128 == 0, "gemm_i4: K=%lld must be a multiple of 128 and >= 256 (INT4 groups + 128 keeper)", (long long)K);
  ATOM_REQUIRE(N % 8 == 0, "gemm_i4: N=%lld must be a multiple of 8", (long long)N);
  ATOM_REQUIRE(!o4 || (N % 128 == 0 && d_scale), "gemm_i4_o4: N=%lld must be a multiple of 128 (one head per scale)", (long long)N);
  ATOM_REQUIRE(aligned16(a) && aligned16(b) && aligned16(a_keeper) && aligned16(b_keeper) && aligned16(d),
               "gemm_i4: operand pointers must be 16-byte aligned");
  if (ar != nullptr) {
    ATOM_REQUIRE(!o4 && M <= 64, "gemm_i4_o16_push: decode batches only (M=%lld <= 64), fp16 output", (long long)M);
    ATOM_REQUIRE(ar->bufs && ar->state && ar->world >= 1 && ar->world <= 32 && ar->rank >= 0 && ar->rank < ar->world && ar->slot % 8 == 0 &&
                 M * N <= ar->slot, "gemm_i4_o16_push: rank=%d world=%d, M x N = %lld must fit a slot of %lld elements", ar->rank, ar->world,
                 (long long)(M * N), (long long)ar->slot);
    flags = (flags | ATOM_GEMM_FORCE_SKINNY) & ~(ATOM_GEMM_FORCE_TALL | ATOM_GEMM_LEGACY_SKINNY);
  }
  ATOM_REQUIRE(aligned16(b_scale) && aligned16(b_keeper_scale), "gemm_i4: weight scale pointers must be 16-byte aligned");
  ATOM_REQUIRE((reinterpret_cast<uintptr_t>(a_scale) & 3) == 0 && (reinterpret_cast<uintptr_t>(a_keeper_scale) & 3) == 0,
               "gemm_i4: activation scale pointers must be 4-byte aligned");
  ATOM_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 24), "gemm_i4: dimension too large");
  GemmOperands op{a, b, a_keeper, b_keeper, M, N, K};
  atom::GemmArgs args{};
  args.a_scale = (const __half*)a_scale; args.b_scale = (const __half*)b_scale;
  args.a_keeper_scale = (const __half*)a_keeper_scale; args.b_keeper_scale = (const __half*)b_keeper_scale;
  args.d = o4 ? nullptr : (__half*)d; args.d4 = o4 ? (uint8_t*)d : nullptr; args.d_scale = (__half2*)d_scale;
  args.M = (int)M; args.N = (int)N; args.G = (int)(K / 128 - 1); args.lda_scale = atom::scale_size((int)M); args.trace = g_trace;
  args.a4 = (const uint8_t*)a; args.a8 = (const int8_t*)a_keeper; args.ldb_scale = (int)N;
  if (ar != nullptr) args.ar = *ar;
  return o4 ? gemm_dispatch<true>(op, args, flags, (cudaStream_t)stream) : gemm_dispatch<false>(op, args, flags, (cudaStream_t)stream);
}

int quant_check(const char* what, int seq_len, int hidden, const void* o8, const void* o4, const void* s8, const void* s4) {
  ATOM_REQUIRE(seq_len > 0, "%s: seq_len=%d must be positive", what, seq_len);
  ATOM_REQUIRE(hidden >= 256 && hidden % 128 == 0 && hidden <= 65536, "%s: hidden_dim=%d must be a multiple of 128 in [256, 65536]", what, hidden);
  ATOM_REQUIRE(o8 && o4 && s8 && s4, "%s: null output pointer", what);
  return ATOM_OK;
}

}  // namespace

extern "C" {

int atom_version(void) { return 100; }
const char* atom_last_error(void) { return g_err.c_str(); }
int atom_scale_index(int row) { return atom::scale_index(row); }
int atom_scale_size(int rows) { return atom::scale_size(rows); }
int atom_gemm_set_trace(void* device_buffer) { g_trace = (unsigned long long*)device_buffer; return ATOM_OK; }
int atom_set_pdl(int enable) { g_pdl = enable ? 1 : 0; return ATOM_OK; }

int atom_reorder_fp16_i4(const void* hidden, const void* reorder_index, int seq_len, int hidden_dim, void* o_outliers,
                         void* o_norms, void* outlier_scales, void* norm_scales, void* stream) {
  int rc = quant_check("reorder_fp16_i4", seq_len, hidden_dim, o_outliers, o_norms, outlier_scales, norm_scales);
  if (rc) return rc;
  ATOM_REQUIRE(hidden && reorder_index && aligned16(hidden), "reorder_fp16_i4: null or misaligned input");
  if ((rc = ensure_dynamic_smem(atom::reorder_quant_kernel, 65536 * 2, "reorder_fp16_i4"))) return rc;   // one fp16 row, hidden <= 65536
  return launch_k("reorder_fp16_i4", atom::reorder_quant_kernel, dim3(seq_len), dim3(atom::QUANT_THREADS), (size_t)hidden_dim * 2,
                  (cudaStream_t)stream, (const __half*)hidden, (const int16_t*)reorder_index, seq_len, hidden_dim, (int8_t*)o_outliers,
                  (uint8_t*)o_norms, (__half*)outlier_scales, (__half*)norm_scales, atom::scale_size(seq_len));
}

int atom_rmsnorm_fp16_i4(const void* hidden, const void* weight, float eps, const void* reorder_index, int seq_len,
                         int hidden_dim, void* o_outliers, void* o_norms, void* outlier_scales, void* norm_scales,
                         void* stream) {
  int rc = quant_check("rmsnorm_fp16_i4", seq_len, hidden_dim, o_outliers, o_norms, outlier_scales, norm_scales);
  if (rc) return rc;
  ATOM_REQUIRE(hidden && weight && reorder_index && aligned16(hidden) && aligned16(weight), "rmsnorm_fp16_i4: null or misaligned input");
  ATOM_REQUIRE(hidden_dim <= 32768, "rmsnorm_fp16_i4: hidden_dim=%d > 32768 unsupported", hidden_dim);
  if ((rc = ensure_dynamic_smem(atom::rmsnorm_quant_kernel<false>, 32768 * 4 + 512, "rmsnorm_fp16_i4"))) return rc;   // row + weight (fp16) + reduction scratch
  return launch_k("rmsnorm_fp16_i4", atom::rmsnorm_quant_kernel<false>, dim3(seq_len), dim3(atom::QUANT_THREADS), (size_t)hidden_dim * 4 + 512,
                  (cudaStream_t)stream, (const __half*)hidden, (const __half*)nullptr, (__half*)nullptr, (const __half*)weight, eps,
                  (const int16_t*)reorder_index, seq_len, hidden_dim, (int8_t*)o_outliers, (uint8_t*)o_norms, (__half*)outlier_scales,
                  (__half*)norm_scales, atom::scale_size(seq_len), atom::ArArgs{});
}

int atom_add_rmsnorm_fp16_i4(const void* hidden, const void* residual, void* sum_out, const void* weight, float eps,
                             const void* reorder_index, int seq_len, int hidden_dim, void* o_outliers, void* o_norms,
                             void* outlier_scales, void* norm_scales, void* stream) {
  int rc = quant_check("add_rmsnorm_fp16_i4", seq_len, hidden_dim, o_outliers, o_norms, outlier_scales, norm_scales);
  if (rc) return rc;
  ATOM_REQUIRE(hidden && residual && sum_out && weight && reorder_index && aligned16(hidden) && aligned16(residual) && aligned16(sum_out) &&
               aligned16(weight), "add_rmsnorm_fp16_i4: null or misaligned input");
  ATOM_REQUIRE(hidden_dim <= 32768, "add_rmsnorm_fp16_i4: hidden_dim=%d > 32768 unsupported", hidden_dim);
  if ((rc = ensure_dynamic_smem(atom::rmsnorm_quant_kernel<false>, 32768 * 4 + 512, "add_rmsnorm_fp16_i4"))) return rc;
  return launch_k("add_rmsnorm_fp16_i4", atom::rmsnorm_quant_kernel<false>, dim3(seq_len), dim3(atom::QUANT_THREADS), (size_t)hidden_dim * 4 + 512,
                  (cudaStream_t)stream, (const __half*)hidden, (const __half*)residual, (__half*)sum_out, (const __half*)weight, eps,
                  (const int16_t*)reorder_index, seq_len, hidden_dim, (int8_t*)o_outliers, (uint8_t*)o_norms, (__half*)outlier_scales,
                  (__half*)norm_scales, atom::scale_size(seq_len), atom::ArArgs{});
}

// reduce half of the fused all-reduce: `hidden` is the sum over the ranks of what atom_gemm_i4_o16_push stored in the receive buffers
int atom_reduce_add_rmsnorm_fp16_i4(const void* peer_buffers, void* state, int64_t slot_elems, int rank, int world, const void* residual,
                                    void* sum_out, const void* weight, float eps, const void* reorder_index, int seq_len, int hidden_dim,
                                    void* o_outliers, void* o_norms, void* outlier_scales, void* norm_scales, void* stream) {
  int rc = quant_check("reduce_add_rmsnorm_fp16_i4", seq_len, hidden_dim, o_outliers, o_norms, outlier_scales, norm_scales);
  if (rc) return rc;
  ATOM_REQUIRE(peer_buffers && state && residual && sum_out && weight && reorder_index && aligned16(residual) && aligned16(sum_out) &&
               aligned16(weight), "reduce_add_rmsnorm_fp16_i4: null or misaligned input");
  ATOM_REQUIRE(hidden_dim <= 32768 && hidden_dim % 1024 == 0, "reduce_add_rmsnorm_fp16_i4: hidden_dim=%d must be a multiple of 1024, at most 32768", hidden_dim);
  ATOM_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world && slot_elems % 8 == 0 && (int64_t)seq_len * hidden_dim <= slot_elems,
               "reduce_add_rmsnorm_fp16_i4: rank=%d world=%d, %d x %d must fit a slot of %lld elements", rank, world, seq_len, hidden_dim, (long long)slot_elems);
  if ((rc = ensure_dynamic_smem(atom::rmsnorm_quant_kernel<true>, 32768 * 4 + 512, "reduce_add_rmsnorm_fp16_i4"))) return rc;
  atom::ArArgs ar{(void* const*)peer_buffers, (uint32_t*)state, (long long)slot_elems, rank, world};
  return launch_k("reduce_add_rmsnorm_fp16_i4", atom::rmsnorm_quant_kernel<true>, dim3(seq_len), dim3(atom::QUANT_THREADS), (size_t)hidden_dim * 4 + 512,
                  (cudaStream_t)stream, (const __half*)nullptr, (const __half*)residual, (__half*)sum_out, (const __half*)weight, eps,
                  (const int16_t*)reorder_index, seq_len, hidden_dim, (int8_t*)o_outliers, (uint8_t*)o_norms, (__half*)outlier_scales,
                  (__half*)norm_scales, atom::scale_size(seq_len), ar);
}

int atom_activate_fp16_i4(const void* a, const void* b, int seq_len, int hidden_dim, void* o_outliers, void* o_norms,
                          void* outlier_scales, void* norm_scales, void* stream) {
  int rc = quant_check("activate_fp16_i4", seq_len, hidden_dim, o_outliers, o_norms, outlier_scales, norm_scales);
  if (rc) return rc;
  ATOM_REQUIRE(a && b && aligned16(a) && aligned16(b), "activate_fp16_i4: null or misaligned input");
  const long long units = (long long)seq_len * (hidden_dim / 128);
  return launch_k("activate_fp16_i4", atom::activate_quant_kernel, dim3((unsigned)((units + 7) / 8)), dim3(256), 0, (cudaStream_t)stream,
                  (const __half*)a, (const __half*)b, seq_len, hidden_dim, (int8_t*)o_outliers, (uint8_t*)o_norms,
                  (__half*)outlier_scales, (__half*)norm_scales, atom::scale_size(seq_len));
}

int atom_gemm_i4_o16(const void* a, const void* b, const void* a_scale, const void* b_scale, const void* a_keeper,
                     const void* b_keeper, const void* a_keeper_scale, const void* b_keeper_scale, void* d, int64_t M,
                     int64_t N, int64_t K, uint32_t flags, void* stream) {
  return gemm_common(a, b, a_scale, b_scale, a_keeper, b_keeper, a_keeper_scale, b_keeper_scale, d, nullptr, M, N, K,
                     flags, stream, false);
}

// push half of the fused all-reduce (row-parallel projection): D is stored into every rank's receive buffer
int atom_gemm_i4_o16_push(const void* a, const void* b, const void* a_scale, const void* b_scale, const void* a_keeper,
                          const void* b_keeper, const void* a_keeper_scale, const void* b_keeper_scale, const void* peer_buffers,
                          void* state, int64_t slot_elems, int rank, int world, int64_t M, int64_t N, int64_t K, uint32_t flags,
                          void* stream) {
  atom::ArArgs ar{(void* const*)peer_buffers, (uint32_t*)state, (long long)slot_elems, rank, world};
  return gemm_common(a, b, a_scale, b_scale, a_keeper, b_keeper, a_keeper_scale, b_keeper_scale, nullptr, nullptr, M, N, K, flags,
                     stream, false, &ar);
}

int atom_gemm_i4_qkv(const void* a, const void* b_qkv, const void* a_scale, const void* b_scale_qkv, const void* a_keeper,
                     const void* b_keeper_qkv, const void* a_keeper_scale, const void* b_keeper_scale_qkv, void* q, void* k,
                     void* k_scale, void* v, void* v_scale, int64_t M, int64_t H, int64_t K, uint32_t flags, void* stream) {
  (void)flags;
  ATOM_REQUIRE(a && b_qkv && a_scale && b_scale_qkv && a_keeper && b_keeper_qkv && a_keeper_scale && b_keeper_scale_qkv && q && k &&
               k_scale && v && v_scale, "gemm_i4_qkv: null pointer argument");
  ATOM_REQUIRE(M > 0 && H > 0 && H % 128 == 0, "gemm_i4_qkv: M=%lld must be positive, H=%lld a positive multiple of 128", (long long)M, (long long)H);
  ATOM_REQUIRE(K >= 256 && K % 128 == 0, "gemm_i4_qkv: K=%lld must be a multiple of 128 and >= 256", (long long)K);
  ATOM_REQUIRE(aligned16(a) && aligned16(b_qkv) && aligned16(a_keeper) && aligned16(b_keeper_qkv) && aligned16(q) && aligned16(b_scale_qkv) &&
               aligned16(b_keeper_scale_qkv), "gemm_i4_qkv: operand pointers must be 16-byte aligned");
  ATOM_REQUIRE(M < (1ll << 31) && 3 * H < (1ll << 31) && K < (1ll << 24), "gemm_i4_qkv: dimension too large");
  const int64_t N = 3 * H, kp = (K - 128) / 2;
  atom::GemmArgs args{};
  args.a_scale = (const __half*)a_scale; args.a_keeper_scale = (const __half*)a_keeper_scale;
  args.M = (int)M; args.G = (int)(K / 128 - 1); args.lda_scale = atom::scale_size((int)M); args.trace = g_trace;
  args.a4 = (const uint8_t*)a; args.a8 = (const int8_t*)a_keeper; args.ldb_scale = (int)N;
  if (M <= 64) {
    // one launch: 3H/128 channel tiles, the tile index selects the epilogue (q: FP16, k / v: asymmetric INT4 per head)
    GemmOperands op{a, b_qkv, a_keeper, b_keeper_qkv, M, N, K};
    args.b_scale = (const __half*)b_scale_qkv; args.b_keeper_scale = (const __half*)b_keeper_scale_qkv;
    args.N = (int)N; args.seg_tiles = (int)(H / 128);
    args.d = (__half*)q; args.d4 = (uint8_t*)k; args.d_scale = (__half2*)k_scale; args.d4_v = (uint8_t*)v; args.d_scale_v = (__half2*)v_scale;
    if (M <= 16) return launch_skinny<16, 1, atom::EPI_QKV>(op, args, (cudaStream_t)stream);
    if (M <= 32) return launch_skinny<32, 1, atom::EPI_QKV>(op, args, (cudaStream_t)stream);
    return launch_skinny<64, 1, atom::EPI_QKV>(op, args, (cudaStream_t)stream);
  }
  // prefill sizes: three launches of the tall kernel on the row slices of the fused matrices
  for (int seg = 0; seg < 3; ++seg) {
    GemmOperands op{a, (const uint8_t*)b_qkv + (size_t)seg * H * kp, a_keeper, (const int8_t*)b_keeper_qkv + (size_t)seg * H * 128, M, H, K};
    atom::GemmArgs sa = args;
    sa.b_scale = (const __half*)b_scale_qkv + (size_t)seg * H; sa.b_keeper_scale = (const __half*)b_keeper_scale_qkv + (size_t)seg * H;
    sa.N = (int)H;
    int rc;
    if (seg == 0) { sa.d = (__half*)q; rc = launch_tall<false>(op, sa, (cudaStream_t)stream); }
    else { sa.d4 = (uint8_t*)(seg == 1 ? k : v); sa.d_scale = (__half2*)(seg == 1 ? k_scale : v_scale); rc = launch_tall<true>(op, sa, (cudaStream_t)stream); }
    if (rc) return rc;
  }
  return ATOM_OK;
}

int atom_gemm_i4_gateup_act(const void* a, const void* b_gu, const void* a_scale, const void* b_scale_gu, const void* a_keeper,
                            const void* b_keeper_gu, const void* a_keeper_scale, const void* b_keeper_scale_gu, void* o_outliers,
                            void* o_norms, void* outlier_scales, void* norm_scales, int64_t M, int64_t I, int64_t K, uint32_t flags,
                            void* stream) {
  (void)flags;
  ATOM_REQUIRE(a && b_gu && a_scale && b_scale_gu && a_keeper && b_keeper_gu && a_keeper_scale && b_keeper_scale_gu && o_outliers &&
               o_norms && outlier_scales && norm_scales, "gemm_i4_gateup_act: null pointer argument");
  ATOM_REQUIRE(M > 0 && I >= 256 && I % 128 == 0, "gemm_i4_gateup_act: M=%lld must be positive, I=%lld a multiple of 128 >= 256", (long long)M, (long long)I);
  ATOM_REQUIRE(K >= 256 && K % 128 == 0, "gemm_i4_gateup_act: K=%lld must be a multiple of 128 and >= 256", (long long)K);
  ATOM_REQUIRE(aligned16(a) && aligned16(b_gu) && aligned16(a_keeper) && aligned16(b_keeper_gu) && aligned16(b_scale_gu) &&
               aligned16(b_keeper_scale_gu), "gemm_i4_gateup_act: operand pointers must be 16-byte aligned");
  ATOM_REQUIRE(2 * I < (1ll << 31) && K < (1ll << 24), "gemm_i4_gateup_act: dimension too large");
  if (M > 64) return fail(ATOM_E_UNSUPPORTED, "gemm_i4_gateup_act: fused epilogue exists for decode batches (M <= 64) only; "
                                              "run the two projections and activate_fp16_i4 for M=%lld", (long long)M);
  GemmOperands op{a, b_gu, a_keeper, b_keeper_gu, M, 2 * I, K};
  atom::GemmArgs args{};
  { const char* e = getenv("ATOM_B200_GU_MODE"); args.dbg = e ? atoi(e) : 0; }
  args.a_scale = (const __half*)a_scale; args.a_keeper_scale = (const __half*)a_keeper_scale;
  args.b_scale = (const __half*)b_scale_gu; args.b_keeper_scale = (const __half*)b_keeper_scale_gu;
  args.M = (int)M; args.N = (int)(2 * I); args.G = (int)(K / 128 - 1); args.lda_scale = atom::scale_size((int)M); args.trace = g_trace;
  args.a4 = (const uint8_t*)a; args.a8 = (const int8_t*)a_keeper; args.ldb_scale = (int)(2 * I); args.gu_rows = (int)I;
  args.q8_out = (int8_t*)o_outliers; args.q4_out = (uint8_t*)o_norms; args.q8_scale = (__half*)outlier_scales; args.q4_scale = (__half*)norm_scales;
  if (M <= 16) return launch_skinny<16, 2, atom::EPI_GATEUP>(op, args, (cudaStream_t)stream, I);
  if (M <= 32) return launch_skinny<32, 2, atom::EPI_GATEUP>(op, args, (cudaStream_t)stream, I);
  return launch_skinny<64, 2, atom::EPI_GATEUP>(op, args, (cudaStream_t)stream, I);
}

int atom_gemm_i4_o4(const void* a, const void* b, const void* a_scale, const void* b_scale, const void* a_keeper,
                    const void* b_keeper, const void* a_keeper_scale, const void* b_keeper_scale, void* d,
                    void* d_scale, int64_t M, int64_t N, int64_t K, uint32_t flags, void* stream) {
  return gemm_common(a, b, a_scale, b_scale, a_keeper, b_keeper, a_keeper_scale, b_keeper_scale, d, d_scale, M, N, K,
                     flags, stream, true);
}

int atom_prefill_attention_i4(const void* q, const void* k, const void* k_param, const void* v, const void* v_param,
                              const void* seqlen_indptr, const void* pos_of_token, const void* rope_table, void* k_f16, void* v_f16,
                              void* out, int total_tokens, int batch_size, int max_len, int num_heads, void* stream) {
  ATOM_REQUIRE(q && k && k_param && v && v_param && seqlen_indptr && pos_of_token && rope_table && k_f16 && v_f16 && out,
               "prefill_attention_i4: null pointer argument");
  ATOM_REQUIRE(batch_size > 0 && num_heads > 0 && max_len > 0, "prefill_attention_i4: batch_size=%d num_heads=%d max_len=%d must be positive",
               batch_size, num_heads, max_len);
  ATOM_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(k_f16) && aligned16(v_f16) && aligned16(out),
               "prefill_attention_i4: pointers must be 16-byte aligned");
  if (total_tokens <= 0) return ATOM_OK;
  const long long th = (long long)total_tokens * num_heads;
  atom::kv_dequant_rope_kernel<<<(unsigned)((th + 3) / 4), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)k, (const __half2*)k_param, (const uint8_t*)v, (const __half2*)v_param, (const int32_t*)pos_of_token,
      (const float2*)rope_table, (__half*)k_f16, (__half*)v_f16, th, num_heads);
  int rc = check_launch("prefill_attention_i4 (dequant + RoPE)");
  if (rc) return rc;
  const dim3 grid((unsigned)((max_len + atom::PF_BQ - 1) / atom::PF_BQ), (unsigned)batch_size, (unsigned)num_heads);
  atom::prefill_attn_kernel<<<grid, atom::PF_THREADS, 0, (cudaStream_t)stream>>>(
      (const __half*)q, (const __half*)k_f16, (const __half*)v_f16, (const int32_t*)seqlen_indptr, (const float2*)rope_table,
      (__half*)out, num_heads, 0.08838834764831845f * 1.4426950408889634f);
  return check_launch("prefill_attention_i4");
}

int atom_allreduce_push_f16(const void* in, void* out, const void* peer_buffers, void* state, int64_t numel, int64_t slot_elems,
                            int rank, int world, void* stream) {
  ATOM_REQUIRE(in && out && peer_buffers && state, "allreduce_push_f16: null pointer argument");
  ATOM_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "allreduce_push_f16: rank=%d world=%d", rank, world);
  ATOM_REQUIRE(numel > 0 && numel % 8 == 0 && numel <= slot_elems && slot_elems % 8 == 0,
               "allreduce_push_f16: numel=%lld must be a positive multiple of 8 and fit a slot of %lld elements", (long long)numel, (long long)slot_elems);
  ATOM_REQUIRE(aligned16(in) && aligned16(out), "allreduce_push_f16: in / out must be 16-byte aligned");
  atom::ArArgs ar{(void* const*)peer_buffers, (uint32_t*)state, (long long)slot_elems, rank, world};
  atom::allreduce_push_kernel<<<atom::AR_CTAS, atom::AR_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, ar, numel / 8);
  return check_launch("allreduce_push_f16");
}

int atom_allreduce_state_words(void) { return atom::AR_STATE_WORDS; }

static int kv_check(const char* what, const void* data, const void* param, const void* indptr, const void* indices,
                    const void* last, int L, int layer, int H, int P, int B) {
  ATOM_REQUIRE(data && param && indptr && indices && last, "%s: null pointer argument", what);
  ATOM_REQUIRE(L > 0 && layer >= 0 && layer < L, "%s: layer_idx=%d out of range [0,%d)", what, layer, L);
  ATOM_REQUIRE(H > 0 && P > 0 && B > 0, "%s: num_heads=%d page_size=%d batch_size=%d must be positive", what, H, P, B);
  return ATOM_OK;
}

int atom_batch_decode_i4(void* o, const void* q, const void* kv_data, const void* kv_param, const void* kv_indptr,
                         const void* kv_indices, const void* last_page_offset, int num_layers, int layer_idx,
                         int num_heads, int page_size, int batch_size, void* stream) {
  int rc = kv_check("batch_decode_i4", kv_data, kv_param, kv_indptr, kv_indices, last_page_offset, num_layers, layer_idx,
                    num_heads, page_size, batch_size);
  if (rc) return rc;
  ATOM_REQUIRE(o && q, "batch_decode_i4: null q/o");
  ATOM_REQUIRE(page_size % 8 == 0 && page_size <= 64, "batch_decode_i4: page_size=%d must be a multiple of 8, at most 64", page_size);
  ATOM_REQUIRE(aligned16(kv_data) && aligned16(kv_param), "batch_decode_i4: KV pool must be 16-byte aligned");
  atom::KvArgs kv{(uint8_t*)kv_data, (__half2*)kv_param, (const int32_t*)kv_indptr, (const int32_t*)kv_indices,
                  (const int32_t*)last_page_offset, num_layers, layer_idx, num_heads, page_size, batch_size};
  const size_t smem = atom::batch_decode_smem_bytes(page_size);
#define ATOM_DECODE(TPL, PG)                                                                                              \
  do {                                                                                                                    \
    if ((rc = ensure_dynamic_smem(atom::batch_decode_kernel<TPL, PG>, 100 * 1024, "batch_decode_i4"))) return rc;         \
    return launch_k("batch_decode_i4", atom::batch_decode_kernel<TPL, PG>, dim3(batch_size, num_heads),                   \
                    dim3(atom::DEC_THREADS), smem, (cudaStream_t)stream, (__half*)o, (const __half*)q, kv);               \
  } while (0)
  if (page_size == 16) ATOM_DECODE(2, 16);      // the two page sizes of the harness / the reference benchmarks: strides as immediates
  if (page_size == 32) ATOM_DECODE(4, 32);
  if (page_size <= 32) ATOM_DECODE(4, 0);
  ATOM_DECODE(8, 0);
#undef ATOM_DECODE
}

int atom_append_kv_i4(void* kv_data, void* kv_param, const void* kv_indptr, const void* kv_indices,
                      const void* last_page_offset, const void* k, const void* v, const void* k_param,
                      const void* v_param, int num_layers, int layer_idx, int num_heads, int page_size, int batch_size,
                      void* stream) {
  int rc = kv_check("append_kv_i4", kv_data, kv_param, kv_indptr, kv_indices, last_page_offset, num_layers, layer_idx,
                    num_heads, page_size, batch_size);
  if (rc) return rc;
  ATOM_REQUIRE(k && v && k_param && v_param, "append_kv_i4: null k/v");
  atom::KvArgs kv{(uint8_t*)kv_data, (__half2*)kv_param, (const int32_t*)kv_indptr, (const int32_t*)kv_indices,
                  (const int32_t*)last_page_offset, num_layers, layer_idx, num_heads, page_size, batch_size};
  const long long threads = (long long)batch_size * num_heads * 16;
  return launch_k("append_kv_i4", atom::append_kv_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, kv,
                  (const uint8_t*)k, (const uint8_t*)v, (const __half2*)k_param, (const __half2*)v_param,
                  (const int32_t*)nullptr, batch_size);
}

int atom_init_kv_i4(void* kv_data, void* kv_param, const void* kv_indptr, const void* kv_indices,
                    const void* last_page_offset, const void* k, const void* v, const void* k_param,
                    const void* v_param, const void* seqlen_indptr, int total_tokens, int num_layers, int layer_idx,
                    int num_heads, int page_size, int batch_size, void* stream) {
  int rc = kv_check("init_kv_i4", kv_data, kv_param, kv_indptr, kv_indices, last_page_offset, num_layers, layer_idx,
                    num_heads, page_size, batch_size);
  if (rc) return rc;
  ATOM_REQUIRE(k && v && k_param && v_param && seqlen_indptr, "init_kv_i4: null k/v/seqlen_indptr");
  if (total_tokens <= 0) return ATOM_OK;
  atom::KvArgs kv{(uint8_t*)kv_data, (__half2*)kv_param, (const int32_t*)kv_indptr, (const int32_t*)kv_indices,
                  (const int32_t*)last_page_offset, num_layers, layer_idx, num_heads, page_size, batch_size};
  const long long threads = (long long)total_tokens * num_heads * 16;
  return launch_k("init_kv_i4", atom::append_kv_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, kv,
                  (const uint8_t*)k, (const uint8_t*)v, (const __half2*)k_param, (const __half2*)v_param,
                  (const int32_t*)seqlen_indptr, total_tokens);
}

}  // extern "C"
