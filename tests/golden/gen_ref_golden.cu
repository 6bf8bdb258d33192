// gen_ref_golden.cu -- runs the REFERENCE's own CPU golden functions
// (run_cpu_reorder_fp16_i4 / run_cpu_rmsnorm_fp16_i4 / run_cpu_activate_fp16_i4 from
// e2e/punica-atom/punica/ops/csrc/{Reorder,Norm,Activate}/test_*.cu) on seeded inputs and dumps
// inputs + outputs as raw binary.  Built and run only in the build container by
// tests/golden/make_golden.py; the resulting fixtures are committed as .npz.
// No GPU is touched: the reference's main()/perf_gpu() are compiled but never called.
#define main atom_unused_reference_main
#if defined(GEN_REORDER)
#include "Reorder/test_Reorder.cu"
#elif defined(GEN_RMSNORM)
#include "Norm/test_RMSNorm.cu"
#elif defined(GEN_ACTIVATE)
#include "Activate/test_activate.cu"
#endif
#undef main
#include <cstdio>
#include <vector>

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static float lcg_uniform() {  // (-1, 1)
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return ((float)((lcg_state >> 40) & 0xFFFFFF) / 8388608.0f) - 1.0f;
}
template <typename T> static void dump(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
  const int seq_len = atoi(argv[1]), hidden = atoi(argv[2]);
  FILE* f = fopen(argv[3], "wb");
  const int ldm = SCALE_SIZE_A(seq_len);
  std::vector<half> x((size_t)seq_len * hidden), x2((size_t)seq_len * hidden), w(hidden);
  std::vector<int16_t> idx(hidden);
  for (auto& v : x) v = __float2half(3.0f * lcg_uniform() * (lcg_uniform() > 0.9f ? 8.f : 1.f));
  for (auto& v : x2) v = __float2half(2.0f * lcg_uniform());
  for (auto& v : w) v = __float2half(1.0f + 0.25f * lcg_uniform());
  for (int i = 0; i < hidden; ++i) idx[i] = (int16_t)i;
  for (int i = hidden - 1; i > 0; --i) {  // Fisher-Yates
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    int j = (int)((lcg_state >> 33) % (uint64_t)(i + 1));
    std::swap(idx[i], idx[j]);
  }
  std::vector<int8_t> o8((size_t)seq_len * 128), o4((size_t)seq_len * (hidden - 128) / 2);
  std::vector<half> s8(ldm, __float2half(0.f)), s4((size_t)(hidden / 128 - 1) * ldm, __float2half(0.f));
#if defined(GEN_REORDER)
  run_cpu_reorder_fp16_i4(x.data(), 128, hidden, seq_len, idx.data(), o8.data(), o4.data(), s8.data(), s4.data());
#elif defined(GEN_RMSNORM)
  run_cpu_rmsnorm_fp16_i4(x.data(), w.data(), 1e-5f, 128, hidden, seq_len, idx.data(), o8.data(), o4.data(),
                          s8.data(), s4.data());
#elif defined(GEN_ACTIVATE)
  run_cpu_activate_fp16_i4(x.data(), x2.data(), 128, hidden, seq_len, o8.data(), o4.data(), s8.data(), s4.data());
#endif
  dump(f, x); dump(f, x2); dump(f, w); dump(f, idx); dump(f, o8); dump(f, o4); dump(f, s8); dump(f, s4);
  fclose(f);
  return 0;
}
