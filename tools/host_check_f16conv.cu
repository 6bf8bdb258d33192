// host_check_f16conv.cu -- CPU unit test of the FP16-path operand conversion (atom_b200/csrc/w4_f16_convert.cuh): the very
// functions the kernel runs, executed on the host (cuda_fp16 intrinsics are host-callable).  No GPU involved.
//   nvcc -std=c++17 -o /tmp/host_check_f16conv tools/host_check_f16conv.cu && /tmp/host_check_f16conv
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../atom_b200/csrc/w4_f16_convert.cuh"

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state; }
static uint16_t bits(__half h) { uint16_t u; memcpy(&u, &h, 2); return u; }

int main() {
  int bad = 0;
  // lop3 emulation == the boolean expression it stands for
  for (int it = 0; it < 100000; ++it) {
    const uint32_t w = rnd();
    const uint32_t expect = (((w & 0x000F000Fu) ^ 0x00080008u) | 0x64006400u);
    if (atom::lop3<0x6A>(w, 0x000F000Fu, 0x64086408u) != expect) { ++bad; if (bad < 5) printf("lop3 mismatch %08x\n", w); }
  }
  // nibbles: every element equals the correctly rounded fp16 product, in the documented order
  for (int it = 0; it < 200000; ++it) {
    const uint32_t w = rnd();
    const float sf = ldexpf(1.0f + (rnd() & 1023) / 1024.0f, -(int)(rnd() % 12) - 1);        // scales in [2^-13, 1)
    const __half s = __float2half_rn(sf);
    __half2 out[4];
    atom::nib8_to_f16(w, __half2half2(s), out);
    const int order[8] = {0, 4, 1, 5, 2, 6, 3, 7};
    for (int p = 0; p < 8; ++p) {
      int e = (w >> (4 * order[p])) & 0xF; if (e >= 8) e -= 16;
      const __half expect = __float2half_rn((float)e * __half2float(s));
      const __half got = (p & 1) ? __high2half(out[p >> 1]) : __low2half(out[p >> 1]);
      if (bits(expect) != bits(got) && !(__half2float(expect) == 0.f && __half2float(got) == 0.f)) {
        ++bad; if (bad < 10) printf("nib mismatch w=%08x p=%d e=%d s=%g got=%g expect=%g\n", w, p, e, __half2float(s), __half2float(got), __half2float(expect));
      }
    }
  }
  // int8
  for (int it = 0; it < 200000; ++it) {
    const uint32_t w = rnd();
    const __half s = __float2half_rn(ldexpf(1.0f + (rnd() & 1023) / 1024.0f, -(int)(rnd() % 10) - 1));
    __half2 out[2];
    atom::i8x4_to_f16(w, __half2half2(s), out);
    for (int p = 0; p < 4; ++p) {
      const int b = (int8_t)((w >> (8 * p)) & 0xFF);
      const __half expect = __float2half_rn((float)b * __half2float(s));
      const __half got = (p & 1) ? __high2half(out[p >> 1]) : __low2half(out[p >> 1]);
      if (bits(expect) != bits(got) && !(__half2float(expect) == 0.f && __half2float(got) == 0.f)) {
        ++bad; if (bad < 10) printf("i8 mismatch w=%08x p=%d b=%d got=%g expect=%g\n", w, p, b, __half2float(got), __half2float(expect));
      }
    }
  }
  // swizzle: a bijection of (row, chunk) onto the tile, every 8 rows one 1024-B atom, chunk index XOR row%8
  {
    static unsigned char seen[256 * 128];
    memset(seen, 0, sizeof(seen));
    for (int r = 0; r < 256; ++r)
      for (int c = 0; c < 8; ++c) {
        const uint32_t o = atom::sw128_chunk_offset(r, c);
        if (o % 16 || o >= sizeof(seen) || seen[o]) { ++bad; printf("swizzle clash r=%d c=%d\n", r, c); }
        seen[o] = 1;
        if (o / 1024 != (uint32_t)(r / 8) || (o % 1024) / 128 != (uint32_t)(r % 8) || ((o % 128) / 16) != (uint32_t)(c ^ (r % 8))) ++bad;
      }
  }
  printf(bad ? "FAILED: %d mismatches\n" : "ok: lop3 truth table, INT4 and INT8 -> fp16 conversion (bit exact vs correctly rounded products), swizzle map\n", bad);
  return bad != 0;
}
