// Byte arithmetic of LiquidFixedLenByteArray entries (liquid_cache_b200/csrc/fixed_math.cuh) compiled for the HOST and looped
// over arrays: the functions k_fixed_to_ordered / k_fixed_from_var run per thread, checked on the CPU
// (tests/test_fixed_math_cpu.py). No CUDA in this file.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/fixed_math.cuh"

extern "C" {

void fx_to_ordered(const uint8_t* le, uint32_t n, uint32_t width, uint8_t* out) {
  std::memcpy(out, le, static_cast<size_t>(n) * width);
  for (uint32_t i = 0; i < n; ++i) lc::fixed_to_ordered_inplace(out + static_cast<size_t>(i) * width, width);
}

void fx_needle(uint64_t lit_u64, int64_t lit_i64, const uint8_t* le, uint32_t width, uint8_t* out) {
  lc::fixed_needle(lit_u64, lit_i64, le, width, out);
}

void fx_from_ordered(const uint8_t* stored, uint32_t n, uint32_t width, uint32_t* out) {
  const uint32_t wpr = width / 4;
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t w = 0; w < wpr; ++w) out[static_cast<size_t>(i) * wpr + w] = lc::fixed_le_word(stored + static_cast<size_t>(i) * width, width, w);
}
}
