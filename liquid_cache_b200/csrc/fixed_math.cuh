// fixed_math.cuh — byte arithmetic of LiquidFixedLenByteArray entries, host + device, so the same code is exercised on the
// CPU (tests/cpp/fixed_math_host.cc, tests/test_fixed_math_cpu.py). Values are kept in ORDER-PRESERVING form: the
// little-endian two's complement integer byte-reversed (big-endian) with the sign bit flipped, so that unsigned
// lexicographic byte order — what the byte-view comparison kernels implement — is the numeric order of the decimals.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define LC_FX_HD __host__ __device__ __forceinline__
#else
#define LC_FX_HD inline
#endif

namespace lc {

// in place: little-endian integer of `width` bytes -> order-preserving form
LC_FX_HD void fixed_to_ordered_inplace(uint8_t* p, uint32_t width) {
  for (uint32_t i = 0; i < width / 2u; ++i) {
    const uint8_t a = p[i], b = p[width - 1u - i];
    p[i] = b;
    p[width - 1u - i] = a;
  }
  p[0] ^= 0x80u;
}

// 32-bit word `wdx` of the little-endian integer whose order-preserving form starts at `stored`:
// output bytes 4*wdx .. 4*wdx+3 are stored bytes width-1-4*wdx .. width-4-4*wdx; the sign bit lives in stored byte 0
LC_FX_HD uint32_t fixed_le_word(const uint8_t* stored, uint32_t width, uint32_t wdx) {
  const uint8_t* p = stored + (width - 4u - 4u * wdx);
  uint32_t v = static_cast<uint32_t>(p[3]) | (static_cast<uint32_t>(p[2]) << 8) | (static_cast<uint32_t>(p[1]) << 16) |
               (static_cast<uint32_t>(p[0]) << 24);
  if (wdx == (width >> 2) - 1u) v ^= 0x80000000u;
  return v;
}

// out of place (the host's twin of fixed_to_ordered_inplace)
LC_FX_HD void fixed_to_ordered(const uint8_t* le, uint32_t width, uint8_t* out) {
  for (uint32_t i = 0; i < width; ++i) out[i] = le[width - 1u - i];
  out[0] ^= 0x80u;
}

// The needle of `decimal_col <op> literal` on a LiquidFixedLenByteArray entry: the literal — LC_LIT_I128 halves (sign-extended
// to the column's width) or, when `le` is given, the column's own little-endian integer — in order-preserving form.
LC_FX_HD void fixed_needle(uint64_t lit_u64, int64_t lit_i64, const uint8_t* le, uint32_t width, uint8_t* out) {
  uint8_t tmp[32];
  if (le) {
    for (uint32_t i = 0; i < width; ++i) tmp[i] = le[i];
  } else {
    for (uint32_t i = 0; i < 8; ++i) {
      tmp[i] = static_cast<uint8_t>(lit_u64 >> (8u * i));
      tmp[8 + i] = static_cast<uint8_t>(static_cast<uint64_t>(lit_i64) >> (8u * i));
    }
    for (uint32_t i = 16; i < 32; ++i) tmp[i] = lit_i64 < 0 ? 0xFFu : 0x00u;
  }
  fixed_to_ordered(tmp, width, out);
}

}  // namespace lc
