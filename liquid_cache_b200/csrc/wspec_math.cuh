// wspec_math.cuh — where the width-specialised integer scan (k_int.cu int_bits_fast_w) finds the packed value of
// (storage-order step j, lane) inside one FastLanes chunk, and which output word / bit that value belongs to. Host + device:
// the kernel reads shared memory through `Reader`, the CPU test (tests/cpp/wspec_host.cc, tests/test_wspec_cpu.py) reads a
// byte buffer, and both run THIS code — so every (T, W) the kernel is instantiated for is checked against a plain FastLanes
// unpack before a GPU sees it.
// Geometry: T = 32: step j = packed row j, word k = j*W/32 of lane `lane` at byte 128*k + 4*lane.
// T = 64: 16 lanes of 64-bit words; half-warp h takes row r + 32, i.e. the same shift and 32-bit word index x + W. 32-bit
// word x of lane L sits at byte (x/2)*128 + (x%2)*4 + 8*L, so moving by W words is a constant byte distance when W is even
// and one of two constants (by the parity of x) when W is odd: two per-lane bases cover both.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define LC_WS_HD __host__ __device__ __forceinline__
#else
#define LC_WS_HD inline
#endif

namespace lc {

// output word of step j's ballot inside the chunk's 32 mask words (the unified transposed order of fastlanes 0.5.0)
LC_WS_HD uint32_t wspec_out_word(uint32_t j) {
  const uint32_t hi = j >> 3;  // 0..3, bit-reversed over two bits
  return (j & 7u) * 4u + (((hi & 1u) << 1) | (hi >> 1));
}

template <uint32_t T, uint32_t W>
struct WspecBases {
  uint32_t lbase, baseE, baseO;
};

template <uint32_t T, uint32_t W>
LC_WS_HD WspecBases<T, W> wspec_bases(uint32_t chunk_base, uint32_t lane) {
  static_assert(T == 32 || T == 64, "32- and 64-bit columns");
  static_assert(W >= 1 && W <= 32, "fields of at most 32 bits");
  // byte distance of "W 32-bit words further down the lane's stream" for an even / an odd word index (T = 64 only)
  constexpr uint32_t dE = (W % 2u == 0u) ? (W / 2u) * 128u : ((W - 1u) / 2u) * 128u + 4u;
  constexpr uint32_t dO = (W % 2u == 0u) ? (W / 2u) * 128u : ((W + 1u) / 2u) * 128u - 4u;
  const uint32_t half = T == 64 ? (lane >> 4) : 0u;
  WspecBases<T, W> b;
  b.lbase = chunk_base + (T == 64 ? (lane & 15u) * 8u : lane * 4u);
  b.baseE = b.lbase + (half ? dE : 0u);
  b.baseO = b.lbase + (half ? dO : 0u);
  return b;
}

// the W-bit packed value this lane holds at step j (j a compile-time constant in the unrolled kernel loop)
template <uint32_t T, uint32_t W, typename Reader>
LC_WS_HD uint32_t wspec_value(const WspecBases<T, W>& bs, uint32_t j, Reader rd) {
  constexpr uint32_t mask32 = W >= 32u ? 0xffffffffu : ((1u << (W & 31u)) - 1u);
  const uint32_t b = j * W, x = b >> 5, sh = b & 31u;  // half-warp 0's packed row of this step is j itself
  const bool two = sh + W > 32u;
  const uint32_t a0 = T == 64 ? ((x & 1u) ? bs.baseO : bs.baseE) + (x >> 1) * 128u + (x & 1u) * 4u : bs.lbase + x * 128u;
  const uint32_t w0 = rd(a0);
  if (two) {
    const uint32_t y = x + 1u;
    const uint32_t a1 = T == 64 ? ((y & 1u) ? bs.baseO : bs.baseE) + (y >> 1) * 128u + (y & 1u) * 4u : bs.lbase + y * 128u;
    const uint32_t w1 = rd(a1);
#ifdef __CUDA_ARCH__
    return __funnelshift_r(w0, w1, sh) & mask32;
#else
    const uint64_t both = (static_cast<uint64_t>(w1) << 32) | w0;  // the same funnel shift right
    return static_cast<uint32_t>(both >> sh) & mask32;
#endif
  }
  return (w0 >> sh) & mask32;
}

}  // namespace lc
