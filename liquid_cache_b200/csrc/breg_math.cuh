// breg_math.cuh — the register-resident FastLanes unpack of k_int_bits (k_int_bits.cu): which words of one 1024-value
// chunk a thread of the warp loads, how they line up as the thread's LOCAL bit stream, and which W-bit field of that stream
// belongs to (step s, lane) — plus the mask word step s's ballot is. Host + device: the kernel loads from global memory
// through `Loader`, the CPU test (tests/cpp/breg_host.cc, tests/test_breg_cpu.py) from a byte buffer, and both run THIS
// code, so every (T, W) the kernel is instantiated for is checked against a plain FastLanes unpack before a GPU sees it.
//
// fastlanes 0.5.0 unified transposed order (bit_pack_array.rs:71-169 calls it): a chunk of a T-bit column has LANES =
// 1024/T lanes; lane L's W-bit fields of rows r = 0..T-1 are concatenated into W T-bit words, word k stored at
// chunk[LANES*k + L]; logical index of (r, L) = (r%8)*128 + FL_ORDER[r/8]*16 + L, FL_ORDER = {0,4,2,6,1,5,3,7}.
// A warp covers the chunk in 32 steps; in step s the 32 threads hold the 32 CONSECUTIVE logical rows of mask word
// out_word(s), thread `lane` = bit `lane`:
//   T = 32  thread = lane L, step s = row s                        -> W coalesced 4-byte loads (one 128-byte row each)
//   T = 64  16 lanes: threads 0-15 take rows 0..31, threads 16-31 rows 32..63 of lane L = lane%16 (FL_ORDER[o+4] =
//           FL_ORDER[o]+1 makes the two halves the low / high 16 bits of one word). Rows 0..31 are bits [0, 32W) of the
//           lane's stream, rows 32..63 bits [32W, 64W): EXACTLY W 32-bit words each, one 4-byte load per word (a warp
//           instruction covers the even or the odd words of two 128-byte lines; its pair fetches the other half from L1).
//   T = 16  64 lanes: thread takes lanes `lane` and 32+lane (sub-lanes 0, 1), 16 rows each, step s = row s/2, sub s%2
//   T = 8   128 lanes: four sub-lanes, 8 rows each, step s = row s/4, sub s%4
// Every index below is a compile-time constant once the step loop is unrolled, so the stream lives in registers.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define LC_BR_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define LC_BR_UNROLL _Pragma("unroll")
#else
#define LC_BR_UNROLL
#endif
#else
#define LC_BR_HD inline
#define LC_BR_UNROLL
#endif

namespace lc {

template <uint32_t T, uint32_t W>
struct BregGeom {
  static_assert(T == 8 || T == 16 || T == 32 || T == 64, "column widths");
  static_assert(W >= 1 && W <= 32 && W <= T, "fields of at most 32 bits (64-bit columns with wider fields take k_int_scan)");
  static constexpr uint32_t SUB = T >= 32 ? 1u : 32u / T;       // sub-lanes a thread decodes
  static constexpr uint32_t ROWS = T >= 32 ? 32u : T;            // local rows per sub-lane
  static constexpr uint32_t NW = (ROWS * W + 31u) / 32u;         // 32-bit words of a sub-lane's local stream
};

// mask word (0..31 inside the chunk) that step s's ballot is
template <uint32_t T>
LC_BR_HD uint32_t breg_out_word(uint32_t s) {
  if (T >= 32) {
    const uint32_t hi = s >> 3;  // 0..3, bit-reversed over two bits
    return (s & 7u) * 4u + (((hi & 1u) << 1) | (hi >> 1));
  }
  if (T == 16) {
    const uint32_t r = s >> 1;
    return (r & 7u) * 4u + (r >> 3) * 2u + (s & 1u);
  }
  return s;
}

// Loader: ld8/ld16/ld32(byte offset) -> uint32_t, ld64(byte offset, &lo, &hi)
template <uint32_t T, uint32_t W, typename Loader>
LC_BR_HD void breg_load(uint32_t lane, uint32_t (&a)[BregGeom<T, W>::SUB][BregGeom<T, W>::NW], Loader ld) {
  using G = BregGeom<T, W>;
  if (T == 32) {
LC_BR_UNROLL
    for (uint32_t k = 0; k < W; ++k) a[0][k] = ld.ld32(128u * k + 4u * lane);
  } else if (T == 64) {
    // 32-bit word x of lane l sits at byte (x/2)*128 + (x%2)*4 + 8*l. The lower half-warp needs x = i, the upper x = W + i
    // (i = 0..W-1): a constant byte distance for even W, and one of two constants (by the parity of i) for odd W — two
    // per-thread bases make every load `[base + immediate]`.
    constexpr uint32_t dE = (W % 2u == 0u) ? (W / 2u) * 128u : ((W - 1u) / 2u) * 128u + 4u;
    constexpr uint32_t dO = (W % 2u == 0u) ? (W / 2u) * 128u : ((W + 1u) / 2u) * 128u - 4u;
    const uint32_t hh = lane >> 4, l = lane & 15u;
    const uint32_t base_e = 8u * l + (hh ? dE : 0u), base_o = 8u * l + (hh ? dO : 0u);
LC_BR_UNROLL
    for (uint32_t i = 0; i < W; ++i) a[0][i] = ld.ld32(((i & 1u) ? base_o : base_e) + (i >> 1) * 128u + (i & 1u) * 4u);
  } else if (T == 16) {
LC_BR_UNROLL
    for (uint32_t h = 0; h < 2u; ++h) {
LC_BR_UNROLL
      for (uint32_t i = 0; i < G::NW; ++i) {
        uint32_t v = ld.ld16(128u * (2u * i) + 2u * (32u * h + lane));
        if (2u * i + 1u < W) v |= ld.ld16(128u * (2u * i + 1u) + 2u * (32u * h + lane)) << 16;
        a[h][i] = v;
      }
    }
  } else {
LC_BR_UNROLL
    for (uint32_t h = 0; h < 4u; ++h) {
LC_BR_UNROLL
      for (uint32_t i = 0; i < G::NW; ++i) {
        uint32_t v = 0;
LC_BR_UNROLL
        for (uint32_t b = 0; b < 4u; ++b)
          if (4u * i + b < W) v |= ld.ld8(128u * (4u * i + b) + (32u * h + lane)) << (8u * b);
        a[h][i] = v;
      }
    }
  }
}

// the packed value this thread holds at step s (s a compile-time constant in the unrolled loop)
template <uint32_t T, uint32_t W>
LC_BR_HD uint32_t breg_value(const uint32_t (&a)[BregGeom<T, W>::SUB][BregGeom<T, W>::NW], uint32_t s) {
  constexpr uint32_t mask32 = W >= 32u ? 0xffffffffu : ((1u << (W & 31u)) - 1u);
  const uint32_t h = T >= 32 ? 0u : (T == 16 ? (s & 1u) : (s & 3u));
  const uint32_t rr = T >= 32 ? s : (T == 16 ? (s >> 1) : (s >> 2));
  const uint32_t b = rr * W, k = b >> 5, sh = b & 31u;
  const uint32_t w0 = a[h][k];
  if (sh + W > 32u) {
    const uint32_t w1 = a[h][k + 1u < BregGeom<T, W>::NW ? k + 1u : k];  // a field that straddles never sits in the last word
#ifdef __CUDA_ARCH__
    return __funnelshift_r(w0, w1, sh) & mask32;
#else
    return static_cast<uint32_t>(((static_cast<uint64_t>(w1) << 32) | w0) >> sh) & mask32;
#endif
  }
  return (w0 >> sh) & mask32;
}

}  // namespace lc
