// alp_math.cuh — per-value arithmetic of the ALP float encoding and of float comparisons, device side.
//
// Reference: trait LiquidFloatType (/root/reference/src/core/src/liquid_array/float_array.rs:70-225):
//   encode_single_unchecked  fast_round(val * F10[e] * IF10[f]),  fast_round(x) = ((x + SWEET) - SWEET) as iN
//   decode_single            (val as fN) * F10[f] * IF10[e]
// Every product and sum below is a single correctly rounded IEEE operation issued through the _rn intrinsics, so
// nvcc cannot contract a multiply and the following add into an FMA (the reference rounds after each step).
// The power-of-ten tables are the reference's decimal literals (float_array.rs:135-160, 170-222) written as hex
// floats of their correctly rounded values (generated from exact rationals; tests/test_oracle_num.py compares them with the checker's own).
// Float comparisons follow arrow-ord's total order (`f64::total_cmp`, the ordering `cmp::{eq,lt,..}` of arrow-rs
// documents for floating point arrays): sign-magnitude bits mapped to a two's complement key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lc {

// constant-bank tables: (e, f) are uniform per entry, so every lane of a warp reads the same word
static __constant__ double kAlpF10d[24] = {0x1.0000000000000p+0, 0x1.4000000000000p+3, 0x1.9000000000000p+6, 0x1.f400000000000p+9, 0x1.3880000000000p+13, 0x1.86a0000000000p+16, 0x1.e848000000000p+19, 0x1.312d000000000p+23, 0x1.7d78400000000p+26, 0x1.dcd6500000000p+29, 0x1.2a05f20000000p+33, 0x1.74876e8000000p+36, 0x1.d1a94a2000000p+39, 0x1.2309ce5400000p+43, 0x1.6bcc41e900000p+46, 0x1.c6bf526340000p+49, 0x1.1c37937e08000p+53, 0x1.6345785d8a000p+56, 0x1.bc16d674ec800p+59, 0x1.158e460913d00p+63, 0x1.5af1d78b58c40p+66, 0x1.b1ae4d6e2ef50p+69, 0x1.0f0cf064dd592p+73, 0x1.52d02c7e14af6p+76};
static __constant__ double kAlpIF10d[24] = {0x1.0000000000000p+0, 0x1.999999999999ap-4, 0x1.47ae147ae147bp-7, 0x1.0624dd2f1a9fcp-10, 0x1.a36e2eb1c432dp-14, 0x1.4f8b588e368f1p-17, 0x1.0c6f7a0b5ed8dp-20, 0x1.ad7f29abcaf48p-24, 0x1.5798ee2308c3ap-27, 0x1.12e0be826d695p-30, 0x1.b7cdfd9d7bdbbp-34, 0x1.5fd7fe1796495p-37, 0x1.19799812dea11p-40, 0x1.c25c268497682p-44, 0x1.6849b86a12b9bp-47, 0x1.203af9ee75616p-50, 0x1.cd2b297d889bcp-54, 0x1.70ef54646d497p-57, 0x1.2725dd1d243acp-60, 0x1.d83c94fb6d2acp-64, 0x1.79ca10c924223p-67, 0x1.2e3b40a0e9b4fp-70, 0x1.e392010175ee6p-74, 0x1.82db34012b251p-77};
static __constant__ float kAlpF10f[11] = {0x1.0000000000000p+0f, 0x1.4000000000000p+3f, 0x1.9000000000000p+6f, 0x1.f400000000000p+9f, 0x1.3880000000000p+13f, 0x1.86a0000000000p+16f, 0x1.e848000000000p+19f, 0x1.312d000000000p+23f, 0x1.7d78400000000p+26f, 0x1.dcd6500000000p+29f, 0x1.2a05f20000000p+33f};
static __constant__ float kAlpIF10f[11] = {0x1.0000000000000p+0f, 0x1.99999a0000000p-4f, 0x1.47ae140000000p-7f, 0x1.0624de0000000p-10f, 0x1.a36e2e0000000p-14f, 0x1.4f8b580000000p-17f, 0x1.0c6f7a0000000p-20f, 0x1.ad7f2a0000000p-24f, 0x1.5798ee0000000p-27f, 0x1.12e0be0000000p-30f, 0x1.b7cdfe0000000p-34f};

template <typename F>
struct Alp;

template <>
struct Alp<double> {
  using I = long long;
  using U = unsigned long long;
  static constexpr uint32_t kMaxExponent = 18;  // MAX_EXPONENT: e in 0..18, f in 0..e
  static __device__ __forceinline__ double f10(uint32_t i) {
    return kAlpF10d[i];
  }
  static __device__ __forceinline__ double if10(uint32_t i) {
    return kAlpIF10d[i];
  }
  static __device__ __forceinline__ I encode(double v, uint32_t e, uint32_t f) {
    const double sweet = 6755399441055744.0;  // 2^52 + 2^51
    const double x = __dmul_rn(__dmul_rn(v, f10(e)), if10(f));
    const double r = __dadd_rn(__dadd_rn(x, sweet), -sweet);
    // Rust `as i64`: NaN -> 0, saturating at the ends
    if (r != r) return 0;
    if (r >= 9223372036854775808.0) return 0x7fffffffffffffffLL;
    if (r <= -9223372036854775808.0) return static_cast<I>(0x8000000000000000ULL);
    return __double2ll_rz(r);
  }
  static __device__ __forceinline__ double decode(I enc, uint32_t e, uint32_t f) {
    return __dmul_rn(__dmul_rn(__ll2double_rn(enc), f10(f)), if10(e));
  }
  static __device__ __forceinline__ I bits(double v) { return __double_as_longlong(v); }
  static __device__ __forceinline__ double from_bits(I b) { return __longlong_as_double(b); }
  // key(a) < key(b) as signed integers  <=>  a.total_cmp(b) == Less
  static __device__ __forceinline__ I order_key(I b) { return b ^ static_cast<I>(static_cast<U>(b >> 63) >> 1); }
};

template <>
struct Alp<float> {
  using I = int;
  using U = unsigned int;
  static constexpr uint32_t kMaxExponent = 10;
  static __device__ __forceinline__ float f10(uint32_t i) {
    return kAlpF10f[i];
  }
  static __device__ __forceinline__ float if10(uint32_t i) {
    return kAlpIF10f[i];
  }
  static __device__ __forceinline__ I encode(float v, uint32_t e, uint32_t f) {
    const float sweet = 12582912.0f;  // 2^23 + 2^22
    const float x = __fmul_rn(__fmul_rn(v, f10(e)), if10(f));
    const float r = __fadd_rn(__fadd_rn(x, sweet), -sweet);
    if (r != r) return 0;
    if (r >= 2147483648.0f) return 0x7fffffff;
    if (r <= -2147483648.0f) return static_cast<I>(0x80000000u);
    return __float2int_rz(r);
  }
  static __device__ __forceinline__ float decode(I enc, uint32_t e, uint32_t f) {
    return __fmul_rn(__fmul_rn(__int2float_rn(enc), f10(f)), if10(e));
  }
  static __device__ __forceinline__ I bits(float v) { return __float_as_int(v); }
  static __device__ __forceinline__ float from_bits(I b) { return __int_as_float(b); }
  static __device__ __forceinline__ I order_key(I b) { return b ^ static_cast<I>(static_cast<U>(b >> 31) >> 1); }
};

// get_bit_width (utils/mod.rs:24-32)
__device__ __forceinline__ uint32_t bit_width_of_u64(unsigned long long max_value) {
  return max_value == 0 ? 1u : 64u - static_cast<uint32_t>(__clzll(static_cast<long long>(max_value)));
}

// Number of (e, f) pairs get_best_exponents walks: e in 0..MAX, f in 0..e  (float_array.rs:730-746)
template <typename F>
__host__ __device__ constexpr uint32_t alp_n_combos() {
  return Alp<F>::kMaxExponent * (Alp<F>::kMaxExponent - 1) / 2;
}

}  // namespace lc
