// The per-value arithmetic of the squeeze kernels (liquid_cache_b200/csrc/squeeze_math.cuh) compiled for the HOST and looped
// over arrays, so that the exact functions k_date_component / k_date_lossy / k_squeeze_map execute per thread are checked
// on the CPU against the oracle (tests/test_squeeze_math_cpu.py). No CUDA anywhere in this file.
#include <cstdint>

#include "liquid_cache_b200/csrc/squeeze_math.cuh"

extern "C" {

void sq_date_component(const int64_t* in, uint32_t n, uint32_t field, long long ticks_per_day, int32_t* out) {
  for (uint32_t i = 0; i < n; ++i) {
    const int32_t days = ticks_per_day ? lc::days_of_ticks(in[i], ticks_per_day) : static_cast<int32_t>(in[i]);
    out[i] = lc::date_component(field, days);
  }
}

void sq_lossy_days(const int32_t* comp, uint32_t n, uint32_t field, int32_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = lc::lossy_days(field, comp[i]);
}

void sq_civil(const int32_t* days, uint32_t n, long long* y, long long* m, long long* d) {
  for (uint32_t i = 0; i < n; ++i) lc::civil_from_days(days[i], &y[i], &m[i], &d[i]);
}

int32_t sq_days_from_civil(long long y, long long m, long long d) { return lc::days_from_civil(y, m, d); }

void sq_codes(const uint64_t* offsets, uint32_t n, uint32_t quantize, unsigned long long limit, unsigned long long bucket_width, uint64_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = lc::squeeze_code(offsets[i], quantize, limit, bucket_width);
}
}
