// squeeze_math.cuh — the per-value arithmetic of the squeeze kernels (k_num.cu), as host + device functions so that the
// very same code is exercised on the CPU against the oracle (tests/cpp/squeeze_math_host.cu, tests/test_squeeze_math_cpu.py).
// Reference: liquid_array/squeezed_date32_array.rs:360-427 (calendar), :326-356 (lossy dates),
// liquid_array/primitive_array.rs:427-481 (clamp / quantize codes) — all under /root/reference/src/core/src.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define LC_SQ_HD __host__ __device__ __forceinline__
#else
#define LC_SQ_HD inline
#endif

namespace lc {

// ymd_from_epoch_days (:360-379) / ymd_to_epoch_days (:416-427): Hinnant's civil_from_days / days_from_civil in 64-bit
// integers, `/` truncating like Rust's.
LC_SQ_HD void civil_from_days(int32_t days, long long* y, long long* m, long long* d) {
  const long long z = static_cast<long long>(days) + 719468ll;
  const long long era = (z >= 0 ? z : z - 146096ll) / 146097ll;
  const long long doe = z - era * 146097ll;
  const long long yoe = (doe - doe / 1460ll + doe / 36524ll - doe / 146096ll) / 365ll;
  const long long doy = doe - (365ll * yoe + yoe / 4ll - yoe / 100ll);
  const long long mp = (5ll * doy + 2ll) / 153ll;
  *d = doy - (153ll * mp + 2ll) / 5ll + 1ll;
  *m = mp + (mp < 10 ? 3ll : -9ll);
  *y = yoe + era * 400ll + (*m <= 2 ? 1ll : 0ll);
}
LC_SQ_HD int32_t days_from_civil(long long year, long long m, long long d) {
  const long long y = year - (m <= 2 ? 1ll : 0ll);
  const long long era = (y >= 0 ? y : y - 399ll) / 400ll;
  const long long yoe = y - era * 400ll;
  const long long mp = m + (m > 2 ? -3ll : 9ll);
  const long long doy = (153ll * mp + 2ll) / 5ll + d - 1ll;
  const long long doe = yoe * 365ll + yoe / 4ll - yoe / 100ll + doy;
  return static_cast<int32_t>(static_cast<unsigned long long>(era * 146097ll + doe - 719468ll));  // `as i32`
}
// timestamp_to_days_since_epoch (:402-410): value.div_euclid(ticks_per_day) as i32
LC_SQ_HD int32_t days_of_ticks(long long v, long long ticks_per_day) {
  long long q = v / ticks_per_day;
  if (v % ticks_per_day < 0) --q;
  return static_cast<int32_t>(static_cast<unsigned long long>(q));
}
// component_from_days (:381-393); field: 0 year, 1 month, 2 day, 3 day of week (Sunday = 0)
LC_SQ_HD int32_t date_component(uint32_t field, int32_t days) {
  if (field == 3u) {
    const int32_t r = static_cast<int32_t>((static_cast<long long>(days) + 4ll) % 7ll);
    return r < 0 ? r + 7 : r;  // rem_euclid
  }
  long long y, m, d;
  civil_from_days(days, &y, &m, &d);
  return static_cast<int32_t>(field == 0u ? y : field == 1u ? m : d);
}
// to_arrow_date32_lossy (:326-356): a date whose component is the stored one — Year -> (y,1,1), Month -> (1970,m,1),
// Day -> (1970,1,d), DayOfWeek -> 1970-01-04 + dow (saturating)
LC_SQ_HD int32_t lossy_days(uint32_t field, int32_t c) {
  if (field == 0u) return days_from_civil(c, 1, 1);
  if (field == 1u) return days_from_civil(1970, static_cast<long long>(static_cast<uint32_t>(c)), 1);
  if (field == 2u) return days_from_civil(1970, 1, static_cast<long long>(static_cast<uint32_t>(c)));
  const long long t = 3ll + static_cast<long long>(c);  // 1970-01-04 is day 3
  return t > 2147483647ll ? 2147483647 : t < -2147483648ll ? static_cast<int32_t>(-2147483647 - 1) : static_cast<int32_t>(t);
}
// LiquidPrimitiveArray::squeeze (primitive_array.rs:427-481): offset -> clamped offset (limit = sentinel) or bucket index
// (limit = bucket_count - 1)
LC_SQ_HD unsigned long long squeeze_code(unsigned long long off, uint32_t quantize, unsigned long long limit, unsigned long long bucket_width) {
  if (quantize) {
    const unsigned long long code = off / bucket_width;
    return code > limit ? limit : code;
  }
  return off >= limit ? limit : off;
}

}  // namespace lc
