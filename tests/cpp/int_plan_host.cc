// The predicate planner of k_int_scan (liquid_cache_b200/csrc/int_plan.cuh: plan_int_pred + make_range) compiled for the
// HOST: given an entry header and `col <op> literal`, which of a list of packed values pass — evaluated exactly as the
// scan loops do, ((u - lo) <= span) != neg in 32 or 64 bits. Checked on the CPU in tests/test_int_plan_cpu.py.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/int_plan.cuh"

extern "C" {

// header fields: tbits, bit_width, is_signed, reference (raw bits, zero extended), squeeze_kind, bucket_width
// predicate: op (lc_op EQ..GE), lit_kind (0 I64, 1 U64, 7 above-all, 8 sentinel), lit_i, lit_u
// packed[n] -> out[n] (0 / 1); returns the UCmp kind the planner chose, *thr_out its threshold
int ip_eval(uint32_t tbits, uint32_t bit_width, uint32_t is_signed, uint64_t reference, uint32_t squeeze_kind, uint64_t bucket_width,
            int32_t op, int32_t lit_kind, int64_t lit_i, uint64_t lit_u, const uint64_t* packed, uint32_t n, uint8_t* out, uint64_t* thr_out) {
  lc::IntHeader h;
  std::memset(&h, 0, sizeof(h));
  h.tbits = static_cast<uint8_t>(tbits);
  h.bit_width = static_cast<uint8_t>(bit_width);
  h.is_signed = is_signed;
  h.reference = reference;
  h.squeeze_kind = static_cast<uint8_t>(squeeze_kind);
  lc::set_int_bucket_width(&h, bucket_width);
  lc::IntPredDesc p{op, lit_kind, lit_i, lit_u};
  int32_t kind = 0;
  uint64_t thr = 0;
  lc::plan_int_pred(&h, p, &kind, &thr);
  *thr_out = thr;
  if (tbits == 64 && bit_width > 32) {  // int_scan_entry: 64-bit packed domain only for wide fields of 64-bit columns
    const lc::URange<uint64_t> g = lc::make_range<uint64_t>(kind, thr);
    for (uint32_t i = 0; i < n; ++i) out[i] = (((packed[i] - g.lo) <= g.span) != g.neg) ? 1 : 0;
  } else {
    const lc::URange<uint32_t> g = lc::make_range<uint32_t>(kind, thr);
    for (uint32_t i = 0; i < n; ++i) out[i] = (((static_cast<uint32_t>(packed[i]) - g.lo) <= g.span) != g.neg) ? 1 : 0;
  }
  return kind;
}
}
