// The decision squeeze_host.cc takes for `col <op> literal` on a squeezed integer entry (liquid_cache_b200/csrc/squeeze_plan.h),
// compiled for the HOST. tests/test_squeeze_plan_cpu.py checks it against the restated reference arrays. No CUDA here.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/squeeze_plan.h"

extern "C" {

// returns 0: the codes decide; 1: they decide unless the sentinel probe finds a row; 2: unless `= k` finds a row;
//         3: the literal does not fit the column's type (the backing is read)
int sp_doubt(uint32_t tbits, uint32_t bit_width, uint32_t is_signed, uint64_t reference, int32_t squeeze_kind, uint64_t bucket_width, int32_t op,
             int32_t lit_kind, int64_t lit_i, uint64_t lit_u) {
  lc::SqueezeFacts f;
  std::memset(&f, 0, sizeof(f));
  f.ih.tbits = static_cast<uint8_t>(tbits);
  f.ih.bit_width = static_cast<uint8_t>(bit_width);
  f.ih.is_signed = is_signed;
  f.ih.reference = reference;
  f.squeeze_kind = squeeze_kind;
  f.bucket_width = bucket_width;
  lc_predicate p;
  std::memset(&p, 0, sizeof(p));
  p.op = op;
  p.lit_kind = lit_kind;
  p.lit_i64 = lit_i;
  p.lit_u64 = lit_u;
  __int128 k = 0;
  if (!lc::literal_of(&f, &p, &k)) return 3;
  const lc::Doubt d = lc::doubt_of(&f, op, k);
  if (!d.possible) return 0;
  return d.probe.lit_kind == lc::kLitSentinelPublic ? 1 : 2;
}
}
