// squeeze_plan.h — what the half-width codes of a squeezed integer entry can say about `col <op> literal`, as plain C++
// without CUDA: the decision squeeze_host.cc takes before it launches anything (run on the codes / probe first / read the
// backing). Exercised on the CPU against the restated reference arrays (tests/cpp/squeeze_plan_host.cc,
// tests/test_squeeze_plan_cpu.py) together with the kernel's planner (int_plan.cuh).
// Reference: hybrid_primitive_array.rs:160-288 (clamp), :487-667 (quantize) under /root/reference/src/core/src/liquid_array.
#pragma once
#include <cstdint>

#include "../../include/lc_gpu.h"
#include "entry_layout.h"

namespace lc {

constexpr int32_t kLitSentinelPublic = 1000;  // lc_predicate.lit_kind used by squeeze_host.cc alone (-> kLitSentinel)

struct SqueezeFacts {  // of one squeezed entry
  IntHeader ih;          // tbits, is_signed, reference, bit_width (of the codes)
  int32_t squeeze_kind;  // 1 clamp, 2 quantize
  uint64_t bucket_width; // quantize
};

inline __int128 reference_of(const SqueezeFacts* e) {
  const uint32_t tbits = e->ih.tbits;
  if (!e->ih.is_signed) return static_cast<__int128>(e->ih.reference);
  const uint64_t sign = 1ull << (tbits - 1);
  const uint64_t raw = e->ih.reference;  // zero-extended raw bits
  return static_cast<__int128>(static_cast<int64_t>((raw ^ sign) - sign));
}

// T::Native::from_i64 / from_u64 (hybrid_primitive_array.rs:167-184): the literal as a value of the column's type
inline bool literal_of(const SqueezeFacts* e, const lc_predicate* pred, __int128* k) {
  __int128 v;
  if (pred->lit_kind == LC_LIT_I64) v = pred->lit_i64;
  else if (pred->lit_kind == LC_LIT_U64) v = static_cast<__int128>(pred->lit_u64);
  else return false;
  const uint32_t tbits = e->ih.tbits;
  const __int128 one = 1;
  const __int128 lo = e->ih.is_signed ? -(one << (tbits - 1)) : 0;
  const __int128 hi = e->ih.is_signed ? (one << (tbits - 1)) - 1 : (one << tbits) - 1;
  if (v < lo || v > hi) return false;
  *k = v;
  return true;
}

inline lc_predicate int_predicate(const SqueezeFacts* e, int32_t op, __int128 lit) {
  lc_predicate p{};
  p.op = op;
  if (e->ih.is_signed) {
    p.lit_kind = LC_LIT_I64;
    p.lit_i64 = static_cast<int64_t>(lit);
  } else {
    p.lit_kind = LC_LIT_U64;
    p.lit_u64 = static_cast<uint64_t>(lit);
  }
  return p;
}

struct Doubt {
  bool possible = false;
  lc_predicate probe{};
};

inline Doubt doubt_of(const SqueezeFacts* sq, int32_t op, __int128 k) {
  Doubt d;
  const __int128 ref = reference_of(sq);
  const uint64_t last = (1ull << sq->ih.bit_width) - 1ull;  // the sentinel / the last bucket
  if (sq->squeeze_kind == LC_SQUEEZE_CLAMP + 1) {
    const __int128 sent_abs = ref + static_cast<__int128>(last);
    const bool strict = op == LC_OP_EQ || op == LC_OP_NE || op == LC_OP_GT || op == LC_OP_LE;
    d.possible = !(strict ? k < sent_abs : k <= sent_abs);
    d.probe.op = LC_OP_EQ;
    d.probe.lit_kind = kLitSentinelPublic;
    return d;
  }
  if (k < ref) return d;  // below the minimum: constants (:537-560)
  const unsigned __int128 rel = static_cast<unsigned __int128>(k - ref);
  const uint64_t bw = sq->bucket_width;
  if (rel / bw > last) return d;  // every bucket index is below the literal's
  const uint64_t r = static_cast<uint64_t>(rel % bw);
  bool known = false;
  switch (op) {
    case LC_OP_LT: case LC_OP_GE: known = r == 0; break;
    case LC_OP_LE: case LC_OP_GT: known = r + 1 == bw; break;
    default: break;
  }
  d.possible = !known;
  d.probe = int_predicate(sq, LC_OP_EQ, k);
  return d;
}


}  // namespace lc
