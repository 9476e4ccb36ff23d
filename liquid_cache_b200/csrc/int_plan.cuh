// int_plan.cuh — how `col <op> literal` becomes a test on the unsigned PACKED value of an integer entry: the planner and the
// range form every scan loop of k_int.cu evaluates. Host + device, so that the same code is exercised on the CPU against
// plain comparisons and the squeezed-array oracle (tests/cpp/int_plan_host.cc, tests/test_int_plan_cpu.py).
#pragma once
#include <cstdint>

#include "entry_layout.h"
#include "kernels.h"

#ifdef __CUDACC__
#define LC_PL_HD __host__ __device__ __forceinline__
#else
#define LC_PL_HD inline
#endif

namespace lc {

// Every comparison of the unsigned packed value against the threshold is one range test:
//   cmp(u) = ((u - lo) <= span) != neg
template <typename C>
struct URange {
  C lo, span;
  bool neg;
};

template <typename C>
LC_PL_HD URange<C> make_range(int32_t kind, uint64_t thr64) {
  const C mx = static_cast<C>(~static_cast<C>(0));
  const C thr = static_cast<C>(thr64);
  URange<C> g;
  g.lo = 0;
  g.span = mx;
  g.neg = false;
  switch (kind) {
    case UC_FALSE: g.neg = true; break;
    case UC_TRUE: break;
    case UC_EQ: g.lo = thr; g.span = 0; break;
    case UC_NE: g.lo = thr; g.span = 0; g.neg = true; break;
    case UC_LT: if (thr == 0) g.neg = true; else g.span = static_cast<C>(thr - 1); break;
    case UC_LE: g.span = thr; break;
    case UC_GT: if (thr == mx) g.neg = true; else { g.lo = static_cast<C>(thr + 1); g.span = static_cast<C>(mx - g.lo); } break;
    default: g.lo = thr; g.span = static_cast<C>(mx - thr); break;
  }
  return g;
}

// (op, literal) -> compare in the unsigned packed domain u = v - reference. All valid values satisfy
// reference <= v <= reference + (2^W - 1) in the column's own ordering, so a literal outside that window
// folds to a constant and one inside becomes an unsigned threshold. No 128-bit arithmetic needed:
// once lit >= reference is known, (lit - reference) fits in 64 unsigned bits.
LC_PL_HD void plan_int_pred(const IntHeader* h, const IntPredDesc& p, int32_t* ucmp, uint64_t* thr) {
  *thr = 0;
  if (h->bit_width == 0) {  // all null: values never matter
    *ucmp = UC_FALSE;
    return;
  }
  const uint32_t W = h->bit_width;
  const uint64_t umax = W == 64 ? ~0ull : ((1ull << W) - 1ull);
  if (p.lit_kind == kLitSentinel) {  // which rows of a clamped entry sit at the sentinel (squeeze_host.cc)
    *thr = umax;
    *ucmp = UC_EQ;
    return;
  }
  bool below, above = false;
  uint64_t d = 0;
  if (p.lit_kind == kLitAboveAll) {  // decimal literal beyond u64::MAX (scan_host.cc make_int_pred)
    below = false;
    above = true;
  } else if (h->is_signed) {
    const int sh = 64 - h->tbits;
    const long long ref = static_cast<long long>(h->reference << sh) >> sh;
    if (p.lit_kind == 1 /*U64*/ && p.lit_u > 0x7fffffffffffffffull) {
      below = false;
      above = true;
    } else {
      const long long lit = p.lit_kind == 1 ? static_cast<long long>(p.lit_u) : p.lit_i;
      below = lit < ref;
      if (!below) {
        d = static_cast<uint64_t>(lit) - static_cast<uint64_t>(ref);
        above = d > umax;
      }
    }
  } else {
    const uint64_t ref = h->reference;
    if (p.lit_kind == 0 /*I64*/ && p.lit_i < 0) {
      below = true;
    } else {
      const uint64_t lit = p.lit_kind == 0 ? static_cast<uint64_t>(p.lit_i) : p.lit_u;
      below = lit < ref;
      if (!below) {
        d = lit - ref;
        above = d > umax;
      }
    }
  }
  if (h->squeeze_kind == 2 && !below && p.lit_kind != kLitAboveAll) {
    // quantized entry (hybrid_primitive_array.rs:564-650): the words are bucket indices b = offset / bucket_width; compare
    // them with the literal's bucket q. b < q / b > q are the operator's two sides, exactly what `b <op> q` gives; inside
    // bucket q the same expression is right whenever the host let the call through (the literal sits on the bucket edge
    // that decides the operator, or no selected row is in bucket q — squeeze_host.cc checks that first with `= literal`,
    // which lands here as b == q).
    d = d / int_bucket_width(*h);  // d was the literal's offset from the reference
    above = d > umax;              // (the test above compared that offset with the code range: redo it for the bucket)
  }
  const int op = p.op;
  if (below) {
    *ucmp = (op == 1 || op == 4 || op == 5) ? UC_TRUE : UC_FALSE;  // NE, GT, GE
  } else if (above) {
    *ucmp = (op == 1 || op == 2 || op == 3) ? UC_TRUE : UC_FALSE;  // NE, LT, LE
  } else {
    *thr = d;
    *ucmp = op == 0 ? UC_EQ : op == 1 ? UC_NE : op == 2 ? UC_LT : op == 3 ? UC_LE : op == 4 ? UC_GT : UC_GE;
  }
}

}  // namespace lc
