// like_math.cuh — LIKE '%needle%' on FSST codes as Shift-And steps, host + device (the kernels in k_str.cu use these
// functions; tests/test_like_math_cpu.py runs them on the CPU against a plain substring search).
//
// Shift-And over the needle (m <= 31): state bit j <=> needle[0..j] matches the text ending here; one text byte b maps
// S -> ((S << 1) | 1) & M[b]. That map is linear over OR, so the effect of a whole FSST symbol (1..8 bytes) collapses into
//     S' = ((S << L) & A) | B        and   "the needle completed inside this symbol"  <=>  (S & H) != 0
// Bit 31 of the state is a constant 1 (needles are <= 31 bytes on this path, B always re-sets it), so "completed at the
// symbol's first bytes regardless of the state" is just bit 31 of H and the hit test is one AND.
// Two steps in a row are again one step of the same form (step_then), which is what lets a warp take 32 codes of a value
// at once and combine them with a shuffle tree instead of one lane walking them one after the other.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LC_LM_HD __host__ __device__ __forceinline__
#else
#define LC_LM_HD inline
#endif

namespace lc {

struct SymStep {
  uint32_t A, B, H;  // B has bit 31 set; H bit 31 = "hits whatever the state"
  uint32_t L;        // symbol length = shift
};
constexpr uint32_t kStateOne = 0x80000000u;

// The step of one symbol: `sym` holds its L bytes little-endian, M[b] has bit j set iff needle[j] == b, m = needle length.
// L == 0 (the escape marker's table entry) is the identity.
LC_LM_HD SymStep like_sym_step(uint64_t sym, uint32_t L, const uint32_t* M, uint32_t m) {
  const uint32_t acc = 1u << (m - 1u);
  uint32_t A = 0xffffffffu, B = 0, H = 0, hit0 = 0;
  for (uint32_t k = 0; k < L; ++k) {
    const uint32_t Mb = M[static_cast<uint32_t>(sym & 0xffu)];
    sym >>= 8;
    A = (A << 1) & Mb;
    B = ((B << 1) | 1u) & Mb;
    H |= (A & acc) >> (k + 1u);
    hit0 |= (B & acc) ? 1u : 0u;
  }
  SymStep st;
  st.A = A;
  st.B = B | kStateOne;
  st.H = H | (hit0 << 31);
  st.L = L;
  return st;
}

// One step applied to the running state; returns whether the needle completed inside it.
LC_LM_HD bool like_apply(uint32_t& S, const SymStep& e) {
  const bool hit = (S & e.H) != 0u;
  S = ((e.L >= 32u ? 0u : S << e.L) & e.A) | e.B | kStateOne;
  return hit;
}

LC_LM_HD uint32_t shl_sat(uint32_t x, uint32_t n) { return n >= 32u ? 0u : x << n; }
LC_LM_HD uint32_t shr_sat(uint32_t x, uint32_t n) { return n >= 32u ? 0u : x >> n; }

// f first, then g, as ONE step:
//   S2 = ((((S << L1) & A1) | B1) << L2) & A2 | B2  =  ((S << (L1 + L2)) & ((A1 << L2) & A2)) | (((B1 << L2) & A2) | B2)
//   hit = (S & H1) | (S1 & H2),  S1 & H2 = ((S << L1) & A1 & H2) | (B1 & H2):  the first term is S & ((A1 & H2) >> L1), the
//   second does not depend on S: bit 31 (the state's constant one) stands for it.
LC_LM_HD SymStep step_then(const SymStep& f, const SymStep& g) {
  SymStep r;
  r.L = f.L + g.L > 32u ? 32u : f.L + g.L;
  r.A = shl_sat(f.A, g.L) & g.A;
  r.B = (shl_sat(f.B, g.L) & g.A) | g.B;
  r.H = f.H | shr_sat(f.A & g.H, f.L) | ((f.B & g.H) ? kStateOne : 0u);
  return r;
}
LC_LM_HD SymStep step_identity() { return SymStep{0xffffffffu, 0u, 0u, 0u}; }

}  // namespace lc
