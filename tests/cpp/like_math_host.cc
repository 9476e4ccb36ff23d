// LIKE on FSST codes as Shift-And steps (liquid_cache_b200/csrc/like_math.cuh) on the HOST, exactly the arithmetic the
// kernels run: the step table k_like_steps builds per symbol table, the one-code-at-a-time walk of like_trip, and the
// 32-codes-at-once walk of like_candidates_warp (lanes classified into code / escape marker / literal as warp_decode
// does, their steps combined in order by the same shuffle tree, one application to the running state per block).
// Checked in tests/test_like_math_cpu.py against a plain substring search of the decoded bytes.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/like_math.cuh"

using lc::SymStep;

extern "C" {

// symbols[256] little-endian packed, lens[256] (lens[255] unused: the escape marker); out[512]
void lm_table(const uint64_t* symbols, const uint8_t* lens, const uint8_t* needle, uint32_t m, SymStep* out) {
  uint32_t M[256];
  for (uint32_t b = 0; b < 256; ++b) {
    uint32_t bits = 0;
    for (uint32_t j = 0; j < m; ++j) bits |= (needle[j] == b ? 1u : 0u) << j;
    M[b] = bits;
  }
  for (uint32_t c = 0; c < 512; ++c) {
    const uint64_t sym = c < 256 ? symbols[c] : static_cast<uint64_t>(c - 256);
    const uint32_t L = c < 255 ? lens[c] : (c == 255 ? 0u : 1u);
    out[c] = lc::like_sym_step(sym, L, M, m);
  }
}

// one code after the other (like_trip's general path)
int lm_walk_seq(const SymStep* steps, const uint8_t* codes, uint32_t n) {
  uint32_t S = lc::kStateOne, pending = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t b = codes[i];
    if (lc::like_apply(S, steps[b + (pending << 8)])) return 1;
    pending = (pending == 0 && b == 255) ? 1u : 0u;
  }
  return 0;
}

// 32 codes per block, combined by the shuffle tree of like_candidates_warp
int lm_walk_blocks(const SymStep* steps, const uint8_t* codes, uint32_t n) {
  uint32_t S = lc::kStateOne, carry_lit = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    SymStep st[32];
    uint32_t esc_last = 0;
    // classification, as the ballot arithmetic of the kernel does it
    uint32_t F = 0;
    for (uint32_t lane = 0; lane < 32; ++lane)
      if (base + lane < n && codes[base + lane] == 255) F |= 1u << lane;
    const uint32_t Fp = carry_lit ? (F & ~1u) : F;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      const bool in = base + lane < n;
      const uint32_t b = in ? codes[base + lane] : 0u;
      const uint32_t zeros = ~Fp & ((1u << lane) - 1u);
      const uint32_t run = zeros ? (lane - 1u - (31u - static_cast<uint32_t>(__builtin_clz(zeros)))) : lane;
      const bool lit = (lane == 0) ? (carry_lit != 0) : ((run & 1u) != 0);
      const bool esc = in && b == 255 && !lit;
      st[lane] = (in && !esc) ? steps[b + (lit ? 256u : 0u)] : lc::step_identity();
      if (lane == 31) esc_last = esc ? 1u : 0u;
    }
    for (uint32_t d = 1; d < 32; d <<= 1) {  // lane i takes lane i + d's step behind its own, for i % 2d == 0
      SymStep nxt[32];
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const SymStep o = st[lane + d < 32 ? lane + d : lane];  // __shfl_down_sync past the end returns the lane's own
        nxt[lane] = (lane & (2 * d - 1)) == 0 ? lc::step_then(st[lane], o) : st[lane];
      }
      std::memcpy(st, nxt, sizeof(st));
    }
    if (lc::like_apply(S, st[0])) return 1;
    carry_lit = esc_last;
  }
  return 0;
}

}  // extern "C"
