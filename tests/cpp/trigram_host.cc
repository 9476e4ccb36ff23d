// The private substring pre-filter of byte-view entries (liquid_cache_b200/csrc/entry_layout.h trigram_bit, kBloomWords) on the
// HOST: the 256-bit set of a byte string, built exactly as k_uniq_pass1 builds it per dictionary value and as
// prepare_str_pred builds it for a needle. Checked in tests/test_trigram_cpu.py.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/entry_layout.h"

extern "C" void tg_bloom(const uint8_t* p, uint32_t len, uint64_t out[4]) {
  std::memset(out, 0, 32);
  for (uint32_t b = 0; b + 2u < len; ++b) {
    const uint32_t t = lc::trigram_bit(p[b], p[b + 1u], p[b + 2u]);
    out[t >> 6] |= 1ull << (t & 63u);
  }
}
