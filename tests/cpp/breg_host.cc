// The register-resident integer scan's geometry (liquid_cache_b200/csrc/breg_math.cuh) compiled for the HOST: for one
// FastLanes chunk it computes, with the kernel's own code, the packed value every (step, lane) pair holds and the mask word
// each step's ballot is. tests/test_breg_cpu.py compares that with a plain FastLanes unpack for every (T, W).
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/breg_math.cuh"

namespace {
struct HostLoader {
  const uint8_t* p;
  uint32_t ld8(uint32_t o) const { return p[o]; }
  uint32_t ld16(uint32_t o) const {
    uint16_t v;
    std::memcpy(&v, p + o, 2);
    return v;
  }
  uint32_t ld32(uint32_t o) const {
    uint32_t v;
    std::memcpy(&v, p + o, 4);
    return v;
  }
  void ld64(uint32_t o, uint32_t* lo, uint32_t* hi) const {
    std::memcpy(lo, p + o, 4);
    std::memcpy(hi, p + o + 4, 4);
  }
};
template <uint32_t T, uint32_t W>
void run(const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  using G = lc::BregGeom<T, W>;
  for (uint32_t lane = 0; lane < 32; ++lane) {
    uint32_t a[G::SUB][G::NW];
    lc::breg_load<T, W>(lane, a, HostLoader{chunk});
    for (uint32_t s = 0; s < 32; ++s) values[s * 32 + lane] = lc::breg_value<T, W>(a, s);
  }
  for (uint32_t s = 0; s < 32; ++s) out_word[s] = lc::breg_out_word<T>(s);
}
template <uint32_t T>
int dispatch(uint32_t W, const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  switch (W) {
#define LC_W(k) case k: if constexpr (k <= T) { run<T, k>(chunk, values, out_word); return 0; } else return 1;
    LC_W(1) LC_W(2) LC_W(3) LC_W(4) LC_W(5) LC_W(6) LC_W(7) LC_W(8) LC_W(9) LC_W(10) LC_W(11) LC_W(12) LC_W(13) LC_W(14) LC_W(15) LC_W(16)
    LC_W(17) LC_W(18) LC_W(19) LC_W(20) LC_W(21) LC_W(22) LC_W(23) LC_W(24) LC_W(25) LC_W(26) LC_W(27) LC_W(28) LC_W(29) LC_W(30) LC_W(31) LC_W(32)
#undef LC_W
  }
  return 1;
}
}  // namespace

// chunk: 128 * W bytes (one FastLanes block of 1024 values of a T-bit column); values[32 steps][32 lanes]; out_word[32]
extern "C" int br_chunk(uint32_t T, uint32_t W, const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  switch (T) {
    case 8: return dispatch<8>(W, chunk, values, out_word);
    case 16: return dispatch<16>(W, chunk, values, out_word);
    case 32: return dispatch<32>(W, chunk, values, out_word);
    case 64: return dispatch<64>(W, chunk, values, out_word);
  }
  return 2;
}
