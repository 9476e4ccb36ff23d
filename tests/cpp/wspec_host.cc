// The width-specialised integer scan's value extraction (liquid_cache_b200/csrc/wspec_math.cuh) compiled for the HOST: for one
// FastLanes chunk it computes, with the kernel's own code, the packed value every (step, lane) pair reads and the output word
// each step's ballot lands in. tests/test_wspec_cpu.py compares that with a plain FastLanes unpack for every (T, W).
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/wspec_math.cuh"

namespace {
template <uint32_t T, uint32_t W>
void run(const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  auto rd = [chunk](uint32_t a) {
    uint32_t v;
    std::memcpy(&v, chunk + a, 4);
    return v;
  };
  for (uint32_t lane = 0; lane < 32; ++lane) {
    const lc::WspecBases<T, W> bs = lc::wspec_bases<T, W>(0u, lane);
    for (uint32_t j = 0; j < 32; ++j) values[j * 32 + lane] = lc::wspec_value<T, W>(bs, j, rd);
  }
  for (uint32_t j = 0; j < 32; ++j) out_word[j] = lc::wspec_out_word(j);
}
template <uint32_t T>
int dispatch(uint32_t W, const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  switch (W) {
#define LC_W(k) case k: run<T, k>(chunk, values, out_word); return 0;
    LC_W(1) LC_W(2) LC_W(3) LC_W(4) LC_W(5) LC_W(6) LC_W(7) LC_W(8) LC_W(9) LC_W(10) LC_W(11) LC_W(12) LC_W(13) LC_W(14) LC_W(15) LC_W(16)
    LC_W(17) LC_W(18) LC_W(19) LC_W(20) LC_W(21) LC_W(22) LC_W(23) LC_W(24) LC_W(25) LC_W(26) LC_W(27) LC_W(28) LC_W(29) LC_W(30) LC_W(31) LC_W(32)
#undef LC_W
  }
  return 1;
}
}  // namespace

// chunk: 128 * W bytes (one FastLanes block of 1024 values of a T-bit column); values[32 steps][32 lanes]; out_word[32]
extern "C" int ws_chunk(uint32_t T, uint32_t W, const uint8_t* chunk, uint32_t* values, uint32_t* out_word) {
  return T == 64 ? dispatch<64>(W, chunk, values, out_word) : T == 32 ? dispatch<32>(W, chunk, values, out_word) : 2;
}
