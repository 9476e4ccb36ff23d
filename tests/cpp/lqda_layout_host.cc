// The byte-view LQDA layout code of the library (liquid_cache_b200/csrc/lqda_layout.h: the writer's section offsets and the
// reader's parse with its checks) compiled for the HOST, for tests/test_lqda_layout_cpu.py. No CUDA in this file.
#include <cstdint>
#include <cstring>

#include "liquid_cache_b200/csrc/lqda_layout.h"

extern "C" {

// out[21]: bt n U ob n_resid sp_size fp_size comp_bytes nulls_len kvals_len slope intercept file_nulls uncompressed comp_off
//          knulls_off kvals_off resid_src pk_src sp_src fp_src; returns 0, or 1 with the reason in *why
int lq_parse(const uint8_t* b, uint64_t len, int64_t* out, const char** why) {
  lc::StrImageIn in;
  std::memset(&in, 0, sizeof(in));
  const char* e = lc::parse_str_image(b, len, &in);
  if (e) {
    *why = e;
    return 1;
  }
  const int64_t v[21] = {in.bt, in.n, in.n_unique, in.offset_bytes, in.n_resid, in.sp_size, in.fp_size, in.comp_bytes, in.nulls_len, in.kvals_len,
                         in.slope, in.intercept, in.file_nulls, static_cast<int64_t>(in.uncompressed), static_cast<int64_t>(in.comp_off),
                         static_cast<int64_t>(in.knulls_off), static_cast<int64_t>(in.kvals_off), static_cast<int64_t>(in.resid_src),
                         static_cast<int64_t>(in.pk_src), static_cast<int64_t>(in.sp_src), static_cast<int64_t>(in.fp_src)};
  std::memcpy(out, v, sizeof(v));
  return 0;
}

// the writer's layout for an entry with these header facts; out[16]: fsst_off keys_off keys_nulls_off keys_values_off co_off pk_off
// sp_off fp_off total fsst_raw_size keys_size nulls_len keys_values_len co_size sp_size fp_size
void lq_layout(uint32_t n, uint32_t n_unique, uint32_t has_nulls, uint32_t has_fp, uint32_t offset_bytes, uint32_t fsst_bytes,
               uint32_t shared_prefix_len, int64_t* out) {
  lc::StrHeader h;
  std::memset(&h, 0, sizeof(h));
  h.n = n;
  h.n_unique = n_unique;
  h.has_nulls = static_cast<uint8_t>(has_nulls);
  h.has_fp = static_cast<uint8_t>(has_fp);
  h.offset_bytes = static_cast<uint8_t>(offset_bytes);
  h.fsst_bytes = fsst_bytes;
  h.shared_prefix_len = shared_prefix_len;
  const lc::StrImage L = lc::str_image_of(h);
  const int64_t v[16] = {static_cast<int64_t>(L.fsst_off), static_cast<int64_t>(L.keys_off), static_cast<int64_t>(L.keys_nulls_off),
                         static_cast<int64_t>(L.keys_values_off), static_cast<int64_t>(L.co_off), static_cast<int64_t>(L.pk_off),
                         static_cast<int64_t>(L.sp_off), static_cast<int64_t>(L.fp_off), static_cast<int64_t>(L.total), L.fsst_raw_size, L.keys_size,
                         L.nulls_len, L.keys_values_len, L.co_size, L.sp_size, L.fp_size};
  std::memcpy(out, v, sizeof(v));
}
}
