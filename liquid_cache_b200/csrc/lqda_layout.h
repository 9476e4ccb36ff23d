// lqda_layout.h — where the sections of a byte-view LQDA image sit (LiquidByteViewArray::to_bytes / from_bytes,
// /root/reference/src/core/src/liquid_array/byte_view_array/serialization.rs:87-325), as plain C++ without any CUDA, so that
// the writer's layout and the reader's parse (ipc_host.cc) are exercised on the CPU against images the oracle writes
// (tests/cpp/lqda_layout_host.cc, tests/test_lqda_layout_cpu.py).
#pragma once
#include <cstdint>
#include <cstring>

#include "entry_layout.h"

namespace lc {

namespace lqda {
inline uint16_t get_u16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
inline uint32_t get_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t get_u64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint64_t pad8(uint64_t x) { return (x + 7) & ~7ull; }
}  // namespace lqda

// ---- writer: byte offsets inside the image of the entry described by `h` ----
struct StrImage {  // byte offsets inside the LQDA image of a byte-view entry
  uint64_t fsst_off, keys_off, keys_nulls_off, keys_values_off, co_off, pk_off, sp_off, fp_off, total;
  uint32_t fsst_raw_size, keys_size, nulls_len, keys_values_len, co_size, sp_size, fp_size;
};
inline StrImage str_image_of(const StrHeader& h) {
  StrImage L{};
  uint64_t cur = lqda::pad8(36);
  L.fsst_off = cur;
  L.fsst_raw_size = 12 + h.fsst_bytes;
  cur = lqda::pad8(cur + L.fsst_raw_size);
  L.keys_off = cur;
  L.nulls_len = h.has_nulls ? (h.n + 7) / 8 : 0;
  L.keys_values_len = ((h.n + 1023) / 1024) * 2048;
  L.keys_nulls_off = cur + 16;
  L.keys_values_off = cur + lqda::pad8(16ull + L.nulls_len);
  L.keys_size = static_cast<uint32_t>(L.keys_values_off + L.keys_values_len - L.keys_off);
  cur = lqda::pad8(L.keys_off + L.keys_size);
  L.co_off = cur;
  L.co_size = 9 + (h.n_unique + 1) * h.offset_bytes;
  cur = lqda::pad8(cur + L.co_size);
  L.pk_off = cur;
  cur = lqda::pad8(cur + 8ull * h.n_unique);
  L.sp_off = cur;
  L.sp_size = h.shared_prefix_len;
  cur = lqda::pad8(cur + L.sp_size);
  L.fp_off = cur;
  L.fp_size = h.has_fp ? 4 * h.n_unique : 0;
  L.total = cur + L.fp_size;
  return L;
}


// ---- reader: the sections of an image, every bound and the dictionary offsets checked; nullptr = fine, else the reason ----
struct StrImageIn {
  uint32_t bt, n, n_unique, offset_bytes, n_resid, sp_size, fp_size, comp_bytes, nulls_len, kvals_len;
  int32_t slope, intercept;
  bool file_nulls;
  uint64_t uncompressed, comp_off, knulls_off, kvals_off, resid_src, pk_src, sp_src, fp_src;
};
inline const char* parse_str_image(const uint8_t* b, uint64_t len, StrImageIn* out) {
  using namespace lqda;
  const uint32_t bt = get_u16(b + 8);
  if (len < 36 || bt > BT_BINARY_VIEW) {
    return "bad byte-view header";
  }
  const uint32_t keys_size = get_u32(b + 16), co_size = get_u32(b + 20), sp_size = get_u32(b + 24), fsst_size = get_u32(b + 28),
                 fp_size = get_u32(b + 32);
  uint64_t cur = pad8(36);
  if (fsst_size < 12 || len < cur + fsst_size) {
    return "FSST section runs past the image";
  }
  const uint64_t uncompressed = get_u64(b + cur);
  const uint32_t comp_bytes = get_u32(b + cur + 8);
  const uint64_t comp_off = cur + 12;
  if (12ull + comp_bytes > fsst_size) {
    return "FSST values longer than their section";
  }
  cur = pad8(cur + fsst_size);
  if (keys_size < 16 || len < cur + keys_size) {
    return "keys section runs past the image";
  }
  const uint8_t* kb = b + cur;
  const uint32_t n = get_u32(kb);
  const bool file_nulls = kb[5] != 0;
  const uint32_t nulls_len = get_u32(kb + 6), kvals_len = get_u32(kb + 10);
  const uint64_t knulls_off = cur + 16, kvals_off = cur + pad8(16ull + (file_nulls ? nulls_len : 0));
  const uint32_t n_chunks = (n + 1023) / 1024;
  if (n > 0x7fffffffu || (n && kb[4] != 16) || kvals_len != n_chunks * 2048u || kvals_off + kvals_len > cur + keys_size ||
      (file_nulls && nulls_len < (n + 7) / 8)) {
    return "keys are not bit-packed at width 16 (or their section does not hold every row)";
  }
  cur = pad8(cur + keys_size);
  if (len < cur + co_size || (co_size && co_size < 9)) {
    return "offsets section runs past the image";
  }
  int32_t slope = 0, intercept = 0;
  uint32_t ob = 1, n_resid = 0;
  const uint64_t resid_src = cur + 9;
  if (co_size) {
    std::memcpy(&slope, b + cur, 4);
    std::memcpy(&intercept, b + cur + 4, 4);
    ob = b[cur + 8];
    if ((ob != 1 && ob != 2 && ob != 4) || (co_size - 9) % ob) {
      return "bad CompactOffsets header";
    }
    n_resid = (co_size - 9) / ob;
  }
  const uint32_t U = n_resid ? n_resid - 1 : 0;
  cur = pad8(cur + co_size);
  const uint64_t pk_src = cur;
  cur = pad8(cur + 8ull * U);
  const uint64_t sp_src = cur;
  cur = pad8(cur + sp_size);
  const uint64_t fp_src = cur;
  if (U > 65536 || len < fp_src + fp_size || (fp_size && fp_size != 4 * U)) {
    return "dictionary sections run past the image";
  }
  // the offsets the decode kernels will follow: slope * i + intercept + residual[i] (CompactOffsets::get_offset,
  // fsst_buffer.rs:360-383) must start at 0, never step back, and stay inside the compressed values
  if (n_resid) {
    uint64_t prev = 0;
    for (uint32_t i = 0; i < n_resid; ++i) {
      int64_t r = 0;
      const uint8_t* rp = b + resid_src + static_cast<uint64_t>(i) * ob;
      if (ob == 1) r = static_cast<int8_t>(rp[0]);
      else if (ob == 2) r = static_cast<int16_t>(get_u16(rp));
      else r = static_cast<int32_t>(get_u32(rp));
      // the kernels' arithmetic (k_str.cu dict_offset): 32-bit wrapping sum, read as unsigned
      const uint64_t off = static_cast<uint32_t>(static_cast<uint32_t>(slope) * i + static_cast<uint32_t>(intercept) + static_cast<uint32_t>(r));
      if ((i == 0 && off != 0) || off < prev || off > comp_bytes) {
        return "a dictionary offset does not fit the compressed values";
      }
      prev = off;
    }
  }
  out->bt = bt;
  out->n = n;
  out->n_unique = U;
  out->offset_bytes = ob;
  out->n_resid = n_resid;
  out->sp_size = sp_size;
  out->fp_size = fp_size;
  out->comp_bytes = comp_bytes;
  out->nulls_len = nulls_len;
  out->kvals_len = kvals_len;
  out->slope = slope;
  out->intercept = intercept;
  out->file_nulls = file_nulls;
  out->uncompressed = uncompressed;
  out->comp_off = comp_off;
  out->knulls_off = knulls_off;
  out->kvals_off = kvals_off;
  out->resid_src = resid_src;
  out->pk_src = pk_src;
  out->sp_src = sp_src;
  out->fp_src = fp_src;
  return nullptr;
}

}  // namespace lc
