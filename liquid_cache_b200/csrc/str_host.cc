// str_host.cc — byte-view insert side and predicate planning.
// Reference: LiquidByteViewArray::from_dict_array_inner
//   (/root/reference/src/core/src/liquid_array/byte_view_array/conversions.rs:260-373),
//   CheckedDictionaryArray (src/core/src/utils/mod.rs:52-154), PrefixKey / CompactOffsets / fit_line
//   (raw/fsst_buffer.rs:160-187, 267-383), StringFingerprint (byte_view_array/fingerprint.rs:19-26),
//   with_fsst_compressor_or_train (src/core/src/cache/transcode.rs:16-33).
#include <cmath>

#include "host_common.h"

namespace lc {

namespace {

struct RowRef {
  const uint8_t* p;
  uint32_t len;
  bool valid;
};

inline uint64_t hash_bytes(const uint8_t* p, uint32_t len) {
  // 64-bit multiply-xorshift over 8-byte words (dictionary build only; not observable)
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (static_cast<uint64_t>(len) * 0xff51afd7ed558ccdull);
  while (len >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
    p += 8;
    len -= 8;
  }
  if (len) {
    uint64_t w = 0;
    std::memcpy(&w, p, len);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
  }
  return h ^ (h >> 32);
}

// u16 dictionary in first-occurrence order (GenericByteDictionaryBuilder<UInt16Type, _>, utils/mod.rs:141-154)
struct DictBuilder {
  std::vector<uint32_t> slots;  // unique index + 1, 0 = empty
  std::vector<uint64_t> hashes;
  std::vector<const uint8_t*> uptr;
  std::vector<uint32_t> ulen;
  uint32_t mask;
  explicit DictBuilder(uint32_t expect) {
    uint32_t cap = 64;
    while (cap < expect * 2u) cap <<= 1;
    slots.assign(cap, 0);
    mask = cap - 1;
  }
  void grow() {
    std::vector<uint32_t> ns(slots.size() * 2, 0);
    const uint32_t nm = static_cast<uint32_t>(ns.size() - 1);
    for (uint32_t u = 0; u < uptr.size(); ++u) {
      uint32_t s = static_cast<uint32_t>(hashes[u]) & nm;
      while (ns[s]) s = (s + 1) & nm;
      ns[s] = u + 1;
    }
    slots.swap(ns);
    mask = nm;
  }
  uint32_t add(const uint8_t* p, uint32_t len) {
    const uint64_t h = hash_bytes(p, len);
    uint32_t s = static_cast<uint32_t>(h) & mask;
    while (slots[s]) {
      const uint32_t u = slots[s] - 1;
      if (hashes[u] == h && ulen[u] == len && std::memcmp(uptr[u], p, len) == 0) return u;
      s = (s + 1) & mask;
    }
    const uint32_t u = static_cast<uint32_t>(uptr.size());
    slots[s] = u + 1;
    hashes.push_back(h);
    uptr.push_back(p);
    ulen.push_back(len);
    if (uptr.size() * 2 > slots.size()) grow();
    return u;
  }
};

// fit_line (raw/fsst_buffer.rs:267-296): least squares in f64, rounded to i32
void fit_line(const std::vector<uint32_t>& offsets, int32_t* slope, int32_t* intercept) {
  const size_t n = offsets.size();
  if (n <= 1) {
    *slope = 0;
    *intercept = n ? static_cast<int32_t>(offsets[0]) : 0;
    return;
  }
  const double nf = static_cast<double>(n);
  const double sum_x = static_cast<double>(n * (n - 1) / 2);
  double sum_y = 0.0, sum_xy = 0.0;
  for (size_t i = 0; i < n; ++i) sum_y += static_cast<double>(offsets[i]);
  for (size_t i = 0; i < n; ++i) sum_xy += static_cast<double>(i) * static_cast<double>(offsets[i]);
  const double sum_x_sq = static_cast<double>(n * (n - 1) * (2 * n - 1) / 6);
  const double sl = (nf * sum_xy - sum_x * sum_y) / (nf * sum_x_sq - sum_x * sum_x);
  const double ic = (sum_y - sl * sum_x) / nf;
  auto sat = [](double v) -> int32_t {
    const double r = std::round(v);
    if (!(r == r)) return 0;
    if (r >= 2147483647.0) return 2147483647;
    if (r <= -2147483648.0) return -2147483647 - 1;
    return static_cast<int32_t>(r);
  };
  *slope = sat(sl);
  *intercept = sat(ic);
}

}  // namespace

static int get_codec(lc_ctx* ctx, uint64_t scope, const DictBuilder& d, std::shared_ptr<FsstCodec>* out) {
  auto it = ctx->codecs.find(scope);
  if (it != ctx->codecs.end()) {
    *out = it->second;
    return LC_OK;
  }
  // first batch of this column chunk trains (transcode.rs:16-33); training input = the unique values
  auto codec = std::make_shared<FsstCodec>();
  fsst_train(d.uptr.data(), d.ulen.data(), d.uptr.size(), codec.get());
  if (cudaMalloc(reinterpret_cast<void**>(&codec->d_dec), sizeof(FsstTable)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&codec->d_enc), sizeof(FsstEncTable)) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc for FSST tables failed");
    return LC_ERR_OOM;
  }
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_dec, &codec->dec, sizeof(FsstTable), cudaMemcpyHostToDevice, ctx->stream));
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_enc, codec->enc.get(), sizeof(FsstEncTable), cudaMemcpyHostToDevice, ctx->stream));
  LC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->codecs[scope] = codec;
  *out = codec;
  return LC_OK;
}

int str_encode(lc_ctx* ctx, const ArrowIn& in, int32_t hint, uint64_t scope, Entry** out) {
  const uint32_t n = static_cast<uint32_t>(in.length);
  // ---- 1. rows ----
  std::vector<RowRef> rows(n);
  std::vector<uint8_t> inline_store;  // view arrays: inline payloads are addressed in place
  for (uint32_t i = 0; i < n; ++i) {
    const int64_t r = in.offset + i;
    RowRef rr{nullptr, 0, true};
    if (in.validity && in.null_count > 0 && !bit_get(in.validity, r)) rr.valid = false;
    if (rr.valid) {
      if (in.kind == ArrowIn::K_BYTES) {
        const int32_t* off = static_cast<const int32_t*>(in.values);
        rr.p = in.data + off[r];
        rr.len = static_cast<uint32_t>(off[r + 1] - off[r]);
      } else if (in.kind == ArrowIn::K_VIEW) {
        const uint8_t* v = static_cast<const uint8_t*>(in.values) + 16 * r;
        uint32_t len;
        std::memcpy(&len, v, 4);
        rr.len = len;
        if (len <= 12) {
          rr.p = v + 4;
        } else {
          uint32_t bi, bo;
          std::memcpy(&bi, v + 8, 4);
          std::memcpy(&bo, v + 12, 4);
          if (static_cast<int64_t>(bi) >= in.n_view_buffers) {
            set_error("view buffer index out of range");
            return LC_ERR_INVALID;
          }
          rr.p = static_cast<const uint8_t*>(in.view_buffers[bi]) + bo;
        }
      } else {  // K_DICT: a null dictionary value makes the row null (typed dictionary iterator)
        const uint32_t key = in.dict_keys[r];
        const int64_t dv = in.dict_offset + key;
        if (static_cast<int64_t>(key) >= in.dict_len) {
          set_error("dictionary key out of range");
          return LC_ERR_INVALID;
        }
        if (in.dict_validity && !bit_get(in.dict_validity, dv)) {
          rr.valid = false;
        } else {
          rr.p = in.dict_data + in.dict_offsets[dv];
          rr.len = static_cast<uint32_t>(in.dict_offsets[dv + 1] - in.dict_offsets[dv]);
        }
      }
    }
    rows[i] = rr;
  }

  // ---- 2. u16 dictionary, first-occurrence order ----
  DictBuilder dict(n < 1024 ? 1024 : n / 2);
  std::vector<uint16_t> keys(n, 0);
  uint32_t null_count = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (!rows[i].valid) {
      ++null_count;
      continue;
    }
    const uint32_t u = dict.add(rows[i].p ? rows[i].p : reinterpret_cast<const uint8_t*>(""), rows[i].len);
    if (u > 65535u) {
      // the reference's UInt16 dictionary builder overflows (panics) here; we decline the batch
      set_error("more than 65536 distinct values in one batch");
      return LC_ERR_UNSUPPORTED_TYPE;
    }
    keys[i] = static_cast<uint16_t>(u);
  }
  const uint32_t U = static_cast<uint32_t>(dict.uptr.size());

  // ---- 3. shared prefix = LCP of all unique values (conversions.rs:269-307) ----
  uint32_t spl = 0;
  if (U > 0) {
    spl = dict.ulen[0];
    for (uint32_t u = 1; u < U && spl > 0; ++u) {
      const uint32_t m = dict.ulen[u] < spl ? dict.ulen[u] : spl;
      uint32_t c = 0;
      while (c < m && dict.uptr[u][c] == dict.uptr[0][c]) ++c;
      spl = c;
    }
  }

  // ---- 4. symbol table: train on the first batch of the scope, reuse afterwards ----
  std::shared_ptr<FsstCodec> codec;
  LC_TRY(get_codec(ctx, scope, dict, &codec));

  // ---- 5. compress uniques, prefix keys, fingerprints ----
  const bool build_fp = (hint == LC_HINT_SUBSTRING_SEARCH);
  std::vector<uint32_t> offsets(U + 1, 0);
  std::vector<uint8_t> comp;
  uint64_t uncompressed = 0;
  uint32_t max_len = 0;
  {
    uint64_t total = 0;
    for (uint32_t u = 0; u < U; ++u) total += dict.ulen[u];
    comp.resize(2 * total + 16);
  }
  std::vector<uint64_t> pkeys(U);
  std::vector<uint32_t> fps(build_fp ? U : 0);
  uint64_t co = 0;
  for (uint32_t u = 0; u < U; ++u) {
    const uint8_t* p = dict.uptr[u];
    const uint32_t len = dict.ulen[u];
    uncompressed += len;
    if (len > max_len) max_len = len;
    co += fsst_compress_host(*codec, p, len, comp.data() + co);
    if (co > 0xFFFFFFF0ull) {
      set_error("compressed dictionary exceeds 4 GiB");
      return LC_ERR_UNSUPPORTED_TYPE;
    }
    offsets[u + 1] = static_cast<uint32_t>(co);
    // PrefixKey::new(suffix) (fsst_buffer.rs:173-187)
    const uint32_t sl = len > spl ? len - spl : 0;
    uint64_t k = 0;
    const uint32_t cp = sl < 7 ? sl : 7;
    for (uint32_t b = 0; b < cp; ++b) k |= static_cast<uint64_t>(p[spl + b]) << (8 * b);
    k |= static_cast<uint64_t>(sl >= 255 ? 255u : sl) << 56;
    pkeys[u] = k;
    if (build_fp) {
      uint32_t bits = 0;
      for (uint32_t b = 0; b < len; ++b) bits |= 1u << (p[b] & 31u);
      fps[u] = bits;
    }
  }

  // ---- 6. CompactOffsets (fsst_buffer.rs:298-359) ----
  int32_t slope = 0, intercept = 0;
  fit_line(offsets, &slope, &intercept);
  std::vector<int32_t> resid(U + 1);
  int32_t rmin = 2147483647, rmax = -2147483647 - 1;
  for (uint32_t i = 0; i <= U; ++i) {
    const uint32_t predicted = static_cast<uint32_t>(slope) * i + static_cast<uint32_t>(intercept);
    const int32_t r = static_cast<int32_t>(offsets[i] - predicted);
    resid[i] = r;
    if (r < rmin) rmin = r;
    if (r > rmax) rmax = r;
  }
  const uint32_t ob = (rmin >= -128 && rmax <= 127) ? 1u : (rmin >= -32768 && rmax <= 32767) ? 2u : 4u;

  // ---- 7. blob ----
  StrHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicStr;
  h.arrow_type = in.byte_type;
  h.has_nulls = null_count > 0;
  h.has_fp = build_fp;
  h.offset_bytes = static_cast<uint8_t>(ob);
  h.n = n;
  h.n_unique = U;
  h.slope = slope;
  h.intercept = intercept;
  h.shared_prefix_len = spl;
  h.null_count = null_count;
  h.max_value_len = max_len;
  h.uncompressed_bytes = uncompressed;
  h.table_ptr = reinterpret_cast<uint64_t>(codec->d_dec);
  uint64_t o = sizeof(StrHeader);
  h.shared_prefix_off = static_cast<uint32_t>(o);
  o += round_up(spl, 16);
  h.sp_end = static_cast<uint32_t>(o);
  h.fp_off = build_fp ? static_cast<uint32_t>(o) : 0;
  if (build_fp) o += round_up(4ull * U, 16);
  h.resid_off = static_cast<uint32_t>(o);
  o += round_up(static_cast<uint64_t>(ob) * (U + 1), 16);
  h.prefix_keys_off = static_cast<uint32_t>(o);
  o += round_up(8ull * U, 16);
  h.rows_off = static_cast<uint32_t>(o);
  h.validity_off = h.has_nulls ? static_cast<uint32_t>(o) : 0;
  if (h.has_nulls) o += round_up((n + 7) / 8, 16);
  h.keys_off = static_cast<uint32_t>(o);
  o += round_up(2ull * n, 16);
  h.head_bytes = static_cast<uint32_t>(o);
  h.fsst_off = static_cast<uint32_t>(o);
  h.fsst_bytes = static_cast<uint32_t>(co);
  o += round_up(co, 16) + 16;
  if (o > 0xFFFFFFF0ull) {
    set_error("byte-view entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(o);
  if (ctx->budget && ctx->arena.bytes_used() + o > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena.bytes_used(),
              (unsigned long long)o, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  Scratch& sc = ctx->scratch;
  LC_TRY(sc.reserve(0, o + 4096));
  uint8_t* hb = sc.host(o);
  if (!hb) {
    set_error("str_encode: scratch exhausted");
    return LC_ERR_OOM;
  }
  std::memset(hb, 0, o);
  std::memcpy(hb, &h, sizeof(h));
  if (spl) std::memcpy(hb + h.shared_prefix_off, dict.uptr[0], spl);
  if (U) std::memcpy(hb + h.prefix_keys_off, pkeys.data(), 8ull * U);
  if (build_fp && U) std::memcpy(hb + h.fp_off, fps.data(), 4ull * U);
  for (uint32_t i = 0; i <= U; ++i) {
    if (ob == 1) reinterpret_cast<int8_t*>(hb + h.resid_off)[i] = static_cast<int8_t>(resid[i]);
    else if (ob == 2) reinterpret_cast<int16_t*>(hb + h.resid_off)[i] = static_cast<int16_t>(resid[i]);
    else reinterpret_cast<int32_t*>(hb + h.resid_off)[i] = resid[i];
  }
  if (h.has_nulls) {
    uint8_t* vb = hb + h.validity_off;
    for (uint32_t i = 0; i < n; ++i)
      if (rows[i].valid) vb[i >> 3] |= static_cast<uint8_t>(1u << (i & 7));
  }
  if (n) std::memcpy(hb + h.keys_off, keys.data(), 2ull * n);
  if (co) std::memcpy(hb + h.fsst_off, comp.data(), co);

  uint32_t slab = 0;
  uint8_t* d_blob = ctx->arena.alloc(o, &slab);
  if (!d_blob) {
    set_error("HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)o);
    return LC_ERR_OOM;
  }
  LC_CUDA_OK(cudaMemcpyAsync(d_blob, hb, o, cudaMemcpyHostToDevice, ctx->stream));
  LC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->h2d_bytes += o;

  Entry* e = new Entry();
  e->liquid_type = LC_LIQUID_BYTE_VIEW;
  e->d_blob = d_blob;
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->arrow_format = in.format;
  e->dict_value_format = in.dict_value_format;
  e->sh = h;
  e->shared_prefix.assign(hb + h.shared_prefix_off, hb + h.shared_prefix_off + spl);
  e->codec = codec;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

}  // namespace lc
