// str_host.cc — byte-view insert side and predicate planning.
// Reference: LiquidByteViewArray::from_dict_array_inner
//   (/root/reference/src/core/src/liquid_array/byte_view_array/conversions.rs:260-373),
//   CheckedDictionaryArray (src/core/src/utils/mod.rs:52-154), PrefixKey / CompactOffsets / fit_line
//   (raw/fsst_buffer.rs:160-187, 267-383), StringFingerprint (byte_view_array/fingerprint.rs:19-26),
//   with_fsst_compressor_or_train (src/core/src/cache/transcode.rs:16-33).
// The host only turns the Arrow layout into (offset, length) pairs over uploaded buffers and sizes the blob; the
// dictionary, compression, prefix keys, fingerprints and offset fit run on the device (k_str_encode.cu). FSST
// training (once per column chunk, invisible in every result) is the one piece of byte work left here.
#include <algorithm>
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

void set_bits_range(uint8_t* dst, uint32_t n) {  // first n bits := 1
  std::memset(dst, 0xFF, n / 8);
  if (n & 7) dst[n / 8] |= static_cast<uint8_t>((1u << (n & 7)) - 1u);
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

// A codec that came from a stored symbol table (lc_ctx_load_symbol_table): device copies, then the scope's entry.
int register_codec(lc_ctx* ctx, uint64_t scope, const std::shared_ptr<FsstCodec>& codec) {
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
  return LC_OK;
}

// A contiguous piece of caller memory that becomes part of the device byte pool.
struct PoolSeg {
  const uint8_t* p;
  uint64_t bytes;
  uint64_t base;  // offset inside the pool
};

int str_encode(lc_ctx* ctx, const ArrowIn& in, int32_t hint, uint64_t scope, Entry** out) {
  const uint32_t n = static_cast<uint32_t>(in.length);
  // ---- 1. rows as (pool offset, length); the value bytes themselves are never touched on the host (except by the
  //         once-per-column-chunk FSST training below) ----
  std::vector<uint32_t> row_off(n, 0), row_len(n, 0);
  std::vector<uint8_t> valid_bits;  // bit offset 0
  bool has_input_nulls = in.validity && in.null_count != 0;
  std::vector<PoolSeg> segs;
  uint64_t pool_bytes = 0, sum_len = 0;
  auto add_seg = [&](const uint8_t* p, uint64_t bytes) -> uint64_t {
    const uint64_t base = pool_bytes;
    segs.push_back(PoolSeg{p, bytes, base});
    pool_bytes += round_up(bytes, 16);
    return base;
  };
  auto is_valid = [&](uint32_t i) -> bool { return !has_input_nulls || bit_get(in.validity, in.offset + i); };
  if (in.kind == ArrowIn::K_BYTES) {
    const int32_t* off = static_cast<const int32_t*>(in.values) + in.offset;
    const int64_t lo = n ? off[0] : 0, hi = n ? off[n] : 0;
    add_seg(in.data + lo, static_cast<uint64_t>(hi - lo));
    for (uint32_t i = 0; i < n; ++i) {
      row_off[i] = static_cast<uint32_t>(off[i] - lo);
      row_len[i] = static_cast<uint32_t>(off[i + 1] - off[i]);
    }
  } else if (in.kind == ArrowIn::K_VIEW) {
    const uint8_t* views = static_cast<const uint8_t*>(in.values) + 16 * in.offset;
    add_seg(views, 16ull * n);  // inline payloads are addressed in place (view bytes 4..15)
    const int64_t nb = in.n_view_buffers;
    std::vector<uint64_t> lo(nb > 0 ? nb : 0, ~0ull), hi(nb > 0 ? nb : 0, 0);
    for (uint32_t i = 0; i < n; ++i) {
      if (!is_valid(i)) continue;
      uint32_t len, bi, bo;
      std::memcpy(&len, views + 16ull * i, 4);
      if (len <= 12) continue;
      std::memcpy(&bi, views + 16ull * i + 8, 4);
      std::memcpy(&bo, views + 16ull * i + 12, 4);
      if (static_cast<int64_t>(bi) >= nb) {
        set_error("view buffer index out of range");
        return LC_ERR_INVALID;
      }
      lo[bi] = std::min<uint64_t>(lo[bi], bo);
      hi[bi] = std::max<uint64_t>(hi[bi], static_cast<uint64_t>(bo) + len);
    }
    std::vector<uint64_t> base(nb > 0 ? nb : 0, 0);
    for (int64_t b = 0; b < nb; ++b)
      if (hi[b] > lo[b] && lo[b] != ~0ull)
        base[b] = add_seg(static_cast<const uint8_t*>(in.view_buffers[b]) + lo[b], hi[b] - lo[b]) - lo[b];
    for (uint32_t i = 0; i < n; ++i) {
      if (!is_valid(i)) continue;
      uint32_t len, bi, bo;
      std::memcpy(&len, views + 16ull * i, 4);
      row_len[i] = len;
      if (len <= 12) {
        row_off[i] = 16u * i + 4u;
      } else {
        std::memcpy(&bi, views + 16ull * i + 8, 4);
        std::memcpy(&bo, views + 16ull * i + 12, 4);
        const uint64_t o = base[bi] + bo;
        if (o > 0xFFFFFFFFull) {
          set_error("string batch larger than 4 GiB");
          return LC_ERR_UNSUPPORTED_TYPE;
        }
        row_off[i] = static_cast<uint32_t>(o);
      }
    }
  } else {  // K_DICT: a null dictionary value makes the row null (typed dictionary iterator)
    const int32_t* doff = in.dict_offsets + in.dict_offset;
    const int64_t lo = in.dict_len ? doff[0] : 0, hi = in.dict_len ? doff[in.dict_len] : 0;
    add_seg(in.dict_data + lo, static_cast<uint64_t>(hi - lo));
    bool extra_nulls = false;
    for (uint32_t i = 0; i < n; ++i) {
      if (!is_valid(i)) continue;
      const uint32_t key = in.dict_keys[in.offset + i];
      if (static_cast<int64_t>(key) >= in.dict_len) {
        set_error("dictionary key out of range");
        return LC_ERR_INVALID;
      }
      if (in.dict_validity && !bit_get(in.dict_validity, in.dict_offset + key)) {
        if (!extra_nulls) {  // materialise a validity bitmap that also carries the dictionary's nulls
          valid_bits.assign(round_up((n + 7) / 8, 16), 0);
          if (has_input_nulls) copy_bits(in.validity, in.offset, n, valid_bits.data(), valid_bits.size());
          else set_bits_range(valid_bits.data(), n);
          extra_nulls = true;
        }
        valid_bits[i >> 3] &= static_cast<uint8_t>(~(1u << (i & 7)));
        continue;
      }
      row_off[i] = static_cast<uint32_t>(doff[key] - lo);
      row_len[i] = static_cast<uint32_t>(doff[key + 1] - doff[key]);
    }
    if (extra_nulls) has_input_nulls = true;
  }
  if (pool_bytes > 0xFFFFFFF0ull) {
    set_error("string batch larger than 4 GiB");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  if (has_input_nulls && valid_bits.empty()) {
    valid_bits.assign(round_up((n + 7) / 8, 16), 0);
    copy_bits(in.validity, in.offset, n, valid_bits.data(), valid_bits.size());
  }
  auto row_is_valid = [&](uint32_t i) -> bool { return !has_input_nulls || bit_get(valid_bits.data(), i); };
  for (uint32_t i = 0; i < n; ++i) {
    if (!row_is_valid(i)) row_off[i] = row_len[i] = 0;
    sum_len += row_len[i];
  }
  auto row_ptr = [&](uint32_t i) -> const uint8_t* {
    // host address of a row's bytes: the segment that contains its pool offset
    const uint64_t o = row_off[i];
    for (size_t k = segs.size(); k-- > 0;)
      if (o >= segs[k].base) return segs[k].p + (o - segs[k].base);
    return segs[0].p;
  };

  // ---- 2. symbol table: the first batch of a column chunk trains it (transcode.rs:16-33) on its unique values;
  //         training is host work by design (once per chunk, never visible in any result) ----
  std::shared_ptr<FsstCodec> codec;
  {
    auto it = ctx->codecs.find(scope);
    if (it != ctx->codecs.end()) {
      codec = it->second;
    } else {
      DictBuilder dict(n < 1024 ? 1024 : n / 2);
      for (uint32_t i = 0; i < n; ++i)
        if (row_is_valid(i)) dict.add(row_len[i] ? row_ptr(i) : reinterpret_cast<const uint8_t*>(""), row_len[i]);
      LC_TRY(get_codec(ctx, scope, dict, &codec));
    }
  }

  // ---- 3. device pipeline ----
  const bool build_fp = (hint == LC_HINT_SUBSTRING_SEARCH);
  uint32_t cap = 64;
  while (cap < 2u * n) cap <<= 1;
  cudaStream_t s = ctx->stream;
  Scratch& sc = ctx->scratch;
  const uint64_t vbytes = has_input_nulls ? round_up((n + 31) / 32 * 4, 16) : 0;
  const uint64_t up_bytes = round_up(4ull * n, 256) * 2 + round_up(vbytes, 256);
  const uint64_t comp_cap = 2 * sum_len + 64;
  const uint64_t dev_need = round_up(pool_bytes + 64, 256) + up_bytes + round_up(4ull * cap, 256) +
                            6 * round_up(4ull * (n + 1), 256) + round_up(2ull * n, 256) + 2 * round_up(8ull * (n + 1), 256) +
                            round_up(comp_cap, 256) + 4096;
  LC_TRY(sc.reserve(dev_need, up_bytes + 4096));
  uint8_t* h_up = sc.host(up_bytes);
  StrEncResult* h_res = reinterpret_cast<StrEncResult*>(sc.host(sizeof(StrEncResult)));
  uint8_t* d_pool = sc.dev(pool_bytes + 64);
  uint8_t* d_up = sc.dev(up_bytes);
  uint32_t* d_table = reinterpret_cast<uint32_t*>(sc.dev(4ull * cap));
  uint32_t* d_slot = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  uint32_t* d_leader = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  uint32_t* d_uniq = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  uint32_t* d_clen = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  uint32_t* d_offsets = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  uint32_t* d_fps = reinterpret_cast<uint32_t*>(sc.dev(4ull * (n + 1)));
  unsigned long long* d_blooms = reinterpret_cast<unsigned long long*>(sc.dev(8ull * (n + 1)));
  uint8_t* d_resid = sc.dev(4ull * (n + 1));
  uint16_t* d_keys = reinterpret_cast<uint16_t*>(sc.dev(2ull * n + 16));
  unsigned long long* d_pkeys = reinterpret_cast<unsigned long long*>(sc.dev(8ull * (n + 1)));
  uint8_t* d_comp = sc.dev(comp_cap);
  StrEncResult* d_res = reinterpret_cast<StrEncResult*>(sc.dev(sizeof(StrEncResult)));
  if (!h_up || !h_res || !d_pool || !d_up || !d_table || !d_slot || !d_leader || !d_uniq || !d_clen || !d_offsets ||
      !d_fps || !d_blooms || !d_resid || !d_keys || !d_pkeys || !d_comp || !d_res) {
    set_error("str_encode: scratch exhausted");
    return LC_ERR_OOM;
  }
  const uint64_t off_len = round_up(4ull * n, 256);
  if (n) {
    std::memcpy(h_up, row_off.data(), 4ull * n);
    std::memcpy(h_up + off_len, row_len.data(), 4ull * n);
  }
  if (vbytes) {
    std::memset(h_up + 2 * off_len, 0, vbytes);
    std::memcpy(h_up + 2 * off_len, valid_bits.data(), (n + 7) / 8);
  }
  LC_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_bytes, cudaMemcpyHostToDevice, s));
  for (const PoolSeg& sg : segs)
    if (sg.bytes) LC_CUDA_OK(cudaMemcpyAsync(d_pool + sg.base, sg.p, sg.bytes, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_bytes + pool_bytes;

  StrEncIo io{};
  io.pool = d_pool;
  io.row_off = reinterpret_cast<const uint32_t*>(d_up);
  io.row_len = reinterpret_cast<const uint32_t*>(d_up + off_len);
  io.valid = vbytes ? reinterpret_cast<const uint32_t*>(d_up + 2 * off_len) : nullptr;
  io.n = n;
  io.table_mask = cap - 1;
  io.row_slot = d_slot;
  io.table = d_table;
  io.leader = d_leader;
  io.keys = d_keys;
  io.uniq_row = d_uniq;
  io.clen = d_clen;
  io.offsets = d_offsets;
  io.pkeys = d_pkeys;
  io.fps = build_fp ? d_fps : nullptr;
  io.blooms = d_blooms;
  io.comp = d_comp;
  io.resid = d_resid;
  io.enc = codec->d_enc;
  io.res = d_res;
  LC_CUDA_OK(launch_str_encode(io, s));
  ctx->kernel_launches += 5;
  LC_CUDA_OK(cudaMemcpyAsync(h_res, d_res, sizeof(StrEncResult), cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += sizeof(StrEncResult);
  if (h_res->error == 1) {
    // the reference's UInt16 dictionary builder overflows (panics) here; we decline the batch
    set_error("more than 65536 distinct values in one batch");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  if (h_res->error) {
    set_error("compressed dictionary exceeds 4 GiB");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  const uint32_t U = h_res->n_unique, spl = h_res->shared_prefix_len, ob = h_res->offset_bytes;
  const uint32_t null_count = h_res->null_count;
  const uint64_t co = h_res->comp_bytes;

  // ---- 4. blob layout (sizes are known now), sections moved device-to-device ----
  StrHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicStr;
  h.arrow_type = in.byte_type;
  h.has_nulls = null_count > 0;
  h.has_fp = build_fp;
  h.offset_bytes = static_cast<uint8_t>(ob);
  h.n = n;
  h.n_unique = U;
  h.slope = h_res->slope;
  h.intercept = h_res->intercept;
  h.shared_prefix_len = spl;
  h.null_count = null_count;
  h.max_value_len = h_res->max_value_len;
  h.uncompressed_bytes = h_res->uncompressed_bytes;
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
  h.bloom_off = (build_fp && U) ? static_cast<uint32_t>(o) : 0;
  if (h.bloom_off) o += round_up(8ull * U, 16);
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
  uint32_t slab = 0;
  uint8_t* d_blob = ctx->arena.alloc(o, &slab);
  if (!d_blob) {
    set_error("HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)o);
    return LC_ERR_OOM;
  }
  // first valid row = unique 0 = where the shared prefix is read from
  uint32_t first_valid = 0;
  while (first_valid < n && !row_is_valid(first_valid)) ++first_valid;
  LC_CUDA_OK(cudaMemsetAsync(d_blob, 0, o, s));  // padding between sections reads as zero
  std::memcpy(h_up, &h, sizeof(h));              // h_up is pinned and idle after the sync above
  LC_CUDA_OK(cudaMemcpyAsync(d_blob, h_up, sizeof(h), cudaMemcpyHostToDevice, s));
  auto d2d = [&](uint32_t dst_off, const void* src, uint64_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(d_blob + dst_off, src, bytes, cudaMemcpyDeviceToDevice, s);
  };
  if (spl) LC_CUDA_OK(d2d(h.shared_prefix_off, d_pool + row_off[first_valid], spl));
  if (build_fp) LC_CUDA_OK(d2d(h.fp_off, d_fps, 4ull * U));
  if (h.bloom_off) LC_CUDA_OK(d2d(h.bloom_off, d_blooms, 8ull * U));
  LC_CUDA_OK(d2d(h.resid_off, d_resid, static_cast<uint64_t>(ob) * (U + 1)));
  LC_CUDA_OK(d2d(h.prefix_keys_off, d_pkeys, 8ull * U));
  if (h.has_nulls) LC_CUDA_OK(d2d(h.validity_off, d_up + 2 * off_len, (n + 7) / 8));
  LC_CUDA_OK(d2d(h.keys_off, d_keys, 2ull * n));
  LC_CUDA_OK(d2d(h.fsst_off, d_comp, co));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->h2d_bytes += sizeof(h);

  Entry* e = new Entry();
  e->liquid_type = LC_LIQUID_BYTE_VIEW;
  e->d_blob = d_blob;
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->arrow_format = in.format;
  e->dict_value_format = in.dict_value_format;
  e->sh = h;
  if (spl) {
    const uint8_t* p0 = row_ptr(first_valid);
    e->shared_prefix.assign(p0, p0 + spl);
  }
  e->codec = codec;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

}  // namespace lc
