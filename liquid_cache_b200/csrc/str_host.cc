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

#include <chrono>
#include <cstdlib>

#include "host_common.h"
#include "host_pool.h"

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
  // one table per scope, trained by whoever gets here first; a second thread inserting the chunk's next batch waits on the
  // slot and finds the table (with_fsst_compressor_or_train holds a RwLock around the same decision, utils.rs:90-130)
  std::shared_ptr<lc_ctx::CodecSlot> slot = ctx->codec_slot(scope);
  std::lock_guard<std::mutex> train_lock(slot->mu);
  if (slot->codec) {
    *out = slot->codec;
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
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_dec, &codec->dec, sizeof(FsstTable), cudaMemcpyHostToDevice, ctx->L()->stream));
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_enc, codec->enc.get(), sizeof(FsstEncTable), cudaMemcpyHostToDevice, ctx->L()->stream));
  LC_CUDA_OK(cudaStreamSynchronize(ctx->L()->stream));
  slot->codec = codec;
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
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_dec, &codec->dec, sizeof(FsstTable), cudaMemcpyHostToDevice, ctx->L()->stream));
  LC_CUDA_OK(cudaMemcpyAsync(codec->d_enc, codec->enc.get(), sizeof(FsstEncTable), cudaMemcpyHostToDevice, ctx->L()->stream));
  LC_CUDA_OK(cudaStreamSynchronize(ctx->L()->stream));
  std::shared_ptr<lc_ctx::CodecSlot> slot = ctx->codec_slot(scope);
  std::lock_guard<std::mutex> g(slot->mu);
  slot->codec = codec;
  return LC_OK;
}

// A contiguous piece of caller memory that becomes part of the device byte pool.
struct PoolSeg {
  const uint8_t* p;
  uint64_t bytes;
  uint64_t base;  // offset inside the pool
};

// Step 1 of an insert, shared by the one-batch and the batched form: rows as (pool offset, length) over the caller's
// buffers (segments that will become the device byte pool), validity re-aligned to bit offset 0.
struct RowPlan {
  uint32_t n = 0;
  std::vector<uint32_t> row_off, row_len;
  std::vector<uint8_t> valid_bits;  // bit offset 0
  bool has_input_nulls = false;
  std::vector<PoolSeg> segs;
  uint64_t pool_bytes = 0, sum_len = 0;
  bool row_is_valid(uint32_t i) const { return !has_input_nulls || bit_get(valid_bits.data(), i); }
  const uint8_t* row_ptr(uint32_t i) const {
    // host address of a row's bytes: the segment that contains its pool offset
    const uint64_t o = row_off[i];
    for (size_t k = segs.size(); k-- > 0;)
      if (o >= segs[k].base) return segs[k].p + (o - segs[k].base);
    return segs[0].p;
  }
};

static int build_row_plan(const ArrowIn& in, RowPlan* plan) {
  const uint32_t n = static_cast<uint32_t>(in.length);
  plan->n = n;
  std::vector<uint32_t>& row_off = plan->row_off;
  std::vector<uint32_t>& row_len = plan->row_len;
  row_off.assign(n, 0);
  row_len.assign(n, 0);
  std::vector<uint8_t>& valid_bits = plan->valid_bits;
  bool& has_input_nulls = plan->has_input_nulls;
  has_input_nulls = in.validity && in.null_count != 0;
  std::vector<PoolSeg>& segs = plan->segs;
  uint64_t& pool_bytes = plan->pool_bytes;
  uint64_t& sum_len = plan->sum_len;
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
  } else if (in.kind == ArrowIn::K_DECIMAL) {
    // fixed-width values (Decimal128 / Decimal256 little-endian words): row i is bytes [i*w, (i+1)*w) of the values buffer
    const uint32_t w = in.dec_width;
    add_seg(static_cast<const uint8_t*>(in.values) + static_cast<uint64_t>(in.offset) * w, static_cast<uint64_t>(n) * w);
    for (uint32_t i = 0; i < n; ++i) {
      row_off[i] = i * w;
      row_len[i] = w;
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
  auto row_is_valid = [&](uint32_t i) -> bool { return plan->row_is_valid(i); };
  for (uint32_t i = 0; i < n; ++i) {
    if (!row_is_valid(i)) row_off[i] = row_len[i] = 0;
    sum_len += row_len[i];
  }
  return LC_OK;
}

int str_encode(lc_ctx* ctx, const ArrowIn& in, int32_t hint, uint64_t scope, Entry** out) {
  const uint32_t n = static_cast<uint32_t>(in.length);
  RowPlan plan;
  LC_TRY(build_row_plan(in, &plan));
  const std::vector<uint32_t>& row_off = plan.row_off;
  const std::vector<uint32_t>& row_len = plan.row_len;
  const std::vector<uint8_t>& valid_bits = plan.valid_bits;
  const bool has_input_nulls = plan.has_input_nulls;
  const std::vector<PoolSeg>& segs = plan.segs;
  const uint64_t pool_bytes = plan.pool_bytes, sum_len = plan.sum_len;
  auto row_is_valid = [&](uint32_t i) -> bool { return plan.row_is_valid(i); };
  auto row_ptr = [&](uint32_t i) -> const uint8_t* { return plan.row_ptr(i); };

  // ---- 2. symbol table: the first batch of a column chunk trains it (transcode.rs:16-33) on its unique values;
  //         training is host work by design (once per chunk, never visible in any result) ----
  std::shared_ptr<FsstCodec> codec;
  {
    codec = ctx->codec_of(scope);
    if (!codec) {
      DictBuilder dict(n < 1024 ? 1024 : n / 2);
      std::vector<uint8_t> ordered;  // fixed-width values are compressed in their order-preserving form: train on that
      if (in.kind == ArrowIn::K_DECIMAL) {
        const uint32_t w = in.dec_width;
        ordered.resize(static_cast<size_t>(n) * w);
        for (uint32_t i = 0; i < n; ++i) {
          if (!row_is_valid(i)) continue;
          fixed_to_ordered(row_ptr(i), w, ordered.data() + static_cast<size_t>(i) * w);
          dict.add(ordered.data() + static_cast<size_t>(i) * w, w);
        }
      } else {
        for (uint32_t i = 0; i < n; ++i)
          if (row_is_valid(i)) dict.add(row_len[i] ? row_ptr(i) : reinterpret_cast<const uint8_t*>(""), row_len[i]);
      }
      LC_TRY(get_codec(ctx, scope, dict, &codec));
    }
  }

  // ---- 3. device pipeline ----
  const bool build_fp = (hint == LC_HINT_SUBSTRING_SEARCH);
  uint32_t cap = 64;
  while (cap < 2u * n) cap <<= 1;
  cudaStream_t s = ctx->L()->stream;
  Scratch& sc = ctx->L()->scratch;
  const uint64_t vbytes = has_input_nulls ? round_up((n + 31) / 32 * 4, 16) : 0;
  const uint64_t up_bytes = round_up(4ull * n, 256) * 2 + round_up(vbytes, 256);
  const uint64_t comp_cap = 2 * sum_len + 64;
  const uint64_t dev_need = round_up(pool_bytes + 64, 256) + up_bytes + round_up(4ull * cap, 256) +
                            6 * round_up(4ull * (n + 1), 256) + round_up(2ull * n, 256) + round_up(8ull * (n + 1), 256) +
                            round_up(8ull * kBloomWords * (n + 1), 256) +
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
  unsigned long long* d_blooms = reinterpret_cast<unsigned long long*>(sc.dev(8ull * kBloomWords * (n + 1)));
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
  if (in.kind == ArrowIn::K_DECIMAL) {  // the one segment holds the n values back to back
    LC_CUDA_OK(launch_fixed_to_ordered(d_pool, n, in.dec_width, s));
    ctx->kernel_launches++;
  }

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
  if (h.bloom_off) o += bloom_section_bytes(U);
  h.fsst_off = static_cast<uint32_t>(o);
  h.fsst_bytes = static_cast<uint32_t>(co);
  o += round_up(co, 16) + 16;
  if (o > 0xFFFFFFF0ull) {
    set_error("byte-view entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(o);
  if (ctx->budget && ctx->arena_used() + o > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(),
              (unsigned long long)o, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  ArenaBlock block(ctx, o);  // handed back on every early return below
  uint8_t* d_blob = block.p;
  const uint32_t slab = block.slab;
  if (!d_blob) {
    set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget for %llu bytes" : "HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)o);
    return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
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
  if (h.bloom_off) {  // rows of the work area -> planes of the blob (entry_layout.h)
    LC_CUDA_OK(launch_bloom_planes(d_blooms, d_res, reinterpret_cast<uint32_t*>(d_blob + h.bloom_off), s));
    ctx->kernel_launches++;
  }
  LC_CUDA_OK(d2d(h.resid_off, d_resid, static_cast<uint64_t>(ob) * (U + 1)));
  LC_CUDA_OK(d2d(h.prefix_keys_off, d_pkeys, 8ull * U));
  if (h.has_nulls) LC_CUDA_OK(d2d(h.validity_off, d_up + 2 * off_len, (n + 7) / 8));
  LC_CUDA_OK(d2d(h.keys_off, d_keys, 2ull * n));
  LC_CUDA_OK(d2d(h.fsst_off, d_comp, co));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->h2d_bytes += sizeof(h);

  Entry* e = new Entry();
  e->liquid_type = LC_LIQUID_BYTE_VIEW;
  e->d_blob = block.release();
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->arrow_format = in.format;
  e->dict_value_format = in.dict_value_format;
  e->sh = h;
  if (spl) {
    const uint8_t* p0 = row_ptr(first_valid);
    if (in.kind == ArrowIn::K_DECIMAL) {  // the blob holds the order-preserving form
      uint8_t tmp[32];
      fixed_to_ordered(p0, in.dec_width, tmp);
      e->shared_prefix.assign(tmp, tmp + spl);
    } else {
      e->shared_prefix.assign(p0, p0 + spl);
    }
  }
  e->codec = codec;
  e->fixed_width = (in.byte_type == BT_DECIMAL128 || in.byte_type == BT_DECIMAL256) ? in.dec_width : 0;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

// ---- batched form -----------------------------------------------------------------------------------------------
// A list of byte-view batches (a row group's worth, one or several columns) in one pass: the row plans are built and the
// value bytes copied into pinned staging by the host pool, ONE upload carries every batch's rows + byte pool, the five
// encode stages run once over the whole list (k_str_encode.cu *_many), ONE download returns the 48-byte results, and the
// sections of every blob are moved device-to-device. Two stream synchronisations per group of batches instead of three
// per batch. Symbol tables are trained first, on the first batch of every column chunk that has none yet
// (transcode.rs:16-33), exactly as the one-batch path would have done in the same order.
namespace {
struct EncTracer {  // LC_TRACE=1: wall-clock split of a batched insert, printed to stderr
  bool on;
  std::chrono::steady_clock::time_point t0;
  EncTracer() : on(std::getenv("LC_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* stage) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[lc_trace] str_encode_many: %s %.3f ms\n", stage, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

int str_encode_many(lc_ctx* ctx, const std::vector<ArrowIn>& ins, int32_t hint, const uint64_t* scopes, std::vector<Entry*>* out) {
  EncTracer tr;
  const uint64_t nb_all = ins.size();
  out->clear();
  if (nb_all == 0) return LC_OK;
  std::vector<RowPlan> plans(nb_all);
  std::vector<int> rcs(nb_all, LC_OK);
  std::vector<std::string> errs(nb_all);
  parallel_for(nb_all, 4, [&](uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      rcs[i] = build_row_plan(ins[i], &plans[i]);
      if (rcs[i] != LC_OK) errs[i] = get_error();  // the message is thread-local
    }
  });
  for (uint64_t i = 0; i < nb_all; ++i)
    if (rcs[i] != LC_OK) {
      set_error("batch %llu: %s", (unsigned long long)i, errs[i].c_str());
      return rcs[i];
    }
  tr.mark("row plans");
  // Symbol tables: every column chunk without one is trained on ITS first batch of this list (what the one-batch path would
  // have done in the same order). Training is ~2 ms of host work per chunk (measured: 1 695 URL values), so the chunks are
  // trained side by side on the host pool; the tables are uploaded afterwards with one synchronisation.
  std::vector<std::shared_ptr<FsstCodec>> codecs(nb_all);
  std::vector<uint64_t> train_first;  // index of the first batch of every scope that needs a table
  std::vector<std::shared_ptr<lc_ctx::CodecSlot>> train_slot;
  {
    std::unordered_map<uint64_t, uint64_t> seen;
    for (uint64_t i = 0; i < nb_all; ++i) {
      if (!seen.emplace(scopes[i], i).second) continue;
      std::shared_ptr<lc_ctx::CodecSlot> slot = ctx->codec_slot(scopes[i]);
      slot->mu.lock();  // held until the table is in place: another thread inserting into the same chunk waits for it
      if (slot->codec) {
        slot->mu.unlock();
      } else {
        train_first.push_back(i);
        train_slot.push_back(slot);
      }
    }
  }
  struct Unlock {  // whatever happens below, the slots are released
    std::vector<std::shared_ptr<lc_ctx::CodecSlot>>* v;
    ~Unlock() {
      for (auto& sl : *v) sl->mu.unlock();
    }
  } unlock{&train_slot};
  std::vector<std::shared_ptr<FsstCodec>> trained(train_first.size());
  parallel_for(train_first.size(), 1, [&](uint64_t b, uint64_t e) {
    for (uint64_t t = b; t < e; ++t) {
      const RowPlan& pl = plans[train_first[t]];
      DictBuilder dict(pl.n < 1024 ? 1024 : pl.n / 2);
      for (uint32_t r = 0; r < pl.n; ++r)
        if (pl.row_is_valid(r)) dict.add(pl.row_len[r] ? pl.row_ptr(r) : reinterpret_cast<const uint8_t*>(""), pl.row_len[r]);
      auto codec = std::make_shared<FsstCodec>();
      fsst_train(dict.uptr.data(), dict.ulen.data(), dict.uptr.size(), codec.get());
      trained[t] = codec;
    }
  });
  tr.mark("fsst training");
  for (uint64_t t = 0; t < train_first.size(); ++t) {
    std::shared_ptr<FsstCodec>& codec = trained[t];
    if (cudaMalloc(reinterpret_cast<void**>(&codec->d_dec), sizeof(FsstTable)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&codec->d_enc), sizeof(FsstEncTable)) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc for FSST tables failed");
      return LC_ERR_OOM;
    }
    LC_CUDA_OK(cudaMemcpyAsync(codec->d_dec, &codec->dec, sizeof(FsstTable), cudaMemcpyHostToDevice, ctx->L()->stream));
    LC_CUDA_OK(cudaMemcpyAsync(codec->d_enc, codec->enc.get(), sizeof(FsstEncTable), cudaMemcpyHostToDevice, ctx->L()->stream));
  }
  if (!train_first.empty()) LC_CUDA_OK(cudaStreamSynchronize(ctx->L()->stream));  // the sources are pageable host memory
  for (uint64_t t = 0; t < train_first.size(); ++t) train_slot[t]->codec = trained[t];
  for (auto& sl : train_slot) sl->mu.unlock();
  train_slot.clear();
  for (uint64_t i = 0; i < nb_all; ++i) codecs[i] = ctx->codec_of(scopes[i]);
  const bool build_fp = (hint == LC_HINT_SUBSTRING_SEARCH);
  cudaStream_t s = ctx->L()->stream;
  Scratch& sc = ctx->L()->scratch;
  struct Taken {
    uint8_t* blob;
    uint32_t slab;
    uint64_t bytes;
  };
  std::vector<Taken> taken;
  auto give_back = [&]() {
    for (const Taken& t : taken) ctx->arena_free(t.slab, t.blob, t.bytes);
    for (Entry* e : *out) delete e;
    out->clear();
  };
  constexpr uint64_t kGroup = 256;
  for (uint64_t g0 = 0; g0 < nb_all; g0 += kGroup) {
    const uint64_t nb = std::min<uint64_t>(kGroup, nb_all - g0);
    // ---- per-batch offsets: upload block [row_off | row_len | validity | pool] mirrored host/device, then work areas ----
    struct Off {
      uint64_t up, off_len, vbytes, pool, pool_staged_end = 0, table, slot, leader, uniq, clen, offsets, fps, blooms, resid, keys, pkeys, comp, res;
      uint32_t cap;
    };
    std::vector<Off> offs(nb);
    uint64_t cur = 0, max_n = 0;
    // per-row (offset, length) pairs and validity of every batch first, the work items behind them (one upload) ...
    for (uint64_t i = 0; i < nb; ++i) {
      const RowPlan& pl = plans[g0 + i];
      Off& o = offs[i];
      o.off_len = round_up(4ull * pl.n, 256);
      o.vbytes = pl.has_input_nulls ? round_up((pl.n + 31) / 32 * 4, 16) : 0;
      o.up = cur;
      cur += 2 * o.off_len + round_up(o.vbytes, 256);
      max_n = std::max<uint64_t>(max_n, pl.n);
    }
    const uint64_t ios_off = cur;
    cur += round_up(nb * sizeof(StrEncIo), 256);
    // ... then the value bytes. A batch whose Arrow buffers are PAGE-LOCKED is uploaded straight from them (no staging
    // copy on the host: the caller's memory is the DMA source); pageable batches are staged into pinned scratch first.
    std::vector<uint8_t> direct(nb, 0);
    for (uint64_t i = 0; i < nb; ++i) {
      const RowPlan& pl = plans[g0 + i];
      bool pinned = !pl.segs.empty();
      for (const PoolSeg& sg : pl.segs) {
        if (!sg.bytes) continue;
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, sg.p) != cudaSuccess || pa.type != cudaMemoryTypeHost) pinned = false;
      }
      cudaGetLastError();
      direct[i] = pinned;
    }
    const uint64_t pool_begin = cur;
    for (int pass = 0; pass < 2; ++pass)  // staged pools first (they travel with the upload block), direct ones behind
      for (uint64_t i = 0; i < nb; ++i) {
        if ((direct[i] != 0) != (pass == 1)) continue;
        offs[i].pool = cur;
        cur += round_up(plans[g0 + i].pool_bytes + 64, 256);
        if (pass == 0) offs[i].pool_staged_end = cur;
      }
    uint64_t staged_end = pool_begin;
    for (uint64_t i = 0; i < nb; ++i)
      if (!direct[i]) staged_end = std::max(staged_end, offs[i].pool_staged_end);
    const uint64_t up_bytes = cur;
    const uint64_t hdr_off = cur;  // host only: blob headers staged for their uploads
    const uint64_t res_off = cur;  // device: results (the header staging area on the host side is reused after the sync)
    cur += round_up(nb * std::max(sizeof(StrEncResult), sizeof(StrHeader)), 256);
    const uint64_t asm_off = cur;  // blob assembly work items (host staging + device copy)
    cur += round_up(nb * sizeof(StrAsmWork), 256);
    const uint64_t host_total = cur;
    const uint64_t table_off = cur;  // all hash tables back to back: one memset
    uint64_t table_words = 0;
    for (uint64_t i = 0; i < nb; ++i) {
      const RowPlan& pl = plans[g0 + i];
      uint32_t cap = 64;
      while (cap < 2u * pl.n) cap <<= 1;
      offs[i].cap = cap;
      offs[i].table = cur;
      cur += 4ull * cap;
      table_words += cap;
    }
    cur = round_up(cur, 256);
    for (uint64_t i = 0; i < nb; ++i) {
      const RowPlan& pl = plans[g0 + i];
      Off& o = offs[i];
      auto take = [&](uint64_t bytes) {
        const uint64_t at = cur;
        cur += round_up(bytes, 256);
        return at;
      };
      const uint64_t n1 = pl.n + 1ull;
      o.slot = take(4 * n1);
      o.leader = take(4 * n1);
      o.uniq = take(4 * n1);
      o.clen = take(4 * n1);
      o.offsets = take(4 * n1);
      o.fps = take(4 * n1);
      o.blooms = take(8ull * kBloomWords * n1);
      o.resid = take(4 * n1);
      o.keys = take(2ull * pl.n + 16);
      o.pkeys = take(8 * n1);
      o.comp = take(2 * pl.sum_len + 64);
    }
    const uint64_t dev_total = cur;
    LC_TRY(sc.reserve(dev_total + 1024, host_total + 1024));
    uint8_t* h = sc.host(host_total);
    uint8_t* d = sc.dev(dev_total);
    if (!h || !d) {
      give_back();
      set_error("str_encode_many: scratch exhausted");
      return LC_ERR_OOM;
    }
    StrEncIo* h_ios = reinterpret_cast<StrEncIo*>(h + ios_off);
    parallel_for(nb, 2, [&](uint64_t b, uint64_t e) {
      for (uint64_t i = b; i < e; ++i) {
        const RowPlan& pl = plans[g0 + i];
        const Off& o = offs[i];
        if (pl.n) {
          std::memcpy(h + o.up, pl.row_off.data(), 4ull * pl.n);
          std::memcpy(h + o.up + o.off_len, pl.row_len.data(), 4ull * pl.n);
        }
        if (o.vbytes) {
          std::memset(h + o.up + 2 * o.off_len, 0, o.vbytes);
          std::memcpy(h + o.up + 2 * o.off_len, pl.valid_bits.data(), (pl.n + 7) / 8);
        }
        if (!direct[i])
          for (const PoolSeg& sg : pl.segs)
            if (sg.bytes) std::memcpy(h + o.pool + sg.base, sg.p, sg.bytes);
        StrEncIo io{};
        io.pool = d + o.pool;
        io.row_off = reinterpret_cast<const uint32_t*>(d + o.up);
        io.row_len = reinterpret_cast<const uint32_t*>(d + o.up + o.off_len);
        io.valid = o.vbytes ? reinterpret_cast<const uint32_t*>(d + o.up + 2 * o.off_len) : nullptr;
        io.n = pl.n;
        io.table_mask = o.cap - 1;
        io.row_slot = reinterpret_cast<uint32_t*>(d + o.slot);
        io.table = reinterpret_cast<uint32_t*>(d + o.table);
        io.leader = reinterpret_cast<uint32_t*>(d + o.leader);
        io.keys = reinterpret_cast<uint16_t*>(d + o.keys);
        io.uniq_row = reinterpret_cast<uint32_t*>(d + o.uniq);
        io.clen = reinterpret_cast<uint32_t*>(d + o.clen);
        io.offsets = reinterpret_cast<uint32_t*>(d + o.offsets);
        io.pkeys = reinterpret_cast<unsigned long long*>(d + o.pkeys);
        io.fps = build_fp ? reinterpret_cast<uint32_t*>(d + o.fps) : nullptr;
        io.blooms = reinterpret_cast<unsigned long long*>(d + o.blooms);
        io.comp = d + o.comp;
        io.resid = d + o.resid;
        io.enc = codecs[g0 + i]->d_enc;
        io.res = reinterpret_cast<StrEncResult*>(d + res_off) + i;
        h_ios[i] = io;
      }
    });
    tr.mark("stage into pinned memory");
    cudaError_t ce = cudaMemcpyAsync(d, h, staged_end, cudaMemcpyHostToDevice, s);
    for (uint64_t i = 0; i < nb && ce == cudaSuccess; ++i) {
      if (!direct[i]) continue;
      for (const PoolSeg& sg : plans[g0 + i].segs)
        if (sg.bytes && ce == cudaSuccess) ce = cudaMemcpyAsync(d + offs[i].pool + sg.base, sg.p, sg.bytes, cudaMemcpyHostToDevice, s);
    }
    if (ce == cudaSuccess)
      ce = launch_str_encode_many(reinterpret_cast<const StrEncIo*>(d + ios_off), static_cast<uint32_t>(nb), static_cast<uint32_t>(max_n),
                                  reinterpret_cast<uint32_t*>(d + table_off), table_words, s);
    StrEncResult* h_res = reinterpret_cast<StrEncResult*>(h + hdr_off);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_res, d + res_off, nb * sizeof(StrEncResult), cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) {
      give_back();
      set_error("CUDA error in str_encode_many: %s", cudaGetErrorString(ce));
      return LC_ERR_CUDA;
    }
    tr.mark("upload + 5 kernels + results");
    ctx->kernel_launches += 5;
    ctx->h2d_bytes += up_bytes + nb * sizeof(StrEncIo);
    ctx->d2h_bytes += nb * sizeof(StrEncResult);
    // ---- blob layout per batch (sizes are known now), sections moved device-to-device ----
    std::vector<StrEncResult> results(h_res, h_res + nb);  // the pinned area is reused for the headers below
    StrAsmWork* h_asm = reinterpret_cast<StrAsmWork*>(h + asm_off);
    for (uint64_t i = 0; i < nb; ++i) {
      const StrEncResult& r = results[i];
      const RowPlan& pl = plans[g0 + i];
      const ArrowIn& in = ins[g0 + i];
      const Off& of = offs[i];
      if (r.error) {
        give_back();
        set_error(r.error == 1 ? "batch %llu: more than 65536 distinct values in one batch" : "batch %llu: compressed dictionary exceeds 4 GiB",
                  (unsigned long long)(g0 + i));
        return LC_ERR_UNSUPPORTED_TYPE;
      }
      const uint32_t n = pl.n, U = r.n_unique, spl = r.shared_prefix_len, ob = r.offset_bytes;
      const uint64_t co = r.comp_bytes;
      StrHeader hd;
      std::memset(&hd, 0, sizeof(hd));
      hd.magic = kMagicStr;
      hd.arrow_type = in.byte_type;
      hd.has_nulls = r.null_count > 0;
      hd.has_fp = build_fp;
      hd.offset_bytes = static_cast<uint8_t>(ob);
      hd.n = n;
      hd.n_unique = U;
      hd.slope = r.slope;
      hd.intercept = r.intercept;
      hd.shared_prefix_len = spl;
      hd.null_count = r.null_count;
      hd.max_value_len = r.max_value_len;
      hd.uncompressed_bytes = r.uncompressed_bytes;
      hd.table_ptr = reinterpret_cast<uint64_t>(codecs[g0 + i]->d_dec);
      uint64_t o = sizeof(StrHeader);
      hd.shared_prefix_off = static_cast<uint32_t>(o);
      o += round_up(spl, 16);
      hd.sp_end = static_cast<uint32_t>(o);
      hd.fp_off = build_fp ? static_cast<uint32_t>(o) : 0;
      if (build_fp) o += round_up(4ull * U, 16);
      hd.resid_off = static_cast<uint32_t>(o);
      o += round_up(static_cast<uint64_t>(ob) * (U + 1), 16);
      hd.prefix_keys_off = static_cast<uint32_t>(o);
      o += round_up(8ull * U, 16);
      hd.rows_off = static_cast<uint32_t>(o);
      hd.validity_off = hd.has_nulls ? static_cast<uint32_t>(o) : 0;
      if (hd.has_nulls) o += round_up((n + 7) / 8, 16);
      hd.keys_off = static_cast<uint32_t>(o);
      o += round_up(2ull * n, 16);
      hd.head_bytes = static_cast<uint32_t>(o);
      hd.bloom_off = (build_fp && U) ? static_cast<uint32_t>(o) : 0;
      if (hd.bloom_off) o += bloom_section_bytes(U);
      hd.fsst_off = static_cast<uint32_t>(o);
      hd.fsst_bytes = static_cast<uint32_t>(co);
      o += round_up(co, 16) + 16;
      if (o > 0xFFFFFFF0ull) {
        give_back();
        set_error("byte-view entry too large");
        return LC_ERR_UNSUPPORTED_TYPE;
      }
      hd.blob_bytes = static_cast<uint32_t>(o);
      if (ctx->budget && ctx->arena_used() + o > ctx->budget) {
        give_back();
        set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(), (unsigned long long)o,
                  (unsigned long long)ctx->budget);
        return LC_ERR_CACHE_FULL;
      }
      uint32_t slab = 0;
      uint8_t* d_blob = ctx->arena_alloc(o, &slab);
      if (!d_blob) {
        give_back();
        set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget for %llu bytes" : "HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)o);
        return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
      }
      taken.push_back({d_blob, slab, o});
      uint32_t first_valid = 0;
      while (first_valid < n && !pl.row_is_valid(first_valid)) ++first_valid;
      // the blob is laid out by ONE kernel for the whole group (k_str_assemble): sections from the pipeline's work areas,
      // gaps zeroed, header written — instead of a memset and nine device-to-device copies per batch
      StrAsmWork& aw = h_asm[i];
      aw.blob = d_blob;
      aw.blob_bytes = hd.blob_bytes;
      aw.hdr = hd;
      aw.n_segs = 0;
      auto seg = [&](uint32_t dst_off, const void* src, uint64_t bytes) {
        if (!bytes) return;
        aw.segs[aw.n_segs++] = StrAsmSeg{static_cast<const uint8_t*>(src), dst_off, static_cast<uint32_t>(bytes)};
      };
      if (spl) seg(hd.shared_prefix_off, d + of.pool + pl.row_off[first_valid], spl);
      if (build_fp) seg(hd.fp_off, d + of.fps, 4ull * U);
      seg(hd.resid_off, d + of.resid, static_cast<uint64_t>(ob) * (U + 1));
      seg(hd.prefix_keys_off, d + of.pkeys, 8ull * U);
      if (hd.has_nulls) seg(hd.validity_off, d + of.up + 2 * of.off_len, (n + 7) / 8);
      seg(hd.keys_off, d + of.keys, 2ull * n);
      if (hd.bloom_off) seg(hd.bloom_off, d + of.blooms, bloom_section_bytes(U));  // transposed into planes by k_str_assemble
      seg(hd.fsst_off, d + of.comp, co);
      Entry* e = new Entry();
      e->liquid_type = LC_LIQUID_BYTE_VIEW;
      e->d_blob = d_blob;
      e->blob_bytes = hd.blob_bytes;
      e->slab = slab;
      e->n = n;
      e->arrow_format = in.format;
      e->dict_value_format = in.dict_value_format;
      e->sh = hd;
      if (spl) {
        const uint8_t* p0 = pl.row_ptr(first_valid);
        e->shared_prefix.assign(p0, p0 + spl);
      }
      e->codec = codecs[g0 + i];
      out->push_back(e);
    }
    ce = cudaMemcpyAsync(d + asm_off, h_asm, nb * sizeof(StrAsmWork), cudaMemcpyHostToDevice, s);
    if (ce == cudaSuccess) ce = launch_str_assemble(reinterpret_cast<const StrAsmWork*>(d + asm_off), static_cast<uint32_t>(nb), s);
    ctx->kernel_launches++;
    tr.mark("blob layout + assembly launched");
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);  // the group's scratch is reused by the next group
    if (ce != cudaSuccess) {
      give_back();
      set_error("CUDA error in str_encode_many: %s", cudaGetErrorString(ce));
      return LC_ERR_CUDA;
    }
    tr.mark("section copies done");
    ctx->h2d_bytes += nb * sizeof(StrAsmWork);
  }
  ctx->n_entries += out->size();
  return LC_OK;
}

}  // namespace lc
