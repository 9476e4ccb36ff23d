// ipc_host.cc — LQDA, the reference's serialized form of a liquid array, for the integer-shaped entries.
//   LiquidArray::to_bytes            liquid_array/mod.rs:116-121
//   LiquidIPCHeader                  liquid_array/ipc.rs:158-250   (16 bytes: "LQDA", version 1, logical id, physical id)
//   ipc::read_from_bytes             liquid_array/ipc.rs:252-283
//   LiquidPrimitiveArray             primitive_array.rs:603-679    header | reference (T bytes) | pad 8 | BitPackedArray
//   LiquidFloatArray                 float_array.rs:393-600        header | reference | pad 8 | e f pad6 | patch count u64 |
//                                                                  indices u64[] | values T[] | pad 8 | BitPackedArray
//   LiquidDecimalArray               decimal_array.rs:180-251      header | {is256, precision, scale, pad5} | reference u64 | BitPackedArray
//   BitPackedArray                   raw/bit_pack_array.rs:181-334 len u32, width u8, has_nulls u8, nulls_len u32, values_len u32,
//                                                                  pad2 | null bitmap | pad 8 | FastLanes words
// (all paths under /root/reference/src/core/src). The entry's validity and packed sections ARE the LQDA null bitmap and
// values, so both directions are header arithmetic on the host plus copies between the blob and the caller's bytes; the only
// kernels are the u32 <-> u64 conversion of ALP patch indices.
//   LiquidByteViewArray              byte_view_array/serialization.rs:87-325   header | {keys, offsets, prefix, fsst, fingerprint sizes} |
//                                    RawFsstBuffer | BitPackedArray<u16> keys at W = 16 | CompactOffsets | PrefixKeys | shared prefix | fingerprints
//   symbol table                     raw/fsst_buffer.rs:854-932  (count u8, lengths, symbols as u64 LE) — travels beside the arrays, as the
//                                    reference's LiquidIPCContext carries the compressor
// Byte-view entries use the same section copies; their u16 keys are FastLanes-transposed by k_int_pack on the way out and
// brought back by k_int_scan<DECODE> on the way in (a 16-bit integer entry of width 16 is exactly that section).
#include <atomic>

#include "host_common.h"
#include "host_pool.h"
#include "lqda_layout.h"

namespace lc {

namespace {
constexpr uint32_t kLqdaMagic = 0x4C514441u;

void put_u16(uint8_t* p, uint16_t v) { std::memcpy(p, &v, 2); }
void put_u32(uint8_t* p, uint32_t v) { std::memcpy(p, &v, 4); }
void put_u64(uint8_t* p, uint64_t v) { std::memcpy(p, &v, 8); }
uint16_t get_u16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
uint32_t get_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
uint64_t get_u64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
uint64_t pad8(uint64_t x) { return (x + 7) & ~7ull; }

struct PhysRow { const char* fmt; uint8_t tbits; bool sg; };
// physical type id -> Arrow C format, width, signedness (ipc.rs:26-47; PhysType of entry_layout.h has the same numbering)
const PhysRow kPhys[16] = {
    {"c", 8, true},   {"s", 16, true},  {"i", 32, true},   {"l", 64, true},   {"C", 8, false},    {"S", 16, false},
    {"I", 32, false}, {"L", 64, false}, {"f", 32, true},   {"g", 64, true},   {"tdD", 32, true},  {"tdm", 64, true},
    {"tss:", 64, true}, {"tsm:", 64, true}, {"tsu:", 64, true}, {"tsn:", 64, true}};

struct Layout {  // byte offsets inside the LQDA image
  uint64_t ref_off = 16, exp_off = 0, pidx_off = 0, pval_off = 0, bp_off = 0, nulls_off = 0, values_off = 0, total = 0;
  uint32_t nulls_len = 0, values_len = 0;
};

Layout layout_of(int32_t liquid_type, uint32_t tb, uint32_t n, uint32_t width, bool has_nulls, uint32_t n_chunks, uint32_t n_patches) {
  Layout L;
  uint64_t cur;
  if (liquid_type == LC_LIQUID_DECIMAL) {
    L.ref_off = 24;
    cur = 32;
  } else {
    cur = pad8(16 + tb);
  }
  if (liquid_type == LC_LIQUID_FLOAT) {
    L.exp_off = cur;
    cur += 16;  // e, f, pad6, patch count
    L.pidx_off = cur;
    cur += 8ull * n_patches;
    L.pval_off = cur;
    cur = pad8(cur + static_cast<uint64_t>(tb) * n_patches);
  }
  L.bp_off = cur;
  L.nulls_len = has_nulls ? (n + 7) / 8 : 0;
  L.values_len = width ? n_chunks * 128u * width : n * tb;  // entirely null: `n` zero elements (bit_pack_array.rs:43-50)
  L.nulls_off = cur + 16;
  L.values_off = cur + pad8(16ull + L.nulls_len);
  L.total = L.values_off + L.values_len;
  return L;
}
}  // namespace

namespace {
int str_entry_to_bytes(lc_ctx* ctx, const Entry* e, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  const StrHeader& h = e->sh;
  const StrImage L = str_image_of(h);
  if (L.total > 0xFFFFFFFFull) {
    set_error("to_bytes: image over 4 GiB");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  *out_bytes = L.total;
  if (!out) return LC_OK;
  if (cap < L.total) {
    set_error("to_bytes: buffer of %llu bytes, image has %llu", (unsigned long long)cap, (unsigned long long)L.total);
    return LC_ERR_INVALID;
  }
  std::memset(out, 0, L.total);
  put_u32(out, kLqdaMagic);
  put_u16(out + 4, 1);
  put_u16(out + 6, static_cast<uint16_t>(LC_LIQUID_BYTE_VIEW));
  put_u16(out + 8, h.arrow_type);  // ArrowByteType id (byte_view_array/mod.rs:113-122)
  put_u32(out + 16, L.keys_size);
  put_u32(out + 20, L.co_size);
  put_u32(out + 24, L.sp_size);
  put_u32(out + 28, L.fsst_raw_size);
  put_u32(out + 32, L.fp_size);
  put_u64(out + L.fsst_off, h.uncompressed_bytes);
  put_u32(out + L.fsst_off + 8, h.fsst_bytes);
  uint8_t* kb = out + L.keys_off;  // BitPackedArray<UInt16>::from_primitive(keys, 16)
  put_u32(kb, h.n);
  kb[4] = 16;
  kb[5] = h.has_nulls ? 1 : 0;
  put_u32(kb + 6, L.nulls_len);
  put_u32(kb + 10, L.keys_values_len);
  uint8_t* co = out + L.co_off;
  std::memcpy(co, &h.slope, 4);
  std::memcpy(co + 4, &h.intercept, 4);
  co[8] = h.offset_bytes;
  cudaStream_t s = ctx->L()->stream;
  const uint8_t* blob = e->d_blob;
  auto d2h = [&](uint64_t dst, uint32_t src_off, uint64_t bytes) -> cudaError_t {
    return bytes ? cudaMemcpyAsync(out + dst, blob + src_off, bytes, cudaMemcpyDeviceToHost, s) : cudaSuccess;
  };
  LC_CUDA_OK(d2h(L.fsst_off + 12, h.fsst_off, h.fsst_bytes));
  LC_CUDA_OK(d2h(L.keys_nulls_off, h.validity_off, L.nulls_len));
  LC_CUDA_OK(d2h(L.co_off + 9, h.resid_off, static_cast<uint64_t>(h.n_unique + 1) * h.offset_bytes));
  LC_CUDA_OK(d2h(L.pk_off, h.prefix_keys_off, 8ull * h.n_unique));
  LC_CUDA_OK(d2h(L.sp_off, h.shared_prefix_off, L.sp_size));
  LC_CUDA_OK(d2h(L.fp_off, h.fp_off, L.fp_size));
  if (h.n) {
    // keys: FastLanes transposition at W = 16 = what k_int_pack does for a 16-bit column of width 16 with reference 0
    Scratch& sc = ctx->L()->scratch;
    const uint64_t tmp_bytes = 64ull + L.keys_values_len;
    LC_TRY(sc.reserve(tmp_bytes + 1024, 1024));
    IntPackWork* h_pw = reinterpret_cast<IntPackWork*>(sc.host(256));
    uint8_t* d_pw = sc.dev(256);
    uint8_t* d_tmp = sc.dev(tmp_bytes);
    if (!h_pw || !d_pw || !d_tmp) {
      set_error("to_bytes: scratch exhausted");
      return LC_ERR_OOM;
    }
    std::memset(h_pw, 0, sizeof(*h_pw));
    h_pw->values = blob + h.keys_off;
    h_pw->validity = nullptr;  // null rows already hold key 0
    h_pw->blob = d_tmp;
    IntHeader& ih = h_pw->hdr;
    ih.magic = kMagicInt;
    ih.phys = PT_U16;
    ih.tbits = 16;
    ih.bit_width = 16;
    ih.n = h.n;
    ih.n_chunks = (h.n + 1023) / 1024;
    ih.packed_off = 64;
    ih.blob_bytes = static_cast<uint32_t>(tmp_bytes);
    LC_CUDA_OK(cudaMemcpyAsync(d_pw, h_pw, sizeof(IntPackWork), cudaMemcpyHostToDevice, s));
    LC_CUDA_OK(launch_int_pack(reinterpret_cast<const IntPackWork*>(d_pw), 1, s));
    ctx->kernel_launches++;
    LC_CUDA_OK(cudaMemcpyAsync(out + L.keys_values_off, d_tmp + 64, L.keys_values_len, cudaMemcpyDeviceToHost, s));
  }
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += L.total;
  return LC_OK;
}
}  // namespace

int symbol_table_to_bytes(const FsstCodec& c, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  const uint32_t n = c.dec.n_symbols;
  *out_bytes = 1ull + n + 8ull * n;
  if (!out) return LC_OK;
  if (cap < *out_bytes || n > 255) {
    set_error("symbol table: buffer too small");
    return LC_ERR_INVALID;
  }
  out[0] = static_cast<uint8_t>(n);
  for (uint32_t i = 0; i < n; ++i) out[1 + i] = c.dec.lens[i];
  for (uint32_t i = 0; i < n; ++i) put_u64(out + 1 + n + 8ull * i, c.dec.symbols[i]);
  return LC_OK;
}

int symbol_table_from_bytes(const uint8_t* b, uint64_t len, FsstCodec* out) {
  if (len < 1 || len < 1ull + b[0] + 8ull * b[0]) {
    set_error("symbol table: truncated");
    return LC_ERR_INVALID;
  }
  const uint32_t n = b[0];
  uint64_t vals[256];
  uint8_t lens[256];
  for (uint32_t i = 0; i < n; ++i) {
    lens[i] = b[1 + i];
    if (lens[i] < 1 || lens[i] > 8) {
      set_error("symbol table: symbol length %u", lens[i]);
      return LC_ERR_INVALID;
    }
    vals[i] = get_u64(b + 1 + n + 8ull * i);
  }
  fsst_from_symbols(vals, lens, n, out);
  return LC_OK;
}

int entry_to_bytes(lc_ctx* ctx, const Entry* e, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  if (e->liquid_type == LC_LIQUID_BYTE_VIEW) return str_entry_to_bytes(ctx, e, out, cap, out_bytes);
  const IntHeader& h = e->ih;
  const uint32_t tb = h.tbits / 8;
  const Layout L = layout_of(e->liquid_type, tb, h.n, h.bit_width, h.has_nulls != 0, h.n_chunks, h.n_patches);
  if (L.total > 0xFFFFFFFFull) {
    set_error("to_bytes: image over 4 GiB");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  *out_bytes = L.total;
  if (!out) return LC_OK;  // size query
  if (cap < L.total) {
    set_error("to_bytes: buffer of %llu bytes, image has %llu", (unsigned long long)cap, (unsigned long long)L.total);
    return LC_ERR_INVALID;
  }
  std::memset(out, 0, h.bit_width ? L.values_off : L.total);  // headers, paddings; the zero elements of an all-null array
  put_u32(out, kLqdaMagic);
  put_u16(out + 4, 1);
  put_u16(out + 6, static_cast<uint16_t>(e->liquid_type));
  put_u16(out + 8, h.phys);
  if (e->liquid_type == LC_LIQUID_DECIMAL) {
    int precision = 0, scale = 0, bw = 128;
    std::sscanf(e->arrow_format.c_str(), "d:%d,%d,%d", &precision, &scale, &bw);
    out[16] = bw == 256 ? 1 : 0;
    out[17] = static_cast<uint8_t>(precision);
    out[18] = static_cast<uint8_t>(static_cast<int8_t>(scale));
  }
  std::memcpy(out + L.ref_off, &h.reference, e->liquid_type == LC_LIQUID_DECIMAL ? 8 : tb);  // little-endian host
  cudaStream_t s = ctx->L()->stream;
  if (e->liquid_type == LC_LIQUID_FLOAT) {
    out[L.exp_off] = static_cast<uint8_t>(h.alp_ef & 0xffu);
    out[L.exp_off + 1] = static_cast<uint8_t>((h.alp_ef >> 8) & 0xffu);
    put_u64(out + L.exp_off + 8, h.n_patches);
    if (h.n_patches) {
      Scratch& sc = ctx->L()->scratch;
      LC_TRY(sc.reserve(8ull * h.n_patches + 1024, 1024));
      uint8_t* d_wide = sc.dev(8ull * h.n_patches);
      if (!d_wide) {
        set_error("to_bytes: scratch exhausted");
        return LC_ERR_OOM;
      }
      LC_CUDA_OK(launch_widen_u32(reinterpret_cast<const uint32_t*>(e->d_blob + h.patch_idx_off), h.n_patches,
                                  reinterpret_cast<unsigned long long*>(d_wide), s));
      ctx->kernel_launches++;
      LC_CUDA_OK(cudaMemcpyAsync(out + L.pidx_off, d_wide, 8ull * h.n_patches, cudaMemcpyDeviceToHost, s));
      LC_CUDA_OK(cudaMemcpyAsync(out + L.pval_off, e->d_blob + h.patch_val_off, static_cast<uint64_t>(tb) * h.n_patches,
                                 cudaMemcpyDeviceToHost, s));
    }
  }
  uint8_t* bp = out + L.bp_off;
  put_u32(bp, h.n);
  bp[4] = h.bit_width;
  bp[5] = (h.has_nulls || h.bit_width == 0) ? 1 : 0;  // new_null_array always carries a null buffer, also for 0 rows
  put_u32(bp + 6, L.nulls_len);
  put_u32(bp + 10, L.values_len);
  if (L.nulls_len) LC_CUDA_OK(cudaMemcpyAsync(out + L.nulls_off, e->d_blob + h.validity_off, L.nulls_len, cudaMemcpyDeviceToHost, s));
  if (h.bit_width && L.values_len)
    LC_CUDA_OK(cudaMemcpyAsync(out + L.values_off, e->d_blob + h.packed_off, L.values_len, cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += L.nulls_len + (h.bit_width ? L.values_len : 0) + (8ull + tb) * h.n_patches;
  return LC_OK;
}

// LiquidByteViewArray::from_bytes (serialization.rs:222-325). The image is checked section by section; dictionary keys must
// name existing dictionary values (the kernels index the dictionary with them).
static int str_entry_from_bytes(lc_ctx* ctx, const uint8_t* b, uint64_t len, const std::shared_ptr<FsstCodec>& codec, Entry** out) {
  if (!codec) {
    set_error("from_bytes: a byte-view image needs its column chunk's symbol table (lc_from_bytes_scoped)");
    return LC_ERR_INVALID;
  }
  StrImageIn in;
  if (const char* why = parse_str_image(b, len, &in)) {
    set_error("from_bytes: %s", why);
    return LC_ERR_INVALID;
  }
  const uint32_t bt = in.bt, n = in.n, U = in.n_unique, ob = in.offset_bytes, sp_size = in.sp_size, fp_size = in.fp_size, n_resid = in.n_resid,
                 comp_bytes = in.comp_bytes, kvals_len = in.kvals_len, n_chunks = (in.n + 1023) / 1024;
  const int32_t slope = in.slope, intercept = in.intercept;
  const bool file_nulls = in.file_nulls;
  const uint64_t uncompressed = in.uncompressed, comp_off = in.comp_off, knulls_off = in.knulls_off, kvals_off = in.kvals_off,
                 resid_src = in.resid_src, pk_src = in.pk_src, sp_src = in.sp_src, fp_src = in.fp_src;
  (void)n_chunks;
  // every key must name a dictionary value (W = 16: the packed words ARE the keys, transposed)
  {
    const uint64_t n_keys = kvals_len / 2;
    std::atomic<bool> bad{false};
    const uint32_t limit = U ? U : 1;  // an all-null batch holds key 0 everywhere
    parallel_for(n_keys, 1u << 16, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i)
        if (get_u16(b + kvals_off + 2 * i) >= limit) bad.store(true, std::memory_order_relaxed);
    });
    if (bad.load()) {
      set_error("from_bytes: a dictionary key is past the end of the dictionary");
      return LC_ERR_INVALID;
    }
  }
  const uint64_t n_valid = file_nulls ? popcount_bits(b + knulls_off, n) : n;
  const bool build_fp = fp_size != 0;
  StrHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicStr;
  h.arrow_type = static_cast<uint8_t>(bt);
  h.null_count = static_cast<uint32_t>(n - n_valid);
  h.has_nulls = h.null_count > 0;
  h.has_fp = build_fp;
  h.offset_bytes = static_cast<uint8_t>(ob);
  h.n = n;
  h.n_unique = U;
  h.slope = slope;
  h.intercept = intercept;
  h.shared_prefix_len = sp_size;
  h.uncompressed_bytes = uncompressed;
  h.table_ptr = reinterpret_cast<uint64_t>(codec->d_dec);
  uint64_t o = sizeof(StrHeader);
  h.shared_prefix_off = static_cast<uint32_t>(o);
  o += round_up(sp_size, 16);
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
  h.bloom_off = 0;  // the private substring filter is not part of LQDA: LIKE falls back to the reference gate alone
  h.fsst_off = static_cast<uint32_t>(o);
  h.fsst_bytes = comp_bytes;
  o += round_up(comp_bytes, 16) + 16;
  if (o > 0xFFFFFFF0ull) {
    set_error("from_bytes: entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(o);
  if (ctx->budget && ctx->arena_used() + o > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(), (unsigned long long)o,
              (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  Scratch& sc = ctx->L()->scratch;
  const uint64_t stage_bytes = round_up(sizeof(StrHeader) + 64 + round_up((n + 7) / 8, 16) + 64, 256);
  LC_TRY(sc.reserve(1024, stage_bytes + 1024));
  uint8_t* h_stage = sc.host(stage_bytes);
  if (!h_stage) {
    set_error("from_bytes: scratch exhausted");
    return LC_ERR_OOM;
  }
  uint32_t slab = 0, tslab = 0;
  uint8_t* d_blob = ctx->arena_alloc(o, &slab);
  // the keys come back through a temporary 16-bit integer entry of width 16: its packed chunks are the image's key words
  const uint64_t tmp_bytes = 64ull + kvals_len;
  uint8_t* d_tmp = n ? ctx->arena_alloc(tmp_bytes, &tslab) : nullptr;
  if (!d_blob || (n && !d_tmp)) {
    if (d_blob) ctx->arena_free(slab, d_blob, o);
    set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget" : "HBM arena: cudaMalloc failed");
    return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
  }
  cudaStream_t s = ctx->L()->stream;
  std::memset(h_stage, 0, stage_bytes);
  std::memcpy(h_stage, &h, sizeof(h));
  IntHeader ih;
  std::memset(&ih, 0, sizeof(ih));
  ih.magic = kMagicInt;
  ih.phys = PT_U16;
  ih.tbits = 16;
  ih.bit_width = 16;
  ih.n = n;
  ih.n_chunks = n_chunks;
  ih.packed_off = 64;
  ih.blob_bytes = static_cast<uint32_t>(tmp_bytes);
  std::memcpy(h_stage + sizeof(StrHeader), &ih, sizeof(ih));
  uint8_t* h_valid = h_stage + sizeof(StrHeader) + 64;
  if (h.has_nulls) {
    std::memcpy(h_valid, b + knulls_off, (n + 7) / 8);
    if (n & 7) h_valid[(n + 7) / 8 - 1] &= static_cast<uint8_t>((1u << (n & 7)) - 1u);
  }
  cudaError_t ce = cudaMemsetAsync(d_blob, 0, o, s);
  auto h2d = [&](uint8_t* dst, const void* src, uint64_t bytes) {
    if (bytes && ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
  };
  h2d(d_blob, h_stage, sizeof(StrHeader));
  h2d(d_blob + h.shared_prefix_off, b + sp_src, sp_size);
  if (build_fp) h2d(d_blob + h.fp_off, b + fp_src, 4ull * U);
  if (n_resid) h2d(d_blob + h.resid_off, b + resid_src, static_cast<uint64_t>(ob) * n_resid);  // none: the zeroed blob reads as offset 0
  h2d(d_blob + h.prefix_keys_off, b + pk_src, 8ull * U);
  if (h.has_nulls) h2d(d_blob + h.validity_off, h_valid, round_up((n + 7) / 8, 16));
  h2d(d_blob + h.fsst_off, b + comp_off, comp_bytes);
  if (n) {
    h2d(d_tmp, h_stage + sizeof(StrHeader), 64);
    h2d(d_tmp + 64, b + kvals_off, kvals_len);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  int rc = LC_OK;
  if (ce != cudaSuccess) {
    set_error("CUDA error in from_bytes: %s", cudaGetErrorString(ce));
    rc = LC_ERR_CUDA;
  }
  if (rc == LC_OK && n) {
    Entry* tmp = new Entry();
    tmp->liquid_type = LC_LIQUID_INTEGER;
    tmp->d_blob = d_tmp;
    tmp->blob_bytes = static_cast<uint32_t>(tmp_bytes);
    tmp->slab = tslab;
    tmp->n = n;
    tmp->arrow_format = "S";
    tmp->ih = ih;
    ctx->n_entries++;
    uint64_t rows = 0, vbytes = 0, nulls = 0;
    DeviceOut dout{d_blob + h.keys_off, 2ull * n, nullptr, nullptr, &rows, &vbytes, &nulls};
    Entry* list[1] = {tmp};
    rc = to_arrow_batch(ctx, list, 1, nullptr, nullptr, nullptr, nullptr, &dout);  // k_int_scan<DECODE>: keys in row order
    release_entry(ctx, tmp);  // frees the temporary blob and retires the cached entry list
    d_tmp = nullptr;
  }
  if (rc != LC_OK) {
    ctx->arena_free(slab, d_blob, o);
    if (d_tmp) ctx->arena_free(tslab, d_tmp, tmp_bytes);
    return rc;
  }
  ctx->h2d_bytes += o;
  static const char* kFmt[6] = {"u", "vu", "S", "S", "z", "vz"};
  Entry* e = new Entry();
  e->liquid_type = LC_LIQUID_BYTE_VIEW;
  e->d_blob = d_blob;
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->arrow_format = kFmt[bt];
  e->dict_value_format = bt == BT_DICT16_UTF8 ? "u" : bt == BT_DICT16_BINARY ? "z" : "";
  e->sh = h;
  e->shared_prefix.assign(b + sp_src, b + sp_src + sp_size);
  e->codec = codec;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

int entry_from_bytes(lc_ctx* ctx, const uint8_t* b, uint64_t len, const std::shared_ptr<FsstCodec>& codec, Entry** out) {
  if (len < 16 || get_u32(b) != kLqdaMagic || get_u16(b + 4) != 1) {
    set_error("from_bytes: not an LQDA version 1 image");
    return LC_ERR_INVALID;
  }
  const int32_t logical = get_u16(b + 6);
  const uint32_t phys = get_u16(b + 8);
  if (logical == LC_LIQUID_BYTE_VIEW) return str_entry_from_bytes(ctx, b, len, codec, out);
  if (logical != LC_LIQUID_INTEGER && logical != LC_LIQUID_FLOAT && logical != LC_LIQUID_DECIMAL) {
    set_error("from_bytes: logical type %d is not read by this build", logical);
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  const bool is_float = logical == LC_LIQUID_FLOAT, is_dec = logical == LC_LIQUID_DECIMAL;
  if (phys > 15 || (is_float != (phys == PT_F32 || phys == PT_F64)) || (is_dec && phys != PT_U64)) {
    set_error("from_bytes: physical type %u does not fit logical type %d", phys, logical);
    return LC_ERR_INVALID;
  }
  const uint32_t tbits = kPhys[phys].tbits, tb = tbits / 8;
  std::string format = kPhys[phys].fmt;
  uint32_t dec_width = 0, n_patches = 0, alp_ef = 0;
  uint64_t ref_off = 16, cur = pad8(16 + tb), pidx_off = 0, pval_off = 0;
  if (is_dec) {
    if (len < 32 || b[16] > 1) {
      set_error("from_bytes: bad decimal header");
      return LC_ERR_INVALID;
    }
    dec_width = b[16] ? 32 : 16;
    char buf[48];
    if (b[16]) std::snprintf(buf, sizeof(buf), "d:%d,%d,256", b[17], static_cast<int>(static_cast<int8_t>(b[18])));
    else std::snprintf(buf, sizeof(buf), "d:%d,%d", b[17], static_cast<int>(static_cast<int8_t>(b[18])));
    format = buf;
    ref_off = 24;
    cur = 32;
  }
  if (is_float) {
    if (len < cur + 16) {
      set_error("from_bytes: truncated float header");
      return LC_ERR_INVALID;
    }
    alp_ef = static_cast<uint32_t>(b[cur]) | (static_cast<uint32_t>(b[cur + 1]) << 8);
    const uint64_t pc = get_u64(b + cur + 8);
    cur += 16;
    if (pc > 0x7fffffffull || len < cur + pc * (8ull + tb)) {
      set_error("from_bytes: truncated patch list");
      return LC_ERR_INVALID;
    }
    if (b[cur - 16] >= (tbits == 64 ? 24u : 11u) || b[cur - 15] >= (tbits == 64 ? 24u : 11u)) {
      set_error("from_bytes: ALP exponents out of range");
      return LC_ERR_INVALID;
    }
    n_patches = static_cast<uint32_t>(pc);
    pidx_off = cur;
    pval_off = cur + 8ull * n_patches;
    cur = pad8(pval_off + static_cast<uint64_t>(tb) * n_patches);
  }
  if (len < cur + 16) {
    set_error("from_bytes: truncated bit-packed header");
    return LC_ERR_INVALID;
  }
  const uint8_t* bp = b + cur;
  const uint32_t n = get_u32(bp);
  uint32_t width = bp[4];
  const bool file_nulls = bp[5] != 0;
  const uint32_t nulls_len = get_u32(bp + 6), values_len = get_u32(bp + 10);
  const uint64_t nulls_off = cur + 16, values_off = cur + pad8(16ull + (file_nulls ? nulls_len : 0));
  if (n > 0x7fffffffu || len < values_off + values_len || (file_nulls && nulls_len < (n + 7) / 8)) {
    set_error("from_bytes: sections run past the end of the image");
    return LC_ERR_INVALID;
  }
  uint64_t n_valid = n;
  if (file_nulls) n_valid = popcount_bits(b + nulls_off, n);
  const uint32_t n_chunks = (n + 1023) / 1024;
  if (values_len == 0 || n_valid == 0) {
    width = 0;  // BitPackedArray::from_bytes returns new_null_array for both (bit_pack_array.rs:274-277, 319-321)
    n_valid = 0;
  } else if (width == 0 || width > tbits || values_len != n_chunks * 128u * width) {
    set_error("from_bytes: bit width %u / values length %u do not fit %u rows", width, values_len, n);
    return LC_ERR_INVALID;
  }

  IntHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicInt;
  h.phys = static_cast<uint8_t>(phys);
  h.tbits = static_cast<uint8_t>(tbits);
  h.bit_width = static_cast<uint8_t>(width);
  h.n = n;
  h.n_chunks = n_chunks;
  h.is_signed = kPhys[phys].sg;
  h.null_count = static_cast<uint32_t>(n - n_valid);
  h.has_nulls = h.null_count != 0;
  if (width) std::memcpy(&h.reference, b + ref_off, is_dec ? 8 : tb);
  h.alp_ef = width ? alp_ef : 0;
  if (!width) n_patches = 0;
  const uint64_t valid_bytes = h.has_nulls ? round_up((n + 7) / 8, 16) : 0;
  h.validity_off = h.has_nulls ? 64 : 0;
  h.packed_off = static_cast<uint32_t>(64 + valid_bytes);
  uint64_t blob_bytes = round_up(h.packed_off + static_cast<uint64_t>(n_chunks) * 128ull * width, 16);
  if (n_patches) {
    h.n_patches = n_patches;
    h.patch_idx_off = static_cast<uint32_t>(blob_bytes);
    blob_bytes = round_up(blob_bytes + 4ull * n_patches, 16);
    h.patch_val_off = static_cast<uint32_t>(blob_bytes);
    blob_bytes = round_up(blob_bytes + static_cast<uint64_t>(tb) * n_patches, 16);
  }
  if (blob_bytes > 0xFFFFFFF0ull) {
    set_error("from_bytes: entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(blob_bytes);
  if (ctx->budget && ctx->arena_used() + blob_bytes > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(),
              (unsigned long long)blob_bytes, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  Scratch& sc = ctx->L()->scratch;
  const uint64_t head_bytes = 64 + valid_bytes;
  LC_TRY(sc.reserve(8ull * n_patches + 1024, head_bytes + 1024));
  uint8_t* h_head = sc.host(head_bytes);
  uint32_t* h_flag = reinterpret_cast<uint32_t*>(sc.host(64));
  uint8_t* d_idx64 = sc.dev(8ull * n_patches + 16);
  uint8_t* d_flag = sc.dev(64);
  if (!h_head || !h_flag || !d_idx64 || !d_flag) {
    set_error("from_bytes: scratch exhausted");
    return LC_ERR_OOM;
  }
  uint32_t slab = 0;
  uint8_t* d_blob = ctx->arena_alloc(blob_bytes, &slab);
  if (!d_blob) {
    set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget for %llu bytes" : "HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)blob_bytes);
    return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
  }
  std::memset(h_head, 0, head_bytes);
  std::memcpy(h_head, &h, sizeof(h));
  if (h.has_nulls) {
    std::memcpy(h_head + 64, b + nulls_off, (n + 7) / 8);
    if (n & 7) h_head[64 + (n + 7) / 8 - 1] &= static_cast<uint8_t>((1u << (n & 7)) - 1u);  // bits past n stay zero in the entry
  }
  cudaStream_t s = ctx->L()->stream;
  cudaError_t ce = cudaMemcpyAsync(d_blob, h_head, head_bytes, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess && width) ce = cudaMemcpyAsync(d_blob + h.packed_off, b + values_off, values_len, cudaMemcpyHostToDevice, s);
  *h_flag = 0;
  if (ce == cudaSuccess && n_patches) {
    ce = cudaMemsetAsync(d_flag, 0, 4, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_idx64, b + pidx_off, 8ull * n_patches, cudaMemcpyHostToDevice, s);
    if (ce == cudaSuccess)
      ce = launch_narrow_u64(reinterpret_cast<const unsigned long long*>(d_idx64), n_patches, n,
                             reinterpret_cast<uint32_t*>(d_blob + h.patch_idx_off), reinterpret_cast<uint32_t*>(d_flag), s);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(d_blob + h.patch_val_off, b + pval_off, static_cast<uint64_t>(tb) * n_patches, cudaMemcpyHostToDevice, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_flag, d_flag, 4, cudaMemcpyDeviceToHost, s);
    ctx->kernel_launches++;
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess || *h_flag) {
    ctx->arena_free(slab, d_blob, blob_bytes);
    if (ce != cudaSuccess) {
      set_error("CUDA error in from_bytes: %s", cudaGetErrorString(ce));
      return LC_ERR_CUDA;
    }
    set_error("from_bytes: a patch index is past the end of the array");
    return LC_ERR_INVALID;
  }
  ctx->h2d_bytes += head_bytes + (width ? values_len : 0) + (8ull + tb) * n_patches;
  Entry* e = new Entry();
  e->liquid_type = logical;
  e->d_blob = d_blob;
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->dec_width = dec_width;
  e->arrow_format = format;
  e->ih = h;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

}  // namespace lc
