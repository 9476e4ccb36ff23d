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
// kernels are the u32 <-> u64 conversion of ALP patch indices. Byte-view LQDA (byte_view_array/serialization.rs) is not built.
#include "host_common.h"

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

int entry_to_bytes(lc_ctx* ctx, const Entry* e, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  if (!is_int_blob(e->liquid_type)) {
    set_error("to_bytes: byte-view entries are not serialized by this build");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
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
  cudaStream_t s = ctx->stream;
  if (e->liquid_type == LC_LIQUID_FLOAT) {
    out[L.exp_off] = static_cast<uint8_t>(h.alp_ef & 0xffu);
    out[L.exp_off + 1] = static_cast<uint8_t>((h.alp_ef >> 8) & 0xffu);
    put_u64(out + L.exp_off + 8, h.n_patches);
    if (h.n_patches) {
      Scratch& sc = ctx->scratch;
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

int entry_from_bytes(lc_ctx* ctx, const uint8_t* b, uint64_t len, Entry** out) {
  if (len < 16 || get_u32(b) != kLqdaMagic || get_u16(b + 4) != 1) {
    set_error("from_bytes: not an LQDA version 1 image");
    return LC_ERR_INVALID;
  }
  const int32_t logical = get_u16(b + 6);
  const uint32_t phys = get_u16(b + 8);
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
  if (ctx->budget && ctx->arena.bytes_used() + blob_bytes > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena.bytes_used(),
              (unsigned long long)blob_bytes, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  Scratch& sc = ctx->scratch;
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
  uint8_t* d_blob = ctx->arena.alloc(blob_bytes, &slab);
  if (!d_blob) {
    set_error("HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)blob_bytes);
    return LC_ERR_OOM;
  }
  std::memset(h_head, 0, head_bytes);
  std::memcpy(h_head, &h, sizeof(h));
  if (h.has_nulls) {
    std::memcpy(h_head + 64, b + nulls_off, (n + 7) / 8);
    if (n & 7) h_head[64 + (n + 7) / 8 - 1] &= static_cast<uint8_t>((1u << (n & 7)) - 1u);  // bits past n stay zero in the entry
  }
  cudaStream_t s = ctx->stream;
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
    ctx->arena.free(slab, blob_bytes);
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
