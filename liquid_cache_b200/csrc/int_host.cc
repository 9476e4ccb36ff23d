// int_host.cc — host orchestration of the integer encode (insert side) and predicate planning.
// Reference: LiquidPrimitiveArray::from_arrow_array (/root/reference/src/core/src/liquid_array/
// primitive_array.rs:159-206), get_bit_width (src/core/src/utils/mod.rs:24-32).
#include "host_common.h"

namespace lc {

static uint32_t bit_width_of(uint64_t max_value) {
  // get_bit_width: 1 when the range is 0, else 64 - leading_zeros
  if (max_value == 0) return 1;
  return 64u - static_cast<uint32_t>(__builtin_clzll(max_value));
}

int int_encode(lc_ctx* ctx, const ArrowIn& in, Entry** out) {
  const uint32_t n = static_cast<uint32_t>(in.length);
  const uint32_t tb = in.tbits / 8;
  const uint64_t val_bytes = static_cast<uint64_t>(n) * tb;
  const uint32_t n_words = (n + 31) / 32;
  const bool has_nulls = in.null_count > 0;
  Scratch& sc = ctx->scratch;
  LC_TRY(sc.reserve(val_bytes + n_words * 4ull + 4096, val_bytes + n_words * 4ull + 4096));

  // ---- stage values (+ validity re-aligned to bit offset 0), one H2D ----
  uint8_t* h_vals = sc.host(round_up(val_bytes, 256));
  uint8_t* h_valid = sc.host(round_up(n_words * 4ull, 256) + 256);
  IntMinMaxWork* h_mm = reinterpret_cast<IntMinMaxWork*>(sc.host(256));
  IntPackWork* h_pw = reinterpret_cast<IntPackWork*>(sc.host(256));
  uint64_t* h_mmout = reinterpret_cast<uint64_t*>(sc.host(256));
  uint8_t* d_vals = sc.dev(round_up(val_bytes, 256));
  uint8_t* d_valid = sc.dev(round_up(n_words * 4ull, 256) + 256);
  uint8_t* d_mm = sc.dev(256);
  uint8_t* d_pw = sc.dev(256);
  uint8_t* d_mmout = sc.dev(256);
  if (!h_vals || !h_valid || !h_mm || !h_pw || !h_mmout || !d_vals || !d_valid || !d_mm || !d_pw || !d_mmout) {
    set_error("int_encode: scratch exhausted");
    return LC_ERR_OOM;
  }
  if (n) std::memcpy(h_vals, static_cast<const uint8_t*>(in.values) + static_cast<uint64_t>(in.offset) * tb, val_bytes);
  if (has_nulls) copy_bits(in.validity, in.offset, n, h_valid, n_words * 4ull);
  h_mm->values = d_vals;
  h_mm->validity = has_nulls ? reinterpret_cast<const uint32_t*>(d_valid) : nullptr;
  h_mm->out = reinterpret_cast<uint64_t*>(d_mmout);
  h_mm->n = n;
  h_mm->phys = in.phys;
  cudaStream_t s = ctx->stream;
  // the staging areas are contiguous in both scratch spaces: one copy covers values, validity, work
  const uint64_t up_bytes = static_cast<uint64_t>(reinterpret_cast<uint8_t*>(h_mm) + 256 - h_vals);
  LC_CUDA_OK(cudaMemcpyAsync(d_vals, h_vals, up_bytes, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_bytes;
  (void)d_mm;

  // ---- pass 1: min / max / valid count ----
  LC_CUDA_OK(launch_int_minmax(reinterpret_cast<const IntMinMaxWork*>(d_mm), 1, s));
  ctx->kernel_launches++;
  LC_CUDA_OK(cudaMemcpyAsync(h_mmout, d_mmout, 24, cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += 24;
  const uint64_t mn = h_mmout[0], mx = h_mmout[1], n_valid = h_mmout[2];

  IntHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicInt;
  h.phys = in.phys;
  h.tbits = in.tbits;
  h.n = n;
  h.n_chunks = (n + 1023) / 1024;
  h.is_signed = in.is_signed;
  h.has_nulls = has_nulls;
  h.null_count = static_cast<uint32_t>(n - n_valid);
  const uint64_t tmask = in.tbits == 64 ? ~0ull : ((1ull << in.tbits) - 1ull);
  if (n_valid == 0) {
    // entire array null (or empty): BitPackedArray::new_null_array, reference_value = 0
    h.bit_width = 0;
    h.reference = 0;
    h.has_nulls = n > 0;
    h.null_count = n;
  } else {
    const uint64_t sub = (mx - mn) & tmask;  // max.sub_wrapping(min) reinterpreted unsigned
    h.bit_width = static_cast<uint8_t>(bit_width_of(sub));
    h.reference = mn & tmask;
  }
  const uint64_t valid_bytes = h.has_nulls ? round_up((n + 7) / 8, 16) : 0;
  h.validity_off = h.has_nulls ? 64 : 0;
  h.packed_off = static_cast<uint32_t>(64 + valid_bytes);
  const uint64_t packed_bytes = static_cast<uint64_t>(h.n_chunks) * 128ull * h.bit_width;
  const uint64_t blob_bytes = round_up(h.packed_off + packed_bytes, 16);
  if (blob_bytes > 0xFFFFFFF0ull) {
    set_error("int_encode: entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(blob_bytes);
  if (ctx->budget && ctx->arena.bytes_used() + blob_bytes > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena.bytes_used(),
              (unsigned long long)blob_bytes, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  uint32_t slab = 0;
  uint8_t* d_blob = ctx->arena.alloc(blob_bytes, &slab);
  if (!d_blob) {
    set_error("HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)blob_bytes);
    return LC_ERR_OOM;
  }

  // ---- pass 2: subtract reference + FastLanes pack (+ header, validity) ----
  std::memset(h_pw, 0, sizeof(*h_pw));
  h_pw->values = d_vals;
  // an all-null array that came without a validity buffer cannot happen (n_valid==0 implies nulls)
  h_pw->validity = has_nulls ? reinterpret_cast<const uint32_t*>(d_valid) : nullptr;
  h_pw->blob = d_blob;
  h_pw->hdr = h;
  LC_CUDA_OK(cudaMemcpyAsync(d_pw, h_pw, sizeof(IntPackWork), cudaMemcpyHostToDevice, s));
  LC_CUDA_OK(launch_int_pack(reinterpret_cast<const IntPackWork*>(d_pw), 1, s));
  ctx->kernel_launches++;
  LC_CUDA_OK(cudaStreamSynchronize(s));

  Entry* e = new Entry();
  e->liquid_type = LC_LIQUID_INTEGER;
  e->d_blob = d_blob;
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->arrow_format = in.format;
  e->ih = h;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

}  // namespace lc
