// int_host.cc — host orchestration of the encode (insert side) of every integer-shaped entry: integers, ALP floats,
// u64 decimals. The host only sizes the blob from a few scalars read back; every array-sized step is a kernel.
// Reference: LiquidPrimitiveArray::from_arrow_array (/root/reference/src/core/src/liquid_array/
// primitive_array.rs:159-206), get_bit_width (src/core/src/utils/mod.rs:24-32),
// LiquidFloatArray::from_arrow_array (float_array.rs:266-269, 609-751),
// LiquidDecimalArray::{fits_u64, from_decimal_array} (decimal_array.rs:127-178).
#include "host_common.h"
#include "host_pool.h"

namespace lc {

static uint32_t bit_width_of(uint64_t max_value) {
  // get_bit_width: 1 when the range is 0, else 64 - leading_zeros
  if (max_value == 0) return 1;
  return 64u - static_cast<uint32_t>(__builtin_clzll(max_value));
}

int int_encode(lc_ctx* ctx, const ArrowIn& in, Entry** out) {
  const bool is_float = in.kind == ArrowIn::K_FLOAT;
  const bool is_dec = in.kind == ArrowIn::K_DECIMAL;
  const uint32_t n = static_cast<uint32_t>(in.length);
  const uint32_t tb = in.tbits / 8;                       // bytes of the packed integer type
  const uint32_t raw_tb = is_dec ? in.dec_width : tb;     // bytes per Arrow value
  const uint64_t raw_bytes = static_cast<uint64_t>(n) * raw_tb;
  const uint64_t val_bytes = static_cast<uint64_t>(n) * tb;
  const uint32_t n_words = (n + 31) / 32;
  const bool has_nulls = in.null_count > 0;
  Scratch& sc = ctx->L()->scratch;
  const uint64_t extra_dev = is_dec ? round_up(val_bytes, 256) + 256
                             : is_float ? 2 * round_up(val_bytes, 256) + round_up(n_words * 4ull, 256) + round_up(n * 4ull, 256) + 4096
                                        : 0;
  LC_TRY(sc.reserve(raw_bytes + n_words * 4ull + 4096 + extra_dev, raw_bytes + n_words * 4ull + 4096));

  // ---- stage values (+ validity re-aligned to bit offset 0), one H2D ----
  uint8_t* h_vals = sc.host(round_up(raw_bytes, 256));
  uint8_t* h_valid = sc.host(round_up(n_words * 4ull, 256) + 256);
  IntMinMaxWork* h_mm = reinterpret_cast<IntMinMaxWork*>(sc.host(256));
  IntPackWork* h_pw = reinterpret_cast<IntPackWork*>(sc.host(256));
  uint64_t* h_mmout = reinterpret_cast<uint64_t*>(sc.host(256));
  uint8_t* d_vals = sc.dev(round_up(raw_bytes, 256));
  uint8_t* d_valid = sc.dev(round_up(n_words * 4ull, 256) + 256);
  uint8_t* d_mm = sc.dev(256);
  uint8_t* d_pw = sc.dev(256);
  uint8_t* d_mmout = sc.dev(256);
  if (!h_vals || !h_valid || !h_mm || !h_pw || !h_mmout || !d_vals || !d_valid || !d_mm || !d_pw || !d_mmout) {
    set_error("int_encode: scratch exhausted");
    return LC_ERR_OOM;
  }
  // device-only work areas of the float / decimal flavours
  uint8_t *d_pack_src = d_vals, *d_exc = nullptr, *d_pidx = nullptr, *d_pval = nullptr, *d_sizes = nullptr;
  if (is_dec) {
    d_pack_src = sc.dev(round_up(val_bytes, 256) + 256);
    if (!d_pack_src) {
      set_error("int_encode: scratch exhausted");
      return LC_ERR_OOM;
    }
  } else if (is_float) {
    d_pack_src = sc.dev(round_up(val_bytes, 256));        // ALP-encoded integers
    d_pval = sc.dev(round_up(val_bytes, 256));            // patch values (at most n)
    d_exc = sc.dev(round_up(n_words * 4ull, 256));
    d_pidx = sc.dev(round_up(n * 4ull, 256));
    d_sizes = sc.dev(2048);                               // one u64 per (e, f) pair (<= 153), then the result struct
    if (!d_pack_src || !d_pval || !d_exc || !d_pidx || !d_sizes) {
      set_error("int_encode: scratch exhausted");
      return LC_ERR_OOM;
    }
  }
  if (n) std::memcpy(h_vals, static_cast<const uint8_t*>(in.values) + static_cast<uint64_t>(in.offset) * raw_tb, raw_bytes);
  if (has_nulls) copy_bits(in.validity, in.offset, n, h_valid, n_words * 4ull);
  const uint32_t* d_valid_w = has_nulls ? reinterpret_cast<const uint32_t*>(d_valid) : nullptr;
  h_mm->values = d_pack_src;
  h_mm->validity = d_valid_w;
  h_mm->out = reinterpret_cast<uint64_t*>(d_mmout);
  h_mm->n = n;
  h_mm->phys = in.phys;
  cudaStream_t s = ctx->L()->stream;
  // the staging areas are contiguous in both scratch spaces: one copy covers values, validity, work
  const uint64_t up_bytes = static_cast<uint64_t>(reinterpret_cast<uint8_t*>(h_mm) + 256 - h_vals);
  LC_CUDA_OK(cudaMemcpyAsync(d_vals, h_vals, up_bytes, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_bytes;

  IntHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kMagicInt;
  h.phys = in.phys;
  h.tbits = in.tbits;
  h.n = n;
  h.n_chunks = (n + 1023) / 1024;
  h.is_signed = in.is_signed;
  h.has_nulls = has_nulls;
  const uint64_t tmask = in.tbits == 64 ? ~0ull : ((1ull << in.tbits) - 1ull);
  uint64_t mn = 0, mx = 0, n_valid = 0;
  uint32_t n_patches = 0;

  if (is_float) {
    // ---- ALP: exponent search on the sample, encode, patch list, min/max — three launches, one 40-byte D2H ----
    n_valid = n - static_cast<uint64_t>(in.null_count);
    if (n_valid) {  // an all-null (or empty) array never looks at its values (float_array.rs:620-630)
      AlpEncIo io{};
      io.values = d_vals;
      io.validity = d_valid_w;
      io.n = n;
      io.is_f64 = in.tbits == 64;
      io.sample_step = n > 1024 ? n / 1024 : 0;  // NUM_SAMPLES (float_array.rs:58, 719-727)
      io.sample_cnt = io.sample_step ? (n + io.sample_step - 1) / io.sample_step : n;
      io.sizes = reinterpret_cast<unsigned long long*>(d_sizes);
      io.res = reinterpret_cast<AlpEncResult*>(d_sizes + 1536);
      io.enc = d_pack_src;
      io.exc_words = reinterpret_cast<uint32_t*>(d_exc);
      io.patch_idx = reinterpret_cast<uint32_t*>(d_pidx);
      io.patch_val = d_pval;
      LC_CUDA_OK(launch_alp_encode(io, s));
      ctx->kernel_launches += 3;
      AlpEncResult* h_res = reinterpret_cast<AlpEncResult*>(h_mmout);
      LC_CUDA_OK(cudaMemcpyAsync(h_res, io.res, sizeof(AlpEncResult), cudaMemcpyDeviceToHost, s));
      LC_CUDA_OK(cudaStreamSynchronize(s));
      ctx->d2h_bytes += sizeof(AlpEncResult);
      mn = static_cast<uint64_t>(h_res->min);
      mx = static_cast<uint64_t>(h_res->max);
      n_patches = h_res->n_patches;
      h.alp_ef = (h_res->e & 0xffu) | ((h_res->f & 0xffu) << 8);
    }
  } else {
    if (is_dec) {
      // ---- Decimal128/256 -> low u64 words; a valid value outside u64 turns the whole array down ----
      LC_CUDA_OK(cudaMemsetAsync(d_mmout, 0, 32, s));
      LC_CUDA_OK(launch_dec_narrow(d_vals, d_valid_w, n, in.dec_width, reinterpret_cast<unsigned long long*>(d_pack_src),
                                   reinterpret_cast<uint32_t*>(d_mmout + 24), s));
      ctx->kernel_launches++;
    }
    // ---- pass 1: min / max / valid count ----
    LC_CUDA_OK(launch_int_minmax(reinterpret_cast<const IntMinMaxWork*>(d_mm), 1, s));
    ctx->kernel_launches++;
    LC_CUDA_OK(cudaMemcpyAsync(h_mmout, d_mmout, 32, cudaMemcpyDeviceToHost, s));
    LC_CUDA_OK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += 32;
    mn = h_mmout[0];
    mx = h_mmout[1];
    n_valid = h_mmout[2];
    if (is_dec && (h_mmout[3] & 0xffffffffull)) {
      // the reference stores such arrays as LiquidFixedLenByteArray (dictionary + FSST over the 16/32-byte values,
      // transcode.rs:118-131): the byte-view insert path takes over (lc_abi.cc encode_locked)
      return LC_INTERNAL_FIXED_LEN;
    }
  }

  h.null_count = static_cast<uint32_t>(n - n_valid);
  if (n_valid == 0) {
    // entire array null (or empty): BitPackedArray::new_null_array, reference_value = 0
    h.bit_width = 0;
    h.reference = 0;
    h.has_nulls = n > 0;
    h.null_count = n;
    h.alp_ef = 0;
  } else {
    const uint64_t sub = (mx - mn) & tmask;  // max.sub_wrapping(min) reinterpreted unsigned
    h.bit_width = static_cast<uint8_t>(bit_width_of(sub));
    h.reference = mn & tmask;
  }
  const uint64_t valid_bytes = h.has_nulls ? round_up((n + 7) / 8, 16) : 0;
  h.validity_off = h.has_nulls ? 64 : 0;
  h.packed_off = static_cast<uint32_t>(64 + valid_bytes);
  const uint64_t packed_bytes = static_cast<uint64_t>(h.n_chunks) * 128ull * h.bit_width;
  uint64_t blob_bytes = round_up(h.packed_off + packed_bytes, 16);
  if (n_patches) {
    h.n_patches = n_patches;
    h.patch_idx_off = static_cast<uint32_t>(blob_bytes);
    blob_bytes = round_up(blob_bytes + 4ull * n_patches, 16);
    h.patch_val_off = static_cast<uint32_t>(blob_bytes);
    blob_bytes = round_up(blob_bytes + static_cast<uint64_t>(tb) * n_patches, 16);
  }
  if (blob_bytes > 0xFFFFFFF0ull) {
    set_error("int_encode: entry too large");
    return LC_ERR_UNSUPPORTED_TYPE;
  }
  h.blob_bytes = static_cast<uint32_t>(blob_bytes);
  if (ctx->budget && ctx->arena_used() + blob_bytes > ctx->budget) {
    set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(),
              (unsigned long long)blob_bytes, (unsigned long long)ctx->budget);
    return LC_ERR_CACHE_FULL;
  }
  ArenaBlock block(ctx, blob_bytes);  // handed back on every early return below
  uint8_t* d_blob = block.p;
  const uint32_t slab = block.slab;
  if (!d_blob) {
    set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget for %llu bytes" : "HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)blob_bytes);
    return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
  }

  // ---- pass 2: subtract reference + FastLanes pack (+ header, validity) ----
  std::memset(h_pw, 0, sizeof(*h_pw));
  h_pw->values = d_pack_src;
  // an all-null array that came without a validity buffer cannot happen (n_valid==0 implies nulls)
  h_pw->validity = d_valid_w;
  h_pw->blob = d_blob;
  h_pw->pack_null_slots = is_float ? 1 : 0;
  h_pw->hdr = h;
  LC_CUDA_OK(cudaMemcpyAsync(d_pw, h_pw, sizeof(IntPackWork), cudaMemcpyHostToDevice, s));
  LC_CUDA_OK(launch_int_pack(reinterpret_cast<const IntPackWork*>(d_pw), 1, s));
  ctx->kernel_launches++;
  if (n_patches) {
    LC_CUDA_OK(cudaMemcpyAsync(d_blob + h.patch_idx_off, d_pidx, 4ull * n_patches, cudaMemcpyDeviceToDevice, s));
    LC_CUDA_OK(cudaMemcpyAsync(d_blob + h.patch_val_off, d_pval, static_cast<uint64_t>(tb) * n_patches, cudaMemcpyDeviceToDevice, s));
  }
  LC_CUDA_OK(cudaStreamSynchronize(s));

  Entry* e = new Entry();
  e->liquid_type = is_float ? LC_LIQUID_FLOAT : is_dec ? LC_LIQUID_DECIMAL : LC_LIQUID_INTEGER;
  e->d_blob = block.release();
  e->blob_bytes = h.blob_bytes;
  e->slab = slab;
  e->n = n;
  e->dec_width = is_dec ? in.dec_width : 0;
  e->arrow_format = in.format;
  e->ih = h;
  ctx->n_entries++;
  *out = e;
  return LC_OK;
}

// ---- batched form -----------------------------------------------------------------------------------
// Many integer-like batches (any mix of the K_INT types) in ONE pass: every batch is copied once into the pinned
// staging area (split over the host pool), one H2D carries values + validity + the min/max work list, k_int_minmax
// and k_int_pack run with one CTA per batch (they take work LISTS), and the host sizes all blobs from one D2H of
// 32 bytes per batch. Two stream synchronisations per call instead of two per batch.
int int_encode_many(lc_ctx* ctx, const std::vector<ArrowIn>& ins, std::vector<Entry*>* out) {
  const uint64_t nb = ins.size();
  out->clear();
  if (nb == 0) return LC_OK;
  std::vector<uint64_t> voff(nb), moff(nb, ~0ull);
  uint64_t cur = 0;
  for (uint64_t i = 0; i < nb; ++i) {
    const ArrowIn& in = ins[i];
    if (in.kind != ArrowIn::K_INT) {
      set_error("int_encode_many: batch %llu is not an integer-like array", (unsigned long long)i);
      return LC_ERR_INVALID;
    }
    const uint64_t n = static_cast<uint64_t>(in.length);
    voff[i] = cur;
    cur += round_up(n * (in.tbits / 8), 256) + 256;
    if (in.null_count > 0) {
      moff[i] = cur;
      cur += round_up(((n + 31) / 32) * 4, 256) + 256;
    }
  }
  const uint64_t mm_off = cur;
  cur += round_up(nb * sizeof(IntMinMaxWork), 256);
  const uint64_t up_bytes = cur;
  const uint64_t pw_off = cur;
  cur += round_up(nb * sizeof(IntPackWork), 256);
  const uint64_t mmout_off = cur;
  cur += round_up(nb * 32, 256);
  const uint64_t total = cur;
  Scratch& sc = ctx->L()->scratch;
  LC_TRY(sc.reserve(total + 1024, total + 1024));
  uint8_t* h = sc.host(total);
  uint8_t* d = sc.dev(total);
  if (!h || !d) {
    set_error("int_encode_many: scratch exhausted");
    return LC_ERR_OOM;
  }
  IntMinMaxWork* h_mm = reinterpret_cast<IntMinMaxWork*>(h + mm_off);
  parallel_for(nb, 16, [&](uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      const ArrowIn& in = ins[i];
      const uint64_t n = static_cast<uint64_t>(in.length);
      const uint32_t tb = in.tbits / 8;
      if (n) std::memcpy(h + voff[i], static_cast<const uint8_t*>(in.values) + static_cast<uint64_t>(in.offset) * tb, n * tb);
      const bool has_nulls = moff[i] != ~0ull;
      if (has_nulls) copy_bits(in.validity, in.offset, static_cast<int64_t>(n), h + moff[i], ((n + 31) / 32) * 4);
      h_mm[i].values = d + voff[i];
      h_mm[i].validity = has_nulls ? reinterpret_cast<const uint32_t*>(d + moff[i]) : nullptr;
      h_mm[i].out = reinterpret_cast<uint64_t*>(d + mmout_off) + 4 * i;
      h_mm[i].n = static_cast<uint32_t>(n);
      h_mm[i].phys = in.phys;
    }
  });
  cudaStream_t s = ctx->L()->stream;
  LC_CUDA_OK(cudaMemcpyAsync(d, h, up_bytes, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_bytes;
  LC_CUDA_OK(launch_int_minmax(reinterpret_cast<const IntMinMaxWork*>(d + mm_off), static_cast<uint32_t>(nb), s));
  ctx->kernel_launches++;
  LC_CUDA_OK(cudaMemcpyAsync(h + mmout_off, d + mmout_off, nb * 32, cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += nb * 32;

  // ---- size every blob, take arena space, write the pack work list ----
  const uint64_t* h_mmout = reinterpret_cast<const uint64_t*>(h + mmout_off);
  IntPackWork* h_pw = reinterpret_cast<IntPackWork*>(h + pw_off);
  struct Taken {
    uint8_t* blob;
    uint32_t slab;
    uint64_t bytes;
  };
  std::vector<Taken> taken;
  taken.reserve(nb);
  auto give_back = [&]() {
    for (const Taken& t : taken) ctx->arena_free(t.slab, t.blob, t.bytes);
  };
  for (uint64_t i = 0; i < nb; ++i) {
    const ArrowIn& in = ins[i];
    const uint32_t n = static_cast<uint32_t>(in.length);
    const uint64_t mn = h_mmout[4 * i], mx = h_mmout[4 * i + 1], n_valid = h_mmout[4 * i + 2];
    const bool has_nulls = moff[i] != ~0ull;
    IntHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    hd.magic = kMagicInt;
    hd.phys = in.phys;
    hd.tbits = in.tbits;
    hd.n = n;
    hd.n_chunks = (n + 1023) / 1024;
    hd.is_signed = in.is_signed;
    hd.has_nulls = has_nulls;
    hd.null_count = static_cast<uint32_t>(n - n_valid);
    const uint64_t tmask = in.tbits == 64 ? ~0ull : ((1ull << in.tbits) - 1ull);
    if (n_valid == 0) {  // entire array null (or empty): BitPackedArray::new_null_array, reference_value = 0
      hd.bit_width = 0;
      hd.reference = 0;
      hd.has_nulls = n > 0;
      hd.null_count = n;
    } else {
      hd.bit_width = static_cast<uint8_t>(bit_width_of((mx - mn) & tmask));
      hd.reference = mn & tmask;
    }
    const uint64_t valid_bytes = hd.has_nulls ? round_up((n + 7) / 8, 16) : 0;
    hd.validity_off = hd.has_nulls ? 64 : 0;
    hd.packed_off = static_cast<uint32_t>(64 + valid_bytes);
    const uint64_t blob_bytes = round_up(hd.packed_off + static_cast<uint64_t>(hd.n_chunks) * 128ull * hd.bit_width, 16);
    if (blob_bytes > 0xFFFFFFF0ull) {
      give_back();
      set_error("int_encode_many: entry too large");
      return LC_ERR_UNSUPPORTED_TYPE;
    }
    hd.blob_bytes = static_cast<uint32_t>(blob_bytes);
    if (ctx->budget && ctx->arena_used() + blob_bytes > ctx->budget) {
      give_back();
      set_error("cache full: %llu + %llu > budget %llu", (unsigned long long)ctx->arena_used(),
                (unsigned long long)blob_bytes, (unsigned long long)ctx->budget);
      return LC_ERR_CACHE_FULL;
    }
    uint32_t slab = 0;
    uint8_t* d_blob = ctx->arena_alloc(blob_bytes, &slab);
    if (!d_blob) {
      give_back();
      set_error(ctx->arena_at_limit() ? "cache full: the HBM reservation has reached the budget for %llu bytes" : "HBM arena: cudaMalloc failed for %llu bytes", (unsigned long long)blob_bytes);
      return ctx->arena_at_limit() ? LC_ERR_CACHE_FULL : LC_ERR_OOM;
    }
    taken.push_back({d_blob, slab, blob_bytes});
    std::memset(&h_pw[i], 0, sizeof(IntPackWork));
    h_pw[i].values = d + voff[i];
    // an all-null batch without a validity buffer cannot happen (n_valid == 0 implies nulls); an all-null batch WITH
    // one keeps it
    h_pw[i].validity = has_nulls ? reinterpret_cast<const uint32_t*>(d + moff[i]) : nullptr;
    h_pw[i].blob = d_blob;
    h_pw[i].hdr = hd;
  }
  cudaError_t ce = cudaMemcpyAsync(d + pw_off, h_pw, nb * sizeof(IntPackWork), cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = launch_int_pack(reinterpret_cast<const IntPackWork*>(d + pw_off), static_cast<uint32_t>(nb), s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) {
    give_back();
    set_error("CUDA error in int_encode_many: %s", cudaGetErrorString(ce));
    return LC_ERR_CUDA;
  }
  ctx->kernel_launches++;
  ctx->h2d_bytes += nb * sizeof(IntPackWork);
  out->reserve(nb);
  for (uint64_t i = 0; i < nb; ++i) {
    Entry* e = new Entry();
    e->liquid_type = LC_LIQUID_INTEGER;
    e->d_blob = taken[i].blob;
    e->blob_bytes = static_cast<uint32_t>(taken[i].bytes);
    e->slab = taken[i].slab;
    e->n = static_cast<uint32_t>(ins[i].length);
    e->arrow_format = ins[i].format;
    e->ih = h_pw[i].hdr;
    ctx->n_entries++;
    out->push_back(e);
  }
  return LC_OK;
}

}  // namespace lc
