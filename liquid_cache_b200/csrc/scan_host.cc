// scan_host.cc — batched get / eval_predicate over lists of HBM-resident entries.
// Host side does planning (per-entry predicate constants, output offsets) and exactly one H2D of the
// work list + selections and one D2H of the results per call; all per-row work is in the kernels.
// Reference call sites: LiquidCache::read_arrow_array / eval_predicate_internal
// (/root/reference/src/core/src/cache/core.rs:595-634, 862-930).
#include <algorithm>

#include "host_common.h"

namespace lc {

namespace {

bool all_ones(const uint8_t* bits, uint64_t n) {
  const uint64_t full = n / 8;
  for (uint64_t i = 0; i < full; ++i)
    if (bits[i] != 0xFF) return false;
  const uint32_t rem = static_cast<uint32_t>(n & 7);
  if (rem && (bits[full] & ((1u << rem) - 1u)) != ((1u << rem) - 1u)) return false;
  return true;
}

// per-entry selection bookkeeping shared by every batched call
struct SelPlan {
  std::vector<const uint8_t*> bits;  // nullptr = dense
  std::vector<uint32_t> k;           // selected rows
  std::vector<uint64_t> word_off;    // offset (in u32 words) of the entry's selection in the upload area
  uint64_t sel_words = 0;
  uint64_t total_k = 0;
};

void plan_selection(Entry* const* entries, uint64_t n, const uint8_t* const* sel_bits, SelPlan* p,
                    const DevSel* dev = nullptr) {
  p->bits.assign(n, nullptr);
  p->k.assign(n, 0);
  p->word_off.assign(n, 0);
  if (dev) {  // selections are already on the device; only the counts matter here
    for (uint64_t i = 0; i < n; ++i) {
      p->k[i] = dev->all_rows ? entries[i]->n : dev->k[i];
      p->total_k += p->k[i];
    }
    return;
  }
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t rows = entries[i]->n;
    const uint8_t* b = sel_bits ? sel_bits[i] : nullptr;
    if (b && all_ones(b, rows)) b = nullptr;
    p->bits[i] = b;
    p->k[i] = b ? static_cast<uint32_t>(popcount_bits(b, rows)) : rows;
    if (b) {
      p->word_off[i] = p->sel_words;
      p->sel_words += round_up((rows + 31) / 32, 4);
    }
    p->total_k += p->k[i];
  }
}

void fill_selection(const SelPlan& p, Entry* const* entries, uint64_t n, uint32_t* h_sel) {
  for (uint64_t i = 0; i < n; ++i) {
    if (!p.bits[i]) continue;
    const uint32_t rows = entries[i]->n;
    const uint64_t words = round_up((rows + 31) / 32, 4);
    copy_bits(p.bits[i], 0, rows, reinterpret_cast<uint8_t*>(h_sel + p.word_off[i]), words * 4);
  }
}

// KMP failure links of the LIKE needle
void kmp_fail(const uint8_t* nd, uint32_t m, uint16_t* fail) {
  if (!m) return;
  fail[0] = 0;
  uint32_t q = 0;
  for (uint32_t i = 1; i < m; ++i) {
    while (q > 0 && nd[q] != nd[i]) q = fail[q - 1];
    if (nd[q] == nd[i]) ++q;
    fail[i] = static_cast<uint16_t>(q);
  }
}

}  // namespace

// ---- byte-view predicate planning ---------------------------------------------------------------
// Restates the shared-prefix / length / prefix7 case analysis of
// byte_view_array/comparisons.rs:21-82 (equality), :351-405 + :469-501 (ordering), :159-183 (LIKE).
struct StrPlan {
  int32_t kind;
  uint32_t flags;
  uint64_t key_expect;
  uint32_t cmp_len;
};

static int plan_str_entry(const Entry* e, int op, const uint8_t* needle, uint32_t m, StrPlan* out) {
  const uint8_t* sp = e->shared_prefix.data();
  const uint32_t spl = static_cast<uint32_t>(e->shared_prefix.size());
  out->flags = 0;
  out->key_expect = 0;
  out->cmp_len = 0;
  if (op == LC_OP_CONST_TRUE || op == LC_OP_CONST_FALSE) {
    out->kind = SP_CONST;
    out->flags = (op == LC_OP_CONST_TRUE) ? 1u : 0u;
    return LC_OK;
  }
  if (op == LC_OP_EQ || op == LC_OP_NE) {
    const bool neg = (op == LC_OP_NE);
    if (m < spl || std::memcmp(needle, sp, spl) != 0) {
      out->kind = SP_CONST;  // no value can equal the needle
      out->flags = neg ? 1u : 0u;
      return LC_OK;
    }
    const uint8_t* s = needle + spl;
    const uint32_t L = m - spl;
    uint64_t k = 0;
    for (uint32_t b = 0; b < (L < 7 ? L : 7); ++b) k |= static_cast<uint64_t>(s[b]) << (8 * b);
    k |= static_cast<uint64_t>(L >= 255 ? 255u : L) << 56;
    out->key_expect = k;
    out->kind = (L <= 7) ? SP_EQ_SHORT : SP_EQ_LONG;
    out->flags = neg ? 2u : 0u;
    return LC_OK;
  }
  if (op >= LC_OP_LT && op <= LC_OP_GE) {
    const bool less_op = (op == LC_OP_LT || op == LC_OP_LE);
    const uint32_t c_len = m < spl ? m : spl;
    const int c = c_len ? std::memcmp(sp, needle, c_len) : 0;
    if (c != 0 || m < spl) {
      // compare_with_shared_prefix: decided for the whole dictionary
      bool res;
      if (c < 0) res = less_op;
      else if (c > 0) res = !less_op;
      else res = !less_op;  // needle shorter than the shared prefix: every value is greater
      out->kind = SP_CONST;
      out->flags = res ? 1u : 0u;
      return LC_OK;
    }
    const uint8_t* s = needle + spl;
    const uint32_t L7 = (m - spl) < 7 ? (m - spl) : 7;
    if (L7 == 0) {
      out->kind = SP_ORD_EMPTY;
      return LC_OK;
    }
    uint64_t k = 0;
    for (uint32_t b = 0; b < L7; ++b) k |= static_cast<uint64_t>(s[b]) << (8 * (7 - b));
    out->kind = SP_ORD;
    out->key_expect = k;
    out->cmp_len = L7;
    return LC_OK;
  }
  if (op == LC_OP_LIKE || op == LC_OP_NOT_LIKE) {
    out->kind = SP_LIKE;
    out->flags = (op == LC_OP_NOT_LIKE ? 2u : 0u) | (e->sh.has_fp ? 0u : 4u);
    return LC_OK;
  }
  set_error("unsupported operator %d on a byte-view column", op);
  return LC_ERR_UNSUPPORTED_EXPR;
}

// substring_pattern_bytes (byte_view_array/fingerprint.rs:59-73): '%x%' with x non-empty, no % or _.
// A backslash would make arrow's LIKE take the escape-aware regex path, so it is declined as well.
static int like_inner(const uint8_t* pat, uint64_t len, const uint8_t** inner, uint32_t* inner_len) {
  if (len < 3 || pat[0] != '%' || pat[len - 1] != '%') {
    set_error("LIKE pattern is not of the form %%x%%");
    return LC_ERR_UNSUPPORTED_EXPR;
  }
  for (uint64_t i = 1; i + 1 < len; ++i) {
    if (pat[i] == '%' || pat[i] == '_' || pat[i] == '\\') {
      set_error("LIKE pattern has wildcards or escapes inside");
      return LC_ERR_UNSUPPORTED_EXPR;
    }
  }
  *inner = pat + 1;
  *inner_len = static_cast<uint32_t>(len - 2);
  return LC_OK;
}

// Everything a predicate launch over byte-view entries needs, built on the host.
struct StrLaunch {
  StrPredDesc desc;
  std::vector<uint8_t> needle_blob;  // needle padded to 4 + KMP links
  const uint8_t* needle = nullptr;   // bytes the per-entry planning compares against
  uint32_t m = 0;
};

static int prepare_str_pred(const lc_predicate* pred, StrLaunch* L) {
  std::memset(&L->desc, 0, sizeof(L->desc));
  L->desc.op = pred->op;
  const int op = pred->op;
  if (op == LC_OP_CONST_TRUE || op == LC_OP_CONST_FALSE) return LC_OK;
  if (pred->lit_kind != LC_LIT_BYTES || (!pred->lit_bytes && pred->lit_len)) {
    set_error("byte-view column needs a bytes literal");
    return LC_ERR_UNSUPPORTED_EXPR;
  }
  const uint8_t* nd = pred->lit_bytes;
  uint64_t m = pred->lit_len;
  if (op == LC_OP_LIKE || op == LC_OP_NOT_LIKE) {
    uint32_t il;
    LC_TRY(like_inner(pred->lit_bytes, pred->lit_len, &nd, &il));
    m = il;
    uint32_t fp = 0;
    for (uint32_t i = 0; i < il; ++i) fp |= 1u << (nd[i] & 31u);
    L->desc.needle_fp = fp;
  }
  if (m > kMaxNeedle) {
    set_error("needle longer than %u bytes", kMaxNeedle);
    return LC_ERR_UNSUPPORTED_EXPR;
  }
  L->needle = nd;
  L->m = static_cast<uint32_t>(m);
  L->desc.needle_len = L->m;
  const uint32_t padded = (L->m + 3u) & ~3u;
  L->needle_blob.assign(padded + 2u * L->m + 16u, 0);
  if (L->m) std::memcpy(L->needle_blob.data(), nd, L->m);
  if (op == LC_OP_LIKE || op == LC_OP_NOT_LIKE)
    kmp_fail(nd, L->m, reinterpret_cast<uint16_t*>(L->needle_blob.data() + padded));
  return LC_OK;
}

// ---- eval_predicate --------------------------------------------------------------------------------
int eval_predicate_batch(lc_ctx* ctx, Entry* const* entries, uint64_t n, const lc_predicate* pred,
                         const uint8_t* const* sel_bits, const PredOut& out) {
  if (n == 0) return LC_OK;
  const int32_t type = entries[0]->liquid_type;
  for (uint64_t i = 0; i < n; ++i) {
    if (entries[i]->liquid_type != type) {
      set_error("eval_predicate_many: entries of different liquid types in one call");
      return LC_ERR_INVALID;
    }
  }
  SelPlan sp;
  plan_selection(entries, n, sel_bits, &sp);
  const bool is_int = (type == LC_LIQUID_INTEGER);
  StrLaunch sl;
  if (!is_int) LC_TRY(prepare_str_pred(pred, &sl));

  // output layout: counts[2n] | mask words | validity words  (word aligned per entry)
  std::vector<uint64_t> out_word_off(n);
  uint64_t out_words = 0;
  for (uint64_t i = 0; i < n; ++i) {
    out_word_off[i] = out_words;
    out_words += round_up((sp.k[i] + 31) / 32, 4);
  }
  const uint64_t work_sz = is_int ? sizeof(IntScanWork) : sizeof(StrScanWork);
  const uint64_t up_works = round_up(n * work_sz, 256);
  const uint64_t up_needle = is_int ? 0 : round_up(sl.needle_blob.size(), 256);
  const uint64_t up_sel = round_up(sp.sel_words * 4, 256);
  const uint64_t up_total = up_works + up_needle + up_sel;
  const uint64_t dn_counts = round_up(n * 8, 256);
  const uint64_t dn_bits = round_up(out_words * 4, 256);
  const uint64_t dn_total = dn_counts + 2 * dn_bits;
  Scratch& sc = ctx->scratch;
  LC_TRY(sc.reserve(up_total + dn_total + 1024, up_total + dn_total + 1024));
  uint8_t* h_up = sc.host(up_total);
  uint8_t* h_dn = sc.host(dn_total);
  uint8_t* d_up = sc.dev(up_total);
  uint8_t* d_dn = sc.dev(dn_total);
  if (!h_up || !h_dn || !d_up || !d_dn) {
    set_error("eval_predicate: scratch exhausted");
    return LC_ERR_OOM;
  }
  uint32_t* d_sel = reinterpret_cast<uint32_t*>(d_up + up_works + up_needle);
  uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_dn);
  uint32_t* d_mask = reinterpret_cast<uint32_t*>(d_dn + dn_counts);
  uint32_t* d_valid = reinterpret_cast<uint32_t*>(d_dn + dn_counts + dn_bits);
  fill_selection(sp, entries, n, reinterpret_cast<uint32_t*>(h_up + up_works + up_needle));

  uint32_t max_blob = 0, max_head = 0, max_unique = 1;
  if (is_int) {
    IntScanWork* w = reinterpret_cast<IntScanWork*>(h_up);
    for (uint64_t i = 0; i < n; ++i) {
      Entry* e = entries[i];
      int32_t ucmp;
      uint64_t thr;
      LC_TRY(plan_int_predicate(e->ih, pred, &ucmp, &thr));
      w[i].blob = e->d_blob;
      w[i].sel = sp.bits[i] ? d_sel + sp.word_off[i] : nullptr;
      w[i].out_values = d_mask + out_word_off[i];
      w[i].out_validity = d_valid + out_word_off[i];
      w[i].out_counts = d_counts + 2 * i;
      w[i].thr = thr;
      w[i].ucmp = ucmp;
      w[i].blob_bytes = e->blob_bytes;
      max_blob = std::max(max_blob, e->blob_bytes);
    }
  } else {
    std::memcpy(h_up + up_works, sl.needle_blob.data(), sl.needle_blob.size());
    sl.desc.needle = d_up + up_works;
    StrScanWork* w = reinterpret_cast<StrScanWork*>(h_up);
    for (uint64_t i = 0; i < n; ++i) {
      Entry* e = entries[i];
      StrPlan p;
      LC_TRY(plan_str_entry(e, pred->op, sl.needle, sl.m, &p));
      std::memset(&w[i], 0, sizeof(StrScanWork));
      w[i].blob = e->d_blob;
      w[i].sel = sp.bits[i] ? d_sel + sp.word_off[i] : nullptr;
      w[i].out_values = d_mask + out_word_off[i];
      w[i].out_validity = d_valid + out_word_off[i];
      w[i].out_counts = d_counts + 2 * i;
      w[i].key_expect = p.key_expect;
      w[i].kind = p.kind;
      w[i].flags = p.flags;
      w[i].cmp_len = p.cmp_len;
      w[i].blob_bytes = e->blob_bytes;
      w[i].head_bytes = e->sh.head_bytes;
      w[i].meta_bytes = e->sh.meta_bytes;
      max_head = std::max(max_head, e->sh.head_bytes);
      max_unique = std::max(max_unique, e->sh.n_unique);
    }
  }
  cudaStream_t s = ctx->stream;
  LC_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_total, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_total;
  if (is_int) {
    LC_CUDA_OK(launch_int_scan(MODE_PRED, reinterpret_cast<const IntScanWork*>(d_up), static_cast<uint32_t>(n), max_blob, s));
  } else {
    LC_CUDA_OK(launch_str_scan(MODE_PRED, reinterpret_cast<const StrScanWork*>(d_up), static_cast<uint32_t>(n), sl.desc,
                               max_head, max_unique, s));
  }
  ctx->kernel_launches++;
  LC_CUDA_OK(cudaMemcpyAsync(h_dn, d_dn, dn_total, cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += dn_total;

  const uint32_t* h_counts = reinterpret_cast<const uint32_t*>(h_dn);
  const uint8_t* h_mask = h_dn + dn_counts;
  const uint8_t* h_valid = h_dn + dn_counts + dn_bits;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t k = h_counts[2 * i], nulls = h_counts[2 * i + 1];
    if (k != sp.k[i]) {
      set_error("internal: selected-row count mismatch on entry %llu (%u vs %u)", (unsigned long long)i, k, sp.k[i]);
      return LC_ERR_INVALID;
    }
    const uint64_t bytes = static_cast<uint64_t>((k + 31) / 32) * 4;
    const uint64_t bo = out.byte_offsets ? out.byte_offsets[i] : 0;
    std::memcpy(out.values + bo, h_mask + out_word_off[i] * 4, bytes);
    if (out.validity) {
      if (nulls == 0) std::memset(out.validity + bo, 0xFF, bytes);
      else std::memcpy(out.validity + bo, h_valid + out_word_off[i] * 4, bytes);
    }
    if (out.len) out.len[i] = k;
    if (out.null_count) out.null_count[i] = nulls;
  }
  return LC_OK;
}

// ---- device pipeline: selection := selection & valid & predicate (no host round trip of bits) ----
int refine_batch(lc_ctx* ctx, Entry* const* entries, uint64_t n, const lc_predicate* pred, uint32_t* d_sel_base,
                 const uint64_t* word_off, bool all_rows, uint32_t* d_counts) {
  if (n == 0) return LC_OK;
  const int32_t type = entries[0]->liquid_type;
  for (uint64_t i = 0; i < n; ++i) {
    if (entries[i]->liquid_type != type) {
      set_error("scan_filter: entries of different liquid types in one call");
      return LC_ERR_INVALID;
    }
  }
  const bool is_int = (type == LC_LIQUID_INTEGER);
  StrLaunch sl;
  if (!is_int) LC_TRY(prepare_str_pred(pred, &sl));
  const uint64_t work_sz = is_int ? sizeof(IntScanWork) : sizeof(StrScanWork);
  const uint64_t up_works = round_up(n * work_sz, 256);
  const uint64_t up_needle = is_int ? 0 : round_up(sl.needle_blob.size(), 256);
  const uint64_t up_total = up_works + up_needle;
  Scratch& sc = ctx->scratch;
  LC_TRY(sc.reserve(up_total + 1024, up_total + 1024));
  uint8_t* h_up = sc.host(up_total);
  uint8_t* d_up = sc.dev(up_total);
  if (!h_up || !d_up) {
    set_error("scan_filter: scratch exhausted");
    return LC_ERR_OOM;
  }
  uint32_t max_blob = 0, max_head = 0, max_unique = 1;
  if (is_int) {
    IntScanWork* w = reinterpret_cast<IntScanWork*>(h_up);
    for (uint64_t i = 0; i < n; ++i) {
      Entry* e = entries[i];
      int32_t ucmp;
      uint64_t thr;
      LC_TRY(plan_int_predicate(e->ih, pred, &ucmp, &thr));
      w[i].blob = e->d_blob;
      w[i].sel = all_rows ? nullptr : d_sel_base + word_off[i];
      w[i].out_values = d_sel_base + word_off[i];
      w[i].out_validity = nullptr;
      w[i].out_counts = d_counts + 2 * i;
      w[i].thr = thr;
      w[i].ucmp = ucmp;
      w[i].blob_bytes = e->blob_bytes;
      max_blob = std::max(max_blob, e->blob_bytes);
    }
  } else {
    std::memcpy(h_up + up_works, sl.needle_blob.data(), sl.needle_blob.size());
    sl.desc.needle = d_up + up_works;
    StrScanWork* w = reinterpret_cast<StrScanWork*>(h_up);
    for (uint64_t i = 0; i < n; ++i) {
      Entry* e = entries[i];
      StrPlan p;
      LC_TRY(plan_str_entry(e, pred->op, sl.needle, sl.m, &p));
      std::memset(&w[i], 0, sizeof(StrScanWork));
      w[i].blob = e->d_blob;
      w[i].sel = all_rows ? nullptr : d_sel_base + word_off[i];
      w[i].out_values = d_sel_base + word_off[i];
      w[i].out_validity = nullptr;
      w[i].out_counts = d_counts + 2 * i;
      w[i].key_expect = p.key_expect;
      w[i].kind = p.kind;
      w[i].flags = p.flags;
      w[i].cmp_len = p.cmp_len;
      w[i].blob_bytes = e->blob_bytes;
      w[i].head_bytes = e->sh.head_bytes;
      w[i].meta_bytes = e->sh.meta_bytes;
      max_head = std::max(max_head, e->sh.head_bytes);
      max_unique = std::max(max_unique, e->sh.n_unique);
    }
  }
  cudaStream_t s = ctx->stream;
  LC_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_total, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_total;
  if (is_int) {
    LC_CUDA_OK(launch_int_scan(MODE_REFINE, reinterpret_cast<const IntScanWork*>(d_up), static_cast<uint32_t>(n), max_blob, s));
  } else {
    LC_CUDA_OK(launch_str_scan(MODE_REFINE, reinterpret_cast<const StrScanWork*>(d_up), static_cast<uint32_t>(n), sl.desc,
                               max_head, max_unique, s));
  }
  ctx->kernel_launches++;
  // the pinned staging area is reused by the next call: wait for the upload (cheap, the kernel keeps running)
  LC_CUDA_OK(cudaStreamSynchronize(s));
  return LC_OK;
}

// ---- get / filter ----------------------------------------------------------------------------------
static void concat_validity(const uint8_t* src_words, uint32_t k, uint64_t dst_bit, uint8_t* dst) {
  // append k bits (src bit offset 0) at bit position dst_bit of dst (zero-initialised)
  for (uint32_t i = 0; i < k;) {
    const uint64_t d = dst_bit + i;
    if ((d & 7) == 0 && (i & 7) == 0 && k - i >= 8) {
      const uint32_t nb = (k - i) / 8;
      std::memcpy(dst + d / 8, src_words + i / 8, nb);
      i += nb * 8;
    } else {
      if ((src_words[i >> 3] >> (i & 7)) & 1) dst[d >> 3] |= static_cast<uint8_t>(1u << (d & 7));
      ++i;
    }
  }
}

static void set_bits_ones(uint8_t* dst, uint64_t from, uint64_t count) {
  for (uint64_t i = 0; i < count;) {
    const uint64_t d = from + i;
    if ((d & 7) == 0 && count - i >= 8) {
      const uint64_t nb = (count - i) / 8;
      std::memset(dst + d / 8, 0xFF, nb);
      i += nb * 8;
    } else {
      dst[d >> 3] |= static_cast<uint8_t>(1u << (d & 7));
      ++i;
    }
  }
}

static int finish_bytes_array(const Entry* proto, uint64_t rows, uint64_t nulls, HostBuf validity, HostBuf offsets,
                              HostBuf data, ArrowSchema* out_schema, ArrowArray* out_array);

int to_arrow_batch(lc_ctx* ctx, Entry* const* entries, uint64_t n, const uint8_t* const* sel_bits,
                   const DevSel* dev_sel, ArrowSchema* out_schema, ArrowArray* out_array, const DeviceOut* dev_out) {
  if (n == 0) {
    set_error("to_arrow: empty entry list");
    return LC_ERR_INVALID;
  }
  const Entry* proto = entries[0];
  for (uint64_t i = 1; i < n; ++i) {
    if (entries[i]->liquid_type != proto->liquid_type || entries[i]->arrow_format != proto->arrow_format ||
        entries[i]->dict_value_format != proto->dict_value_format) {
      set_error("to_arrow_many: entries have different arrow types");
      return LC_ERR_INVALID;
    }
  }
  SelPlan sp;
  plan_selection(entries, n, sel_bits, &sp, dev_sel);
  auto sel_ptr = [&](uint64_t i, uint32_t* d_upload) -> const uint32_t* {
    if (dev_sel) return dev_sel->all_rows ? nullptr : dev_sel->d_base + dev_sel->word_off[i];
    return sp.bits[i] ? d_upload + sp.word_off[i] : nullptr;
  };
  if (dev_out) {
    set_error("device-resident results are not wired up for this call yet");
    return LC_ERR_INVALID;
  }
  const bool is_int = (proto->liquid_type == LC_LIQUID_INTEGER);
  cudaStream_t s = ctx->stream;
  Scratch& sc = ctx->scratch;

  std::vector<uint64_t> vword_off(n), row_base(n);
  uint64_t vwords = 0, rows = 0;
  for (uint64_t i = 0; i < n; ++i) {
    vword_off[i] = vwords;
    vwords += round_up((sp.k[i] + 31) / 32, 4);
    row_base[i] = rows;
    rows += sp.k[i];
  }
  if (rows > 0x7fffffffull) {
    set_error("result has more than 2^31 rows");
    return LC_ERR_INVALID;
  }
  const uint64_t up_sel = round_up(sp.sel_words * 4, 256);
  const uint64_t dn_counts = round_up(n * 16, 256);
  const uint64_t dn_valid = round_up(vwords * 4, 256);

  if (is_int) {
    const uint32_t tb = proto->ih.tbits / 8;
    const uint64_t up_works = round_up(n * sizeof(IntScanWork), 256);
    const uint64_t up_total = up_works + up_sel;
    const uint64_t dn_total = dn_counts + dn_valid;
    const uint64_t val_bytes = round_up(rows * tb, 256);
    LC_TRY(sc.reserve(up_total + dn_total + val_bytes + 1024, up_total + dn_total + 1024));
    uint8_t* h_up = sc.host(up_total);
    uint8_t* h_dn = sc.host(dn_total);
    uint8_t* d_up = sc.dev(up_total);
    uint8_t* d_dn = sc.dev(dn_total);
    uint8_t* d_vals = sc.dev(val_bytes);
    if (!h_up || !h_dn || !d_up || !d_dn || !d_vals) {
      set_error("to_arrow: scratch exhausted");
      return LC_ERR_OOM;
    }
    uint32_t* d_sel = reinterpret_cast<uint32_t*>(d_up + up_works);
    uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_dn);
    uint32_t* d_valid = reinterpret_cast<uint32_t*>(d_dn + dn_counts);
    fill_selection(sp, entries, n, reinterpret_cast<uint32_t*>(h_up + up_works));
    IntScanWork* w = reinterpret_cast<IntScanWork*>(h_up);
    uint32_t max_blob = 0;
    for (uint64_t i = 0; i < n; ++i) {
      Entry* e = entries[i];
      w[i].blob = e->d_blob;
      w[i].sel = sel_ptr(i, d_sel);
      w[i].out_values = d_vals + row_base[i] * tb;
      w[i].out_validity = d_valid + vword_off[i];
      w[i].out_counts = d_counts + 4 * i;
      w[i].thr = 0;
      w[i].ucmp = UC_TRUE;
      w[i].blob_bytes = e->blob_bytes;
      max_blob = std::max(max_blob, e->blob_bytes);
    }
    LC_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_total, cudaMemcpyHostToDevice, s));
    ctx->h2d_bytes += up_total;
    LC_CUDA_OK(launch_int_scan(MODE_DECODE, reinterpret_cast<const IntScanWork*>(d_up), static_cast<uint32_t>(n), max_blob, s));
    ctx->kernel_launches++;
    HostBuf values{host_alloc(rows * tb), rows * tb};
    if (!values.p) {
      set_error("host allocation of %llu bytes failed", (unsigned long long)(rows * tb));
      return LC_ERR_OOM;
    }
    LC_CUDA_OK(cudaMemcpyAsync(h_dn, d_dn, dn_total, cudaMemcpyDeviceToHost, s));
    if (rows) LC_CUDA_OK(cudaMemcpyAsync(values.p, d_vals, rows * tb, cudaMemcpyDeviceToHost, s));
    LC_CUDA_OK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += dn_total + rows * tb;
    const uint32_t* h_counts = reinterpret_cast<const uint32_t*>(h_dn);
    uint64_t nulls = 0;
    for (uint64_t i = 0; i < n; ++i) nulls += h_counts[4 * i + 1];
    HostBuf validity{nullptr, 0};
    if (nulls) {
      validity.bytes = (rows + 7) / 8;
      validity.p = host_alloc(validity.bytes);
      std::memset(validity.p, 0, round_up(validity.bytes, 64));
      for (uint64_t i = 0; i < n; ++i) {
        if (h_counts[4 * i + 1] == 0) set_bits_ones(validity.p, row_base[i], sp.k[i]);
        else concat_validity(h_dn + dn_counts + vword_off[i] * 4, sp.k[i], row_base[i], validity.p);
      }
    }
    export_schema(proto->arrow_format, "", out_schema);
    std::vector<HostBuf> bufs;
    bufs.push_back(validity);
    bufs.push_back(values);
    export_array(static_cast<int64_t>(rows), static_cast<int64_t>(nulls), std::move(bufs), nullptr, out_array);
    return LC_OK;
  }

  // ---------------- byte-view: pass 1 (lengths), host prefix sums, pass 2 (decode) ----------------
  const uint64_t up_works = round_up(n * sizeof(StrGatherWork), 256);
  const uint64_t up_total = up_works + up_sel;
  uint64_t ulen_words = 0;
  std::vector<uint64_t> ulen_off(n);
  for (uint64_t i = 0; i < n; ++i) {
    ulen_off[i] = ulen_words;
    ulen_words += round_up(entries[i]->sh.n_unique, 4);
  }
  const uint64_t dv_rowoff = round_up((rows + n) * 4, 256);  // k_i + 1 per entry
  const uint64_t dv_rowkey = round_up(rows * 4 + 4, 256);
  const uint64_t dv_ulen = round_up(ulen_words * 4 + 4, 256);
  const uint64_t dn_total = dn_counts + dn_valid;
  const uint64_t up2 = round_up(n * sizeof(StrDecodeWork), 256);
  LC_TRY(sc.reserve(up_total + dn_total + dv_rowoff + dv_rowkey + dv_ulen + up2 + 1024, up_total + dn_total + up2 + 1024));
  uint8_t* h_up = sc.host(up_total);
  uint8_t* h_dn = sc.host(dn_total);
  uint8_t* h_up2 = sc.host(up2);
  uint8_t* d_up = sc.dev(up_total);
  uint8_t* d_dn = sc.dev(dn_total);
  uint8_t* d_rowoff = sc.dev(dv_rowoff);
  uint8_t* d_rowkey = sc.dev(dv_rowkey);
  uint8_t* d_ulen = sc.dev(dv_ulen);
  uint8_t* d_up2 = sc.dev(up2);
  if (!h_up || !h_dn || !h_up2 || !d_up || !d_dn || !d_rowoff || !d_rowkey || !d_ulen || !d_up2) {
    set_error("to_arrow: scratch exhausted");
    return LC_ERR_OOM;
  }
  uint32_t* d_sel = reinterpret_cast<uint32_t*>(d_up + up_works);
  uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_dn);
  uint32_t* d_valid = reinterpret_cast<uint32_t*>(d_dn + dn_counts);
  fill_selection(sp, entries, n, reinterpret_cast<uint32_t*>(h_up + up_works));
  StrGatherWork* w = reinterpret_cast<StrGatherWork*>(h_up);
  uint32_t max_head = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Entry* e = entries[i];
    std::memset(&w[i], 0, sizeof(StrGatherWork));
    w[i].blob = e->d_blob;
    w[i].sel = sel_ptr(i, d_sel);
    w[i].row_off = reinterpret_cast<uint32_t*>(d_rowoff) + row_base[i] + i;
    w[i].row_key = reinterpret_cast<uint32_t*>(d_rowkey) + row_base[i];
    w[i].ulen = reinterpret_cast<uint32_t*>(d_ulen) + ulen_off[i];
    w[i].out_validity = d_valid + vword_off[i];
    w[i].out_counts = d_counts + 4 * i;
    w[i].blob_bytes = e->blob_bytes;
    w[i].head_bytes = e->sh.head_bytes;
    // all decoded lengths up front when most of the dictionary will be touched anyway
    w[i].flags = (static_cast<uint64_t>(sp.k[i]) * 4 >= e->sh.n_unique) ? kGatherPrecompLens : 0u;
    max_head = std::max(max_head, e->sh.head_bytes);
  }
  LC_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_total, cudaMemcpyHostToDevice, s));
  ctx->h2d_bytes += up_total;
  LC_CUDA_OK(launch_str_lengths(reinterpret_cast<const StrGatherWork*>(d_up), static_cast<uint32_t>(n), max_head, s));
  ctx->kernel_launches++;
  LC_CUDA_OK(cudaMemcpyAsync(h_dn, d_dn, dn_total, cudaMemcpyDeviceToHost, s));
  LC_CUDA_OK(cudaStreamSynchronize(s));
  ctx->d2h_bytes += dn_total;
  const uint32_t* h_counts = reinterpret_cast<const uint32_t*>(h_dn);
  uint64_t nulls = 0, total_bytes = 0;
  std::vector<uint64_t> byte_base(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (h_counts[4 * i] != sp.k[i]) {
      set_error("internal: selected-row count mismatch on entry %llu", (unsigned long long)i);
      return LC_ERR_INVALID;
    }
    nulls += h_counts[4 * i + 1];
    byte_base[i] = total_bytes;
    total_bytes += h_counts[4 * i + 2];
  }
  if (total_bytes > 0x7fffffffull) {
    set_error("decoded values exceed 2 GiB (int32 offsets); split the call");
    return LC_ERR_INVALID;
  }
  // device result buffers
  const uint64_t off_bytes = (rows + 1) * 4;
  uint8_t* d_off = nullptr;
  uint8_t* d_bytes = nullptr;
  // scratch is already carved; results go to a separate temporary allocation
  const uint64_t res_bytes = round_up(off_bytes, 256) + round_up(total_bytes + 8, 256);
  uint8_t* d_res = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_res), res_bytes, s) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMallocAsync of %llu result bytes failed", (unsigned long long)res_bytes);
    return LC_ERR_OOM;
  }
  d_off = d_res;
  d_bytes = d_res + round_up(off_bytes, 256);
  StrDecodeWork* w2 = reinterpret_cast<StrDecodeWork*>(h_up2);
  for (uint64_t i = 0; i < n; ++i) {
    w2[i].blob = entries[i]->d_blob;
    w2[i].row_off = reinterpret_cast<uint32_t*>(d_rowoff) + row_base[i] + i;
    w2[i].row_key = reinterpret_cast<uint32_t*>(d_rowkey) + row_base[i];
    w2[i].out_offsets = reinterpret_cast<int32_t*>(d_off) + row_base[i];
    w2[i].out_bytes = d_bytes;
    w2[i].byte_base = static_cast<uint32_t>(byte_base[i]);
    w2[i].k = sp.k[i];
  }
  HostBuf offsets{host_alloc(off_bytes), off_bytes};
  HostBuf data{host_alloc(total_bytes ? total_bytes : 1), total_bytes};
  if (!offsets.p || !data.p) {
    cudaFreeAsync(d_res, s);
    set_error("host allocation failed");
    return LC_ERR_OOM;
  }
  cudaError_t ce = cudaMemcpyAsync(d_up2, h_up2, up2, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = launch_str_decode(reinterpret_cast<const StrDecodeWork*>(d_up2), static_cast<uint32_t>(n), s);
  if (ce == cudaSuccess && rows) ce = cudaMemcpyAsync(offsets.p, d_off, rows * 4, cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess && total_bytes) ce = cudaMemcpyAsync(data.p, d_bytes, total_bytes, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(d_res, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) {
    host_free(offsets.p);
    host_free(data.p);
    set_error("CUDA error in byte-view decode: %s", cudaGetErrorString(ce));
    return LC_ERR_CUDA;
  }
  ctx->kernel_launches++;
  ctx->h2d_bytes += up2;
  ctx->d2h_bytes += rows * 4 + total_bytes;
  reinterpret_cast<int32_t*>(offsets.p)[rows] = static_cast<int32_t>(total_bytes);
  HostBuf validity{nullptr, 0};
  if (nulls) {
    validity.bytes = (rows + 7) / 8;
    validity.p = host_alloc(validity.bytes);
    std::memset(validity.p, 0, round_up(validity.bytes, 64));
    for (uint64_t i = 0; i < n; ++i) {
      if (h_counts[4 * i + 1] == 0) set_bits_ones(validity.p, row_base[i], sp.k[i]);
      else concat_validity(h_dn + dn_counts + vword_off[i] * 4, sp.k[i], row_base[i], validity.p);
    }
  }
  return finish_bytes_array(proto, rows, nulls, validity, offsets, data, out_schema, out_array);
}

// Turn (validity, int32 offsets, bytes) into the ORIGINAL arrow type of the column:
// Utf8 / Binary as is; Utf8View / BinaryView by building 16-byte views over the single data buffer;
// Dictionary<UInt16,_> by re-encoding (what arrow's cast dictionary -> original type leaves the caller with,
// byte_view_array/mod.rs:287-290).
static int finish_bytes_array(const Entry* proto, uint64_t rows, uint64_t nulls, HostBuf validity, HostBuf offsets,
                              HostBuf data, ArrowSchema* out_schema, ArrowArray* out_array) {
  const uint8_t bt = proto->sh.arrow_type;
  const int32_t* off = reinterpret_cast<const int32_t*>(offsets.p);
  if (bt == BT_UTF8 || bt == BT_BINARY) {
    export_schema(bt == BT_UTF8 ? "u" : "z", "", out_schema);
    std::vector<HostBuf> bufs{validity, offsets, data};
    export_array(static_cast<int64_t>(rows), static_cast<int64_t>(nulls), std::move(bufs), nullptr, out_array);
    return LC_OK;
  }
  if (bt == BT_UTF8_VIEW || bt == BT_BINARY_VIEW) {
    HostBuf views{host_alloc(rows * 16 + 16), rows * 16};
    HostBuf sizes{host_alloc(8), 8};
    if (!views.p || !sizes.p) {
      set_error("host allocation failed");
      return LC_ERR_OOM;
    }
    std::memset(views.p, 0, rows * 16 + 16);
    for (uint64_t r = 0; r < rows; ++r) {
      const bool ok = !validity.p || bit_get(validity.p, static_cast<int64_t>(r));
      if (!ok) continue;
      const uint32_t len = static_cast<uint32_t>(off[r + 1] - off[r]);
      uint8_t* v = views.p + 16 * r;
      std::memcpy(v, &len, 4);
      if (len <= 12) {
        std::memcpy(v + 4, data.p + off[r], len);
      } else {
        std::memcpy(v + 4, data.p + off[r], 4);
        const uint32_t bi = 0, bo = static_cast<uint32_t>(off[r]);
        std::memcpy(v + 8, &bi, 4);
        std::memcpy(v + 12, &bo, 4);
      }
    }
    const int64_t sz = static_cast<int64_t>(data.bytes);
    std::memcpy(sizes.p, &sz, 8);
    host_free(offsets.p);
    export_schema(bt == BT_UTF8_VIEW ? "vu" : "vz", "", out_schema);
    std::vector<HostBuf> bufs{validity, views, data, sizes};
    export_array(static_cast<int64_t>(rows), static_cast<int64_t>(nulls), std::move(bufs), nullptr, out_array);
    return LC_OK;
  }
  // Dictionary<UInt16, Utf8|Binary>: first-occurrence re-encode of the decoded rows
  std::unordered_map<std::string, uint16_t> seen;
  std::vector<std::string> order;
  HostBuf keys{host_alloc(rows * 2 + 2), rows * 2};
  if (!keys.p) return LC_ERR_OOM;
  std::memset(keys.p, 0, rows * 2 + 2);
  for (uint64_t r = 0; r < rows; ++r) {
    const bool ok = !validity.p || bit_get(validity.p, static_cast<int64_t>(r));
    if (!ok) continue;
    std::string sv(reinterpret_cast<const char*>(data.p) + off[r], static_cast<size_t>(off[r + 1] - off[r]));
    auto it = seen.find(sv);
    uint16_t key;
    if (it == seen.end()) {
      if (order.size() >= 65536) {
        host_free(keys.p);
        set_error("more than 65536 distinct values in a dictionary result");
        return LC_ERR_UNSUPPORTED_TYPE;
      }
      key = static_cast<uint16_t>(order.size());
      seen.emplace(sv, key);
      order.push_back(std::move(sv));
    } else {
      key = it->second;
    }
    reinterpret_cast<uint16_t*>(keys.p)[r] = key;
  }
  uint64_t dbytes = 0;
  for (auto& sv : order) dbytes += sv.size();
  HostBuf doff{host_alloc((order.size() + 1) * 4), (order.size() + 1) * 4};
  HostBuf ddata{host_alloc(dbytes ? dbytes : 1), dbytes};
  int32_t* dof = reinterpret_cast<int32_t*>(doff.p);
  uint64_t p = 0;
  for (size_t i = 0; i < order.size(); ++i) {
    dof[i] = static_cast<int32_t>(p);
    std::memcpy(ddata.p + p, order[i].data(), order[i].size());
    p += order[i].size();
  }
  dof[order.size()] = static_cast<int32_t>(p);
  host_free(offsets.p);
  host_free(data.p);
  ArrowArray dict_arr;
  std::vector<HostBuf> dbufs{HostBuf{nullptr, 0}, doff, ddata};
  export_array(static_cast<int64_t>(order.size()), 0, std::move(dbufs), nullptr, &dict_arr);
  export_schema("S", bt == BT_DICT16_UTF8 ? "u" : "z", out_schema);
  std::vector<HostBuf> bufs{validity, keys};
  export_array(static_cast<int64_t>(rows), static_cast<int64_t>(nulls), std::move(bufs), &dict_arr, out_array);
  return LC_OK;
}

}  // namespace lc
