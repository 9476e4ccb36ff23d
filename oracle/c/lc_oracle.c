/*
 * lc_oracle.c — plain-C restatement of the reference's CPU path, used as the TIMED CPU BASELINE
 * (bench.py cpu_baseline / --impl reference) and cross-checked against oracle/liquid_oracle.py.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY: nothing under liquid_cache_b200/ links or loads this.
 * The reference is Rust and cannot be compiled here (no cargo/rustc, crates un-vendored), so this is a
 * "port": it keeps the reference's pass structure rather than fusing anything —
 *   integers  decode-all (FastLanes unpack per 1024 block, bit_pack_array.rs:127-169) -> + reference
 *             (primitive_array.rs:350-368) -> arrow filter (370-374) -> compare (liquid_array/mod.rs:265-280)
 *   strings   fingerprint gate -> decompress the candidates (fsst_buffer.rs:642-663) -> substring match
 *             (comparisons.rs:600-651) -> dictionary results broadcast through the u16 keys (325-347)
 *   caller    nulls->false and boolean_buffer_and_then with BMI2 PDEP (datafusion/src/utils.rs:62-236)
 * FSST and FastLanes layouts are restated from their publications (see liquid_oracle.py header: byte layout
 * "parity unpinned", results independent of it).
 *
 * Build: make -C oracle/c   (gcc -O3 -march=native -shared -fPIC -pthread)
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ FastLanes ---- */
static const int FL_ORDER[8] = {0, 4, 2, 6, 1, 5, 3, 7};
static inline int fl_index(int row, int lane) { return FL_ORDER[row >> 3] * 16 + (row & 7) * 128 + lane; }

#define DEF_FL(T, BITS)                                                                              \
  static void fl_unpack_##BITS(const T* packed, int w, T* out) {                                      \
    const int lanes = 1024 / BITS;                                                                   \
    const T mask = (w == BITS) ? (T)~(T)0 : (T)(((T)1 << w) - 1);                                     \
    for (int row = 0; row < BITS; ++row) {                                                           \
      const int b = row * w, k = b / BITS, sh = b % BITS;                                            \
      const int base = FL_ORDER[row >> 3] * 16 + (row & 7) * 128;                                     \
      const T* p0 = packed + lanes * k;                                                              \
      if (sh + w <= BITS) {                                                                          \
        for (int l = 0; l < lanes; ++l) out[base + l] = (T)((p0[l] >> sh) & mask);                    \
      } else {                                                                                       \
        const T* p1 = p0 + lanes;                                                                    \
        for (int l = 0; l < lanes; ++l) out[base + l] = (T)(((p0[l] >> sh) | (p1[l] << (BITS - sh))) & mask); \
      }                                                                                              \
    }                                                                                                \
  }                                                                                                  \
  static void fl_pack_##BITS(const T* in, int w, T* packed) {                                         \
    const int lanes = 1024 / BITS;                                                                   \
    const T mask = (w == BITS) ? (T)~(T)0 : (T)(((T)1 << w) - 1);                                     \
    memset(packed, 0, (size_t)(1024 * w / 8));                                                       \
    for (int row = 0; row < BITS; ++row) {                                                           \
      const int b = row * w, k = b / BITS, sh = b % BITS;                                            \
      const int base = FL_ORDER[row >> 3] * 16 + (row & 7) * 128;                                     \
      T* p0 = packed + lanes * k;                                                                    \
      for (int l = 0; l < lanes; ++l) p0[l] |= (T)((in[base + l] & mask) << sh);                      \
      if (sh + w > BITS) {                                                                           \
        T* p1 = p0 + lanes;                                                                          \
        for (int l = 0; l < lanes; ++l) p1[l] |= (T)((in[base + l] & mask) >> (BITS - sh));           \
      }                                                                                              \
    }                                                                                                \
  }
DEF_FL(uint8_t, 8)
DEF_FL(uint16_t, 16)
DEF_FL(uint32_t, 32)
DEF_FL(uint64_t, 64)

/* ------------------------------------------------------------------ integer entry ---- */
typedef struct {
  int tbits;          /* 8/16/32/64 */
  int is_signed;
  int bit_width;      /* 0 = all null */
  uint32_t n;
  uint64_t reference; /* raw bits */
  void* packed;       /* ceil(n/1024) * 128*W bytes */
  uint8_t* validity;  /* n bits or NULL */
  uint32_t null_count;
} lco_int;

static inline int bit_get(const uint8_t* b, uint64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }

EXPORT void lco_int_free(lco_int* e) {
  if (!e) return;
  free(e->packed);
  free(e->validity);
  free(e);
}

/* LiquidPrimitiveArray::from_arrow_array */
EXPORT lco_int* lco_int_encode(const void* values, const uint8_t* validity, uint32_t n, int tbits, int is_signed) {
  lco_int* e = (lco_int*)calloc(1, sizeof(lco_int));
  e->tbits = tbits;
  e->is_signed = is_signed;
  e->n = n;
  const int tb = tbits / 8;
  /* pass 1+2: arrow min, arrow max over valid values */
  int any = 0;
  int64_t smin = INT64_MAX, smax = INT64_MIN;
  uint64_t umin = UINT64_MAX, umax = 0;
  uint32_t nulls = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (validity && !bit_get(validity, i)) { ++nulls; continue; }
    any = 1;
    if (is_signed) {
      int64_t v = tbits == 8 ? ((const int8_t*)values)[i] : tbits == 16 ? ((const int16_t*)values)[i]
                : tbits == 32 ? ((const int32_t*)values)[i] : ((const int64_t*)values)[i];
      if (v < smin) smin = v;
      if (v > smax) smax = v;
    } else {
      uint64_t v = tbits == 8 ? ((const uint8_t*)values)[i] : tbits == 16 ? ((const uint16_t*)values)[i]
                 : tbits == 32 ? ((const uint32_t*)values)[i] : ((const uint64_t*)values)[i];
      if (v < umin) umin = v;
      if (v > umax) umax = v;
    }
  }
  e->null_count = nulls;
  if (nulls) {
    e->validity = (uint8_t*)calloc((n + 7) / 8 + 8, 1);
    memcpy(e->validity, validity, (n + 7) / 8);
  }
  if (!any) return e;
  const uint64_t tmask = tbits == 64 ? ~0ull : ((1ull << tbits) - 1);
  const uint64_t mn = is_signed ? (uint64_t)smin : umin, mx = is_signed ? (uint64_t)smax : umax;
  const uint64_t sub = (mx - mn) & tmask;
  e->bit_width = sub == 0 ? 1 : 64 - __builtin_clzll(sub);
  e->reference = mn & tmask;
  const uint32_t n_chunks = (n + 1023) / 1024;
  const size_t chunk_bytes = (size_t)128 * e->bit_width;
  e->packed = calloc((size_t)n_chunks * chunk_bytes + 64, 1);
  /* pass 3: subtract reference into the unsigned twin; pass 4: pack per 1024 block (tail zero padded) */
  uint8_t* tmp = (uint8_t*)calloc(1024, tb);
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const uint32_t lo = c * 1024, cnt = (n - lo) < 1024 ? (n - lo) : 1024;
    memset(tmp, 0, (size_t)1024 * tb);
#define SUBPACK(T, BITS)                                                                         \
  {                                                                                              \
    T* t = (T*)tmp;                                                                              \
    const T* v = (const T*)values + lo;                                                          \
    const T r = (T)e->reference;                                                                 \
    for (uint32_t i = 0; i < cnt; ++i) t[i] = (T)(v[i] - r);                                      \
    fl_pack_##BITS(t, e->bit_width, (T*)((uint8_t*)e->packed + c * chunk_bytes));                \
  }
    if (tbits == 8) SUBPACK(uint8_t, 8)
    else if (tbits == 16) SUBPACK(uint16_t, 16)
    else if (tbits == 32) SUBPACK(uint32_t, 32)
    else SUBPACK(uint64_t, 64)
  }
  free(tmp);
  return e;
}

/* to_arrow_array: unpack all chunks into a fresh buffer, truncate, add the reference */
static void* int_decode_all(const lco_int* e) {
  const int tb = e->tbits / 8;
  const uint32_t n_chunks = (e->n + 1023) / 1024;
  uint8_t* out = (uint8_t*)malloc((size_t)(n_chunks ? n_chunks : 1) * 1024 * tb);
  if (e->bit_width == 0) {
    memset(out, 0, (size_t)(n_chunks ? n_chunks : 1) * 1024 * tb);
    return out;
  }
  const size_t chunk_bytes = (size_t)128 * e->bit_width;
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const uint8_t* p = (const uint8_t*)e->packed + c * chunk_bytes;
    if (e->tbits == 8) fl_unpack_8((const uint8_t*)p, e->bit_width, (uint8_t*)out + (size_t)c * 1024);
    else if (e->tbits == 16) fl_unpack_16((const uint16_t*)p, e->bit_width, (uint16_t*)out + (size_t)c * 1024);
    else if (e->tbits == 32) fl_unpack_32((const uint32_t*)p, e->bit_width, (uint32_t*)out + (size_t)c * 1024);
    else fl_unpack_64((const uint64_t*)p, e->bit_width, (uint64_t*)out + (size_t)c * 1024);
  }
  if (e->reference) {
    const uint32_t n = e->n;
    if (e->tbits == 8) { uint8_t* o = out; const uint8_t r = (uint8_t)e->reference; for (uint32_t i = 0; i < n; ++i) o[i] += r; }
    else if (e->tbits == 16) { uint16_t* o = (uint16_t*)out; const uint16_t r = (uint16_t)e->reference; for (uint32_t i = 0; i < n; ++i) o[i] += r; }
    else if (e->tbits == 32) { uint32_t* o = (uint32_t*)out; const uint32_t r = (uint32_t)e->reference; for (uint32_t i = 0; i < n; ++i) o[i] += r; }
    else { uint64_t* o = (uint64_t*)out; const uint64_t r = e->reference; for (uint32_t i = 0; i < n; ++i) o[i] += r; }
  }
  return out;
}

/* arrow filter of a primitive array: values + validity of the selected rows. sel == NULL means all rows. */
static uint32_t filter_prim(const void* vals, const uint8_t* validity, const uint8_t* sel, uint32_t n, int tb,
                            void* out_vals, uint8_t* out_valid, uint32_t* out_nulls) {
  uint32_t k = 0, nulls = 0;
  if (!sel) {
    memcpy(out_vals, vals, (size_t)n * tb);
    if (validity && out_valid) memcpy(out_valid, validity, (n + 7) / 8);
    if (validity) for (uint32_t i = 0; i < n; ++i) nulls += !bit_get(validity, i);
    *out_nulls = nulls;
    return n;
  }
  const uint64_t* sw = (const uint64_t*)sel;
  const uint32_t n64 = (n + 63) / 64;
  for (uint32_t w = 0; w < n64; ++w) {
    uint64_t bits = 0;
    const uint32_t rem = n - w * 64;
    memcpy(&bits, (const uint8_t*)sel + (size_t)w * 8, rem >= 64 ? 8 : (rem + 7) / 8);
    if (rem < 64) bits &= (1ull << rem) - 1;
    (void)sw;
    while (bits) {
      const uint32_t i = w * 64 + (uint32_t)__builtin_ctzll(bits);
      bits &= bits - 1;
      memcpy((uint8_t*)out_vals + (size_t)k * tb, (const uint8_t*)vals + (size_t)i * tb, tb);
      if (out_valid) {
        const int v = validity ? bit_get(validity, i) : 1;
        if (v) out_valid[k >> 3] |= (uint8_t)(1u << (k & 7));
        else ++nulls;
      }
      ++k;
    }
  }
  *out_nulls = nulls;
  return k;
}

/* get().with_selection(): decode all, arrow filter. Returns k; caller provides out buffers sized for n. */
EXPORT uint32_t lco_int_filter(const lco_int* e, const uint8_t* sel, void* out_vals, uint8_t* out_valid,
                               uint32_t* out_nulls) {
  void* all = int_decode_all(e);
  if (out_valid) memset(out_valid, 0, (e->n + 7) / 8 + 1);
  const uint32_t k = filter_prim(all, e->validity, sel, e->n, e->tbits / 8, out_vals, out_valid, out_nulls);
  free(all);
  return k;
}

/* try_eval_predicate: filter, then compare against the literal (arrow-ord cmp on the native type).
 * op: 0 EQ 1 NE 2 LT 3 LE 4 GT 5 GE. out_mask has k bits (value bits), out_valid k bits. */
EXPORT uint32_t lco_int_eval(const lco_int* e, const uint8_t* sel, int op, int64_t lit_i, uint64_t lit_u,
                             uint8_t* out_mask, uint8_t* out_valid, uint32_t* out_nulls) {
  const int tb = e->tbits / 8;
  void* vals = malloc((size_t)(e->n ? e->n : 1) * tb);
  const uint32_t k = lco_int_filter(e, sel, vals, out_valid, out_nulls);
  memset(out_mask, 0, (k + 7) / 8 + 1);
#define CMPLOOP(T, LIT)                                                                         \
  {                                                                                             \
    const T* v = (const T*)vals;                                                                \
    const T l = (T)(LIT);                                                                       \
    for (uint32_t i = 0; i < k; ++i) {                                                          \
      int r;                                                                                    \
      switch (op) {                                                                             \
        case 0: r = v[i] == l; break;                                                           \
        case 1: r = v[i] != l; break;                                                           \
        case 2: r = v[i] < l; break;                                                            \
        case 3: r = v[i] <= l; break;                                                           \
        case 4: r = v[i] > l; break;                                                            \
        default: r = v[i] >= l; break;                                                          \
      }                                                                                         \
      out_mask[i >> 3] |= (uint8_t)(r << (i & 7));                                              \
    }                                                                                           \
  }
  if (e->is_signed) {
    if (e->tbits == 8) CMPLOOP(int8_t, lit_i) else if (e->tbits == 16) CMPLOOP(int16_t, lit_i)
    else if (e->tbits == 32) CMPLOOP(int32_t, lit_i) else CMPLOOP(int64_t, lit_i)
  } else {
    if (e->tbits == 8) CMPLOOP(uint8_t, lit_u) else if (e->tbits == 16) CMPLOOP(uint16_t, lit_u)
    else if (e->tbits == 32) CMPLOOP(uint32_t, lit_u) else CMPLOOP(uint64_t, lit_u)
  }
  free(vals);
  return k;
}

/* ------------------------------------------------------------------ FSST ---- */
typedef struct {
  uint64_t symbols[256];
  uint8_t lens[256];
  int n_symbols;
  /* encoder: symbols grouped by length, probed longest first through small open-addressed tables */
  uint32_t hash_cap;
  uint64_t* hkeys[9];
  int16_t* hvals[9];
} lco_fsst;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static inline uint64_t load_le(const uint8_t* p, size_t avail) {
  uint64_t w = 0;
  memcpy(&w, p, avail >= 8 ? 8 : avail);
  return w;
}
static inline uint64_t len_mask(int l) { return l >= 8 ? ~0ull : ((1ull << (8 * l)) - 1); }

EXPORT void lco_fsst_free(lco_fsst* t) {
  if (!t) return;
  for (int l = 1; l <= 8; ++l) { free(t->hkeys[l]); free(t->hvals[l]); }
  free(t);
}

static void fsst_index(lco_fsst* t) {
  t->hash_cap = 1024;
  for (int l = 1; l <= 8; ++l) {
    t->hkeys[l] = (uint64_t*)calloc(t->hash_cap, 8);
    t->hvals[l] = (int16_t*)malloc(t->hash_cap * 2);
    for (uint32_t i = 0; i < t->hash_cap; ++i) t->hvals[l][i] = -1;
  }
  for (int c = 0; c < t->n_symbols; ++c) {
    const int l = t->lens[c];
    uint32_t s = (uint32_t)mix64(t->symbols[c]) & (t->hash_cap - 1);
    while (t->hvals[l][s] >= 0) s = (s + 1) & (t->hash_cap - 1);
    t->hkeys[l][s] = t->symbols[c];
    t->hvals[l][s] = (int16_t)c;
  }
}

static inline int fsst_find(const lco_fsst* t, uint64_t w, size_t rem, int* len) {
  for (int l = rem >= 8 ? 8 : (int)rem; l >= 1; --l) {
    const uint64_t key = w & len_mask(l);
    uint32_t s = (uint32_t)mix64(key) & (t->hash_cap - 1);
    while (t->hvals[l][s] >= 0) {
      if (t->hkeys[l][s] == key) { *len = l; return t->hvals[l][s]; }
      s = (s + 1) & (t->hash_cap - 1);
    }
  }
  return -1;
}

/* Compressor::train restated: a few generations of "count symbols and adjacent pairs in a greedy parse of the
 * sample, keep the 255 candidates with the largest frequency x length". */
typedef struct { uint64_t val; uint32_t len; uint64_t gain; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->gain != y->gain) return x->gain > y->gain ? -1 : 1;
  if (x->len != y->len) return x->len > y->len ? -1 : 1;
  return x->val < y->val ? -1 : x->val > y->val;
}

EXPORT lco_fsst* lco_fsst_train(const uint8_t* data, const int32_t* offsets, uint32_t n) {
  lco_fsst* t = (lco_fsst*)calloc(1, sizeof(lco_fsst));
  fsst_index(t);
  /* sample: strided strings up to ~16 KiB */
  const size_t target = 16384;
  const uint32_t step = n > 256 ? n / 256 : 1;
  uint32_t* c1 = (uint32_t*)malloc(512 * 4);
  uint32_t* c2 = (uint32_t*)malloc(512 * 512 * 4);
  for (int gen = 0; gen < 5; ++gen) {
    memset(c1, 0, 512 * 4);
    memset(c2, 0, 512 * 512 * 4);
    size_t seen = 0;
    for (uint32_t i = 0; i < n && seen < target; i += step) {
      const uint8_t* p = data + offsets[i];
      size_t rem = (size_t)(offsets[i + 1] - offsets[i]);
      if (rem > 2048) rem = 2048;
      seen += rem;
      int prev = -1;
      while (rem) {
        int l = 1;
        const int s = fsst_find(t, load_le(p, rem), rem, &l);
        const int code = s >= 0 ? 256 + s : p[0];
        if (s < 0) l = 1;
        c1[code]++;
        if (prev >= 0) c2[prev * 512 + code]++;
        prev = code;
        p += l; rem -= l;
      }
    }
    size_t nc = 0, cap = 4096;
    cand_t* cs = (cand_t*)malloc(cap * sizeof(cand_t));
#define SYM_OF(code, V, L) do { if ((code) < 256) { V = (uint64_t)(code); L = 1; } else { V = t->symbols[(code) - 256]; L = t->lens[(code) - 256]; } } while (0)
    for (int a = 0; a < 512; ++a) {
      if (!c1[a]) continue;
      uint64_t va; uint32_t la; SYM_OF(a, va, la);
      if (nc + 513 > cap) { cap *= 2; cs = (cand_t*)realloc(cs, cap * sizeof(cand_t)); }
      cs[nc++] = (cand_t){va, la, (uint64_t)c1[a] * la};
      if (la >= 8 || gen == 4) continue;
      for (int b = 0; b < 512; ++b) {
        const uint32_t cnt = c2[a * 512 + b];
        if (cnt < 2) continue;
        uint64_t vb; uint32_t lb; SYM_OF(b, vb, lb);
        const uint32_t l = la + lb > 8 ? 8 : la + lb;
        cs[nc++] = (cand_t){(va | (vb << (8 * la))) & len_mask((int)l), l, (uint64_t)cnt * l};
      }
    }
    qsort(cs, nc, sizeof(cand_t), cand_cmp);
    for (int l = 1; l <= 8; ++l) { free(t->hkeys[l]); free(t->hvals[l]); }
    int ns = 0;
    for (size_t i = 0; i < nc && ns < 255; ++i) {
      int dup = 0;
      for (int j = 0; j < ns; ++j) if (t->symbols[j] == cs[i].val && t->lens[j] == cs[i].len) { dup = 1; break; }
      if (dup) continue;
      t->symbols[ns] = cs[i].val;
      t->lens[ns] = (uint8_t)cs[i].len;
      ++ns;
    }
    t->n_symbols = ns;
    free(cs);
    fsst_index(t);
  }
  free(c1); free(c2);
  return t;
}

/* Compressor::compress_into: greedy longest match, 0xFF escapes */
EXPORT size_t lco_fsst_compress(const lco_fsst* t, const uint8_t* in, size_t len, uint8_t* out) {
  size_t o = 0;
  while (len) {
    int l = 1;
    const int s = fsst_find(t, load_le(in, len), len, &l);
    if (s >= 0) { out[o++] = (uint8_t)s; in += l; len -= l; }
    else { out[o++] = 255; out[o++] = in[0]; in += 1; len -= 1; }
  }
  return o;
}

/* Decompressor::decompress_into: 8-byte store per code, advance by the symbol length */
EXPORT size_t lco_fsst_decompress(const lco_fsst* t, const uint8_t* in, size_t len, uint8_t* out) {
  uint8_t* o = out;
  const uint8_t* end = in + len;
  while (in < end) {
    const uint8_t c = *in++;
    if (c == 255) { *o++ = *in++; }
    else { memcpy(o, &t->symbols[c], 8); o += t->lens[c]; }
  }
  return (size_t)(o - out);
}

/* ------------------------------------------------------------------ byte-view entry ---- */
typedef struct {
  uint32_t n, n_unique;
  uint16_t* keys;        /* n */
  uint8_t* validity;     /* n bits or NULL */
  uint32_t null_count;
  uint64_t* prefix_keys; /* U x {prefix7, len} */
  uint32_t* fingerprints;/* U or NULL */
  uint32_t* offsets;     /* U+1 byte offsets into comp (CompactOffsets expanded; see lco_str_offset) */
  int32_t slope, intercept; int offset_bytes; void* residuals;
  uint8_t* comp; size_t comp_len;
  uint8_t* shared_prefix; uint32_t shared_prefix_len;
  uint64_t uncompressed_bytes;
  const lco_fsst* fsst;
} lco_str;

EXPORT void lco_str_free(lco_str* e) {
  if (!e) return;
  free(e->keys); free(e->validity); free(e->prefix_keys); free(e->fingerprints); free(e->offsets);
  free(e->residuals); free(e->comp); free(e->shared_prefix); free(e);
}

static inline uint32_t str_offset(const lco_str* e, uint32_t i) {
  /* CompactOffsets::get_offset */
  int32_t r = e->offset_bytes == 1 ? ((const int8_t*)e->residuals)[i]
            : e->offset_bytes == 2 ? ((const int16_t*)e->residuals)[i] : ((const int32_t*)e->residuals)[i];
  return (uint32_t)(e->slope * (int32_t)i + e->intercept + r);
}

/* from_dict_array_inner over a Utf8 batch (offsets/data), nulls via validity bitmap (may be NULL) */
EXPORT lco_str* lco_str_encode(const int32_t* offsets, const uint8_t* data, const uint8_t* validity, uint32_t n,
                               const lco_fsst* fsst, int build_fp) {
  lco_str* e = (lco_str*)calloc(1, sizeof(lco_str));
  e->n = n;
  e->fsst = fsst;
  e->keys = (uint16_t*)calloc(n ? n : 1, 2);
  /* u16 dictionary, first-occurrence order */
  uint32_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  uint32_t* slots = (uint32_t*)calloc(cap, 4);
  uint32_t* ufirst = (uint32_t*)malloc((size_t)(n ? n : 1) * 4); /* row index of the first occurrence */
  uint64_t* uhash = (uint64_t*)malloc((size_t)(n ? n : 1) * 8);
  uint32_t U = 0, nulls = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (validity && !bit_get(validity, i)) { ++nulls; continue; }
    const uint8_t* p = data + offsets[i];
    const uint32_t len = (uint32_t)(offsets[i + 1] - offsets[i]);
    uint64_t h = 0x9E3779B97F4A7C15ull ^ len;
    for (uint32_t b = 0; b + 8 <= len; b += 8) { uint64_t w; memcpy(&w, p + b, 8); h = mix64(h ^ w); }
    if (len & 7) h = mix64(h ^ load_le(p + (len & ~7u), len & 7));
    uint32_t s = (uint32_t)h & (cap - 1);
    int found = -1;
    while (slots[s]) {
      const uint32_t u = slots[s] - 1;
      const uint32_t r = ufirst[u];
      if (uhash[u] == h && (uint32_t)(offsets[r + 1] - offsets[r]) == len && memcmp(data + offsets[r], p, len) == 0) { found = (int)u; break; }
      s = (s + 1) & (cap - 1);
    }
    if (found < 0) { found = (int)U; slots[s] = U + 1; ufirst[U] = i; uhash[U] = h; ++U; }
    e->keys[i] = (uint16_t)found;
  }
  free(slots); free(uhash);
  e->n_unique = U;
  e->null_count = nulls;
  if (nulls) { e->validity = (uint8_t*)calloc((n + 7) / 8 + 8, 1); memcpy(e->validity, validity, (n + 7) / 8); }
  /* shared prefix */
  uint32_t spl = 0;
  if (U) {
    const uint8_t* p0 = data + offsets[ufirst[0]];
    spl = (uint32_t)(offsets[ufirst[0] + 1] - offsets[ufirst[0]]);
    for (uint32_t u = 1; u < U && spl; ++u) {
      const uint8_t* p = data + offsets[ufirst[u]];
      const uint32_t l = (uint32_t)(offsets[ufirst[u] + 1] - offsets[ufirst[u]]);
      uint32_t c = 0, m = l < spl ? l : spl;
      while (c < m && p[c] == p0[c]) ++c;
      spl = c;
    }
    e->shared_prefix = (uint8_t*)malloc(spl + 1);
    memcpy(e->shared_prefix, p0, spl);
  }
  e->shared_prefix_len = spl;
  /* compress uniques, prefix keys, fingerprints */
  size_t total = 0;
  for (uint32_t u = 0; u < U; ++u) total += (size_t)(offsets[ufirst[u] + 1] - offsets[ufirst[u]]);
  e->uncompressed_bytes = total;
  e->comp = (uint8_t*)malloc(2 * total + 16);
  e->offsets = (uint32_t*)malloc((size_t)(U + 1) * 4);
  e->prefix_keys = (uint64_t*)malloc((size_t)(U ? U : 1) * 8);
  if (build_fp) e->fingerprints = (uint32_t*)malloc((size_t)(U ? U : 1) * 4);
  size_t co = 0;
  e->offsets[0] = 0;
  for (uint32_t u = 0; u < U; ++u) {
    const uint8_t* p = data + offsets[ufirst[u]];
    const uint32_t len = (uint32_t)(offsets[ufirst[u] + 1] - offsets[ufirst[u]]);
    co += lco_fsst_compress(fsst, p, len, e->comp + co);
    e->offsets[u + 1] = (uint32_t)co;
    const uint32_t sl = len > spl ? len - spl : 0;
    uint64_t k = 0;
    for (uint32_t b = 0; b < (sl < 7 ? sl : 7); ++b) k |= (uint64_t)p[spl + b] << (8 * b);
    k |= (uint64_t)(sl >= 255 ? 255 : sl) << 56;
    e->prefix_keys[u] = k;
    if (build_fp) { uint32_t bits = 0; for (uint32_t b = 0; b < len; ++b) bits |= 1u << (p[b] & 31); e->fingerprints[u] = bits; }
  }
  e->comp_len = co;
  free(ufirst);
  /* CompactOffsets: fit_line in f64, residuals in 1/2/4 bytes */
  {
    const size_t m = (size_t)U + 1;
    double sum_y = 0, sum_xy = 0;
    for (size_t i = 0; i < m; ++i) sum_y += (double)e->offsets[i];
    for (size_t i = 0; i < m; ++i) sum_xy += (double)i * (double)e->offsets[i];
    int32_t slope = 0, intercept = (int32_t)e->offsets[0];
    if (m > 1) {
      const double nf = (double)m, sum_x = (double)(m * (m - 1) / 2), sum_x_sq = (double)(m * (m - 1) * (2 * m - 1) / 6);
      const double sl = (nf * sum_xy - sum_x * sum_y) / (nf * sum_x_sq - sum_x * sum_x);
      const double ic = (sum_y - sl * sum_x) / nf;
      slope = (int32_t)__builtin_round(sl);
      intercept = (int32_t)__builtin_round(ic);
    }
    e->slope = slope; e->intercept = intercept;
    int32_t* res = (int32_t*)malloc(m * 4);
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (size_t i = 0; i < m; ++i) {
      res[i] = (int32_t)(e->offsets[i] - ((uint32_t)slope * (uint32_t)i + (uint32_t)intercept));
      if (res[i] < lo) lo = res[i];
      if (res[i] > hi) hi = res[i];
    }
    e->offset_bytes = (lo >= -128 && hi <= 127) ? 1 : (lo >= -32768 && hi <= 32767) ? 2 : 4;
    e->residuals = malloc(m * (size_t)e->offset_bytes);
    for (size_t i = 0; i < m; ++i) {
      if (e->offset_bytes == 1) ((int8_t*)e->residuals)[i] = (int8_t)res[i];
      else if (e->offset_bytes == 2) ((int16_t*)e->residuals)[i] = (int16_t)res[i];
      else ((int32_t*)e->residuals)[i] = res[i];
    }
    free(res);
  }
  return e;
}

EXPORT uint64_t lco_str_bytes(const lco_str* e) {
  return (uint64_t)e->n * 2 + (uint64_t)e->n_unique * 8 + (e->fingerprints ? (uint64_t)e->n_unique * 4 : 0) +
         (uint64_t)(e->n_unique + 1) * (uint64_t)e->offset_bytes + e->comp_len + e->shared_prefix_len;
}

/* filter_inner: arrow filter of the u16 keys (+validity) */
static uint32_t filter_keys(const lco_str* e, const uint8_t* sel, uint16_t* out_keys, uint8_t* out_valid,
                            uint32_t* out_nulls) {
  return filter_prim(e->keys, e->validity, sel, e->n, 2, out_keys, out_valid, out_nulls);
}

/* compare_like_substring + map_dictionary_results_to_array_results.
 * pattern = inner bytes of '%x%'. negate = NOT LIKE (inverted only if a candidate exists, as in the reference). */
EXPORT uint32_t lco_str_like(const lco_str* e, const uint8_t* sel, const uint8_t* inner, uint32_t m, int negate,
                             uint8_t* out_mask, uint8_t* out_valid, uint32_t* out_nulls) {
  const uint32_t U = e->n_unique;
  uint16_t* keys = (uint16_t*)malloc((size_t)(e->n ? e->n : 1) * 2);
  if (out_valid) memset(out_valid, 0, (e->n + 7) / 8 + 1);
  const uint32_t k = filter_keys(e, sel, keys, out_valid, out_nulls);
  uint8_t* dict = (uint8_t*)calloc(U ? U : 1, 1);
  /* compute_fingerprint_candidates */
  uint32_t nfp = 0;
  for (uint32_t b = 0; b < m; ++b) nfp |= 1u << (inner[b] & 31);
  uint32_t* cand = (uint32_t*)malloc((size_t)(U ? U : 1) * 4);
  uint32_t nc = 0;
  for (uint32_t u = 0; u < U; ++u)
    if (!e->fingerprints || (e->fingerprints[u] & nfp) == nfp) cand[nc++] = u;
  if (nc) {
    /* to_uncompressed_selected: decompress the candidates back to back with offsets */
    uint8_t* buf = (uint8_t*)malloc(e->uncompressed_bytes + 16);
    uint32_t* off = (uint32_t*)malloc((size_t)(nc + 1) * 4);
    size_t o = 0;
    off[0] = 0;
    for (uint32_t c = 0; c < nc; ++c) {
      const uint32_t s = str_offset(e, cand[c]), t = str_offset(e, cand[c] + 1);
      o += lco_fsst_decompress(e->fsst, e->comp + s, t - s, buf + o);
      off[c + 1] = (uint32_t)o;
    }
    /* arrow LIKE '%x%' == substring search over each value */
    for (uint32_t c = 0; c < nc; ++c)
      if (memmem(buf + off[c], off[c + 1] - off[c], inner, m)) dict[cand[c]] = 1;
    if (negate) for (uint32_t u = 0; u < U; ++u) dict[u] = !dict[u];
    free(buf); free(off);
  }
  memset(out_mask, 0, (k + 7) / 8 + 1);
  for (uint32_t i = 0; i < k; ++i) {
    const int valid = out_valid && *out_nulls ? bit_get(out_valid, i) : 1;
    if (valid && dict[keys[i]]) out_mask[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
  free(cand); free(dict); free(keys);
  return k;
}

/* compare_equals / compare_not_equals (needle suffix <= 7 bytes only needs the prefix keys; longer needles
 * compare compressed bytes) */
EXPORT uint32_t lco_str_eq(const lco_str* e, const uint8_t* sel, const uint8_t* needle, uint32_t m, int negate,
                           uint8_t* out_mask, uint8_t* out_valid, uint32_t* out_nulls) {
  const uint32_t U = e->n_unique, spl = e->shared_prefix_len;
  uint16_t* keys = (uint16_t*)malloc((size_t)(e->n ? e->n : 1) * 2);
  if (out_valid) memset(out_valid, 0, (e->n + 7) / 8 + 1);
  const uint32_t k = filter_keys(e, sel, keys, out_valid, out_nulls);
  uint8_t* dict = (uint8_t*)calloc(U ? U : 1, 1);
  if (m >= spl && memcmp(needle, e->shared_prefix, spl) == 0) {
    const uint8_t* s = needle + spl;
    const uint32_t L = m - spl;
    uint64_t expect = 0;
    for (uint32_t b = 0; b < (L < 7 ? L : 7); ++b) expect |= (uint64_t)s[b] << (8 * b);
    expect |= (uint64_t)(L >= 255 ? 255 : L) << 56;
    if (L <= 7) {
      for (uint32_t u = 0; u < U; ++u) dict[u] = e->prefix_keys[u] == expect;
    } else {
      uint8_t* cn = (uint8_t*)malloc(2 * (size_t)m + 16);
      const size_t cl = lco_fsst_compress(e->fsst, needle, m, cn);
      for (uint32_t u = 0; u < U; ++u) {
        if (e->prefix_keys[u] != expect) continue;
        const uint32_t a = str_offset(e, u), b = str_offset(e, u + 1);
        dict[u] = (b - a == cl) && memcmp(e->comp + a, cn, cl) == 0;
      }
      free(cn);
    }
  }
  if (negate) for (uint32_t u = 0; u < U; ++u) dict[u] = !dict[u];
  memset(out_mask, 0, (k + 7) / 8 + 1);
  for (uint32_t i = 0; i < k; ++i) {
    const int valid = out_valid && *out_nulls ? bit_get(out_valid, i) : 1;
    if (valid && dict[keys[i]]) out_mask[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
  free(dict); free(keys);
  return k;
}

/* filter + to_arrow_array -> Utf8: keys filtered, referenced uniques decompressed (keyed when k < 2048 or
 * k < U, else all), then the dictionary is unpacked row by row (arrow cast Dictionary -> Utf8).
 * out_offsets has k+1 entries, out_data must hold the result (caller sizes it with lco_str_filter_bytes). */
EXPORT uint32_t lco_str_filter(const lco_str* e, const uint8_t* sel, int32_t* out_offsets, uint8_t* out_data,
                               uint8_t* out_valid, uint32_t* out_nulls, uint64_t* out_bytes) {
  const uint32_t U = e->n_unique;
  uint16_t* keys = (uint16_t*)malloc((size_t)(e->n ? e->n : 1) * 2);
  if (out_valid) memset(out_valid, 0, (e->n + 7) / 8 + 1);
  const uint32_t k = filter_keys(e, sel, keys, out_valid, out_nulls);
  uint8_t* used = (uint8_t*)calloc(U ? U : 1, 1);
  const int keyed = k < 2048 || k < U;
  if (keyed) { for (uint32_t i = 0; i < k; ++i) if (!*out_nulls || bit_get(out_valid, i)) used[keys[i]] = 1; }
  else memset(used, 1, U);
  uint8_t* dbuf = (uint8_t*)malloc(e->uncompressed_bytes + 16);
  uint32_t* doff = (uint32_t*)malloc((size_t)(U + 1) * 4);
  uint32_t* dlen = (uint32_t*)malloc((size_t)(U ? U : 1) * 4);
  size_t o = 0;
  for (uint32_t u = 0; u < U; ++u) {
    doff[u] = (uint32_t)o;
    dlen[u] = 0;
    if (!used[u]) continue;
    const uint32_t s = str_offset(e, u), t = str_offset(e, u + 1);
    const size_t l = lco_fsst_decompress(e->fsst, e->comp + s, t - s, dbuf + o);
    dlen[u] = (uint32_t)l;
    o += l;
  }
  uint64_t w = 0;
  for (uint32_t i = 0; i < k; ++i) {
    out_offsets[i] = (int32_t)w;
    if (*out_nulls && !bit_get(out_valid, i)) continue;
    const uint32_t u = keys[i];
    if (out_data) memcpy(out_data + w, dbuf + doff[u], dlen[u]);
    w += dlen[u];
  }
  out_offsets[k] = (int32_t)w;
  *out_bytes = w;
  free(used); free(dbuf); free(doff); free(dlen); free(keys);
  return k;
}

/* ------------------------------------------------------------------ caller side ---- */
/* boolean_buffer_and_then with BMI2 PDEP (datafusion/src/utils.rs:104-236) */
EXPORT void lco_and_then(const uint8_t* left, uint64_t left_len, const uint8_t* right, uint64_t right_len,
                         uint8_t* out) {
  if (left_len == right_len) { memcpy(out, right, (right_len + 7) / 8); return; }
  const uint64_t nw = (left_len + 63) / 64;
  uint64_t rpos = 0;
  for (uint64_t w = 0; w < nw; ++w) {
    uint64_t l = 0;
    const uint64_t rem_bytes = (left_len + 7) / 8 - w * 8;
    memcpy(&l, left + w * 8, rem_bytes >= 8 ? 8 : rem_bytes);
    const int cnt = __builtin_popcountll(l);
    uint64_t r = 0;
    if (cnt) {
      const uint64_t byte = rpos >> 3, sh = rpos & 7;
      const uint64_t avail = (right_len + 7) / 8 > byte ? (right_len + 7) / 8 - byte : 0;
      uint64_t lo = 0, hi = 0;
      memcpy(&lo, right + byte, avail >= 8 ? 8 : avail);
      if (avail > 8) memcpy(&hi, right + byte + 8, avail - 8 >= 8 ? 8 : avail - 8);
      r = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
    }
#ifdef __BMI2__
    const uint64_t o = _pdep_u64(r, l);
#else
    uint64_t o = 0, m = l, rr = r;
    while (m) { const uint64_t b = m & -m; if (rr & 1) o |= b; rr >>= 1; m &= m - 1; }
#endif
    memcpy(out + w * 8, &o, rem_bytes >= 8 ? 8 : rem_bytes);
    rpos += (uint64_t)cnt;
  }
}

/* ------------------------------------------------------------------ threaded scan drivers ---- */
/* The shape being matched: DataFusion partition tasks over LiquidCacheReader
 * (src/datafusion/src/reader/runtime/liquid_cache_reader.rs:297-391) — a fixed set of worker threads created ONCE
 * (the tokio runtime), each taking batches of a scan and running the per-entry path above on them. So the pool
 * below is persistent: threads are created by lco_pool_create, outside every timed region, and a pass only wakes
 * them. Entries are handed out dynamically in grains of a few entries (kinder to the CPU than a static split: no
 * tail imbalance), and every worker keeps its scratch buffers across passes. */
typedef struct {
  const void* const* entries; uint32_t n_entries;
  int kind; /* 0 str like, 1 int range (two conjuncts + and_then), 2 int decode, 3 str like + get,
               5 = 1 + get of the column with the final selection,
               4 two-column batch: range on column A (two conjuncts), third conjunct on column B, get(B, A) */
  const uint8_t* needle; uint32_t m;
  int op1, op2; int64_t lit1, lit2;
  uint32_t grain;
  const void* const* entries2; int op3; int64_t lit3;  /* kind 4: second column of the batch and its conjunct */
} scan_job;

typedef struct {
  uint8_t *mask, *valid, *mask2, *sel; size_t cap;       /* per-entry bitmaps */
  int32_t* off; size_t off_cap; uint8_t* data; size_t data_cap; void* vals; size_t vals_cap; /* get outputs */
} scan_scratch;

static void scratch_free(scan_scratch* s) {
  free(s->mask); free(s->valid); free(s->mask2); free(s->sel); free(s->off); free(s->data); free(s->vals);
  memset(s, 0, sizeof *s);
}

/* the reference's per-batch work on entry i of the job; returns rows whose final mask bit is set */
static uint64_t scan_one(const scan_job* a, uint32_t i, scan_scratch* s, uint64_t* rows) {
  uint64_t matched = 0;
  const uint32_t n = a->kind == 1 || a->kind == 2 || a->kind == 4 || a->kind == 5 ? ((const lco_int*)a->entries[i])->n : ((const lco_str*)a->entries[i])->n;
  if ((size_t)n / 8 + 64 > s->cap) {
    s->cap = (size_t)n / 8 + 64;
    s->mask = (uint8_t*)realloc(s->mask, s->cap); s->valid = (uint8_t*)realloc(s->valid, s->cap);
    s->mask2 = (uint8_t*)realloc(s->mask2, s->cap); s->sel = (uint8_t*)realloc(s->sel, s->cap);
  }
  uint8_t *mask = s->mask, *valid = s->valid, *mask2 = s->mask2, *sel = s->sel;
  uint32_t nulls = 0;
  *rows += n;
  if (a->kind == 0 || a->kind == 3) {
    const lco_str* e = (const lco_str*)a->entries[i];
    const uint32_t k = lco_str_like(e, NULL, a->needle, a->m, 0, mask, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k + 7) / 8; ++b) mask[b] &= valid[b];  /* prep_null_mask_filter */
    uint64_t hit = 0;
    for (uint32_t b = 0; b < (k + 7) / 8; ++b) hit += (uint64_t)__builtin_popcount(mask[b]);
    matched += hit;
    if (a->kind == 3 && hit) {
      const size_t need_off = (size_t)(hit + 1) * 4;
      const size_t need_data = e->uncompressed_bytes + 16 > 1 << 20 ? e->uncompressed_bytes + 16 : 1 << 20;
      if (need_off > s->off_cap) { s->off_cap = need_off; s->off = (int32_t*)realloc(s->off, need_off); }
      if (need_data > s->data_cap) { s->data_cap = need_data; s->data = (uint8_t*)realloc(s->data, need_data); }
      uint64_t bytes = 0;
      lco_str_filter(e, mask, s->off, s->data, valid, &nulls, &bytes);
    }
  } else if (a->kind == 1 || a->kind == 5) {
    const lco_int* e = (const lco_int*)a->entries[i];
    uint32_t k = lco_int_eval(e, NULL, a->op1, a->lit1, (uint64_t)a->lit1, mask, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k + 7) / 8; ++b) mask[b] &= valid[b];
    memcpy(sel, mask, (k + 7) / 8);
    /* second conjunct under the running selection, then boolean_buffer_and_then */
    const uint32_t k2 = lco_int_eval(e, sel, a->op2, a->lit2, (uint64_t)a->lit2, mask2, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k2 + 7) / 8; ++b) mask2[b] &= valid[b];
    lco_and_then(sel, n, mask2, k2, mask);
    uint64_t hit = 0;
    for (uint32_t b = 0; b < (n + 7) / 8; ++b) hit += (uint64_t)__builtin_popcount(mask[b]);
    matched += hit;
    if (a->kind == 5 && hit) {  /* read_from_cache: the projected column with the final selection */
      const size_t need = (size_t)(n ? n : 1) * (e->tbits / 8);
      if (need > s->vals_cap) { s->vals_cap = need; s->vals = realloc(s->vals, need); }
      lco_int_filter(e, mask, s->vals, valid, &nulls);
    }
  } else if (a->kind == 4) {
    /* the reader's loop for one batch (liquid_cache_reader.rs:297-391): each conjunct under the running selection,
     * nulls->false, boolean_buffer_and_then; then every projected column read with the final selection */
    const lco_int* ea = (const lco_int*)a->entries[i];
    const lco_int* eb = (const lco_int*)a->entries2[i];
    uint32_t k = lco_int_eval(ea, NULL, a->op1, a->lit1, (uint64_t)a->lit1, mask, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k + 7) / 8; ++b) mask[b] &= valid[b];
    memcpy(sel, mask, (k + 7) / 8);
    k = lco_int_eval(ea, sel, a->op2, a->lit2, (uint64_t)a->lit2, mask2, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k + 7) / 8; ++b) mask2[b] &= valid[b];
    lco_and_then(sel, n, mask2, k, mask);
    memcpy(sel, mask, ((size_t)n + 7) / 8);
    k = lco_int_eval(eb, sel, a->op3, a->lit3, (uint64_t)a->lit3, mask2, valid, &nulls);
    if (nulls) for (uint32_t b = 0; b < (k + 7) / 8; ++b) mask2[b] &= valid[b];
    lco_and_then(sel, n, mask2, k, mask);
    uint64_t hit = 0;
    for (uint32_t b = 0; b < (n + 7) / 8; ++b) hit += (uint64_t)__builtin_popcount(mask[b]);
    matched += hit;
    const size_t need = (size_t)(n ? n : 1) * 8;
    if (need > s->vals_cap) { s->vals_cap = need; s->vals = realloc(s->vals, need); }
    if (hit) {
      lco_int_filter(eb, mask, s->vals, valid, &nulls);
      lco_int_filter(ea, mask, s->vals, valid, &nulls);
    }
  } else {
    const lco_int* e = (const lco_int*)a->entries[i];
    const size_t need = (size_t)(n ? n : 1) * (e->tbits / 8);
    if (need > s->vals_cap) { s->vals_cap = need; s->vals = realloc(s->vals, need); }
    matched += lco_int_filter(e, NULL, s->vals, valid, &nulls);
  }
  return matched;
}

typedef struct lco_pool {
  pthread_t* th; uint32_t nthreads;
  pthread_mutex_t mu; pthread_cond_t cv_go, cv_done;
  uint64_t generation; uint32_t running; int stop;
  scan_job job;
  const void* const* entries2; int op3; int64_t lit3;   /* lco_pool_set_second */
  volatile uint32_t next;            /* next entry to hand out (atomic fetch-add) */
  uint64_t matched, rows;            /* totals of the current pass (under mu) */
} lco_pool;

static void* pool_worker(void* p) {
  lco_pool* pool = (lco_pool*)p;
  scan_scratch s; memset(&s, 0, sizeof s);
  uint64_t seen = 0;
  for (;;) {
    pthread_mutex_lock(&pool->mu);
    while (!pool->stop && pool->generation == seen) pthread_cond_wait(&pool->cv_go, &pool->mu);
    if (pool->stop) { pthread_mutex_unlock(&pool->mu); break; }
    seen = pool->generation;
    const scan_job job = pool->job;
    pthread_mutex_unlock(&pool->mu);
    uint64_t matched = 0, rows = 0;
    for (;;) {
      const uint32_t b = __atomic_fetch_add(&pool->next, job.grain, __ATOMIC_RELAXED);
      if (b >= job.n_entries) break;
      const uint32_t e = b + job.grain < job.n_entries ? b + job.grain : job.n_entries;
      for (uint32_t i = b; i < e; ++i) matched += scan_one(&job, i, &s, &rows);
    }
    pthread_mutex_lock(&pool->mu);
    pool->matched += matched; pool->rows += rows;
    if (--pool->running == 0) pthread_cond_signal(&pool->cv_done);
    pthread_mutex_unlock(&pool->mu);
  }
  scratch_free(&s);
  return NULL;
}

EXPORT lco_pool* lco_pool_create(uint32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  lco_pool* pool = (lco_pool*)calloc(1, sizeof(lco_pool));
  pool->nthreads = nthreads;
  pool->th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  pthread_mutex_init(&pool->mu, NULL);
  pthread_cond_init(&pool->cv_go, NULL); pthread_cond_init(&pool->cv_done, NULL);
  for (uint32_t t = 0; t < nthreads; ++t) pthread_create(&pool->th[t], NULL, pool_worker, pool);
  return pool;
}

EXPORT void lco_pool_destroy(lco_pool* pool) {
  if (!pool) return;
  pthread_mutex_lock(&pool->mu);
  pool->stop = 1;
  pthread_cond_broadcast(&pool->cv_go);
  pthread_mutex_unlock(&pool->mu);
  for (uint32_t t = 0; t < pool->nthreads; ++t) pthread_join(pool->th[t], NULL);
  pthread_mutex_destroy(&pool->mu); pthread_cond_destroy(&pool->cv_go); pthread_cond_destroy(&pool->cv_done);
  free(pool->th); free(pool);
}

/* kind 4 only: the batch's second column (same length as the entry list) and the conjunct evaluated on it */
EXPORT void lco_pool_set_second(lco_pool* pool, const void* const* entries2, int op3, int64_t lit3) {
  pool->entries2 = entries2; pool->op3 = op3; pool->lit3 = lit3;
}

/* One pass of the scan over all entries on the pool's threads. Nothing is created or joined here. */
EXPORT uint64_t lco_pool_scan(lco_pool* pool, const void* const* entries, uint32_t n_entries, int kind, const uint8_t* needle,
                              uint32_t m, int op1, int64_t lit1, int op2, int64_t lit2, uint32_t grain, uint64_t* out_rows) {
  pthread_mutex_lock(&pool->mu);
  pool->job = (scan_job){entries, n_entries, kind, needle, m, op1, op2, lit1, lit2, grain ? grain : 4,
                         pool->entries2, pool->op3, pool->lit3};
  pool->next = 0; pool->matched = 0; pool->rows = 0;
  pool->running = pool->nthreads;
  pool->generation++;
  pthread_cond_broadcast(&pool->cv_go);
  while (pool->running) pthread_cond_wait(&pool->cv_done, &pool->mu);
  const uint64_t matched = pool->matched;
  if (out_rows) *out_rows = pool->rows;
  pthread_mutex_unlock(&pool->mu);
  return matched;
}

/* Single pass on the calling thread (no pool): the single-thread figure reported next to the pooled one. */
EXPORT uint64_t lco_scan_serial(const void* const* entries, uint32_t n_entries, int kind, const uint8_t* needle, uint32_t m,
                                int op1, int64_t lit1, int op2, int64_t lit2, uint64_t* out_rows) {
  const scan_job job = {entries, n_entries, kind, needle, m, op1, op2, lit1, lit2, 1, NULL, 0, 0};
  if (kind == 4) return 0; /* two-column batches go through a pool (one thread if need be) */
  scan_scratch s; memset(&s, 0, sizeof s);
  uint64_t matched = 0, rows = 0;
  for (uint32_t i = 0; i < n_entries; ++i) matched += scan_one(&job, i, &s, &rows);
  scratch_free(&s);
  if (out_rows) *out_rows = rows;
  return matched;
}

/* Convenience for the tests: a throw-away pool around one pass. NOT used by any timed path. */
EXPORT uint64_t lco_scan(const void* const* entries, uint32_t n_entries, int kind, const uint8_t* needle, uint32_t m,
                         int op1, int64_t lit1, int op2, int64_t lit2, uint32_t nthreads, uint64_t* out_rows) {
  lco_pool* pool = lco_pool_create(nthreads);
  const uint64_t matched = lco_pool_scan(pool, entries, n_entries, kind, needle, m, op1, lit1, op2, lit2, 4, out_rows);
  lco_pool_destroy(pool);
  return matched;
}
