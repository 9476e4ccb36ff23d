/*
 * lc_synth.c — deterministic synthetic ClickBench-hits / TPC-H shaped columns.
 * Bench / test infrastructure only. One call produces ONE 8192-row Arrow batch (entry) so 100 M rows never
 * have to exist on the host at once.
 *
 *   URL        Utf8: scheme + host token + path/query tokens (Cyrillic percent-escapes as in nano_hits), length
 *              ~ lognormal(mean 76 B, cap 500), ~1 900 distinct values per 8192-row entry drawn Zipf(1.2) from a
 *              per-entry pool, token "google" injected into a fraction `inject_p` of the rows (config 2)
 *   EventTime  Int64: 1373832014 + uniform in an 86 400 s window (W = 17)                        (config 3)
 *   UserID     Int64: 2^17 distinct values spread over the full i64 range (W = 64)              (config 3)
 *   l_shipdate Date32: uniform in [8036, 10556] (1992-01-02 .. 1998-12-01, W = 12)              (config 4)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t* s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t below(uint64_t* s, uint32_t n) { return (uint32_t)(((splitmix(s) >> 32) * (uint64_t)n) >> 32); }

static const char* HOST_WORDS[] = {"auto", "news", "mail", "yandex", "rambler", "vk", "ok", "avito", "market", "kino",
                                   "forum", "blog", "shop", "moto", "dom", "tv", "map", "foto", "job", "game",
                                   "travel", "med", "bank", "real", "sport", "love", "soft", "mob", "book", "music"};
static const char* TLDS[] = {".ru", ".com", ".net", ".org", ".ua", ".by", ".kz", ".info"};
static const char* PATH_TOK[] = {
    "index", "catalog", "item", "view", "search", "page", "news", "article", "user", "profile", "photo", "video",
    "list", "category", "product", "cart", "order", "forum", "topic", "thread", "post", "tag", "archive", "2013",
    "07", "15", "id", "ru", "en", "main", "mobile", "api", "v1", "img", "static", "files", "download", "rss",
    "%D0%BA%D1%83%D0%BF%D0%B8%D1%82%D1%8C", "%D0%B0%D0%B2%D1%82%D0%BE", "%D0%BC%D0%BE%D1%81%D0%BA%D0%B2%D0%B0",
    "%D0%BD%D0%BE%D0%B2%D0%BE%D1%81%D1%82%D0%B8", "%D1%84%D0%BE%D1%82%D0%BE", "%D1%86%D0%B5%D0%BD%D0%B0",
    "%D0%BE%D1%82%D0%B7%D1%8B%D0%B2%D1%8B", "%D1%81%D0%BA%D0%B0%D1%87%D0%B0%D1%82%D1%8C"};
static const char* QUERY_KEYS[] = {"id", "page", "q", "ref", "utm_source", "sid", "cat", "sort", "from", "lang"};
#define NHOSTW (sizeof(HOST_WORDS) / sizeof(HOST_WORDS[0]))
#define NTLD (sizeof(TLDS) / sizeof(TLDS[0]))
#define NPATH (sizeof(PATH_TOK) / sizeof(PATH_TOK[0]))
#define NQK (sizeof(QUERY_KEYS) / sizeof(QUERY_KEYS[0]))

static uint32_t put(char* dst, uint32_t pos, uint32_t cap, const char* s) {
  while (*s && pos < cap) dst[pos++] = *s++;
  return pos;
}

/* one URL into dst (cap bytes), returns its length */
static uint32_t make_url(uint64_t* rng, char* dst, uint32_t cap, int with_google) {
  /* lognormal target length: mean 76 -> mu = ln(76) - sigma^2/2, sigma = 0.6 */
  const double z = sqrt(-2.0 * log(u01(rng) + 1e-12)) * cos(6.283185307179586 * u01(rng));
  double target = exp(4.1507 + 0.6 * z);
  if (target > 500.0) target = 500.0;
  if (target < 12.0) target = 12.0;
  uint32_t pos = 0;
  pos = put(dst, pos, cap, (splitmix(rng) & 3) ? "http://" : "https://");
  if (splitmix(rng) & 1) pos = put(dst, pos, cap, "www.");
  if (with_google) {
    pos = put(dst, pos, cap, "google");
  } else {
    char host[48];
    /* ~2 000 host tokens: word + 0..66 */
    snprintf(host, sizeof(host), "%s%u", HOST_WORDS[below(rng, NHOSTW)], below(rng, 67));
    pos = put(dst, pos, cap, host);
  }
  pos = put(dst, pos, cap, TLDS[below(rng, NTLD)]);
  while (pos < (uint32_t)target && pos + 40 < cap) {
    pos = put(dst, pos, cap, "/");
    pos = put(dst, pos, cap, PATH_TOK[below(rng, NPATH)]);
    if ((splitmix(rng) & 7) == 0) {
      char num[16];
      snprintf(num, sizeof(num), "%u", below(rng, 1000000));
      pos = put(dst, pos, cap, num);
    }
  }
  if ((splitmix(rng) & 3) == 0 && pos + 24 < cap) {
    char q[40];
    snprintf(q, sizeof(q), "?%s=%u", QUERY_KEYS[below(rng, NQK)], below(rng, 100000));
    pos = put(dst, pos, cap, q);
  }
  if (pos > 500) pos = 500;
  return pos;
}

/* Zipf(1.2) cumulative weights over a pool of `pool` ranks, built once */
static double* g_zipf = NULL;
static uint32_t g_zipf_n = 0;
static void zipf_init(uint32_t pool) {
  if (g_zipf && g_zipf_n == pool) return;  /* call lcs_init() once before generating from several threads */
  free(g_zipf);
  g_zipf = (double*)malloc(sizeof(double) * pool);
  double acc = 0;
  for (uint32_t r = 0; r < pool; ++r) {
    acc += 1.0 / pow((double)(r + 1), 1.2);
    g_zipf[r] = acc;
  }
  for (uint32_t r = 0; r < pool; ++r) g_zipf[r] /= acc;
  g_zipf_n = pool;
}
static uint32_t zipf_draw(uint64_t* rng) {
  const double u = u01(rng);
  uint32_t lo = 0, hi = g_zipf_n - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) / 2;
    if (g_zipf[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

EXPORT void lcs_init(uint32_t pool) { zipf_init(pool); }

/* URL column entry. offsets: rows+1 int32; data: caller buffer of data_cap bytes (rows*512 is always enough).
 * Returns bytes written. pool = size of the per-entry unique pool (8000 gives ~1 900 distinct per 8192 rows). */
EXPORT uint64_t lcs_url_entry(uint64_t seed, uint64_t entry_idx, uint32_t rows, uint32_t pool, double inject_p,
                              int32_t* offsets, uint8_t* data, uint64_t data_cap) {
  zipf_init(pool);
  uint64_t rng = seed * 0x9E3779B97F4A7C15ull + entry_idx * 0xD1B54A32D192ED03ull + 1;
  /* pool of unique URLs, generated lazily */
  char* pool_buf = (char*)malloc((size_t)pool * 504);
  uint16_t* pool_len = (uint16_t*)calloc(pool, 2);
  uint64_t pos = 0;
  for (uint32_t i = 0; i < rows; ++i) {
    offsets[i] = (int32_t)pos;
    if (pos + 512 > data_cap) break;
    if (inject_p > 0 && u01(&rng) < inject_p) {
      pos += make_url(&rng, (char*)data + pos, 504, 1);
      continue;
    }
    const uint32_t r = zipf_draw(&rng);
    if (!pool_len[r]) {
      uint64_t prng = seed ^ (entry_idx * 0xA24BAED4963EE407ull + r * 0x9FB21C651E98DF25ull);
      pool_len[r] = (uint16_t)make_url(&prng, pool_buf + (size_t)r * 504, 504, 0);
    }
    memcpy(data + pos, pool_buf + (size_t)r * 504, pool_len[r]);
    pos += pool_len[r];
  }
  offsets[rows] = (int32_t)pos;
  free(pool_buf);
  free(pool_len);
  return pos;
}

/* kind 0: EventTime  kind 1: UserID  kind 2: l_shipdate (int32 out)  kind 3: small-range Int16 flag column */
EXPORT void lcs_int_entry(uint64_t seed, uint64_t entry_idx, uint32_t rows, int kind, void* out) {
  uint64_t rng = seed * 0x9E3779B97F4A7C15ull + entry_idx * 0xD1B54A32D192ED03ull + 7 + (uint64_t)kind;
  if (kind == 0) {
    int64_t* o = (int64_t*)out;
    for (uint32_t i = 0; i < rows; ++i) o[i] = 1373832014ll + (int64_t)below(&rng, 86400);
  } else if (kind == 1) {
    int64_t* o = (int64_t*)out;
    for (uint32_t i = 0; i < rows; ++i) {
      uint64_t id = below(&rng, 1u << 17);
      uint64_t h = id * 0x9E3779B97F4A7C15ull;  /* spread the 2^17 ids over the whole i64 range */
      h ^= h >> 29;
      o[i] = (int64_t)(h * 0xBF58476D1CE4E5B9ull);
    }
  } else if (kind == 2) {
    int32_t* o = (int32_t*)out;
    for (uint32_t i = 0; i < rows; ++i) o[i] = 8036 + (int32_t)below(&rng, 10556 - 8036 + 1);
  } else {
    int16_t* o = (int16_t*)out;
    for (uint32_t i = 0; i < rows; ++i) o[i] = (splitmix(&rng) & 15) ? 0 : (int16_t)(1 + below(&rng, 15));
  }
}
