// fsst_host.cc — FSST symbol-table construction (host) for the byte-view encoding.
//
// The reference trains with the un-vendored crate fsst-rs 0.5.10 (Cargo.lock:3066-3069; call sites
// /root/reference/src/core/src/liquid_array/raw/fsst_buffer.rs:391-397 train, :73 compress_into,
// :104 decompress_into). This file restates the PUBLISHED algorithm (Boncz, Neumann, Leis: "FSST:
// Fast Random Access String Compression", VLDB 2020): <= 255 symbols of 1..8 bytes, code 255 =
// escape + literal byte, table built bottom-up over a few generations by counting symbols and
// adjacent symbol pairs in a greedy parse of a sample and keeping the candidates with the highest
// gain = frequency x length. Table contents never influence decoded output; every result the
// reference exposes on this path (get / eval_predicate) is independent of them.
//
// Training runs once per column chunk (compressor scope) and is output-invisible, so it stays on
// the host; compression of the unique values at insert time runs on the device (k_fsst.cu) against
// the lookup tables built here, and fsst_compress_host below is the same greedy matcher used while
// training and for cross-checking the kernel.
#include <algorithm>

#include "host_common.h"

namespace lc {

static inline uint32_t fsst_hash3(uint64_t w) {
  uint64_t h = (w & 0xFFFFFFull) * 2971215073ull;
  return static_cast<uint32_t>((h ^ (h >> 15)) & 2047u);
}

static inline uint64_t load_le(const uint8_t* p, size_t avail) {
  uint64_t w = 0;
  if (avail >= 8) {
    std::memcpy(&w, p, 8);
  } else {
    std::memcpy(&w, p, avail);
  }
  return w;
}

static inline uint64_t len_mask(uint32_t len) { return len >= 8 ? ~0ull : ((1ull << (8 * len)) - 1ull); }

namespace {

struct Sym {
  uint64_t val;
  uint32_t len;
};

// Greedy matcher over a set of symbols: long symbols (3..8 bytes) through a lossy hash on the first
// three bytes (one symbol per bucket), then the 2-byte table, then the 1-byte table, else escape.
// Used while training; the final table has the same structure (FsstEncTable).
struct Matcher {
  std::vector<uint64_t> hash_sym;
  std::vector<uint16_t> hash_meta;   // sym | len<<8 ; 0 = empty (len >= 3 makes a used slot non-zero)
  std::vector<uint16_t> short_meta;  // sym | len<<8 ; len 0 = no symbol
  uint16_t one_byte[256];
  Matcher() : hash_sym(2048, 0), hash_meta(2048, 0), short_meta(65536, 0) { std::memset(one_byte, 0, sizeof(one_byte)); }

  bool can_insert(const Sym& s) const {
    if (s.len >= 3) return hash_meta[fsst_hash3(s.val)] == 0;
    return true;
  }
  void insert(const Sym& s, uint32_t sym) {
    const uint16_t m = static_cast<uint16_t>(sym | (s.len << 8));
    if (s.len >= 3) {
      uint32_t h = fsst_hash3(s.val);
      hash_sym[h] = s.val;
      hash_meta[h] = m;
    } else if (s.len == 2) {
      short_meta[s.val & 0xFFFF] = m;
    } else {
      one_byte[s.val & 0xFF] = m;
      for (uint32_t x = 0; x < 256; ++x) {
        uint32_t k = static_cast<uint32_t>(s.val & 0xFF) | (x << 8);
        if ((short_meta[k] >> 8) != 2) short_meta[k] = m;
      }
    }
  }
  // returns symbol number or -1 (escape); *len = bytes consumed
  inline int match(uint64_t w, size_t remaining, uint32_t* len) const {
    if (remaining >= 3) {
      const uint32_t h = fsst_hash3(w);
      const uint16_t m = hash_meta[h];
      const uint32_t l = m >> 8;
      if (l && l <= remaining && (w & len_mask(l)) == hash_sym[h]) {
        *len = l;
        return m & 0xFF;
      }
    }
    uint16_t m = short_meta[w & 0xFFFF];
    if ((m >> 8) == 2 && remaining < 2) m = one_byte[w & 0xFF];
    if (m >> 8) {
      *len = m >> 8;
      return m & 0xFF;
    }
    *len = 1;
    return -1;
  }
};

}  // namespace

static void build_enc(const std::vector<Sym>& syms, FsstCodec* out) {
  out->enc.reset(new FsstEncTable());
  FsstEncTable& e = *out->enc;
  std::memset(&e, 0, sizeof(e));
  std::memset(&out->dec, 0, sizeof(out->dec));
  out->dec.n_symbols = static_cast<uint32_t>(syms.size());
  for (uint32_t i = 0; i < syms.size(); ++i) {
    out->dec.symbols[i] = syms[i].val;
    out->dec.lens[i] = static_cast<uint8_t>(syms[i].len);
  }
  // short table: len-2 symbols first, then len-1 fill
  for (uint32_t i = 0; i < syms.size(); ++i)
    if (syms[i].len == 2) e.short_code[syms[i].val & 0xFFFF] = static_cast<uint16_t>(i | (2u << 8));
  for (uint32_t i = 0; i < syms.size(); ++i) {
    if (syms[i].len != 1) continue;
    for (uint32_t x = 0; x < 256; ++x) {
      uint32_t k = static_cast<uint32_t>(syms[i].val & 0xFF) | (x << 8);
      if ((e.short_code[k] >> 8) != 2) e.short_code[k] = static_cast<uint16_t>(i | (1u << 8));
    }
    e.one_byte[syms[i].val & 0xFF] = static_cast<uint16_t>(i | (1u << 8));
  }
  for (uint32_t i = 0; i < syms.size(); ++i) {
    if (syms[i].len < 3) continue;
    uint32_t h = fsst_hash3(syms[i].val);
    e.hash_sym[h] = syms[i].val;
    e.hash_meta[h] = static_cast<uint16_t>(i | (syms[i].len << 8));
  }
}

// A codec from a stored symbol table (load_symbol_table, raw/fsst_buffer.rs:886-932). The decode view is exact; the encode
// lookup keeps one symbol per 3-byte hash bucket like the trained tables do (a table written by fsst-rs may have more, which
// only costs compression of LATER inserts under this scope, never a result).
void fsst_from_symbols(const uint64_t* vals, const uint8_t* lens, size_t n, FsstCodec* out) {
  std::vector<Sym> syms(n);
  for (size_t i = 0; i < n; ++i) {
    syms[i].len = lens[i];
    syms[i].val = vals[i] & len_mask(lens[i]);
  }
  build_enc(syms, out);
}

void fsst_train(const uint8_t* const* strs, const uint32_t* lens, size_t n, FsstCodec* out) {
  // ---- sample (~16 KiB, evenly strided over the input) ----
  constexpr size_t kSampleTarget = 16384;
  size_t total = 0;
  for (size_t i = 0; i < n; ++i) total += lens[i];
  std::vector<std::pair<const uint8_t*, uint32_t>> sample;
  if (total <= kSampleTarget * 2) {
    for (size_t i = 0; i < n; ++i)
      if (lens[i]) sample.emplace_back(strs[i], lens[i]);
  } else {
    const double avg = static_cast<double>(total) / static_cast<double>(n);
    size_t want = static_cast<size_t>(static_cast<double>(kSampleTarget) / (avg > 1 ? avg : 1)) + 1;
    size_t step = n / want ? n / want : 1;
    size_t got = 0;
    for (size_t i = 0; i < n && got < kSampleTarget; i += step) {
      if (!lens[i]) continue;
      uint32_t l = lens[i] > 2048 ? 2048 : lens[i];
      sample.emplace_back(strs[i], l);
      got += l;
    }
  }

  std::vector<Sym> syms;  // current real symbols
  std::vector<uint32_t> count1(512);
  std::vector<uint32_t> count2(512 * 512);
  for (int gen = 0; gen < 5; ++gen) {
    Matcher m;
    for (uint32_t i = 0; i < syms.size(); ++i) m.insert(syms[i], i);
    std::fill(count1.begin(), count1.end(), 0);
    std::fill(count2.begin(), count2.end(), 0);
    // greedy parse of the sample, counting codes and adjacent pairs; codes < 256 are raw bytes
    for (auto& s : sample) {
      const uint8_t* p = s.first;
      size_t rem = s.second;
      int prev = -1;
      while (rem) {
        uint32_t l;
        int idx = m.match(load_le(p, rem), rem, &l);
        int code = idx >= 0 ? 256 + idx : static_cast<int>(p[0]);
        count1[code]++;
        if (prev >= 0) count2[prev * 512 + code]++;
        prev = code;
        p += l;
        rem -= l;
      }
    }
    auto sym_of = [&](int code) -> Sym {
      if (code < 256) return Sym{static_cast<uint64_t>(code), 1};
      return syms[code - 256];
    };
    // candidates: every used code, and every adjacent pair concatenated (truncated to 8 bytes)
    struct Cand {
      uint64_t val;
      uint32_t len;
      uint64_t gain;
    };
    std::unordered_map<uint64_t, Cand> cands;  // key mixes val and len
    auto add = [&](const Sym& s, uint64_t cnt) {
      if (!cnt) return;
      uint64_t key = s.val * 0x9E3779B97F4A7C15ull + s.len;
      auto it = cands.find(key);
      uint64_t gain = cnt * s.len;
      if (it == cands.end()) cands.emplace(key, Cand{s.val, s.len, gain});
      else if (it->second.val == s.val && it->second.len == s.len) it->second.gain += gain;
    };
    for (int c1 = 0; c1 < 512; ++c1) {
      if (!count1[c1]) continue;
      const Sym s1 = sym_of(c1);
      add(s1, count1[c1]);
      if (s1.len >= 8 || gen == 4) continue;  // last generation: keep, do not grow
      for (int c2 = 0; c2 < 512; ++c2) {
        uint32_t cnt = count2[c1 * 512 + c2];
        if (!cnt) continue;
        const Sym s2 = sym_of(c2);
        uint32_t l = s1.len + s2.len > 8 ? 8 : s1.len + s2.len;
        uint64_t v = (s1.val | (s2.val << (8 * s1.len))) & len_mask(l);
        add(Sym{v, l}, cnt);
      }
    }
    std::vector<Cand> order;
    order.reserve(cands.size());
    for (auto& kv : cands) order.push_back(kv.second);
    std::sort(order.begin(), order.end(), [](const Cand& a, const Cand& b) {
      if (a.gain != b.gain) return a.gain > b.gain;
      if (a.len != b.len) return a.len > b.len;
      return a.val < b.val;
    });
    std::vector<Sym> next;
    Matcher probe;
    for (auto& c : order) {
      if (next.size() >= 255) break;
      if (c.gain < 2 && c.len > 1) continue;  // a multi-byte symbol seen once is noise
      Sym s{c.val, c.len};
      if (!probe.can_insert(s)) continue;  // lossy hash: one long symbol per bucket
      probe.insert(s, static_cast<uint32_t>(next.size()));
      next.push_back(s);
    }
    syms.swap(next);
  }
  build_enc(syms, out);
}

// Greedy compressor over the final table; identical decisions to the device kernel (k_fsst.cu).
size_t fsst_compress_host(const FsstCodec& c, const uint8_t* in, size_t len, uint8_t* out) {
  const FsstEncTable& e = *c.enc;
  size_t o = 0;
  const uint8_t* p = in;
  size_t rem = len;
  while (rem) {
    const uint64_t w = load_le(p, rem);
    uint32_t l = 0;
    int code = -1;
    if (rem >= 3) {
      const uint32_t h = fsst_hash3(w);
      const uint16_t m = e.hash_meta[h];
      const uint32_t ml = m >> 8;
      if (ml && ml <= rem && (w & len_mask(ml)) == e.hash_sym[h]) {
        code = m & 0xFF;
        l = ml;
      }
    }
    if (code < 0) {
      uint16_t m = e.short_code[w & 0xFFFF];
      if ((m >> 8) == 2 && rem < 2) m = e.one_byte[w & 0xFF];
      const uint32_t ml = m >> 8;
      if (ml) {
        code = m & 0xFF;
        l = ml;
      }
    }
    if (code >= 0) {
      out[o++] = static_cast<uint8_t>(code);
      p += l;
      rem -= l;
    } else {
      out[o++] = 255;
      out[o++] = p[0];
      p += 1;
      rem -= 1;
    }
  }
  return o;
}

size_t fsst_decompress_host(const FsstTable& t, const uint8_t* in, size_t len, uint8_t* out, size_t cap) {
  size_t o = 0;
  for (size_t i = 0; i < len; ++i) {
    uint8_t code = in[i];
    if (code == 255) {
      if (i + 1 >= len) break;
      if (o < cap) out[o] = in[i + 1];
      ++o;
      ++i;
    } else {
      uint32_t l = t.lens[code];
      uint64_t s = t.symbols[code];
      for (uint32_t k = 0; k < l; ++k) {
        if (o < cap) out[o] = static_cast<uint8_t>(s >> (8 * k));
        ++o;
      }
    }
  }
  return o;
}

}  // namespace lc
