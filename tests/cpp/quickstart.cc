// The reference's quick-start examples through the C++ mirror of its API (liquid_cache_b200/csrc/liquid_cache.hpp):
//   /root/reference/README.md:43-88            UInt64 [10..15]: get, get with selection, `col > 12`
//   /root/reference/src/core/README.md:17-104  strings ["apple","banana",NULL,"apple","cherry"]: `= "apple"` under a selection
// plus the Option::None / Err(array) conventions of the builders. Exit codes: 0 all answers as published, 3 no CUDA device
// (the library refuses to run without one — there is no CPU path), 1 a wrong answer.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "liquid_cache_b200/csrc/liquid_cache.hpp"

namespace lc = liquid_cache;

namespace {

// ---- minimal Arrow C Data Interface producers for the two input shapes of the examples ----
struct Owned {
  std::vector<const void*> bufs;
  std::vector<uint8_t> validity, data;
  std::vector<int32_t> offsets;
  std::string format;
};
void release_schema(ArrowSchema* s) {
  delete static_cast<Owned*>(s->private_data);
  s->release = nullptr;
}
void release_array(ArrowArray* a) {
  delete static_cast<Owned*>(a->private_data);
  a->release = nullptr;
}
void make_schema(const char* format, ArrowSchema* out) {
  Owned* o = new Owned();
  o->format = format;
  std::memset(out, 0, sizeof(*out));
  out->format = o->format.c_str();
  out->name = "";
  out->flags = ARROW_FLAG_NULLABLE;
  out->release = release_schema;
  out->private_data = o;
}
void make_u64(const std::vector<uint64_t>& v, ArrowSchema* s, ArrowArray* a) {
  make_schema("L", s);
  Owned* o = new Owned();
  o->data.resize(v.size() * 8);
  std::memcpy(o->data.data(), v.data(), o->data.size());
  o->bufs = {nullptr, o->data.data()};
  std::memset(a, 0, sizeof(*a));
  a->length = static_cast<int64_t>(v.size());
  a->n_buffers = 2;
  a->buffers = o->bufs.data();
  a->release = release_array;
  a->private_data = o;
}
void make_utf8(const std::vector<const char*>& v, ArrowSchema* s, ArrowArray* a) {  // nullptr = NULL
  make_schema("u", s);
  Owned* o = new Owned();
  o->validity.assign((v.size() + 7) / 8, 0);
  o->offsets.push_back(0);
  int64_t nulls = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    if (v[i]) {
      o->validity[i / 8] |= static_cast<uint8_t>(1u << (i % 8));
      o->data.insert(o->data.end(), v[i], v[i] + std::strlen(v[i]));
    } else {
      ++nulls;
    }
    o->offsets.push_back(static_cast<int32_t>(o->data.size()));
  }
  if (o->data.empty()) o->data.push_back(0);
  o->bufs = {nulls ? o->validity.data() : nullptr, o->offsets.data(), o->data.data()};
  std::memset(a, 0, sizeof(*a));
  a->length = static_cast<int64_t>(v.size());
  a->null_count = nulls;
  a->n_buffers = 3;
  a->buffers = o->bufs.data();
  a->release = release_array;
  a->private_data = o;
}

int failures = 0;
void expect(bool ok, const char* what) {
  if (!ok) {
    std::fprintf(stderr, "WRONG: %s\n", what);
    ++failures;
  }
}
std::vector<uint64_t> u64_values(const ArrowArray& a) {
  const uint64_t* p = static_cast<const uint64_t*>(a.buffers[1]) + a.offset;
  return std::vector<uint64_t>(p, p + a.length);
}
std::string mask_string(const lc::BooleanArray& m) {  // 'T' / 'F' / 'N' per row
  std::string s;
  for (uint64_t i = 0; i < m.len; ++i) {
    const bool valid = m.null_count == 0 || ((m.validity[i / 8] >> (i % 8)) & 1);
    s += !valid ? 'N' : ((m.values[i / 8] >> (i % 8)) & 1) ? 'T' : 'F';
  }
  return s;
}
uint8_t bits_of(const char* tf) {  // "TFTFTF" -> LSB-first byte
  uint8_t b = 0;
  for (size_t i = 0; tf[i]; ++i)
    if (tf[i] == 'T') b |= static_cast<uint8_t>(1u << i);
  return b;
}

}  // namespace

int main() {
  std::unique_ptr<lc::LiquidCache> cache;
  try {
    cache.reset(lc::LiquidCacheBuilder().with_batch_size(8192).build());
  } catch (const lc::GpuError& e) {
    std::fprintf(stderr, "no device: %s\n", e.what());
    return 3;
  }
  try {
    // ---- README.md:43-88 ----
    ArrowSchema s;
    ArrowArray a;
    make_u64({10, 11, 12, 13, 14, 15}, &s, &a);
    const lc::EntryID id = 1;
    cache->insert(id, &s, &a).run();
    a.release(&a);
    s.release(&s);
    expect(cache->is_cached(id), "entry 1 is cached after insert");
    ArrowSchema os;
    ArrowArray oa;
    expect(cache->get(id).read(&os, &oa), "get(1) finds the entry");
    expect(std::string(os.format) == "L" && u64_values(oa) == std::vector<uint64_t>({10, 11, 12, 13, 14, 15}), "get(1) = [10..15]");
    oa.release(&oa);
    os.release(&os);
    const uint8_t sel = bits_of("TFTFTF");
    expect(cache->get(id).with_selection({&sel, 6}).read(&os, &oa), "get(1) with a selection");
    expect(u64_values(oa) == std::vector<uint64_t>({10, 12, 14}), "selection T F T F T F -> [10, 12, 14]");
    oa.release(&oa);
    os.release(&os);
    lc::BooleanArray mask;
    expect(cache->eval_predicate(id, lc::LiquidExpr::compare_u64(LC_OP_GT, 12)).read(&mask), "eval_predicate(1) finds the entry");
    expect(mask_string(mask) == "FFFTTT", "col > 12 -> [F, F, F, T, T, T]");

    // ---- src/core/README.md:17-104 ----
    make_utf8({"apple", "banana", nullptr, "apple", "cherry"}, &s, &a);
    const lc::EntryID sid = lc::parquet_array_id(1, 0, 3, 0);
    cache->insert(sid, &s, &a).with_squeeze_hint(LC_HINT_SUBSTRING_SEARCH).run();
    a.release(&a);
    s.release(&s);
    const uint8_t ssel = bits_of("TTFTT");
    lc::LiquidExpr eq = lc::LiquidExpr::compare_bytes(LC_OP_EQ, "apple");
    lc::LiquidExpr moved = std::move(eq);  // the literal travels with the expression
    expect(cache->eval_predicate(sid, moved).with_selection({&ssel, 5}).read(&mask), "eval_predicate on the string entry");
    expect(mask_string(mask) == "TFTF", "= 'apple' under [T, T, F, T, T] -> [T, F, T, F]");
    expect(cache->eval_predicate(sid, lc::LiquidExpr::compare_bytes(LC_OP_NE, "apple")).read(&mask) && mask_string(mask) == "FTNFT",
           "!= 'apple' keeps the null");
    expect(cache->eval_predicate(sid, lc::LiquidExpr::like("%an%")).read(&mask) && mask_string(mask) == "FTNFF", "LIKE '%an%'");
    expect(cache->get(sid).with_selection({&ssel, 5}).read(&os, &oa) && oa.length == 4 && oa.null_count == 0 && std::string(os.format) == "u",
           "string get with a selection drops the null row");
    oa.release(&oa);
    os.release(&os);

    // ---- conventions ----
    expect(!cache->get(999).read(&os, &oa) && !cache->eval_predicate(999, moved).read(&mask), "absent entry -> None");
    make_schema("b", &s);  // Boolean: transcode gives the array back (cache/transcode.rs:282-289)
    Owned* o = new Owned();
    o->data = {0x05};
    o->bufs = {nullptr, o->data.data()};
    std::memset(&a, 0, sizeof(a));
    a.length = 3;
    a.n_buffers = 2;
    a.buffers = o->bufs.data();
    a.release = release_array;
    a.private_data = o;
    bool declined = false;
    try {
      cache->insert(7, &s, &a).run();
    } catch (const lc::UnsupportedType&) {
      declined = true;
    }
    a.release(&a);
    s.release(&s);
    expect(declined && !cache->is_cached(7), "Boolean arrays are declined, the caller keeps the Arrow array");
    bool refused = false;
    try {
      cache->eval_predicate(id, lc::LiquidExpr::like("%1%")).read(&mask);
    } catch (const lc::UnsupportedExpr&) {
      refused = true;
    }
    expect(refused, "LIKE on an integer column is not a supported expression");
    cache->reset();
    expect(!cache->is_cached(id) && !cache->is_cached(sid), "reset empties the cache");
    const lc_stats st = cache->stats();
    std::printf("cpp quickstart: %d wrong answers, %llu kernel launches\n", failures, (unsigned long long)st.kernel_launches);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return failures ? 1 : 0;
}
