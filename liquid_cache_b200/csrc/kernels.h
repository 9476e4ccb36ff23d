// kernels.h — work-item structs and launchers of the sm_100a kernels (host-visible side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "entry_layout.h"

namespace lc {

// Shared-memory budget of the scan kernels: a fixed control area + the staged entry blob.
constexpr uint32_t kScanFixedSmem = 4608;
constexpr uint32_t kStageCap = 96 * 1024;  // entries larger than this are read straight from global

enum ScanMode : int32_t {
  MODE_DECODE = 0,  // to_arrow_array / filter: values (+validity) of the selected rows
  MODE_PRED = 1,    // try_eval_predicate: compact mask (+validity) over the selected rows
  MODE_REFINE = 2,  // device pipeline: selection := selection & valid & cmp, full length, in place
};

// One entry of a batched integer scan. 64 bytes.
struct alignas(16) IntScanWork {
  const uint8_t* blob;      // IntHeader + sections, in HBM
  const uint32_t* sel;      // selection words (bit i of word i/32 = row i) or nullptr = all rows
  void* out_values;         // DECODE: native T[k]; PRED: mask words; REFINE: selection words (may alias sel)
  uint32_t* out_validity;   // DECODE/PRED: validity words of the output, or nullptr
  uint32_t* out_counts;     // [0] = k (selected rows / surviving rows for REFINE), [1] = nulls among selected
  uint64_t thr;             // PRED/REFINE: threshold in the unsigned (v - reference) domain
  int32_t ucmp;             // PRED/REFINE: UCmp
  uint32_t blob_bytes;
};
static_assert(sizeof(IntScanWork) == 64, "IntScanWork must be 64 bytes");

struct alignas(16) IntMinMaxWork {  // 32 bytes
  const void* values;         // native T[n] in device scratch
  const uint32_t* validity;   // bit-offset-0 validity words or nullptr
  uint64_t* out;              // [0]=min, [1]=max (sign- or zero-extended), [2]=valid count
  uint32_t n;
  uint32_t phys;              // PhysType
};

struct alignas(16) IntPackWork {  // 96 bytes
  const void* values;
  const uint32_t* validity;
  uint8_t* blob;
  uint64_t pad;
  IntHeader hdr;
};
static_assert(sizeof(IntPackWork) == 96, "IntPackWork must be 96 bytes");

cudaError_t launch_int_scan(int mode, const IntScanWork* d_works, uint32_t n_works, uint32_t max_blob_bytes,
                            cudaStream_t s);
cudaError_t launch_int_minmax(const IntMinMaxWork* d_works, uint32_t n_works, cudaStream_t s);
cudaError_t launch_int_pack(const IntPackWork* d_works, uint32_t n_works, cudaStream_t s);

// ---- byte-view (string) path -------------------------------------------------------------------
enum StrPredKind : int32_t {
  SP_CONST = 0,     // every unique gets `const_result`
  SP_EQ_SHORT = 1,  // needle suffix <= 7 bytes: decided on PrefixKey alone (comparisons.rs:33-49)
  SP_EQ_LONG = 2,   // length + prefix7 gates, then full compare of the decoded value (comparisons.rs:51-79)
  SP_ORD = 3,       // prefix7 compare, ties decoded and compared in full (comparisons.rs:114-151,351-405)
  SP_ORD_EMPTY = 4, // needle suffix empty: decided on PrefixKey.len (comparisons.rs:371-381)
  SP_LIKE = 5,      // fingerprint gate + substring match on the encoded bytes (comparisons.rs:159-183,600-651)
};

constexpr uint32_t kMaxNeedle = 1024;  // needle bytes staged into shared memory

// Predicate descriptor shared by every entry of a launch whose shared prefix agrees with `sp_case`;
// the per-entry part lives in StrScanWork.
struct alignas(16) StrPredDesc {
  int32_t op;            // lc_op (EQ..GE, LIKE, NOT_LIKE)
  uint32_t needle_len;   // full needle (for LIKE: the inner pattern without the % signs)
  uint32_t needle_fp;    // fingerprint of the LIKE needle (fingerprint.rs:19-26)
  uint32_t pad;
  const uint8_t* needle; // device: needle bytes padded to 4, then needle_len x u16 KMP failure links
};

struct alignas(16) StrScanWork {  // 80 bytes
  const uint8_t* blob;
  const uint32_t* sel;
  void* out_values;        // PRED: mask words; REFINE: selection words
  uint32_t* out_validity;
  uint32_t* out_counts;
  uint64_t key_expect;     // EQ_SHORT/EQ_LONG: the PrefixKey a match must equal (prefix7 | len<<56);
                           // ORD: first min(7,len) suffix bytes of the needle, big-endian in the top bytes
  int32_t kind;            // StrPredKind after the host looked at the entry's shared prefix
  uint32_t flags;          // bit0 const_result, bit1 negate (NE / NOT LIKE), bit2 LIKE without fingerprints
  uint32_t cmp_len;        // ORD: number of prefix bytes compared (1..7)
  uint32_t blob_bytes;
  uint32_t head_bytes;
  uint32_t meta_bytes;
  uint32_t pad[2];
};
static_assert(sizeof(StrScanWork) == 80, "StrScanWork must be 80 bytes");

cudaError_t launch_str_scan(int mode, const StrScanWork* d_works, uint32_t n_works, const StrPredDesc& pred,
                            uint32_t max_head_bytes, uint32_t max_unique, cudaStream_t s);

// get()/filter() for byte-view entries: pass 1 (lengths + local offsets), host prefix sums, pass 2 (decode).
constexpr uint32_t kGatherPrecompLens = 1u;  // compute decoded lengths of ALL long uniques up front

struct alignas(16) StrGatherWork {  // 80 bytes
  const uint8_t* blob;
  const uint32_t* sel;       // nullptr = all rows
  uint32_t* row_off;         // scratch, k+1: exclusive prefix of decoded lengths of the selected rows
  uint32_t* row_key;         // scratch, k: dictionary key of each selected row (0xFFFFFFFF = null row)
  uint32_t* ulen;            // scratch, U: decoded length per unique (kGatherPrecompLens)
  uint32_t* out_validity;    // validity words of this entry's slice (word aligned scratch), or nullptr
  uint32_t* out_counts;      // [0]=k, [1]=nulls among selected, [2]=sum of decoded bytes
  uint32_t blob_bytes;
  uint32_t head_bytes;
  uint32_t flags;
  uint32_t pad[3];
};
static_assert(sizeof(StrGatherWork) == 80, "StrGatherWork must be 80 bytes");

struct alignas(16) StrDecodeWork {  // 48 bytes
  const uint8_t* blob;
  const uint32_t* row_off;   // from pass 1
  const uint32_t* row_key;
  int32_t* out_offsets;      // this entry's first slot in the concatenated offsets buffer
  uint8_t* out_bytes;        // concatenated value buffer (base)
  uint32_t byte_base;        // where this entry's bytes start in out_bytes
  uint32_t k;                // selected rows
};
static_assert(sizeof(StrDecodeWork) == 48, "StrDecodeWork must be 48 bytes");

cudaError_t launch_str_lengths(const StrGatherWork* d_works, uint32_t n_works, uint32_t max_head_bytes,
                               cudaStream_t s);
cudaError_t launch_str_decode(const StrDecodeWork* d_works, uint32_t n_works, cudaStream_t s);

// ---- FSST compression at insert ----------------------------------------------------------------
struct alignas(16) FsstEncTable {
  // Greedy longest-match lookup used by the compress kernel. Built on the host at training time.
  //   long symbols (3..8 bytes): open-addressed hash on the first 3 bytes, one symbol per slot
  //   short symbols: 65536-entry table indexed by the next two bytes -> code | len<<8 (len 1 or 2),
  //                  or 0xFFFF... escape marker
  uint64_t hash_sym[2048];
  uint16_t hash_meta[2048];   // code | len<<8 ; 0 = empty slot
  uint16_t short_code[65536]; // code | len<<8 ; len==0 => escape the byte
  uint16_t one_byte[256];     // code | 1<<8 of the 1-byte symbol for this byte, 0 = none (last-byte fallback)
};

struct alignas(16) FsstCompressWork {  // 48 bytes
  const uint8_t* values;      // concatenated unique values (device scratch)
  const uint32_t* offsets;    // U+1 offsets into values
  uint8_t* out;               // worst-case 2x scratch: unique i compressed at out + 2*offsets[i]
  uint32_t* out_lens;         // compressed length of each unique
  const FsstEncTable* table;
  uint32_t n_unique;
  uint32_t pad;
};
cudaError_t launch_fsst_compress(const FsstCompressWork* d_works, uint32_t n_works, cudaStream_t s);

// ---- bit utilities -----------------------------------------------------------------------------
// boolean_buffer_and_then: out[p] = left[p] & right[rank_left(p)]  (datafusion/src/utils.rs:62-236)
cudaError_t launch_and_then(const uint32_t* d_left, uint32_t left_bits, const uint32_t* d_right, uint32_t* d_out,
                            cudaStream_t s);

}  // namespace lc
