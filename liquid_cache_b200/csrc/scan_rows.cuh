// scan_rows.cuh — the selection machinery shared by the integer and byte-view scan kernels.
//
// One 256-thread CTA walks the rows of ONE entry in logical order, 32 rows per warp step
// (lane == row % 32), and produces one of:
//   MODE_DECODE  emit(row, dst): values of the selected rows, compacted   (LiquidArray::filter)
//   MODE_PRED    compact mask + validity over the selected rows           (try_eval_predicate contract,
//                liquid_array/mod.rs:123-130: result length = popcount(selection))
//   MODE_REFINE  selection := selection & valid & cmp, full length        (predicate + nulls->false +
//                boolean_buffer_and_then fused: cache/column.rs:134-137, datafusion/src/utils.rs:62-83)
// Selection -> write offset: popcount per 32-bit selection word, block-wide exclusive scan
// (warp shuffles), rank inside the word via __popc(sel & lanemask_lt); the mask bits of the selected
// lanes are squeezed together with a warp ballot (warp_pext2) and appended to a bit stream kept in
// shared memory, flushed one 8192-row tile at a time.
#pragma once
#include "device_utils.cuh"
#include "kernels.h"

namespace lc {

struct ScanSmem {
  uint64_t bar[2];
  uint32_t counts[2];
  uint32_t misc[2];
  uint32_t warp_tot[8];
  uint32_t sel[256];
  uint32_t off[256];
  uint32_t maskbuf[264];
  uint32_t validbuf[264];
  uint8_t scratch[8][32];
  uint64_t io_slot[2][4];  // persistent kernels: {sel_off, out_off, valid_off} of this / the next entry
  uint64_t ref_slot[2];    // persistent kernels: {blob, blob_bytes} of the entry after the next one
  uint32_t fcnt[2];        // persistent kernels: survivor count of this / the previous entry (flushed one round late)
};
static_assert(sizeof(ScanSmem) <= kScanFixedSmem, "fixed smem area too small");

__device__ __forceinline__ void scan_smem_init(ScanSmem* sm) {
  if (threadIdx.x == 0) {
    sm->counts[0] = 0;
    sm->counts[1] = 0;
    sm->misc[0] = 0;
    sm->misc[1] = 0;
    sm->maskbuf[0] = 0;
    sm->validbuf[0] = 0;
  }
}

struct IdentityOrder {
  __device__ __forceinline__ uint32_t operator()(uint32_t j) const { return j; }
};

// cmp(row, c, j) -> bool (only called for row < n); emit(row, dst, c, j) writes the decoded value (MODE_DECODE).
// A warp covers one 1024-row chunk c per pass in 32 steps j; step j handles the 32 rows of logical word
// order(j) of that chunk. Byte-view columns walk the words in order; bit-packed integers walk them in FastLanes
// STORAGE order, where step j needs exactly one packed word per lane (see k_int.cu) — the selection machinery
// does not care, every word's write offset comes from the prefix-sum table.
// `valid` may be nullptr (no nulls). Requires scan_smem_init + __syncthreads() before the call.
template <int MODE, typename Cmp, typename Emit, typename Order = IdentityOrder>
__device__ __forceinline__ void scan_entry_rows(const uint32_t* __restrict__ sel, uint32_t n,
                                                const uint32_t* __restrict__ valid, uint32_t entry_null_count,
                                                uint32_t* __restrict__ out_bits, uint32_t* __restrict__ out_valid,
                                                uint32_t* __restrict__ out_counts, ScanSmem* sm, Cmp cmp,
                                                Emit emit, Order order = Order()) {
  const uint32_t n_words = (n + 31u) >> 5;
  const uint32_t tail = n & 31u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool want_valid = (valid != nullptr) && (out_valid != nullptr);

  if (sel == nullptr) {
    // ---------------- dense: every row selected, rank == row ----------------
    uint32_t survivors = 0;
    for (uint32_t w0 = warp * 32u; w0 < n_words; w0 += 256u) {
      const uint32_t c = w0 >> 5;
#pragma unroll 4
      for (uint32_t j = 0; j < 32; ++j) {
        const uint32_t wi = w0 + order(j);
        if (wi >= n_words) continue;
        const uint32_t row = wi * 32u + lane;
        const bool in = row < n;
        uint32_t vw = valid ? valid[wi] : kFullMask;
        if (wi == n_words - 1u && tail) vw &= (1u << tail) - 1u;
        if (MODE == MODE_DECODE) {
          if (in) emit(row, row, c, j);
          if (want_valid && lane == 0) out_valid[wi] = vw;
        } else {
          const uint32_t cw = __ballot_sync(kFullMask, in && cmp(row, c, j)) & vw;
          if (lane == 0) {
            out_bits[wi] = cw;
            if (MODE == MODE_PRED && want_valid) out_valid[wi] = vw;
          }
          survivors += __popc(cw);
        }
      }
    }
    if (out_counts) {
      if (MODE != MODE_DECODE) {
        if (lane == 0 && survivors) atomicAdd(&sm->counts[0], survivors);
        __syncthreads();
        if (threadIdx.x == 0) {
          if (MODE == MODE_REFINE) {
            out_counts[0] = sm->counts[0];
            out_counts[1] = 0;
          } else {
            out_counts[0] = n;
            out_counts[1] = entry_null_count;
            out_counts[2] = sm->counts[0];  // set bits of the mask (nulls count as false)
          }
        }
      } else if (threadIdx.x == 0) {
        out_counts[0] = n;
        out_counts[1] = entry_null_count;
      }
    }
    return;
  }

  // ---------------- general: selection -> write offsets ----------------
  uint32_t goff = 0;      // selected rows before this tile
  uint32_t warp_acc = 0;  // lane 0: REFINE survivors, else nulls among the selected rows
  uint32_t true_acc = 0;  // lane 0, PRED: set bits of the compact mask
  for (uint32_t tile_w0 = 0; tile_w0 < n_words; tile_w0 += 256u) {
    const uint32_t my_wi = tile_w0 + threadIdx.x;
    uint32_t sw = 0;
    if (my_wi < n_words) {
      sw = sel[my_wi];
      if (my_wi == n_words - 1u && tail) sw &= (1u << tail) - 1u;
    }
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan_256(__popc(sw), sm->warp_tot, &tile_total);
    sm->sel[threadIdx.x] = sw;
    sm->off[threadIdx.x] = excl;
    if (MODE != MODE_REFINE) {
      for (uint32_t i = threadIdx.x + 1u; i < 264u; i += 256u) {
        sm->maskbuf[i] = 0;
        sm->validbuf[i] = 0;
      }
    }
    __syncthreads();
    const uint32_t pbase = goff & 31u;

    const uint32_t c = (tile_w0 >> 5) + warp;
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t lw = warp * 32u + order(j);
      const uint32_t wi = tile_w0 + lw;
      if (wi >= n_words) continue;
      const uint32_t selw = sm->sel[lw];
      if (selw == 0) {
        if (MODE == MODE_REFINE && lane == 0) out_bits[wi] = 0;
        continue;
      }
      const uint32_t row = wi * 32u + lane;
      const uint32_t vw = valid ? valid[wi] : kFullMask;
      const uint32_t k = __popc(selw);
      const uint32_t off = sm->off[lw];
      const bool mine = (selw >> lane) & 1u;
      if (MODE == MODE_REFINE) {
        const uint32_t cw = __ballot_sync(kFullMask, mine && cmp(row, c, j)) & vw;
        if (lane == 0) {
          out_bits[wi] = cw;
          warp_acc += __popc(cw);
        }
      } else {
        const uint32_t vbit = (vw >> lane) & 1u;
        if (MODE == MODE_DECODE) {
          if (mine) emit(row, goff + off + __popc(selw & lanemask_lt()), c, j);
          if (want_valid) {
            uint32_t b0, b1;
            warp_pext2(vbit, selw, lane, sm->scratch[warp], &b0, &b1);
            if (lane == 0) bits_append(sm->validbuf, pbase + off, k, b0);
          }
        } else {
          const uint32_t cbit = (mine && cmp(row, c, j)) ? vbit : 0u;
          uint32_t b0, b1;
          warp_pext2(cbit | (vbit << 1), selw, lane, sm->scratch[warp], &b0, &b1);
          if (lane == 0) {
            bits_append(sm->maskbuf, pbase + off, k, b0);
            if (want_valid) bits_append(sm->validbuf, pbase + off, k, b1);
            true_acc += __popc(b0);
          }
        }
        if (lane == 0) warp_acc += __popc(selw & ~vw);
      }
    }
    __syncthreads();

    if (MODE != MODE_REFINE) {
      const uint32_t total_bits = pbase + tile_total;
      const uint32_t nfull = total_bits >> 5;
      const uint32_t gword0 = goff >> 5;
      for (uint32_t i = threadIdx.x; i < nfull; i += 256u) {
        if (MODE == MODE_PRED) out_bits[gword0 + i] = sm->maskbuf[i];
        if (want_valid) out_valid[gword0 + i] = sm->validbuf[i];
      }
      const uint32_t carry_m = sm->maskbuf[nfull], carry_v = sm->validbuf[nfull];
      __syncthreads();
      if (threadIdx.x == 0) {
        sm->maskbuf[0] = carry_m;
        sm->validbuf[0] = carry_v;
      }
    }
    goff += tile_total;
  }
  if (MODE != MODE_REFINE) {
    __syncthreads();
    if (threadIdx.x == 0 && (goff & 31u)) {
      if (MODE == MODE_PRED) out_bits[goff >> 5] = sm->maskbuf[0];
      if (want_valid) out_valid[goff >> 5] = sm->validbuf[0];
    }
  }
  if (out_counts) {
    if (lane == 0 && warp_acc) atomicAdd(&sm->counts[1], warp_acc);
    if (MODE == MODE_PRED && lane == 0 && true_acc) atomicAdd(&sm->counts[0], true_acc);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (MODE == MODE_REFINE) {
        out_counts[0] = sm->counts[1];
        out_counts[1] = 0;
      } else {
        out_counts[0] = goff;
        out_counts[1] = sm->counts[1];
        if (MODE == MODE_PRED) out_counts[2] = sm->counts[0];
      }
    }
  }
}

}  // namespace lc
