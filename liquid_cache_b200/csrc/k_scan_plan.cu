// k_scan_plan.cu — the bookkeeping of a get-with-selection done ON THE DEVICE, so that reading the survivors of a scan
// costs the host one synchronisation instead of three.
//
// read_from_cache (src/datafusion/src/reader/runtime/liquid_cache_reader.rs:342-391) reads every projected column with the
// final selection; the reference learns each batch's row count from the BooleanBuffer it holds. Here the selection lives in
// HBM, and what the decode kernels need before they can run — where each entry's rows, validity words, dictionary-length
// scratch and bytes start in the concatenated result — are prefix sums over per-entry counts that are themselves on the
// device. Round 1 fetched the counts, summed on the host and uploaded the offsets (twice for byte views: rows, then bytes).
// These two single-CTA kernels do the sums in place; the host only reads a 64-byte header together with the result.
#include "device_utils.cuh"
#include "kernels.h"

namespace lc {

// 1024 threads; thread t owns a contiguous run of entries, so the scan is: serial over the run, block-wide over the runs.
template <int NV>
__device__ __forceinline__ void block_scan_runs(uint64_t (&v)[NV], uint64_t (&excl)[NV], uint64_t (&total)[NV], uint64_t* smem /*[NV][32]*/) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    uint64_t x = v[q];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t y = __shfl_up_sync(kFullMask, x, d);
      if (lane >= static_cast<uint32_t>(d)) x += y;
    }
    if (lane == 31u) smem[q * 32 + warp] = x;
    excl[q] = x - v[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const uint64_t w = smem[q * 32 + lane];
      uint64_t x = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(kFullMask, x, d);
        if (lane >= static_cast<uint32_t>(d)) x += y;
      }
      smem[q * 32 + lane] = x - w;  // exclusive base of each warp
      if (lane == 31u) smem[NV * 32 + q] = x;  // grand total
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    excl[q] += smem[q * 32 + warp];
    total[q] = smem[NV * 32 + q];
  }
}

// counts2[2i] = rows of entry i that survived (MODE_REFINE's count). Writes, for every entry: row_base (rows before it),
// vword_off (validity words before it, 4-word aligned per entry as the decode kernels lay them out), ulen_off (dictionary
// length scratch before it; only entries with survivors take space). Refuses (hdr->overflow) when a capacity is too small.
__global__ void __launch_bounds__(1024) k_scan_plan_rows(const uint32_t* __restrict__ counts2, const uint32_t* __restrict__ n_unique,
                                                         uint32_t n, uint64_t cap_rows, uint64_t cap_ulen, uint64_t* row_base,
                                                         uint64_t* vword_off, uint64_t* ulen_off, ScanPlanHdr* hdr) {
  __shared__ uint64_t smem[3 * 32 + 3];
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t b = threadIdx.x * per, e = b + per < n ? b + per : n;
  uint64_t v[3] = {0, 0, 0};
  uint32_t hits = 0;
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t k = counts2[2u * i];
    v[0] += k;
    v[1] += ((k + 31u) / 32u + 3u) & ~3u;
    if (k) {
      ++hits;
      if (n_unique) v[2] += (n_unique[i] + 3u) & ~3u;
    }
  }
  uint64_t excl[3], total[3];
  block_scan_runs<3>(v, excl, total, smem);
  uint64_t r = excl[0], w = excl[1], u = excl[2];
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t k = counts2[2u * i];
    row_base[i] = r;
    vword_off[i] = w;
    ulen_off[i] = u;
    r += k;
    w += ((k + 31u) / 32u + 3u) & ~3u;
    if (k && n_unique) u += (n_unique[i] + 3u) & ~3u;
  }
  // entries with survivors (for the record; the decode kernels skip the others by their count)
  uint32_t h = hits;
  for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(kFullMask, h, d);
  __shared__ uint32_t s_hits;
  if (threadIdx.x == 0) s_hits = 0;
  __syncthreads();
  if ((threadIdx.x & 31u) == 0 && h) atomicAdd(&s_hits, h);
  __syncthreads();
  if (threadIdx.x == 0) {
    hdr->n_hit = s_hits;
    hdr->rows = total[0];
    hdr->vwords = total[1];
    hdr->ulen_words = total[2];
    hdr->bytes = 0;
    hdr->nulls = 0;
    hdr->overflow = (total[0] > cap_rows || total[2] > cap_ulen) ? 1u : 0u;
  }
}

// counts4[4i + 2] = decoded bytes of entry i's surviving rows (k_str_lengths; 0 where no row survived),
// counts4[4i + 1] = nulls among them. Writes byte_base and the closing offset of the concatenated array.
__global__ void __launch_bounds__(1024) k_scan_plan_bytes(const uint32_t* __restrict__ counts4, uint32_t n, uint64_t cap_bytes,
                                                          uint64_t* byte_base, int32_t* out_offsets, ScanPlanHdr* hdr) {
  __shared__ uint64_t smem[2 * 32 + 2];
  if (hdr->overflow) return;
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t b = threadIdx.x * per, e = b + per < n ? b + per : n;
  uint64_t v[2] = {0, 0};
  for (uint32_t i = b; i < e; ++i) {
    v[0] += counts4[4u * i + 2u];
    v[1] += counts4[4u * i + 1u];
  }
  uint64_t excl[2], total[2];
  block_scan_runs<2>(v, excl, total, smem);
  uint64_t r = excl[0];
  for (uint32_t i = b; i < e; ++i) {
    byte_base[i] = r;
    r += counts4[4u * i + 2u];
  }
  if (threadIdx.x == 0) {
    hdr->bytes = total[0];
    hdr->nulls = total[1];
    if (total[0] > cap_bytes || total[0] > 0x7fffffffull) hdr->overflow = 2u;
    else out_offsets[hdr->rows] = static_cast<int32_t>(total[0]);
  }
}

cudaError_t launch_scan_plan_rows(const uint32_t* d_counts2, const uint32_t* d_n_unique, uint32_t n, uint64_t cap_rows, uint64_t cap_ulen,
                                  uint64_t* d_row_base, uint64_t* d_vword_off, uint64_t* d_ulen_off, ScanPlanHdr* d_hdr, cudaStream_t s) {
  k_scan_plan_rows<<<1, 1024, 0, s>>>(d_counts2, d_n_unique, n, cap_rows, cap_ulen, d_row_base, d_vword_off, d_ulen_off, d_hdr);
  return cudaGetLastError();
}

cudaError_t launch_scan_plan_bytes(const uint32_t* d_counts4, uint32_t n, uint64_t cap_bytes, uint64_t* d_byte_base,
                                   int32_t* d_out_offsets, ScanPlanHdr* d_hdr, cudaStream_t s) {
  k_scan_plan_bytes<<<1, 1024, 0, s>>>(d_counts4, n, cap_bytes, d_byte_base, d_out_offsets, d_hdr);
  return cudaGetLastError();
}

}  // namespace lc
