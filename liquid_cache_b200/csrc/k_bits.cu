// k_bits.cu — boolean_buffer_and_then on the device.
// Reference: /root/reference/src/datafusion/src/utils.rs:17-83 (semantics), :104-236 (the BMI2 PDEP
// routine, the only hand-written intrinsic in the reference). out bit p = left[p] AND the
// rank_left(p)-th bit of right, where right has popcount(left) bits. On the GPU the "deposit" is a
// rank computed from a prefix sum of per-word popcounts plus __popc(word & lanemask_lt).
#include "device_utils.cuh"
#include "kernels.h"

namespace lc {

__global__ void __launch_bounds__(256) k_and_then(const uint32_t* __restrict__ left, uint32_t left_bits,
                                                  const uint32_t* __restrict__ right, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_tot[8];
  __shared__ uint32_t s_word[256];
  __shared__ uint32_t s_off[256];
  const uint32_t n_words = (left_bits + 31u) >> 5;
  const uint32_t tail = left_bits & 31u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (uint32_t w0 = 0; w0 < n_words; w0 += 256u) {
    const uint32_t wi = w0 + threadIdx.x;
    uint32_t lw = 0;
    if (wi < n_words) {
      lw = left[wi];
      if (wi == n_words - 1u && tail) lw &= (1u << tail) - 1u;
    }
    uint32_t tot;
    const uint32_t excl = block_excl_scan_256(__popc(lw), s_tot, &tot);
    s_word[threadIdx.x] = lw;
    s_off[threadIdx.x] = carry + excl;
    __syncthreads();
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t lwi = warp * 32u + j;
      if (w0 + lwi >= n_words) break;
      const uint32_t word = s_word[lwi];
      bool bit = false;
      if ((word >> lane) & 1u) {
        const uint32_t r = s_off[lwi] + __popc(word & lanemask_lt());
        bit = (right[r >> 5] >> (r & 31u)) & 1u;
      }
      const uint32_t o = __ballot_sync(kFullMask, bit);
      if (lane == 0) out[w0 + lwi] = o;
    }
    carry += tot;
    __syncthreads();
  }
}

cudaError_t launch_and_then(const uint32_t* d_left, uint32_t left_bits, const uint32_t* d_right, uint32_t* d_out,
                            cudaStream_t s) {
  if (left_bits == 0) return cudaSuccess;
  k_and_then<<<1, 256, 0, s>>>(d_left, left_bits, d_right, d_out);
  return cudaGetLastError();
}

}  // namespace lc
