// host_pool.h — a small fork-join helper for the host-side preparation loops of the batched calls
// (staging selection bitmaps, popcounts). The GPU path is fast enough that a single core walking a few thousand
// 1 KB bitmaps fresh out of a DMA shows up in the end-to-end time; these loops are embarrassingly parallel.
#pragma once
#include <cstdint>
#include <functional>

namespace lc {

// Runs fn(begin, end) over [0, n) split into contiguous ranges of at least `min_grain` items on the calling thread
// plus the pool's workers; returns when every range is done. Thread count: LC_HOST_THREADS (default
// min(8, hardware threads / 2)); 1 disables the pool. One parallel_for at a time per process (internally serialised).
void parallel_for(uint64_t n, uint64_t min_grain, const std::function<void(uint64_t, uint64_t)>& fn);
// Threads parallel_for runs on (the caller included).
unsigned host_pool_threads();

}  // namespace lc
