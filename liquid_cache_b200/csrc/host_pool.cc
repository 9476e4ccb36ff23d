// host_pool.cc — see host_pool.h
#include "host_pool.h"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

namespace lc {
namespace {

class Pool {
 public:
  Pool() {
    unsigned want = 0;
    if (const char* e = std::getenv("LC_HOST_THREADS")) want = static_cast<unsigned>(std::atoi(e));
    if (want == 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      want = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 1);
    }
    n_threads_ = want;
    for (unsigned i = 1; i < want; ++i) workers_.emplace_back([this] { worker(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  unsigned threads() const { return n_threads_; }

  void run(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)>& fn) {
    std::lock_guard<std::mutex> serial(run_mu_);
    fn_ = &fn;
    n_ = n;
    grain_ = grain;
    next_.store(0, std::memory_order_relaxed);
    pending_.store(static_cast<int>(workers_.size()), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> l(mu_);
      ++generation_;
    }
    cv_.notify_all();
    drain();
    // wait for the workers to leave this generation (they only spin through an empty queue by now)
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [this] { return pending_.load(std::memory_order_acquire) == 0; });
    fn_ = nullptr;
  }

 private:
  void drain() {
    for (;;) {
      const uint64_t b = next_.fetch_add(grain_, std::memory_order_relaxed);
      if (b >= n_) return;
      const uint64_t e = b + grain_ < n_ ? b + grain_ : n_;
      (*fn_)(b, e);
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      drain();
      if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> l(mu_);
        done_cv_.notify_one();
      }
    }
  }

  unsigned n_threads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  uint64_t generation_ = 0;
  bool stop_ = false;
  const std::function<void(uint64_t, uint64_t)>* fn_ = nullptr;
  uint64_t n_ = 0, grain_ = 1;
  std::atomic<uint64_t> next_{0};
  std::atomic<int> pending_{0};
};

// Worker threads do not survive fork(): the child gets a pool object whose workers are gone and would wait for them for
// ever. The child therefore starts over with a fresh pool on its first parallel_for (the parent's object is leaked there).
Pool* g_pool = nullptr;
std::once_flag g_atfork_once;
void forget_pool_in_child() { g_pool = nullptr; }

Pool& pool() {
  std::call_once(g_atfork_once, [] { pthread_atfork(nullptr, nullptr, forget_pool_in_child); });
  static std::mutex make_mu;
  std::lock_guard<std::mutex> l(make_mu);
  if (!g_pool) g_pool = new Pool();  // leaked on purpose: worker threads must not be joined from a static destructor
  return *g_pool;
}

}  // namespace

unsigned host_pool_threads() { return pool().threads(); }

void parallel_for(uint64_t n, uint64_t min_grain, const std::function<void(uint64_t, uint64_t)>& fn) {
  if (n == 0) return;
  if (min_grain == 0) min_grain = 1;
  Pool& p = pool();
  if (p.threads() <= 1 || n <= min_grain) {
    fn(0, n);
    return;
  }
  // ~4 ranges per thread for balance, never below the caller's grain
  uint64_t grain = (n + p.threads() * 4 - 1) / (p.threads() * 4);
  if (grain < min_grain) grain = min_grain;
  p.run(n, grain, fn);
}

}  // namespace lc
