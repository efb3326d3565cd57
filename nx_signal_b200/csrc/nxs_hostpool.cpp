// nxs_hostpool.cpp -- host-side worker threads of the "_host" entry points.
//
// The reference returns the full two-sided spectrum of a real signal (lib/nx_signal.ex:49,
// :102, :129), whose bins nfft/2+1 .. nfft-1 are the exact conjugate mirror of bins
// nfft/2-1 .. 1.  The GPU computes the spectrum; moving that redundant half over PCIe would
// double the device->host time, which is what bounds an NIF call.  nxs_stft_f32_host therefore
// copies only bins 0 .. nfft/2 of every frame into the caller's rows and these threads fill in
// z[f][nfft-k] = conj(z[f][k]) (a bit copy with one sign flip -- no arithmetic), slab by slab
// while later slabs are still in flight.  The result is bit-identical to the two-sided device
// output (tests/test_stft_gpu.py::test_host_mirror_bit_identical).
#if defined(__SSE2__)
#include <emmintrin.h>
#define NXS_HAVE_SSE2 1
#else
#define NXS_HAVE_SSE2 0  // aarch64 hosts (Grace): portable scalar paths below
#endif
#include <sched.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "nxs_hostpool.h"

namespace nxs {

void cpu_relax() {
#if NXS_HAVE_SSE2
  _mm_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#endif
}

struct HostPool::Impl {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  uint64_t generation = 0;
  bool stop = false;
  // current job
  void (*fn)(void*, int64_t) = nullptr;
  void* arg = nullptr;
  int64_t n = 0;
  std::atomic<int64_t> next{0};
  int pending = 0;  // workers still inside the current job

  // gated jobs: item i may run only once *gate > i (the caller raises the gate as data lands)
  const std::atomic<int64_t>* gate = nullptr;
  std::atomic<bool> abort{false};

  void drain() {
    for (;;) {
      const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) break;
      if (gate) {
        int spins = 0;
        while (gate->load(std::memory_order_acquire) <= i && !abort.load(std::memory_order_relaxed)) {
          cpu_relax();
          if (++spins == 4096) {
            spins = 0;
            sched_yield();
          }
        }
      }
      if (abort.load(std::memory_order_relaxed)) continue;
      fn(arg, i);
    }
  }

  void worker_main() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
      }
      drain();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
};

int HostPool::default_threads() {
  const char* e = getenv("NXS_HOST_THREADS");
  if (e && atoi(e) > 0) return atoi(e);
  cpu_set_t set;
  int n = 0;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  if (n <= 0) n = 1;
  return n > 32 ? 32 : n;
}

HostPool::HostPool(int nthreads) : impl_(new Impl()), nthreads_(nthreads < 1 ? 1 : nthreads) {
  for (int i = 1; i < nthreads_; ++i) impl_->workers.emplace_back([this] { impl_->worker_main(); });
}

HostPool::~HostPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_work.notify_all();
  for (auto& t : impl_->workers) t.join();
  delete impl_;
}

void HostPool::begin(int64_t n, void (*fn)(void*, int64_t), void* arg, const std::atomic<int64_t>* gate) {
  Impl& s = *impl_;
  {
    std::lock_guard<std::mutex> lk(s.mu);
    s.fn = fn;
    s.arg = arg;
    s.n = n < 0 ? 0 : n;
    s.gate = gate;
    s.abort.store(false, std::memory_order_relaxed);
    s.next.store(0, std::memory_order_relaxed);
    s.pending = (int)s.workers.size();
    ++s.generation;
  }
  s.cv_work.notify_all();
}

void HostPool::finish(bool abort) {
  Impl& s = *impl_;
  if (abort) s.abort.store(true, std::memory_order_relaxed);
  s.drain();  // the caller helps (and is the only runner when there are no workers)
  std::unique_lock<std::mutex> lk(s.mu);
  s.cv_done.wait(lk, [&] { return s.pending == 0; });
  s.gate = nullptr;
}

const std::atomic<bool>* HostPool::abort_flag() const { return &impl_->abort; }

void HostPool::parallel_for(int64_t n, void (*fn)(void*, int64_t), void* arg) {
  if (n <= 0) return;
  begin(n, fn, arg, nullptr);
  finish(false);
}

// z[nfft - k] = conj(z[k]) for k = 1 .. nfft - (nfft/2 + 1), rows [row0, row1) of a
// [rows][nfft] interleaved c64 matrix.
void mirror_rows_c64(float* z, int64_t nfft, int64_t row0, int64_t row1) {
  const int64_t nout = nfft / 2 + 1;
  const int64_t kmax = nfft - nout;  // last k whose mirror lies outside the stored half
#if NXS_HAVE_SSE2
  const __m128 sign = _mm_castsi128_ps(_mm_set_epi32((int)0x80000000u, 0, (int)0x80000000u, 0));
#endif
  for (int64_t r = row0; r < row1; ++r) {
    float* row = z + 2 * r * nfft;
    int64_t k = 1;
#if NXS_HAVE_SSE2
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0 && (nfft & 1) == 0) {
      // pairs (k, k+1), k odd: destination index nfft-k-1 is even -> 16-byte aligned, streamed
      for (; k + 1 <= kmax; k += 2) {
        __m128 v = _mm_loadu_ps(row + 2 * k);              // [re k, im k, re k+1, im k+1]
        v = _mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 3, 2));  // [re k+1, im k+1, re k, im k]
        v = _mm_xor_ps(v, sign);
        _mm_stream_ps(row + 2 * (nfft - k - 1), v);
      }
    }
#endif
    for (; k <= kmax; ++k) {
      row[2 * (nfft - k)] = row[2 * k];
      row[2 * (nfft - k) + 1] = -row[2 * k + 1];
    }
  }
#if NXS_HAVE_SSE2
  _mm_sfence();
#endif
}

// rows [row0, row1) of the caller's [rows][nfft] c64 result from a staged slab: src holds bins
// 0 .. nout-1 of row (row0 + i) at src + 2 * i * src_pitch.  The stored bins are copied; with
// `mirror` (nout == nfft/2 + 1) the bins above nfft/2 are written as conj(src[nfft - j]) in the same
// pass, read from the (cache-resident) slab rather than from the row just written.
void unstage_rows_c64(float* z, int64_t nfft, int64_t nout, const float* src, int64_t src_pitch, int64_t row0,
                      int64_t row1, bool mirror) {
#if NXS_HAVE_SSE2
  const __m128 sign = _mm_castsi128_ps(_mm_set_epi32((int)0x80000000u, 0, (int)0x80000000u, 0));
#endif
  for (int64_t r = row0; r < row1; ++r) {
    float* row = z + 2 * r * nfft;
    const float* s = src + 2 * (r - row0) * src_pitch;
#if NXS_HAVE_SSE2
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {  // streamed: the caller's rows are written once, never read here
      int64_t i = 0;
      for (; i + 4 <= 2 * nout; i += 4) _mm_stream_ps(row + i, _mm_loadu_ps(s + i));
      for (; i < 2 * nout; ++i) row[i] = s[i];
    } else
#endif
      memcpy(row, s, size_t(nout) * 2 * sizeof(float));
    if (!mirror) continue;
    int64_t j = nout;  // destination bin; source bin nfft - j
#if NXS_HAVE_SSE2
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
      if (j & 1) {  // reach an even (16-byte aligned) destination bin
        row[2 * j] = s[2 * (nfft - j)];
        row[2 * j + 1] = -s[2 * (nfft - j) + 1];
        ++j;
      }
      for (; j + 2 <= nfft; j += 2) {
        // dest bins j, j+1 <- conj(src bins nfft-j, nfft-j-1): load src [nfft-j-1, nfft-j], swap halves
        __m128 v = _mm_loadu_ps(s + 2 * (nfft - j - 1));
        v = _mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 3, 2));
        v = _mm_xor_ps(v, sign);
        _mm_stream_ps(row + 2 * j, v);
      }
    }
#endif
    for (; j < nfft; ++j) {
      row[2 * j] = s[2 * (nfft - j)];
      row[2 * j + 1] = -s[2 * (nfft - j) + 1];
    }
  }
#if NXS_HAVE_SSE2
  _mm_sfence();
#endif
}

}  // namespace nxs
