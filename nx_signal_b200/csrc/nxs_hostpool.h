// nxs_hostpool.h -- host worker threads used by the "_host" entry points (see nxs_hostpool.cpp).
#pragma once
#include <stdint.h>

#include <atomic>

namespace nxs {

class HostPool {
 public:
  explicit HostPool(int nthreads);
  ~HostPool();
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;
  // fn(arg, i) for every i in [0, n), on the workers and the calling thread; returns when done
  void parallel_for(int64_t n, void (*fn)(void*, int64_t), void* arg);
  // gated form: begin() starts the workers on items 0 .. n-1 in order, item i waiting until
  // *gate > i (gate == nullptr: no waiting); the caller raises the gate as data becomes ready
  // and then calls finish(), which joins the work and returns when every item is done.
  // finish(true) makes the remaining items no-ops (error paths).
  void begin(int64_t n, void (*fn)(void*, int64_t), void* arg, const std::atomic<int64_t>* gate);
  void finish(bool abort);
  int threads() const { return nthreads_; }
  // set by finish(true): items that wait on their own conditions must give up when it is raised
  const std::atomic<bool>* abort_flag() const;
  // NXS_HOST_THREADS, else the CPUs this process may run on (capped at 32)
  static int default_threads();

 private:
  struct Impl;
  Impl* impl_;
  int nthreads_;
};

// z[r][nfft - k] = conj(z[r][k]) for the bins above nfft/2, rows [row0, row1) of a
// [rows][nfft] interleaved c64 matrix (bit copy + sign flip)
void mirror_rows_c64(float* z, int64_t nfft, int64_t row0, int64_t row1);
// the same result from a staged slab of bins 0 .. nout-1 per row (see nxs_hostpool.cpp)
void unstage_rows_c64(float* z, int64_t nfft, int64_t nout, const float* src, int64_t src_pitch, int64_t row0,
                      int64_t row1, bool mirror);
void cpu_relax();

}  // namespace nxs
