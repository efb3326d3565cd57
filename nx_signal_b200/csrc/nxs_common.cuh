// nxs_common.cuh -- shared declarations of the library's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nxsignal_b200.h"
#include "nxs_hostpool.h"

namespace nxs {

// One device-resident table set per FFT plan (built on first use, lives with the ctx).
struct PlanTables {
  float2* tw = nullptr;    // per-pass twiddles, Plan::TW_TOTAL entries
  float2* post = nullptr;  // r2c post-pass: -i * exp(-i pi k / N), k in [0, N/2]   (N = nfft/2)
};

// Bin-major form of a triangular filterbank for the fused STFT epilogue (nxs_stft.cu): every FFT
// bin k feeds at most two consecutive filters, jl[k] (weight w2[k].x) and jl[k] + 1 (w2[k].y), and
// jl is non-decreasing, so the bins split into segments of equal jl.  A thread owns B consecutive
// bins and emits one partial sum ("piece") per segment it touches; filter j is then the sum of the
// .x pieces of segment j + 1 and the .y pieces of segment j (pieces are numbered in bin order).
struct MelLayout {
  int B = 0;                // bins per thread
  bool ok = false;          // false: the bank is not of that form (or too many pieces) -> gather path
  float2* d_w2 = nullptr;   // [half] (weight for filter jl, weight for filter jl + 1), load-major (nxs_mel.cu)
  int* d_desc = nullptr;    // [half / B] piece base | boundary mask << 12 | skip << 31
  int* d_ps = nullptr;      // [mel_bins + 3] first piece of each segment
};

// sparse mel filterbank (nxs_mel.cu) of one parameter set, resident on the device
struct MelBank {
  int64_t nfft = 0, mel_bins = 0;
  double sr = 0, max_mel = 0, f_sp = 0;
  float* d_wts = nullptr;  // packed nonzero weights
  int nw = 0;              // their count
  int* d_idx = nullptr;    // [3][mel_bins]: start, count, offset
  std::vector<MelLayout> layouts;
};

// fused log-mel epilogue of the STFT kernel (launch_stft): the spectrum is never stored
struct MelEpilogue {
  MelBank* bank;
  float* out;  // [channels * num_frames][mel_bins]
  int* chmax;  // [channels]
};

}  // namespace nxs

struct nxs_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  uint64_t launches = 0;
  // scratch: prepared (scaled, zero-extended) window and small coefficient blocks
  float* d_coef = nullptr;
  size_t d_coef_bytes = 0;
  // generic device scratch (FIR spectra, etc.)
  void* d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  // work buffer of composite device entries (complex-input STFT: real planes + their spectra)
  void* d_work = nullptr;
  size_t d_work_bytes = 0;
  // staging for the _host entry points
  void* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  void* d_stage_in = nullptr;
  size_t d_stage_in_bytes = 0;
  void* d_stage_out = nullptr;
  size_t d_stage_out_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t out_stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // _host STFT: one event per device->host slab, and the threads that mirror the slabs
  std::vector<cudaEvent_t> slab_events;
  nxs::HostPool* pool = nullptr;
  double host_t[4] = {0, 0, 0, 0};  // nxs_ctx_host_timeline
  // nxs_stft_f32_host on pinned results: 0 = both spectrum halves over PCIe, 1 = lower half + host mirror,
  // 3 = mixed (three chunks of four as 1, one as 0);
  // measured cost (seconds per result byte) of each, and the mode pinned by nxs_ctx_set_host_mode (-1: auto)
  double host_cost[4] = {0, 0, 0, 0};
  int host_mode_forced = -1, host_mode_last = 0;
  uint64_t host_calls = 0;
  // stream ordering of the context's shared device state (d_coef, d_scratch, tables under construction):
  // the last stream that used them and an event recorded behind that use (StreamOrder)
  cudaStream_t order_stream = nullptr;
  cudaEvent_t order_event = nullptr;
  bool order_valid = false;
  // optional per-kernel timing (nxs_ctx_profile)
  bool prof_enabled = false;
  std::vector<cudaEvent_t> prof_events;  // start/stop pairs
  size_t prof_used = 0;
  // twiddle tables keyed by (kind << 32 | N)
  std::unordered_map<uint64_t, nxs::PlanTables> tables;
  std::unordered_map<uint64_t, float2*> dft_tables;  // generic DFT: exp(-2 pi i m / n), m < n
  std::vector<nxs::MelBank> mel_banks;
};

namespace nxs {

int set_cuda_error(nxs_ctx* ctx, cudaError_t e, const char* where);
#define NXS_CUDA(ctx, call)                                            \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return nxs::set_cuda_error(ctx, e__, #call); \
  } while (0)

struct PadGeom;
// RAII device switch of every entry point
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// The prepared window (d_coef) and the scratch block (d_scratch) are per context, and every compute
// entry rewrites them on the stream it runs on.  A call arriving on a different stream than the
// previous one first waits (on the device) for that call's work; the previous call's stream is
// never blocked and the host never waits.  Same-stream sequences cost one event record per call.
struct StreamOrder {
  nxs_ctx* ctx;
  cudaStream_t st;
  StreamOrder(nxs_ctx* c, cudaStream_t s) : ctx(c), st(s) {
    if (ctx->order_valid && ctx->order_stream != st) cudaStreamWaitEvent(st, ctx->order_event, 0);
  }
  ~StreamOrder() {
    if (!ctx->order_event && cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    if (cudaEventRecord(ctx->order_event, st) == cudaSuccess) {
      ctx->order_stream = st;
      ctx->order_valid = true;
    } else {
      cudaGetLastError();
    }
  }
};

int stft_check(int64_t channels, int64_t length, int64_t x_ld, int64_t frame_length, int64_t hop, int64_t fft_length,
               int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate, PadGeom* g, int64_t* M);

void prof_begin(nxs_ctx* ctx, cudaStream_t st);
void prof_end(nxs_ctx* ctx, cudaStream_t st);
int ensure_coef(nxs_ctx* ctx, size_t bytes);
int ensure_scratch(nxs_ctx* ctx, size_t bytes);
// exp(+2 pi i m / n), m < n, in double (exact at multiples of pi/2); cached per context
int get_dft_table_f64(nxs_ctx* ctx, int64_t n, double2** out);

// padding geometry shared by stft / as_windowed (lib/nx_signal.ex:303-331, 343-349)
struct PadGeom {
  int64_t lo, hi;  // samples added in front / behind
  int reflect;     // 1: numpy-style reflect, 0: zeros
};
int resolve_padding(int64_t length, int64_t window_length, int pad_mode, int64_t pad_lo, int64_t pad_hi,
                    PadGeom* g);
int64_t frames_for(int64_t length, int64_t window_length, int64_t stride, const PadGeom& g);

// cudaFuncSetAttribute + occupancy query once per (kernel instantiation, device), not per call
struct LaunchCache {
  int occ[16] = {0};
  size_t smem[16] = {0};
  template <class K>
  int get(nxs_ctx* ctx, K kern, int threads, size_t smem_bytes, int* out) {
    const int d = ctx->device & 15;
    if (occ[d] == 0 || smem[d] != smem_bytes) {
      NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      int o = 1;
      NXS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem_bytes));
      smem[d] = smem_bytes;
      occ[d] = o < 1 ? 1 : o;
    }
    *out = occ[d];
    return NXS_OK;
  }
};

// kernels' host launchers (each returns NXS_* and bumps ctx->launches)
int launch_stft(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                const PadGeom& g, int64_t num_frames, int scaling, double sampling_rate, float2* z,
                int64_t z_ld, int onesided, cudaStream_t st, const MelEpilogue* mel = nullptr);
// true when launch_stft serves fft_length with a kernel whose upper half-spectrum is the exact
// conjugate mirror of the lower half (the r2c kernels; not the generic DFT)
bool stft_has_exact_mirror(int64_t fft_length);
// complex `data` (nxs_cplx.cu): two real-plane transforms and a combine pass
int launch_stft_c64(nxs_ctx* ctx, const float2* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, const PadGeom& g,
                    int64_t num_frames, int scaling, double sampling_rate, float2* z, cudaStream_t st);
int launch_istft(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_len,
                 const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                 double sampling_rate, float2* y, cudaStream_t st);
int launch_istft_c2r(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                     double sampling_rate, float* y, cudaStream_t st);
int launch_as_windowed(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                       int64_t x_ld, int64_t window_length, int64_t stride, const PadGeom& g,
                       int64_t num_frames, void* out, cudaStream_t st);
int launch_overlap_and_add(nxs_ctx* ctx, const float* t, int complex_, int64_t batch, int64_t num_frames,
                           int64_t frame_length, int64_t overlap, float* out, cudaStream_t st);
int launch_fir(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
               const float* taps, int64_t num_taps, int mode, float* y, int64_t y_ld, cudaStream_t st);
int launch_stft_to_mel(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                       int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel, double f_sp,
                       float* out, cudaStream_t st);
int launch_stft_mel(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, const PadGeom& g,
                    int64_t num_frames, int scaling, double sampling_rate, int64_t mel_bins, double max_mel,
                    double f_sp, float* out, cudaStream_t st);
// bin-major layout of `bank` for threads owning B consecutive bins (built on first use)
int get_mel_layout(nxs_ctx* ctx, MelBank* bank, int B, const MelLayout** out);
int launch_convolve_nd(nxs_ctx* ctx, const float* a, const int64_t* as, const float* b, const int64_t* bs,
                       int is_complex, int mode, float* out, cudaStream_t st);

int launch_median(nxs_ctx* ctx, const float* t, const int64_t shape[3], const int64_t kernel[3], float* out,
                  cudaStream_t st);
int launch_wiener(nxs_ctx* ctx, const void* t, int is_f64, const int64_t shape[3], const int64_t kernel[3],
                  int has_noise, double noise, void* out, cudaStream_t st);
int launch_argrelextrema(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis, int order, int cmp,
                         int* indices, int64_t* valid_dev, cudaStream_t st);

// division of a non-negative 32-bit numerator by a divisor fixed at launch: one multiply-high and a
// shift instead of the ~20-instruction emulated division (valid for numerators below 2^31)
struct FastDiv {
  unsigned mul = 0, shr = 0;
  int d = 1;
  FastDiv() {}
  explicit FastDiv(int denom) : d(denom) {
    if (denom > 1) {
      unsigned lg = 0;
      while ((1u << lg) < (unsigned)denom) ++lg;  // ceil(log2 denom)
      const unsigned p = 31 + lg;
      mul = (unsigned)((((uint64_t)1 << p) + (uint64_t)denom - 1) / (uint64_t)denom);
      shr = p - 32;
    }
  }
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((unsigned)n, mul) >> shr); }
};

// numpy-style reflect of index i into [0, L)
__host__ __device__ inline int64_t reflect_index(int64_t i, int64_t L) {
  if (L <= 1) return 0;
  const int64_t per = 2 * (L - 1);
  int64_t r = i % per;
  if (r < 0) r += per;
  return r < L ? r : per - r;
}

}  // namespace nxs
