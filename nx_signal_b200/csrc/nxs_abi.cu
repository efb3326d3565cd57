// nxs_abi.cu -- context object and the extern "C" entry points declared in
// include/nxsignal_b200.h.  Argument checking mirrors the reference's ArgumentError
// conditions (cited per entry); no entry point ever falls back to the CPU.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>

#include "nxs_hostio.cuh"

namespace nxs {

int set_cuda_error(nxs_ctx* ctx, cudaError_t e, const char* where) {
  if (ctx) {
    ctx->last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
  }
  cudaGetLastError();  // clear sticky-less errors
  return e == cudaErrorMemoryAllocation ? NXS_ENOMEM : NXS_ECUDA;
}

int grow_buf(nxs_ctx* ctx, void** p, size_t* have, size_t need, bool host) {
  if (*have >= need) return NXS_OK;
  if (*p) {
    if (host) cudaFreeHost(*p);
    else cudaFree(*p);
    *p = nullptr;
    *have = 0;
  }
  size_t sz = need + need / 8;
  cudaError_t e = host ? cudaMallocHost(p, sz) : cudaMalloc(p, sz);
  if (e != cudaSuccess) {
    sz = need;
    e = host ? cudaMallocHost(p, sz) : cudaMalloc(p, sz);
  }
  if (e != cudaSuccess) return set_cuda_error(ctx, e, host ? "cudaMallocHost" : "cudaMalloc");
  *have = sz;
  return NXS_OK;
}

void prof_begin(nxs_ctx* ctx, cudaStream_t st) {
  if (!ctx->prof_enabled) return;
  if (ctx->prof_used + 2 > ctx->prof_events.size()) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    ctx->prof_events.push_back(a);
    ctx->prof_events.push_back(b);
  }
  cudaEventRecord(ctx->prof_events[ctx->prof_used], st);
}

void prof_end(nxs_ctx* ctx, cudaStream_t st) {
  if (!ctx->prof_enabled || ctx->prof_used + 2 > ctx->prof_events.size()) return;
  cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], st);
  ctx->prof_used += 2;
}

int ensure_coef(nxs_ctx* ctx, size_t bytes) {
  void* p = ctx->d_coef;
  int rc = grow_buf(ctx, &p, &ctx->d_coef_bytes, bytes < 4096 ? 4096 : bytes, false);
  ctx->d_coef = (float*)p;
  return rc;
}

int ensure_scratch(nxs_ctx* ctx, size_t bytes) {
  return grow_buf(ctx, &ctx->d_scratch, &ctx->d_scratch_bytes, bytes, false);
}

int resolve_padding(int64_t length, int64_t window_length, int pad_mode, int64_t pad_lo, int64_t pad_hi,
                    PadGeom* g) {
  (void)length;
  g->reflect = 0;
  switch (pad_mode) {
    case NXS_PAD_VALID: g->lo = g->hi = 0; return NXS_OK;
    case NXS_PAD_SAME: {  // lib/nx_signal.ex:308-312
      int64_t total = window_length - 1;
      if (total < 0) total = 0;
      g->lo = total / 2;
      g->hi = total - g->lo;
      return NXS_OK;
    }
    case NXS_PAD_REFLECT:  // lib/nx_signal.ex:257-264, 343-349
      g->lo = g->hi = window_length / 2;
      g->reflect = 1;
      return NXS_OK;
    case NXS_PAD_EXPLICIT: g->lo = pad_lo; g->hi = pad_hi; return NXS_OK;
    default: return NXS_EINVAL;  // lib/nx_signal.ex:325-329
  }
}

int64_t frames_for(int64_t length, int64_t window_length, int64_t stride, const PadGeom& g) {
  const int64_t padded = length + g.lo + g.hi;
  return padded < window_length ? 0 : (padded - window_length) / stride + 1;
}

// _dev entries run on the caller's stream; NULL is CUDA's (legacy) default stream, as everywhere in CUDA
static cudaStream_t pick(nxs_ctx*, void* stream) { return (cudaStream_t)stream; }

// host-pointer helper: stage in -> run -> stage out, on the context's stream
template <class F>
static int host_roundtrip(nxs_ctx* ctx, const void* in, size_t in_bytes, const void* in2, size_t in2_bytes,
                          void* out, size_t out_bytes, F&& run) {
  int rc = grow_buf(ctx, &ctx->d_stage_in, &ctx->d_stage_in_bytes, in_bytes + in2_bytes + 512, false);
  if (rc) return rc;
  rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, out_bytes + 256, false);
  if (rc) return rc;
  char* d_in = (char*)ctx->d_stage_in;
  const size_t off2 = (in_bytes + 255) / 256 * 256;
  char* d_in2 = d_in + off2;
  if (in_bytes) NXS_CUDA(ctx, cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (in2_bytes) NXS_CUDA(ctx, cudaMemcpyAsync(d_in2, in2, in2_bytes, cudaMemcpyHostToDevice, ctx->stream));
  rc = run(d_in, d_in2, ctx->d_stage_out);
  if (rc) return rc;
  if (out_bytes) NXS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NXS_OK;
}

int stft_check(int64_t channels, int64_t length, int64_t x_ld, int64_t frame_length, int64_t hop,
                      int64_t fft_length, int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling,
                      double sampling_rate, PadGeom* g, int64_t* M) {
  if (channels < 0 || length < 1 || x_ld < length || frame_length < 1 || fft_length < 1) return NXS_ESHAPE;
  if (hop < 1) return NXS_EINVAL;  // stride must be an integer >= 1 (lib/nx_signal.ex:279-284)
  if (scaling != NXS_SCALE_NONE && scaling != NXS_SCALE_SPECTRUM && scaling != NXS_SCALE_PSD)
    return NXS_EINVAL;  // lib/nx_signal.ex:124-126
  if (!(sampling_rate == sampling_rate)) return NXS_EINVAL;
  int rc = resolve_padding(length, frame_length, pad_mode, pad_lo, pad_hi, g);
  if (rc) return rc;
  *M = frames_for(length, frame_length, hop, *g);
  return NXS_OK;
}

}  // namespace nxs

using namespace nxs;

extern "C" {

int nxs_abi_version(void) { return NXS_ABI_VERSION; }

const char* nxs_strerror(int code) {
  switch (code) {
    case NXS_OK: return "ok";
    case NXS_EINVAL: return "invalid argument";
    case NXS_ESHAPE: return "incompatible shapes";
    case NXS_EUNSUPPORTED: return "unsupported configuration";
    case NXS_ECUDA: return "CUDA error";
    case NXS_ENCCL: return "NCCL error";
    case NXS_ENOMEM: return "out of memory";
    case NXS_ENODEVICE: return "no CUDA device (this backend has no CPU path)";
    default: return "unknown error";
  }
}

int nxs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nxs_ctx_create(int device, nxs_ctx** out) {
  if (!out) return NXS_EINVAL;
  *out = nullptr;
  const int n = nxs_device_count();
  if (n <= 0) return NXS_ENODEVICE;
  if (device < 0 || device >= n) return NXS_EINVAL;
  nxs_ctx* ctx = new nxs_ctx();
  ctx->device = device;
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  if (!guard.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    cudaGetLastError();
    return NXS_ECUDA;
  }
  for (auto& e : ctx->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return NXS_OK;
}

int nxs_ctx_destroy(nxs_ctx* ctx) {
  if (!ctx) return NXS_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->tables) {
    cudaFree(kv.second.tw);
    cudaFree(kv.second.post);
  }
  for (auto& kv : ctx->dft_tables) cudaFree(kv.second);
  for (auto& b : ctx->mel_banks) {
    cudaFree(b.d_wts);
    cudaFree(b.d_idx);
    for (auto& l : b.layouts) {
      cudaFree(l.d_w2);
      cudaFree(l.d_desc);
      cudaFree(l.d_ps);
    }
  }
  cudaFree(ctx->d_coef);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_work);
  cudaFree(ctx->d_stage_in);
  cudaFree(ctx->d_stage_out);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->prof_events) cudaEventDestroy(e);
  for (auto& e : ctx->slab_events) cudaEventDestroy(e);
  if (ctx->order_event) cudaEventDestroy(ctx->order_event);
  delete ctx->pool;
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
  delete ctx;
  return NXS_OK;
}

const char* nxs_last_error(const nxs_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int nxs_ctx_synchronize(nxs_ctx* ctx) {
  if (!ctx) return NXS_EINVAL;
  DeviceGuard guard(ctx->device);
  NXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NXS_OK;
}

uint64_t nxs_ctx_launch_count(const nxs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int nxs_ctx_profile(nxs_ctx* ctx, int enable) {
  if (!ctx) return NXS_EINVAL;
  ctx->prof_enabled = enable != 0;
  return NXS_OK;
}

int nxs_ctx_profile_read(nxs_ctx* ctx, double* total_ms, int64_t* launches) {
  if (!ctx || !total_ms || !launches) return NXS_EINVAL;
  DeviceGuard guard(ctx->device);
  double sum = 0.0;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    NXS_CUDA(ctx, cudaEventSynchronize(ctx->prof_events[i + 1]));
    float ms = 0.f;
    NXS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *launches = (int64_t)(ctx->prof_used / 2);
  ctx->prof_used = 0;
  return NXS_OK;
}

int nxs_ctx_host_timeline(const nxs_ctx* ctx, double out_seconds[4]) {
  if (!ctx || !out_seconds) return NXS_EINVAL;
  for (int i = 0; i < 4; ++i) out_seconds[i] = ctx->host_t[i];
  return NXS_OK;
}

// ---- STFT -----------------------------------------------------------------------------------
int nxs_stft_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                     int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                     float* z, void* stream) {
  if (!ctx || !x || !window || !z) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi,
                      scaling, sampling_rate, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_stft(ctx, x, channels, length, x_ld, window, frame_length, hop, fft_length, g, M, scaling,
                     sampling_rate, reinterpret_cast<float2*>(z), fft_length, 0, pick(ctx, stream));
}

int nxs_stft_onesided_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                              const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                              int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                              float* z, int64_t z_ld, void* stream) {
  if (!ctx || !x || !window || !z) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi,
                      scaling, sampling_rate, &g, &M);
  if (rc) return rc;
  if (z_ld < fft_length / 2 + 1) return NXS_ESHAPE;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_stft(ctx, x, channels, length, x_ld, window, frame_length, hop, fft_length, g, M, scaling,
                     sampling_rate, reinterpret_cast<float2*>(z), z_ld, 1, pick(ctx, stream));
}

// nxs_stft_f32_host: nxs_hostio.cu (channel-chunk pipeline, one-sided transfer + host mirror, pinned rings)

// ---- ISTFT ----------------------------------------------------------------------------------
static int istft_check(int64_t channels, int64_t num_frames, int64_t z_len, int64_t frame_length, int64_t hop,
                       int64_t fft_length, int scaling) {
  if (channels < 0 || num_frames < 1 || z_len < 1 || frame_length < 1 || fft_length < 1) return NXS_ESHAPE;
  if (fft_length != frame_length) return NXS_ESHAPE;  // `frames * window` must broadcast (lib/nx_signal.ex:628)
  if (hop < 1 || hop > frame_length) return NXS_EINVAL;  // overlap_length >= 0 and < window (lib/nx_signal.ex:692-695)
  if (scaling != NXS_SCALE_NONE && scaling != NXS_SCALE_SPECTRUM && scaling != NXS_SCALE_PSD)
    return NXS_EINVAL;  // lib/nx_signal.ex:622-624
  return NXS_OK;
}

int nxs_istft_c64_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_len,
                      const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                      int scaling, double sampling_rate, float* y, void* stream) {
  if (!ctx || !z || !window || !y) return NXS_EINVAL;
  int rc = istft_check(channels, num_frames, z_len, frame_length, hop, fft_length, scaling);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_istft(ctx, reinterpret_cast<const float2*>(z), channels, num_frames, z_len, window,
                      frame_length, hop, fft_length, scaling, sampling_rate, reinterpret_cast<float2*>(y),
                      pick(ctx, stream));
}

int nxs_istft_c64_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_len,
                       const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                       int scaling, double sampling_rate, float* y) {
  if (!ctx || !z || !window || !y) return NXS_EINVAL;
  int rc = istft_check(channels, num_frames, z_len, frame_length, hop, fft_length, scaling);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  // channels are independent: H2D of the spectrum | kernels | D2H of the signal, chunk by chunk
  PipeSpec ps;
  ps.in = z;
  ps.in_row_bytes = ps.in_pitch = size_t(num_frames) * z_len * sizeof(float2);
  ps.out = y;
  ps.out_row_bytes = ps.out_pitch = (size_t(num_frames) * hop + (frame_length - hop)) * sizeof(float2);
  ps.rows = channels;
  ps.aux = window;
  ps.aux_bytes = size_t(frame_length) * sizeof(float);
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dz, void* dw, void* dy) {
    return launch_istft(ctx, (const float2*)dz, n, num_frames, z_len, (const float*)dw, frame_length, hop, fft_length,
                        scaling, sampling_rate, (float2*)dy, ctx->stream);
  });
}

static int istft_c2r_check(int64_t channels, int64_t num_frames, int64_t z_ld, int64_t frame_length, int64_t hop,
                           int64_t fft_length, int scaling) {
  int rc = istft_check(channels, num_frames, z_ld, frame_length, hop, fft_length, scaling);
  if (rc) return rc;
  if ((fft_length & 1) || z_ld < fft_length / 2 + 1) return NXS_ESHAPE;
  return NXS_OK;
}

int nxs_istft_c2r_f32_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                          const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                          int scaling, double sampling_rate, float* y, void* stream) {
  if (!ctx || !z || !window || !y) return NXS_EINVAL;
  int rc = istft_c2r_check(channels, num_frames, z_ld, frame_length, hop, fft_length, scaling);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_istft_c2r(ctx, reinterpret_cast<const float2*>(z), channels, num_frames, z_ld, window,
                          frame_length, hop, fft_length, scaling, sampling_rate, y, pick(ctx, stream));
}

int nxs_istft_c2r_f32_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                           const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                           int scaling, double sampling_rate, float* y) {
  if (!ctx || !z || !window || !y) return NXS_EINVAL;
  int rc = istft_c2r_check(channels, num_frames, z_ld, frame_length, hop, fft_length, scaling);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  PipeSpec ps;
  ps.in = z;
  ps.in_row_bytes = ps.in_pitch = size_t(num_frames) * z_ld * sizeof(float2);
  ps.out = y;
  ps.out_row_bytes = ps.out_pitch = (size_t(num_frames) * hop + (frame_length - hop)) * sizeof(float);
  ps.rows = channels;
  ps.aux = window;
  ps.aux_bytes = size_t(frame_length) * sizeof(float);
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dz, void* dw, void* dy) {
    return launch_istft_c2r(ctx, (const float2*)dz, n, num_frames, z_ld, (const float*)dw, frame_length, hop,
                            fft_length, scaling, sampling_rate, (float*)dy, ctx->stream);
  });
}

// ---- stft_to_mel --------------------------------------------------------------------------------
static int mel_check(int64_t channels, int64_t num_frames, int64_t z_ld, int64_t fft_length, int64_t mel_bins,
                     double sampling_rate, double f_sp) {
  if (channels < 0 || num_frames < 1 || fft_length < 2 || mel_bins < 1) return NXS_ESHAPE;
  if (z_ld < fft_length / 2) return NXS_ESHAPE;  // slice_along_axis(0, div(fft_length, 2)) must fit (lib/nx_signal.ex:496)
  if (!(sampling_rate == sampling_rate) || !(f_sp > 0)) return NXS_EINVAL;
  return NXS_OK;
}

int nxs_stft_to_mel_f32_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                            int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel,
                            double mel_frequency_spacing, float* out, void* stream) {
  if (!ctx || !z || !out) return NXS_EINVAL;
  int rc = mel_check(channels, num_frames, z_ld, fft_length, mel_bins, sampling_rate, mel_frequency_spacing);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_stft_to_mel(ctx, reinterpret_cast<const float2*>(z), channels, num_frames, z_ld, fft_length, mel_bins,
                            sampling_rate, max_mel, mel_frequency_spacing, out, pick(ctx, stream));
}

int nxs_stft_to_mel_f32_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                             int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel,
                             double mel_frequency_spacing, float* out) {
  if (!ctx || !z || !out) return NXS_EINVAL;
  int rc = mel_check(channels, num_frames, z_ld, fft_length, mel_bins, sampling_rate, mel_frequency_spacing);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  PipeSpec ps;  // the clamp's maximum is per channel, so channels are independent rows
  ps.in = z;
  ps.in_row_bytes = ps.in_pitch = size_t(num_frames) * z_ld * sizeof(float2);
  ps.out = out;
  ps.out_row_bytes = ps.out_pitch = size_t(num_frames) * mel_bins * sizeof(float);
  ps.rows = channels;
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dz, void*, void* dout) {
    return launch_stft_to_mel(ctx, (const float2*)dz, n, num_frames, z_ld, fft_length, mel_bins, sampling_rate,
                              max_mel, mel_frequency_spacing, (float*)dout, ctx->stream);
  });
}

int nxs_stft_mel_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                         const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int pad_mode,
                         int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate, int64_t mel_bins,
                         double max_mel, double mel_frequency_spacing, float* out, void* stream) {
  if (!ctx || !x || !window || !out) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling,
                      sampling_rate, &g, &M);
  if (rc) return rc;
  if (M > 0) {
    rc = mel_check(channels, M, fft_length, fft_length, mel_bins, sampling_rate, mel_frequency_spacing);
    if (rc) return rc;
  }
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_stft_mel(ctx, x, channels, length, x_ld, window, frame_length, hop, fft_length, g, M, scaling,
                         sampling_rate, mel_bins, max_mel, mel_frequency_spacing, out, pick(ctx, stream));
}

int nxs_stft_mel_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                          const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int pad_mode,
                          int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate, int64_t mel_bins,
                          double max_mel, double mel_frequency_spacing, float* out) {
  if (!ctx || !x || !window || !out) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling,
                      sampling_rate, &g, &M);
  if (rc) return rc;
  if (M <= 0 || channels <= 0) return NXS_OK;
  rc = mel_check(channels, M, fft_length, fft_length, mel_bins, sampling_rate, mel_frequency_spacing);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t in_bytes = (size_t(channels - 1) * x_ld + length) * sizeof(float);
  const size_t w_bytes = size_t(frame_length) * sizeof(float);
  const size_t row_out = size_t(M) * mel_bins;  // floats per channel
  const size_t out_bytes = size_t(channels) * row_out * sizeof(float);
  rc = grow_buf(ctx, &ctx->d_stage_in, &ctx->d_stage_in_bytes, in_bytes + w_bytes + 512, false);
  if (rc) return rc;
  rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, out_bytes + 256, false);
  if (rc) return rc;
  float* dx = (float*)ctx->d_stage_in;
  float* dw = (float*)((char*)ctx->d_stage_in + (in_bytes + 255) / 256 * 256);
  float* dout = (float*)ctx->d_stage_out;
  NXS_CUDA(ctx, cudaMemcpyAsync(dw, window, w_bytes, cudaMemcpyHostToDevice, ctx->stream));

  // Channels are independent (the clamp's maximum is per channel), so the call is pipelined over
  // channel chunks: H2D of chunk k+1 | kernels of chunk k | D2H of chunk k-1 on three streams.
  const int64_t nchunks = channels < 8 ? channels : 8;
  const int64_t per = (channels + nchunks - 1) / nchunks;
  bool chained = false;
  for (int64_t c0 = 0, k = 0; c0 < channels && !chained; c0 += per, ++k) {
    const int64_t nc = channels - c0 < per ? channels - c0 : per;
    while ((size_t)(2 * k + 2) > ctx->slab_events.size()) {
      cudaEvent_t e;
      NXS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->slab_events.push_back(e);
    }
    const size_t cb = (size_t(nc - 1) * x_ld + length) * sizeof(float);
    NXS_CUDA(ctx, cudaMemcpyAsync(dx + c0 * x_ld, x + c0 * x_ld, cb, cudaMemcpyHostToDevice, ctx->copy_stream));
    NXS_CUDA(ctx, cudaEventRecord(ctx->slab_events[2 * k], ctx->copy_stream));
    NXS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->slab_events[2 * k], 0));
    rc = launch_stft_mel(ctx, dx + c0 * x_ld, nc, length, x_ld, dw, frame_length, hop, fft_length, g, M, scaling,
                         sampling_rate, mel_bins, max_mel, mel_frequency_spacing, dout + c0 * row_out, ctx->stream);
    if (rc == NXS_EUNSUPPORTED && k == 0) {
      chained = true;  // the fused kernel does not serve this configuration: one unpipelined chained pass below
      break;
    }
    if (rc) break;
    NXS_CUDA(ctx, cudaEventRecord(ctx->slab_events[2 * k + 1], ctx->stream));
    NXS_CUDA(ctx, cudaStreamWaitEvent(ctx->out_stream, ctx->slab_events[2 * k + 1], 0));
    NXS_CUDA(ctx, cudaMemcpyAsync(out + c0 * row_out, dout + c0 * row_out, size_t(nc) * row_out * sizeof(float),
                                  cudaMemcpyDeviceToHost, ctx->out_stream));
  }
  if (chained) {
    NXS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    NXS_CUDA(ctx, cudaMemcpyAsync(dx, x, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    // spectrum (one-sided when the length allows) to a temporary, then the mel kernel
    const bool one = stft_has_exact_mirror(fft_length);
    const int64_t z_ld = one ? fft_length / 2 + 1 : fft_length;
    float2* z = nullptr;
    cudaError_t e = cudaMalloc(&z, size_t(channels) * M * z_ld * sizeof(float2));
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaMalloc(stft_mel spectrum)");
    rc = launch_stft(ctx, dx, channels, length, x_ld, dw, frame_length, hop, fft_length, g, M, scaling, sampling_rate, z,
                     z_ld, one ? 1 : 0, ctx->stream);
    if (rc == NXS_OK)
      rc = launch_stft_to_mel(ctx, z, channels, M, z_ld, fft_length, mel_bins, sampling_rate, max_mel,
                              mel_frequency_spacing, dout, ctx->stream);
    if (rc == NXS_OK) {
      cudaError_t e2 = cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
      if (e2 != cudaSuccess) rc = set_cuda_error(ctx, e2, "cudaMemcpyAsync(mel result)");
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(z);
    return rc;
  }
  // nothing of ours may still read or write the caller's buffers when the call returns
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->out_stream);
  if (rc) return rc;
  NXS_CUDA(ctx, e1);
  NXS_CUDA(ctx, e2);
  NXS_CUDA(ctx, e3);
  return NXS_OK;
}

// ---- median / wiener / argrelextrema (nxs_post.cu) ------------------------------------------------
// rank <= 3 tensors are viewed as 3-d with leading dimensions of size 1
static int pad3(int rank, const int64_t* shape, const int64_t* kernel, int64_t s3[3], int64_t k3[3]) {
  if (rank < 0 || rank > 3 || (rank > 0 && (!shape || !kernel))) return NXS_ESHAPE;
  for (int i = 0; i < 3; ++i) s3[i] = k3[i] = 1;
  for (int i = 0; i < rank; ++i) {
    s3[3 - rank + i] = shape[i];
    k3[3 - rank + i] = kernel[i];
    if (shape[i] < 0 || kernel[i] < 1 || shape[i] >= (int64_t(1) << 31)) return NXS_ESHAPE;
  }
  return NXS_OK;
}

static int median_check(int rank, const int64_t* shape, const int64_t* kernel, int64_t s3[3], int64_t k3[3]) {
  int rc = pad3(rank, shape, kernel, s3, k3);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i)
    if (s3[i] > 0 && k3[i] > s3[i]) return NXS_ESHAPE;  // Nx.slice: the window must fit inside the tensor
  return NXS_OK;
}

int nxs_median_f32_dev(nxs_ctx* ctx, const float* t, int rank, const int64_t* shape, const int64_t* kernel_shape,
                       float* out, void* stream) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int64_t s3[3], k3[3];
  int rc = median_check(rank, shape, kernel_shape, s3, k3);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_median(ctx, t, s3, k3, out, pick(ctx, stream));
}

int nxs_median_f32_host(nxs_ctx* ctx, const float* t, int rank, const int64_t* shape, const int64_t* kernel_shape,
                        float* out) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int64_t s3[3], k3[3];
  int rc = median_check(rank, shape, kernel_shape, s3, k3);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t bytes = size_t(s3[0] * s3[1] * s3[2]) * sizeof(float);
  if (k3[0] == 1 && s3[0] > 1) {  // the window does not span the leading axis: its slices are independent rows
    PipeSpec ps;
    ps.in = t;
    ps.out = out;
    ps.in_row_bytes = ps.in_pitch = ps.out_row_bytes = ps.out_pitch = size_t(s3[1] * s3[2]) * sizeof(float);
    ps.rows = s3[0];
    return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dt, void*, void* dout) {
      const int64_t sn[3] = {n, s3[1], s3[2]};
      return launch_median(ctx, (const float*)dt, sn, k3, (float*)dout, ctx->stream);
    });
  }
  return host_roundtrip(ctx, t, bytes, nullptr, 0, out, bytes, [&](void* dt, void*, void* dout) {
    return launch_median(ctx, (const float*)dt, s3, k3, (float*)dout, ctx->stream);
  });
}

int nxs_wiener_dev(nxs_ctx* ctx, const void* t, int is_f64, int rank, const int64_t* shape, const int64_t* kernel_size,
                   int has_noise, double noise, void* out, void* stream) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int64_t s3[3], k3[3];
  int rc = pad3(rank, shape, kernel_size, s3, k3);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_wiener(ctx, t, is_f64 != 0, s3, k3, has_noise != 0, noise, out, pick(ctx, stream));
}

int nxs_wiener_host(nxs_ctx* ctx, const void* t, int is_f64, int rank, const int64_t* shape, const int64_t* kernel_size,
                    int has_noise, double noise, void* out) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int64_t s3[3], k3[3];
  int rc = pad3(rank, shape, kernel_size, s3, k3);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t bytes = size_t(s3[0] * s3[1] * s3[2]) * (is_f64 ? sizeof(double) : sizeof(float));
  return host_roundtrip(ctx, t, bytes, nullptr, 0, out, bytes, [&](void* dt, void*, void* dout) {
    return launch_wiener(ctx, dt, is_f64 != 0, s3, k3, has_noise != 0, noise, dout, ctx->stream);
  });
}

static int argrel_check(int rank, const int64_t* shape, int axis, int order, int comparator, int64_t* total) {
  if (rank < 1 || rank > 8 || !shape) return NXS_ESHAPE;
  if (axis < 0 || axis >= rank || order < 1) return NXS_EINVAL;
  if (comparator < NXS_CMP_LESS || comparator > NXS_CMP_GREATER_EQUAL) return NXS_EINVAL;
  int64_t n = 1;
  for (int i = 0; i < rank; ++i) {
    if (shape[i] < 0) return NXS_ESHAPE;
    n *= shape[i];
    if (n >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  }
  *total = n;
  return NXS_OK;
}

int nxs_argrelextrema_f32_dev(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis, int order,
                              int comparator, int32_t* indices, int64_t* valid_count, void* stream) {
  if (!ctx || !data || !indices || !valid_count) return NXS_EINVAL;
  int64_t total = 0;
  int rc = argrel_check(rank, shape, axis, order, comparator, &total);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_argrelextrema(ctx, data, rank, shape, axis, order, comparator, indices, valid_count, pick(ctx, stream));
}

int nxs_argrelextrema_f32_host(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis, int order,
                               int comparator, int32_t* indices, int64_t* valid_count) {
  if (!ctx || !data || !indices || !valid_count) return NXS_EINVAL;
  int64_t total = 0;
  int rc = argrel_check(rank, shape, axis, order, comparator, &total);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t idx_bytes = size_t(total) * rank * sizeof(int32_t);
  const size_t idx_pad = (idx_bytes + 255) / 256 * 256;
  // staged result = indices followed by the count
  rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, idx_pad + 256, false);
  if (rc) return rc;
  rc = host_roundtrip(ctx, data, size_t(total) * sizeof(float), nullptr, 0, indices, idx_bytes,
                      [&](void* dd, void*, void* dout) {
                        return launch_argrelextrema(ctx, (const float*)dd, rank, shape, axis, order, comparator,
                                                    (int*)dout, (int64_t*)((char*)dout + idx_pad), ctx->stream);
                      });
  if (rc) return rc;
  NXS_CUDA(ctx, cudaMemcpy(valid_count, (char*)ctx->d_stage_out + idx_pad, sizeof(int64_t), cudaMemcpyDeviceToHost));
  return NXS_OK;
}

// ---- as_windowed ------------------------------------------------------------------------------
static int aw_check(int elem_size, int64_t channels, int64_t length, int64_t x_ld, int64_t window_length,
                    int64_t stride, int pad_mode, int64_t pad_lo, int64_t pad_hi, PadGeom* g, int64_t* M) {
  if (elem_size != 4 && elem_size != 8) return NXS_EUNSUPPORTED;
  if (channels < 0 || length < 1 || x_ld < length || window_length < 1) return NXS_ESHAPE;
  if (stride < 1) return NXS_EINVAL;  // lib/nx_signal.ex:282-284
  int rc = resolve_padding(length, window_length, pad_mode, pad_lo, pad_hi, g);
  if (rc) return rc;
  *M = frames_for(length, window_length, stride, *g);
  return NXS_OK;
}

int nxs_as_windowed_dev(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                        int64_t x_ld, int64_t window_length, int64_t stride, int pad_mode, int64_t pad_lo,
                        int64_t pad_hi, void* out, void* stream) {
  if (!ctx || !x || !out) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = aw_check(elem_size, channels, length, x_ld, window_length, stride, pad_mode, pad_lo, pad_hi, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_as_windowed(ctx, x, elem_size, channels, length, x_ld, window_length, stride, g, M, out,
                            pick(ctx, stream));
}

int nxs_as_windowed_host(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                         int64_t x_ld, int64_t window_length, int64_t stride, int pad_mode, int64_t pad_lo,
                         int64_t pad_hi, void* out) {
  if (!ctx || !x || !out) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = aw_check(elem_size, channels, length, x_ld, window_length, stride, pad_mode, pad_lo, pad_hi, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (channels == 0 || M == 0) return NXS_OK;
  StreamOrder stream_order(ctx, ctx->stream);
  PipeSpec ps;
  ps.in = x;
  ps.in_row_bytes = size_t(length) * elem_size;
  ps.in_pitch = size_t(x_ld) * elem_size;
  ps.out = out;
  ps.out_row_bytes = ps.out_pitch = size_t(M) * window_length * elem_size;
  ps.rows = channels;
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dx, void*, void* dout) {
    return launch_as_windowed(ctx, dx, elem_size, n, length, x_ld, window_length, stride, g, M, dout, ctx->stream);
  });
}

// ---- overlap_and_add --------------------------------------------------------------------------
static int ola_check(int64_t batch, int64_t num_frames, int64_t frame_length, int64_t overlap) {
  if (batch < 0 || num_frames < 1 || frame_length < 1) return NXS_ESHAPE;
  if (overlap >= frame_length || overlap < 0) return NXS_EINVAL;  // lib/nx_signal.ex:692-695
  return NXS_OK;
}

static int ola_dev(nxs_ctx* ctx, const float* t, int cplx, int64_t batch, int64_t num_frames,
                   int64_t frame_length, int64_t overlap, float* out, void* stream) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int rc = ola_check(batch, num_frames, frame_length, overlap);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_overlap_and_add(ctx, t, cplx, batch, num_frames, frame_length, overlap, out, pick(ctx, stream));
}

static int ola_host(nxs_ctx* ctx, const float* t, int cplx, int64_t batch, int64_t num_frames,
                    int64_t frame_length, int64_t overlap, float* out) {
  if (!ctx || !t || !out) return NXS_EINVAL;
  int rc = ola_check(batch, num_frames, frame_length, overlap);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t es = cplx ? 8 : 4;
  PipeSpec ps;
  ps.in = t;
  ps.in_row_bytes = ps.in_pitch = size_t(num_frames) * frame_length * es;
  ps.out = out;
  ps.out_row_bytes = ps.out_pitch = (size_t(num_frames) * (frame_length - overlap) + overlap) * es;
  ps.rows = batch;
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dt, void*, void* dout) {
    return launch_overlap_and_add(ctx, (const float*)dt, cplx, n, num_frames, frame_length, overlap, (float*)dout,
                                  ctx->stream);
  });
}

int nxs_overlap_and_add_f32_dev(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                int64_t frame_length, int64_t overlap_length, float* out, void* stream) {
  return ola_dev(ctx, t, 0, batch, num_frames, frame_length, overlap_length, out, stream);
}
int nxs_overlap_and_add_c64_dev(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                int64_t frame_length, int64_t overlap_length, float* out, void* stream) {
  return ola_dev(ctx, t, 1, batch, num_frames, frame_length, overlap_length, out, stream);
}
int nxs_overlap_and_add_f32_host(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                 int64_t frame_length, int64_t overlap_length, float* out) {
  return ola_host(ctx, t, 0, batch, num_frames, frame_length, overlap_length, out);
}
int nxs_overlap_and_add_c64_host(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                 int64_t frame_length, int64_t overlap_length, float* out) {
  return ola_host(ctx, t, 1, batch, num_frames, frame_length, overlap_length, out);
}

// ---- FIR --------------------------------------------------------------------------------------
static int fir_check(int64_t channels, int64_t length, int64_t x_ld, int64_t num_taps, int mode,
                     int64_t y_ld, int64_t* out_len) {
  if (channels < 0 || length < 1 || x_ld < length || num_taps < 1) return NXS_ESHAPE;
  int rc = nxs_fir_out_len(length, num_taps, mode, out_len);  // convolution.ex:41-44
  if (rc) return rc;
  if (y_ld < *out_len) return NXS_ESHAPE;
  return NXS_OK;
}

int nxs_fir_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* taps, int64_t num_taps, int mode, float* y, int64_t y_ld, void* stream) {
  if (!ctx || !x || !taps || !y) return NXS_EINVAL;
  int64_t out_len = 0;
  int rc = fir_check(channels, length, x_ld, num_taps, mode, y_ld, &out_len);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_fir(ctx, x, channels, length, x_ld, taps, num_taps, mode, y, y_ld, pick(ctx, stream));
}

int nxs_fir_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* taps, int64_t num_taps, int mode, float* y, int64_t y_ld) {
  if (!ctx || !x || !taps || !y) return NXS_EINVAL;
  int64_t out_len = 0;
  int rc = fir_check(channels, length, x_ld, num_taps, mode, y_ld, &out_len);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  // channels are independent: H2D | overlap-save kernels | D2H, chunk by chunk -- the call costs
  // max(H2D, D2H) instead of their sum (PCIe is full duplex)
  PipeSpec ps;
  ps.in = x;
  ps.in_row_bytes = size_t(length) * sizeof(float);
  ps.in_pitch = size_t(x_ld) * sizeof(float);
  ps.out = y;
  ps.out_row_bytes = size_t(out_len) * sizeof(float);
  ps.out_pitch = size_t(y_ld) * sizeof(float);
  ps.rows = channels;
  ps.aux = taps;
  ps.aux_bytes = size_t(num_taps) * sizeof(float);
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dx, void* dt, void* dy) {
    return launch_fir(ctx, (const float*)dx, n, length, x_ld, (const float*)dt, num_taps, mode, (float*)dy, y_ld,
                      ctx->stream);
  });
}

// ---- N-d convolution ----------------------------------------------------------------------------
static int conv_check(const int64_t* as, const int64_t* bs, int mode, int64_t* os) {
  if (mode != NXS_MODE_FULL && mode != NXS_MODE_SAME && mode != NXS_MODE_VALID) return NXS_EINVAL;
  bool ok1 = true, ok2 = true;
  for (int d = 0; d < 3; ++d) {
    if (as[d] < 1 || bs[d] < 1) return NXS_ESHAPE;
    ok1 = ok1 && as[d] >= bs[d];
    ok2 = ok2 && as[d] <= bs[d];
  }
  for (int d = 0; d < 3; ++d) {
    if (mode == NXS_MODE_FULL) os[d] = as[d] + bs[d] - 1;
    else if (mode == NXS_MODE_SAME) os[d] = as[d];
    else {
      if (!ok1 && !ok2) return NXS_ESHAPE;  // convolution.ex:131-134, 342-345
      os[d] = (ok1 ? as[d] - bs[d] : bs[d] - as[d]) + 1;
    }
  }
  return NXS_OK;
}

int nxs_convolve_nd_dev(nxs_ctx* ctx, const float* a, const int64_t a_shape[3], const float* b,
                        const int64_t b_shape[3], int is_complex, int mode, float* out, void* stream) {
  if (!ctx || !a || !b || !out || !a_shape || !b_shape) return NXS_EINVAL;
  int64_t os[3];
  int rc = conv_check(a_shape, b_shape, mode, os);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  return launch_convolve_nd(ctx, a, a_shape, b, b_shape, is_complex, mode, out, pick(ctx, stream));
}

int nxs_convolve_nd_host(nxs_ctx* ctx, const float* a, const int64_t a_shape[3], const float* b,
                         const int64_t b_shape[3], int is_complex, int mode, float* out) {
  if (!ctx || !a || !b || !out || !a_shape || !b_shape) return NXS_EINVAL;
  int64_t os[3];
  int rc = conv_check(a_shape, b_shape, mode, os);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, ctx->stream);
  const size_t es = is_complex ? 8 : 4;
  const size_t na = size_t(a_shape[0]) * a_shape[1] * a_shape[2], nb = size_t(b_shape[0]) * b_shape[1] * b_shape[2];
  const size_t no = size_t(os[0]) * os[1] * os[2];
  return host_roundtrip(ctx, a, na * es, b, nb * es, out, no * es, [&](void* da, void* db, void* dout) {
    return launch_convolve_nd(ctx, (const float*)da, a_shape, (const float*)db, b_shape, is_complex, mode,
                              (float*)dout, ctx->stream);
  });
}

// ---- coefficient broadcast (the path's only collective) -------------------------------------------
int nxs_bcast_coeffs_dev(nxs_ctx* ctx, void* comm, float* buf, int64_t count, int root, void* stream) {
  if (!ctx || !comm || !buf || count < 0) return NXS_EINVAL;
  // NCCL is resolved from whatever libnccl the process already carries (torch's bundled one in
  // this repo's harness), so this library has no link-time NCCL dependency.
  typedef int (*bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static bcast_fn fn = nullptr;
  if (!fn) {
    fn = (bcast_fn)dlsym(RTLD_DEFAULT, "ncclBroadcast");
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (bcast_fn)dlsym(h, "ncclBroadcast");
    }
  }
  if (!fn) {
    ctx->last_error = "ncclBroadcast not found (libnccl.so.2 not loadable)";
    return NXS_ENCCL;
  }
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, pick(ctx, stream));
  const int rc = fn(buf, buf, (size_t)count, /*ncclFloat32*/ 7, root, comm, pick(ctx, stream));
  if (rc != 0) {
    ctx->last_error = "ncclBroadcast failed with code " + std::to_string(rc);
    return NXS_ENCCL;
  }
  return NXS_OK;
}

}  // extern "C"
