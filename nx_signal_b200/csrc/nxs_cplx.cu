// nxs_cplx.cu -- NxSignal.stft/3 on COMPLEX data.
//
// The reference's graph takes any numeric tensor: as_windowed -> Nx.multiply(window) -> Nx.fft
// (lib/nx_signal.ex:94-102) is a complex transform when `data` is c64.  The window is real, so the
// STFT is linear over the real field: stft(xr + i xi) = stft(xr) + i stft(xi).  The signal is split
// into its two real planes, both go through the real-input kernels as 2 C channels, and one pass
// combines the two spectra, z = (zr.re - zi.im, zr.im + zi.re).  A rare path (complex baseband
// signals): it costs one extra read + write of the spectrum; channel chunks bound the work buffer.
#include "nxs_common.cuh"
#include "nxs_hostio.cuh"

namespace nxs {

__global__ void __launch_bounds__(256) split_planes_kernel(const float2* __restrict__ x, int64_t channels,
                                                          int64_t length, int64_t x_ld, int64_t p_ld,
                                                          float* __restrict__ planes) {
  const int64_t total = channels * length;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / length, s = i - c * length;
    const float2 v = x[c * x_ld + s];
    planes[c * p_ld + s] = v.x;
    planes[(channels + c) * p_ld + s] = v.y;
  }
}

__global__ void __launch_bounds__(256) combine_spectra_kernel(const float2* __restrict__ zr,
                                                             const float2* __restrict__ zi, int64_t n,
                                                             float2* __restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 a = zr[i], b = zi[i];
    z[i] = make_float2(a.x - b.y, a.y + b.x);
  }
}

int launch_stft_c64(nxs_ctx* ctx, const float2* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, const PadGeom& g,
                    int64_t num_frames, int scaling, double sampling_rate, float2* z, cudaStream_t st) {
  if (channels <= 0 || num_frames <= 0) return NXS_OK;
  const int64_t p_ld = (length + 3) / 4 * 4;  // 16-byte aligned rows: the TMA-staged kernels serve the planes
  const size_t per_ch = 2 * (size_t(p_ld) * sizeof(float) + size_t(num_frames) * fft_length * sizeof(float2));
  int64_t cc = int64_t((size_t(1) << 30) / (per_ch ? per_ch : 1));
  if (cc < 1) cc = 1;
  if (cc > channels) cc = channels;
  int rc = grow_buf(ctx, &ctx->d_work, &ctx->d_work_bytes, size_t(cc) * per_ch + 512, false);
  if (rc) return rc;
  float* planes = (float*)ctx->d_work;
  float2* z2 = (float2*)((char*)ctx->d_work + (size_t(2 * cc) * p_ld * sizeof(float) + 255) / 256 * 256);
  const int64_t grid_cap = int64_t(ctx->sm_count) * 16;
  for (int64_t c0 = 0; c0 < channels; c0 += cc) {
    const int64_t n = channels - c0 < cc ? channels - c0 : cc;
    int64_t grid = (n * length + 255) / 256;
    split_planes_kernel<<<(unsigned)(grid < grid_cap ? grid : grid_cap), 256, 0, st>>>(x + c0 * x_ld, n, length, x_ld, p_ld,
                                                                                    planes);
    ctx->launches++;
    NXS_CUDA(ctx, cudaGetLastError());
    rc = launch_stft(ctx, planes, 2 * n, length, p_ld, window, frame_length, hop, fft_length, g, num_frames, scaling,
                     sampling_rate, z2, fft_length, 0, st);
    if (rc) return rc;
    const int64_t cnt = n * num_frames * fft_length;
    grid = (cnt + 255) / 256;
    combine_spectra_kernel<<<(unsigned)(grid < grid_cap ? grid : grid_cap), 256, 0, st>>>(z2, z2 + cnt, cnt,
                                                                                       z + c0 * num_frames * fft_length);
    ctx->launches++;
    NXS_CUDA(ctx, cudaGetLastError());
  }
  return NXS_OK;
}

}  // namespace nxs

using namespace nxs;

extern "C" {

int nxs_stft_c64_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int pad_mode,
                     int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate, float* z, void* stream) {
  if (!ctx || !x || !window || !z) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling,
                      sampling_rate, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  StreamOrder stream_order(ctx, (cudaStream_t)stream);
  return launch_stft_c64(ctx, reinterpret_cast<const float2*>(x), channels, length, x_ld, window, frame_length, hop,
                         fft_length, g, M, scaling, sampling_rate, reinterpret_cast<float2*>(z), (cudaStream_t)stream);
}

int nxs_stft_c64_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                      const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int pad_mode,
                      int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate, float* z) {
  if (!ctx || !x || !window || !z) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling,
                      sampling_rate, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (channels == 0 || M == 0) return NXS_OK;
  StreamOrder stream_order(ctx, ctx->stream);
  PipeSpec ps;
  ps.in = x;
  ps.in_row_bytes = size_t(length) * sizeof(float2);
  ps.in_pitch = size_t(x_ld) * sizeof(float2);
  ps.out = z;
  ps.out_row_bytes = ps.out_pitch = size_t(M) * fft_length * sizeof(float2);
  ps.rows = channels;
  ps.aux = window;
  ps.aux_bytes = size_t(frame_length) * sizeof(float);
  return host_pipeline(ctx, ps, [&](int64_t, int64_t n, void* dx, void* dw, void* dz) {
    return launch_stft_c64(ctx, (const float2*)dx, n, length, x_ld, (const float*)dw, frame_length, hop, fft_length, g,
                           M, scaling, sampling_rate, (float2*)dz, ctx->stream);
  });
}

}  // extern "C"
