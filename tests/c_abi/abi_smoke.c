/*
 * abi_smoke.c -- the C ABI used from plain C, the way a NIF uses it (no Python, no torch): malloc'ed
 * (pageable) buffers, nxs_ctx_create, nxs_window_f32, nxs_num_frames, nxs_stft_f32_host, nxs_istft_c64_host,
 * nxs_fir_f32_host, checked against naive double-precision sums computed here.
 *
 * TEST INFRASTRUCTURE (tests/test_c_abi.py builds and runs it).  Exit codes: 0 = all checks passed,
 * 3 = no CUDA device (nxs_ctx_create returned NXS_ENODEVICE: the library has no CPU path), 1 = a check failed.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "nxsignal_b200.h"

#define PI 3.14159265358979323846
#define CHECK(rc, what)                                                                 \
  do {                                                                                  \
    if ((rc) != NXS_OK) {                                                               \
      fprintf(stderr, "%s failed: %s (%s)\n", what, nxs_strerror(rc), nxs_last_error(ctx)); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

int main(void) {
  printf("abi %d, %s\n", nxs_abi_version(), nxs_build_info());
  nxs_ctx* ctx = NULL;
  int rc = nxs_ctx_create(0, &ctx);
  if (rc == NXS_ENODEVICE) {
    printf("no CUDA device: %s\n", nxs_strerror(rc));
    return 3;
  }
  CHECK(rc, "nxs_ctx_create");

  /* ---- stft: 3 channels x 5000 samples, hann(256), hop 64, nfft 256 ---- */
  enum { C = 3, L = 5000, N = 256, H = 64 };
  float* x = (float*)malloc(sizeof(float) * C * L);
  float w[N];
  unsigned s = 12345u;
  for (int i = 0; i < C * L; ++i) {
    s = s * 1664525u + 1013904223u;
    x[i] = (float)((double)(s >> 8) / 16777216.0 - 0.5) + 0.5f * (float)sin(2 * PI * 440.0 * (i % L) / 48000.0);
  }
  CHECK(nxs_window_f32(NXS_WIN_HANN, N, 1, 0.0, 0.0, w), "nxs_window_f32");
  int64_t M = 0;
  CHECK(nxs_num_frames(L, N, H, NXS_PAD_VALID, 0, 0, &M), "nxs_num_frames");
  if (M != (L - N) / H + 1) return 1;
  float* z = (float*)malloc(sizeof(float) * 2 * C * M * N);
  CHECK(nxs_stft_f32_host(ctx, x, C, L, L, w, N, H, N, NXS_PAD_VALID, 0, 0, NXS_SCALE_NONE, 48000.0, z), "nxs_stft_f32_host");
  double worst = 0.0;
  for (int c = 0; c < C; ++c)
    for (int64_t m = 0; m < M; m += 7) { /* every 7th frame against a naive f64 DFT of the f32-rounded windowed frame */
      double scale = 0.0, err = 0.0;
      for (int k = 0; k < N; ++k) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < N; ++n) {
          const double v = (double)(float)(x[c * L + m * H + n] * w[n]);
          re += v * cos(2 * PI * k * n / N);
          im -= v * sin(2 * PI * k * n / N);
        }
        const float* g = z + 2 * ((c * M + m) * N + k);
        const double d = hypot(g[0] - re, g[1] - im), a = hypot(re, im);
        if (d > err) err = d;
        if (a > scale) scale = a;
      }
      if (err / scale > worst) worst = err / scale;
    }
  printf("stft: %lld frames per channel, worst frame-relative error %.2e\n", (long long)M, worst);
  if (!(worst <= 1e-5)) return 1;

  /* ---- istft(stft(x)) reproduces x away from the edges ---- */
  const int64_t out_len = M * H + N - H;
  float* y = (float*)malloc(sizeof(float) * 2 * C * out_len);
  CHECK(nxs_istft_c64_host(ctx, z, C, M, N, w, N, H, N, NXS_SCALE_NONE, 48000.0, y), "nxs_istft_c64_host");
  worst = 0.0;
  for (int c = 0; c < C; ++c)
    for (int64_t i = N; i < out_len - N; ++i) {
      const double d = fabs((double)y[2 * (c * out_len + i)] - (double)x[c * L + i]);
      if (d > worst) worst = d;
    }
  printf("istft round trip: max |y - x| = %.2e\n", worst);
  if (!(worst <= 2e-5)) return 1;

  /* ---- FIR: 33-tap moving average, mode :same, against a direct f64 sum ---- */
  enum { K = 33 };
  float taps[K], *f = (float*)malloc(sizeof(float) * C * L);
  for (int k = 0; k < K; ++k) taps[k] = 1.0f / K;
  CHECK(nxs_fir_f32_host(ctx, x, C, L, L, taps, K, NXS_MODE_SAME, f, L), "nxs_fir_f32_host");
  worst = 0.0;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < L; i += 3) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        const int j = i + (K - 1) / 2 - k;
        if (j >= 0 && j < L) acc += (double)taps[k] * (double)x[c * L + j];
      }
      const double d = fabs((double)f[c * L + i] - acc);
      if (d > worst) worst = d;
    }
  printf("fir: max |y - direct| = %.2e\n", worst);
  if (!(worst <= 1e-5)) return 1;

  /* ---- error convention: bad arguments come back as codes, nothing aborts ---- */
  if (nxs_stft_f32_host(ctx, x, C, L, L, w, N, 0, N, NXS_PAD_VALID, 0, 0, NXS_SCALE_NONE, 48000.0, z) != NXS_EINVAL) return 1;
  if (nxs_istft_c64_host(ctx, z, C, M, N, w, N, H, 2 * N, NXS_SCALE_NONE, 48000.0, y) != NXS_ESHAPE) return 1;
  free(x); free(z); free(y); free(f);
  CHECK(nxs_ctx_destroy(ctx), "nxs_ctx_destroy");
  printf("c abi smoke ok\n");
  return 0;
}
