/*
 * nxs_oracle.c -- C port of the oracle's STFT (oracle/nxsignal_oracle.py :: stft / nx_fft).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-checked against the numpy oracle) and by
 * bench.py's cpu_baseline / --impl reference legs as the timed CPU stand-in for the
 * reference.  Never linked into or called from nx_signal_b200/.
 *
 * It follows the same algorithm the reference executes on Nx.BinaryBackend (nx 0.11.0,
 * /root/reference/mix.lock:10; not vendored, restated): frames = as_windowed(x)
 * (lib/nx_signal.ex:249-364), f32 multiply by the window (:101), then Nx.fft (:102) as a
 * recursive radix-2 decimation in time in double with twiddles exp(-i*(2*pi/n)*k), a naive
 * DFT for odd n, |re|,|im| <= 1e-10 snapped to zero and one rounding to c64; optional
 * scaling (:113-127).  Twiddles are tabulated per level (same values as evaluating them per
 * butterfly).  Frames are distributed over OpenMP threads.
 *
 * "parity pinned": the numpy oracle this port mirrors reproduces the reference's doctest
 * vectors bit-exactly (tests/test_oracle_golden.py); tests/test_oracle_c.py checks this
 * port against it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

typedef struct { double re, im; } cplx;

typedef struct {
  int levels;
  int n[40];
  cplx* tw[40]; /* tw[l][k] = exp(sign*i*(2*pi/n[l])*k), k < n[l]/2, for even n[l] > 2 */
  cplx* dft;    /* naive DFT matrix for the odd leaf size (m*m), or NULL */
  int leaf;
} plan_t;

static void plan_init(plan_t* p, int n, double sign) {
  p->levels = 0;
  p->dft = NULL;
  p->leaf = 1;
  int m = n;
  while (m > 2 && m % 2 == 0) {
    const int l = p->levels++;
    p->n[l] = m;
    p->tw[l] = (cplx*)malloc(sizeof(cplx) * (size_t)(m / 2));
    const double t = sign * 2.0 * PI / m;
    for (int k = 0; k < m / 2; ++k) {
      const double ang = t * k;
      p->tw[l][k].re = cos(ang);
      p->tw[l][k].im = sin(ang);
    }
    m /= 2;
  }
  p->leaf = m;
  if (m > 2) { /* odd leaf: naive DFT */
    p->dft = (cplx*)malloc(sizeof(cplx) * (size_t)m * m);
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < m; ++k) {
        const double ang = sign * 2.0 * PI * ((double)j * (double)k) / m;
        p->dft[(size_t)j * m + k].re = cos(ang);
        p->dft[(size_t)j * m + k].im = sin(ang);
      }
  }
}

static void plan_free(plan_t* p) {
  for (int l = 0; l < p->levels; ++l) free(p->tw[l]);
  free(p->dft);
}

/* out[0..n) = FFT of in[0], in[stride], ...; scratch holds >= n elements per level */
static void fft_rec(const plan_t* p, int level, const cplx* in, int stride, int n, cplx* out, cplx* scratch) {
  if (n == 1) { out[0] = in[0]; return; }
  if (n == 2) {
    const cplx a = in[0], b = in[stride];
    out[0].re = a.re + b.re; out[0].im = a.im + b.im;
    out[1].re = a.re - b.re; out[1].im = a.im - b.im;
    return;
  }
  if (n % 2 == 1) {
    for (int k = 0; k < n; ++k) { out[k].re = 0.0; out[k].im = 0.0; }
    for (int j = 0; j < n; ++j) {
      const cplx x = in[(size_t)j * stride];
      const cplx* row = p->dft + (size_t)j * n;
      for (int k = 0; k < n; ++k) {
        out[k].re += x.re * row[k].re - x.im * row[k].im;
        out[k].im += x.re * row[k].im + x.im * row[k].re;
      }
    }
    return;
  }
  const int h = n / 2;
  cplx* even = scratch;
  cplx* odd = scratch + h;
  fft_rec(p, level + 1, in, stride * 2, h, even, scratch + n);
  fft_rec(p, level + 1, in + stride, stride * 2, h, odd, scratch + n);
  const cplx* tw = p->tw[level];
  for (int k = 0; k < h; ++k) {
    const double br = tw[k].re * odd[k].re - tw[k].im * odd[k].im;
    const double bi = tw[k].re * odd[k].im + tw[k].im * odd[k].re;
    out[k].re = even[k].re + br; out[k].im = even[k].im + bi;
    out[k + h].re = even[k].re - br; out[k + h].im = even[k].im - bi;
  }
}

static int64_t reflect_index(int64_t i, int64_t L) {
  if (L <= 1) return 0;
  const int64_t per = 2 * (L - 1);
  int64_t r = i % per;
  if (r < 0) r += per;
  return r < L ? r : per - r;
}

int nxs_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* z[c][m][nfft] (interleaved re,im f32).  pad_reflect: 1 = reflect, 0 = zeros; pad_lo = samples in front.
 * scaling: 0 none, 1 spectrum, 2 psd.  threads <= 0: all. Returns 0. */
int nxs_oracle_stft_f32(const float* x, int64_t channels, int64_t length, int64_t x_ld, const float* window,
                        int64_t frame_length, int64_t hop, int64_t nfft, int64_t pad_lo, int pad_reflect,
                        int64_t num_frames, int scaling, double sampling_rate, float* z, int threads) {
  plan_t plan;
  plan_init(&plan, (int)nfft, -1.0);
  double S = 1.0;
  if (scaling == 1) {
    double acc = 0.0;
    for (int64_t i = 0; i < frame_length; ++i) acc += (double)window[i];
    S = (double)(float)acc;
  } else if (scaling == 2) {
    double acc = 0.0;
    for (int64_t i = 0; i < frame_length; ++i) acc += (double)(float)((double)window[i] * (double)window[i]);
    S = (double)(float)sqrt((double)(float)((double)(float)sampling_rate * (double)(float)acc));
  }
  const int64_t nload = frame_length < nfft ? frame_length : nfft;
  const int64_t total = channels * num_frames;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    cplx* in = (cplx*)malloc(sizeof(cplx) * (size_t)nfft);
    cplx* out = (cplx*)malloc(sizeof(cplx) * (size_t)nfft);
    cplx* scratch = (cplx*)malloc(sizeof(cplx) * (size_t)nfft * 2 + 64);
#pragma omp for schedule(static)
    for (int64_t f = 0; f < total; ++f) {
      const int64_t c = f / num_frames, m = f % num_frames;
      const float* xr = x + c * x_ld;
      const int64_t src0 = m * hop - pad_lo;
      for (int64_t s = 0; s < nfft; ++s) {
        double v = 0.0;
        if (s < nload) {
          const int64_t src = src0 + s;
          float xv = 0.f;
          if (src >= 0 && src < length) xv = xr[src];
          else if (pad_reflect) xv = xr[reflect_index(src, length)];
          v = (double)(float)((double)xv * (double)window[s]); /* Nx.multiply -> f32 */
        }
        in[s].re = v; in[s].im = 0.0;
      }
      fft_rec(&plan, 0, in, 1, (int)nfft, out, scratch);
      float* zf = z + (size_t)f * (size_t)nfft * 2;
      for (int64_t k = 0; k < nfft; ++k) {
        double re = out[k].re, im = out[k].im;
        if (fabs(re) <= 1.0e-10) re = 0.0;
        if (fabs(im) <= 1.0e-10) im = 0.0;
        if (scaling != 0) { /* spectrum (c64) / S: round to c64 first, then divide in double */
          re = (double)(float)re / S; im = (double)(float)im / S;
        }
        zf[2 * k] = (float)re; zf[2 * k + 1] = (float)im;
      }
    }
    free(in); free(out); free(scratch);
  }
  plan_free(&plan);
  return 0;
}
