// nxs_host.cpp -- host-side closed forms of the path: window generators, firwin,
// fft_frequencies, frame times, frame counts.  O(n) each, no GPU involved.
//
// Arithmetic follows what Nx.BinaryBackend does for the reference's defn graphs: every
// tensor op is evaluated in double and rounded once to f32, float literals are f32 scalars
// (see DESIGN.md "numerics").  Written from the reference's formulas:
//   lib/nx_signal/windows.ex:33-386, lib/nx_signal/filters.ex:147-279,
//   lib/nx_signal/waveforms.ex:451-457, lib/nx_signal.ex:108-111, 154-166, 289-331.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/nxsignal_b200.h"

namespace {

inline float r32(double x) { return (float)x; }
inline double mul(double a, double b) { return (double)r32(a * b); }
inline double dvd(double a, double b) { return (double)r32(a / b); }
inline double add(double a, double b) { return (double)r32(a + b); }
inline double sub(double a, double b) { return (double)r32(a - b); }
inline double lit(double x) { return (double)r32(x); }

const double kPi = 3.14159265358979323846;

// cos-sum windows share: a = (f32(c*pi) * n) / (l - 1)
inline double phase(double cpi, int64_t n, int64_t l) { return dvd(mul(lit(cpi), (double)n), (double)(l - 1)); }

void hann_like(int64_t n, bool periodic, float* out, int kind) {
  const int64_t l = periodic ? n + 1 : n;
  for (int64_t i = 0; i < n; ++i) {
    const double c = (double)r32(cos(phase(2 * kPi, i, l)));
    if (kind == NXS_WIN_HANN) out[i] = r32(lit(0.5) * sub(1.0, c));                  // windows.ex:298
    else out[i] = r32(lit(0.54) - mul(lit(0.46), c));                                // windows.ex:245
  }
}

void blackman(int64_t n, bool periodic, float* out) {  // windows.ex:165-199
  const int64_t l = periodic ? n + 1 : n;
  const int64_t m = (l + 1) / 2;
  std::vector<float> left(m);
  for (int64_t i = 0; i < m; ++i) {
    const double c1 = (double)r32(cos(phase(2 * kPi, i, l)));
    const double c2 = (double)r32(cos(phase(4 * kPi, i, l)));
    left[i] = r32(sub(lit(0.42), mul(lit(0.5), c1)) + mul(lit(0.08), c2));
  }
  std::vector<float> w;
  w.reserve(l + 1);
  for (int64_t i = 0; i < m; ++i) w.push_back(left[i]);
  if (l % 2 == 0) for (int64_t i = m - 1; i >= 0; --i) w.push_back(left[i]);
  else for (int64_t i = m - 2; i >= 0; --i) w.push_back(left[i]);
  for (int64_t i = 0; i < n; ++i) out[i] = w[i];
}

void bartlett(int64_t n, float* out) {  // windows.ex:62-76
  const int64_t half = n / 2, left = half + n % 2;
  for (int64_t i = 0; i < left; ++i) out[i] = r32(dvd(mul((double)i, 2.0), (double)n));
  for (int64_t i = 0; i < half; ++i) {
    const double idx = add((double)i, (double)left);
    out[left + i] = r32(2.0 - dvd(mul(idx, 2.0), (double)n));
  }
}

void triangular(int64_t n, float* out) {  // windows.ex:103-127
  const int64_t h = (n + 1) / 2;
  std::vector<float> left(h);
  for (int64_t i = 0; i < h; ++i) {
    const double idx = add((double)i, 1.0);
    if (n % 2 == 1) left[i] = r32(dvd(mul(idx, 2.0), (double)(n + 1)));
    else left[i] = r32(dvd(sub(mul(2.0, idx), 1.0), (double)n));
  }
  for (int64_t i = 0; i < h; ++i) out[i] = left[i];
  if (n % 2 == 1) for (int64_t i = 0; i + 1 < h; ++i) out[h + i] = left[h - 2 - i];
  else for (int64_t i = 0; i < h; ++i) out[h + i] = left[h - 1 - i];
}

double kaiser_i0(double x) {  // windows.ex:371-386 (x already f32-valued)
  const double ax = (double)r32(fabs(x));
  auto p = [&](int e) { return (double)r32(pow(ax, (double)e)); };
  if ((float)ax < 3.75f) {
    return add(add(add(add(1.0, dvd(p(2), 4.0)), dvd(p(4), 64.0)), dvd(p(6), 2304.0)), dvd(p(8), 147456.0));
  }
  const double ex = (double)r32(exp(ax));
  const double den = (double)r32(sqrt(mul(lit(2 * kPi), ax)));
  const double poly = add(1.0, add(dvd(1.0, mul(8.0, ax)), dvd(9.0, mul(128.0, p(2)))));
  return mul(dvd(ex, den), poly);
}

void linspace(double start, double stop, int64_t n, bool endpoint, float* out) {
  // Nx.linspace as f32 tensor ops: step = f32(f32(stop - start) / div); out = f32(f32(i*step) + start)
  const double div = (double)(endpoint ? n - 1 : n);
  const double step = dvd(sub(stop, start), div);
  for (int64_t i = 0; i < n; ++i) out[i] = r32(mul((double)i, step) + start);
}

void kaiser(int64_t n, bool periodic, double beta, double eps, float* out) {  // windows.ex:348-369
  const int64_t l = periodic ? n + 1 : n;
  std::vector<float> ratio(l);
  linspace(-1.0, 1.0, l, true, ratio.data());
  const double i0b = kaiser_i0(lit(beta));
  for (int64_t i = 0; i < n; ++i) {
    const double r2 = (double)r32((double)ratio[i] * (double)ratio[i]);
    double arg = sub(1.0, r2);
    if ((float)arg < (float)lit(eps)) arg = lit(eps);
    const double r = mul(lit(beta), (double)r32(sqrt(arg)));
    out[i] = r32(kaiser_i0(r) / i0b);
  }
}

int window_into(int kind, int64_t n, int periodic, double beta, double eps, float* out) {
  if (n < 0 || (!out && n > 0)) return NXS_EINVAL;
  if (n == 0) return NXS_OK;
  switch (kind) {
    case NXS_WIN_RECTANGULAR: for (int64_t i = 0; i < n; ++i) out[i] = 1.0f; return NXS_OK;
    case NXS_WIN_BARTLETT: bartlett(n, out); return NXS_OK;
    case NXS_WIN_TRIANGULAR: triangular(n, out); return NXS_OK;
    case NXS_WIN_BLACKMAN: blackman(n, periodic != 0, out); return NXS_OK;
    case NXS_WIN_HAMMING:
    case NXS_WIN_HANN: hann_like(n, periodic != 0, out, kind); return NXS_OK;
    case NXS_WIN_KAISER: kaiser(n, periodic != 0, beta, eps, out); return NXS_OK;
    default: return NXS_EINVAL;
  }
}

double sinc32(double t) {  // waveforms.ex:451-457
  const double tp = mul(t, lit(kPi));
  if ((float)tp == 0.0f) return 1.0;
  return dvd((double)r32(sin(tp)), tp);
}

}  // namespace

extern "C" {

int nxs_window_f32(int kind, int64_t n, int periodic, double beta, double eps, float* out) {
  return window_into(kind, n, periodic, beta, eps, out);
}

int nxs_firwin_f32(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta,
                   int pass_zero, int scale, double sampling_rate, float* out) {
  if (num_taps < 1 || ncut < 1 || !cutoffs || !out) return NXS_EINVAL;
  const double nyq = sampling_rate / 2.0;
  std::vector<double> cl(cutoffs, cutoffs + ncut);
  for (auto& c : cl) c /= nyq;
  std::sort(cl.begin(), cl.end());
  if (cl.front() <= 0.0 || cl.back() >= 1.0) return NXS_EINVAL;  // filters.ex:170-178
  const bool even_cuts = (ncut % 2) == 0;
  const bool nyq_gain = (pass_zero && even_cuts) || (!pass_zero && !even_cuts);
  if (nyq_gain && num_taps % 2 == 0) return NXS_EINVAL;  // filters.ex:189-193
  switch (window_kind) {
    case NXS_WIN_HAMMING: case NXS_WIN_HANN: case NXS_WIN_BLACKMAN: case NXS_WIN_BARTLETT:
    case NXS_WIN_RECTANGULAR: case NXS_WIN_KAISER: break;
    default: return NXS_EINVAL;  // filters.ex:274-277
  }
  const double m = lit((double)(num_taps - 1) / 2.0);
  std::vector<double> alpha(num_taps), h(num_taps, 0.0);
  for (int64_t i = 0; i < num_taps; ++i) alpha[i] = sub((double)i, m);
  std::vector<double> freqs;
  freqs.push_back(0.0);
  for (double c : cl) freqs.push_back(c);
  freqs.push_back(1.0);
  for (size_t i = 0; i + 1 < freqs.size(); ++i) {
    const bool take = pass_zero ? (i % 2 == 0) : (i % 2 == 1);
    if (!take) continue;
    const double a = lit(freqs[i]), b = lit(freqs[i + 1]);
    for (int64_t k = 0; k < num_taps; ++k) {
      const double ca = mul(a, sinc32(mul(a, alpha[k])));
      const double cb = mul(b, sinc32(mul(b, alpha[k])));
      h[k] = sub(add(h[k], cb), ca);  // filters.ex:223-227
    }
  }
  std::vector<float> w(num_taps);
  int rc = window_into(window_kind, num_taps, 0, beta, 1.0e-7, w.data());
  if (rc) return rc;
  for (int64_t k = 0; k < num_taps; ++k) h[k] = mul(h[k], (double)w[k]);
  if (scale) {  // filters.ex:229-252
    double sf;
    if (pass_zero) sf = 0.0;
    else if (ncut == 1) sf = 1.0;
    else sf = (cl[0] + cl[1]) / 2.0;
    double dot = 0.0;
    for (int64_t k = 0; k < num_taps; ++k) dot += h[k] * (double)r32(cos(mul(alpha[k], lit(kPi * sf))));
    const double s = (double)r32(fabs(dot));
    for (int64_t k = 0; k < num_taps; ++k) h[k] = dvd(h[k], s);
  }
  for (int64_t k = 0; k < num_taps; ++k) out[k] = (float)h[k];
  return NXS_OK;
}

int nxs_fft_frequencies_f32(double sampling_rate, int64_t fft_length, float* out) {
  if (fft_length < 1 || !out) return NXS_EINVAL;
  const double sr = lit(sampling_rate);
  const double step = dvd(sr, (double)fft_length);
  linspace(0.0, mul(step, (double)fft_length), fft_length, false, out);
  return NXS_OK;
}

// NxSignal.mel_filters/4 (lib/nx_signal.ex:397-445): Slaney-style triangular filters,
// out [mel_bins][fft_length], every op rounded to f32 as the reference's defn graph does.
int nxs_mel_filters_f32(int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel,
                        double mel_frequency_spacing, float* out) {
  if (fft_length < 1 || mel_bins < 1 || !out || !(mel_frequency_spacing > 0)) return NXS_EINVAL;
  const double f_sp = mel_frequency_spacing;
  std::vector<float> fftfreqs(fft_length), mels(mel_bins + 2), mel_f(mel_bins + 2);
  int rc = nxs_fft_frequencies_f32(sampling_rate, fft_length, fftfreqs.data());
  if (rc) return rc;
  linspace(0.0, lit(max_mel / f_sp), mel_bins + 2, true, mels.data());  // :412
  const double min_log_hz = 1000.0, min_log_mel = lit(min_log_hz / f_sp);
  const double logstep = dvd((double)r32(log(lit(6.4))), 27.0);  // :419
  for (int64_t i = 0; i < mel_bins + 2; ++i) {
    const double m = (double)mels[i];
    if ((float)m >= (float)min_log_mel) {  // :421-426
      const double e = (double)r32(exp(mul(logstep, sub(m, min_log_mel))));
      mel_f[i] = r32(mul(min_log_hz, e));
    } else {
      mel_f[i] = r32(mul(lit(f_sp), m));
    }
  }
  for (int64_t j = 0; j < mel_bins; ++j) {
    const double fd0 = sub((double)mel_f[j + 1], (double)mel_f[j]);
    const double fd1 = sub((double)mel_f[j + 2], (double)mel_f[j + 1]);
    const double enorm = dvd(2.0, sub((double)mel_f[j + 2], (double)mel_f[j]));  // :436
    for (int64_t k = 0; k < fft_length; ++k) {
      const double r0 = sub((double)mel_f[j], (double)fftfreqs[k]);
      const double r2 = sub((double)mel_f[j + 2], (double)fftfreqs[k]);
      const float lower = r32(dvd(-r0, fd0));  // :431
      const float upper = r32(dvd(r2, fd1));   // :432
      float wgt = lower < upper ? lower : upper;
      if (lower != lower || upper != upper) wgt = NAN;  // Nx.min propagates NaN (0/0 when two edges coincide)
      if (!(wgt > 0.0f) && wgt == wgt) wgt = 0.0f;
      out[j * fft_length + k] = r32(mul((double)wgt, enorm));
    }
  }
  return NXS_OK;
}

int nxs_stft_times_f32(int64_t frame_length, double sampling_rate, int64_t num_frames, float* out) {
  if (num_frames < 0 || (!out && num_frames > 0)) return NXS_EINVAL;
  if (num_frames == 0) return NXS_OK;
  const double sr = lit(sampling_rate);
  const double ts = dvd((double)frame_length, mul(2.0, sr));
  const double last = mul(ts, (double)num_frames);
  linspace(ts, last, num_frames, true, out);
  return NXS_OK;
}

int nxs_num_frames(int64_t length, int64_t window_length, int64_t stride, int pad_mode, int64_t pad_lo,
                   int64_t pad_hi, int64_t* num_frames) {
  if (!num_frames || window_length < 1 || stride < 1 || length < 0) return NXS_EINVAL;
  int64_t lo = 0, hi = 0;
  switch (pad_mode) {
    case NXS_PAD_VALID: break;
    case NXS_PAD_SAME: {
      int64_t total = window_length - 1;
      if (total < 0) total = 0;
      lo = total / 2;
      hi = total - lo;
      break;
    }
    case NXS_PAD_REFLECT: lo = hi = window_length / 2; break;
    case NXS_PAD_EXPLICIT: lo = pad_lo; hi = pad_hi; break;
    default: return NXS_EINVAL;
  }
  const int64_t padded = length + lo + hi;
  *num_frames = padded < window_length ? 0 : (padded - window_length) / stride + 1;
  return NXS_OK;
}

int nxs_fir_out_len(int64_t length, int64_t num_taps, int mode, int64_t* out_len) {
  if (!out_len || length < 1 || num_taps < 1) return NXS_EINVAL;
  switch (mode) {
    case NXS_MODE_FULL: *out_len = length + num_taps - 1; return NXS_OK;
    case NXS_MODE_SAME: *out_len = length; return NXS_OK;
    case NXS_MODE_VALID: *out_len = std::max(length, num_taps) - std::min(length, num_taps) + 1; return NXS_OK;
    default: return NXS_EINVAL;
  }
}

}  // extern "C"
