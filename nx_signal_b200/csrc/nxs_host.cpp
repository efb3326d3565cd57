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

// T = float: Nx.BinaryBackend's f32 tensors (compute in double, round once per op); T = double: the
// same graphs with `type: :f64` (windows.ex:58,161,226,279,342; filters.ex:153) -- plain double arithmetic
template <class T> inline T rt(double x) { return (T)x; }
template <class T> inline double mul(double a, double b) { return (double)rt<T>(a * b); }
template <class T> inline double dvd(double a, double b) { return (double)rt<T>(a / b); }
template <class T> inline double add(double a, double b) { return (double)rt<T>(a + b); }
template <class T> inline double sub(double a, double b) { return (double)rt<T>(a - b); }
template <class T> inline double lit(double x) { return (double)rt<T>(x); }

const double kPi = 3.14159265358979323846;

// cos-sum windows share: a = (f32(c*pi) * n) / (l - 1)
template <class T> inline double phase(double cpi, int64_t n, int64_t l) { return dvd<T>(mul<T>(lit<T>(cpi), (double)n), (double)(l - 1)); }

template <class T> void hann_like(int64_t n, bool periodic, T* out, int kind) {
  const int64_t l = periodic ? n + 1 : n;
  for (int64_t i = 0; i < n; ++i) {
    const double c = (double)rt<T>(cos(phase<T>(2 * kPi, i, l)));
    if (kind == NXS_WIN_HANN) out[i] = rt<T>(lit<T>(0.5) * sub<T>(1.0, c));                  // windows.ex:298
    else out[i] = rt<T>(lit<T>(0.54) - mul<T>(lit<T>(0.46), c));                                // windows.ex:245
  }
}

template <class T> void blackman(int64_t n, bool periodic, T* out) {  // windows.ex:165-199
  const int64_t l = periodic ? n + 1 : n;
  const int64_t m = (l + 1) / 2;
  std::vector<T> left(m);
  for (int64_t i = 0; i < m; ++i) {
    const double c1 = (double)rt<T>(cos(phase<T>(2 * kPi, i, l)));
    const double c2 = (double)rt<T>(cos(phase<T>(4 * kPi, i, l)));
    left[i] = rt<T>(sub<T>(lit<T>(0.42), mul<T>(lit<T>(0.5), c1)) + mul<T>(lit<T>(0.08), c2));
  }
  std::vector<T> w;
  w.reserve(l + 1);
  for (int64_t i = 0; i < m; ++i) w.push_back(left[i]);
  if (l % 2 == 0) for (int64_t i = m - 1; i >= 0; --i) w.push_back(left[i]);
  else for (int64_t i = m - 2; i >= 0; --i) w.push_back(left[i]);
  for (int64_t i = 0; i < n; ++i) out[i] = w[i];
}

template <class T> void bartlett(int64_t n, T* out) {  // windows.ex:62-76
  const int64_t half = n / 2, left = half + n % 2;
  for (int64_t i = 0; i < left; ++i) out[i] = rt<T>(dvd<T>(mul<T>((double)i, 2.0), (double)n));
  for (int64_t i = 0; i < half; ++i) {
    const double idx = add<T>((double)i, (double)left);
    out[left + i] = rt<T>(2.0 - dvd<T>(mul<T>(idx, 2.0), (double)n));
  }
}

template <class T> void triangular(int64_t n, T* out) {  // windows.ex:103-127
  const int64_t h = (n + 1) / 2;
  std::vector<T> left(h);
  for (int64_t i = 0; i < h; ++i) {
    const double idx = add<T>((double)i, 1.0);
    if (n % 2 == 1) left[i] = rt<T>(dvd<T>(mul<T>(idx, 2.0), (double)(n + 1)));
    else left[i] = rt<T>(dvd<T>(sub<T>(mul<T>(2.0, idx), 1.0), (double)n));
  }
  for (int64_t i = 0; i < h; ++i) out[i] = left[i];
  if (n % 2 == 1) for (int64_t i = 0; i + 1 < h; ++i) out[h + i] = left[h - 2 - i];
  else for (int64_t i = 0; i < h; ++i) out[h + i] = left[h - 1 - i];
}

template <class T> double kaiser_i0(double x) {  // windows.ex:371-386 (x already f32-valued)
  const double ax = (double)rt<T>(fabs(x));
  auto p = [&](int e) { return (double)rt<T>(pow(ax, (double)e)); };
  if ((T)ax < (T)3.75) {
    return add<T>(add<T>(add<T>(add<T>(1.0, dvd<T>(p(2), 4.0)), dvd<T>(p(4), 64.0)), dvd<T>(p(6), 2304.0)), dvd<T>(p(8), 147456.0));
  }
  const double ex = (double)rt<T>(exp(ax));
  const double den = (double)rt<T>(sqrt(mul<T>(lit<float>(2 * kPi), ax)));  // 2 * Nx.Constants.pi(): pi is f32 by default
  const double poly = add<T>(1.0, add<T>(dvd<T>(1.0, mul<T>(8.0, ax)), dvd<T>(9.0, mul<T>(128.0, p(2)))));
  return mul<T>(dvd<T>(ex, den), poly);
}

template <class T> void linspace(double start, double stop, int64_t n, bool endpoint, T* out) {
  // Nx.linspace as f32 tensor ops: step = f32(f32(stop - start) / div); out = f32(f32(i*step) + start)
  // start / stop are numbers (f32 scalars) whatever `type:` is, so the step is an f32 value; only the
  // iota * step + start part runs in the requested type
  const double div = (double)(endpoint ? n - 1 : n);
  const double step = dvd<float>(sub<float>(stop, start), div);
  for (int64_t i = 0; i < n; ++i) out[i] = rt<T>(mul<T>((double)i, step) + start);
}

template <class T> void kaiser(int64_t n, bool periodic, double beta, double eps, T* out) {  // windows.ex:348-369
  const int64_t l = periodic ? n + 1 : n;
  std::vector<T> ratio(l);
  linspace<T>(-1.0, 1.0, l, true, ratio.data());
  const double i0b = kaiser_i0<float>(lit<float>(beta));  // beta is a number: its I0 is an f32 scalar for any `type`
  for (int64_t i = 0; i < n; ++i) {
    const double r2 = (double)rt<T>((double)ratio[i] * (double)ratio[i]);
    double arg = sub<T>(1.0, r2);
    if ((T)arg < (T)lit<T>(eps)) arg = lit<T>(eps);
    const double r = mul<T>(lit<T>(beta), (double)rt<T>(sqrt(arg)));
    out[i] = rt<T>(kaiser_i0<T>(r) / i0b);
  }
}

template <class T> int window_into(int kind, int64_t n, int periodic, double beta, double eps, T* out) {
  if (n < 0 || (!out && n > 0)) return NXS_EINVAL;
  if (n == 0) return NXS_OK;
  switch (kind) {
    case NXS_WIN_RECTANGULAR: for (int64_t i = 0; i < n; ++i) out[i] = (T)1; return NXS_OK;
    case NXS_WIN_BARTLETT: bartlett<T>(n, out); return NXS_OK;
    case NXS_WIN_TRIANGULAR: triangular<T>(n, out); return NXS_OK;
    case NXS_WIN_BLACKMAN: blackman<T>(n, periodic != 0, out); return NXS_OK;
    case NXS_WIN_HAMMING:
    case NXS_WIN_HANN: hann_like<T>(n, periodic != 0, out, kind); return NXS_OK;
    case NXS_WIN_KAISER: kaiser<T>(n, periodic != 0, beta, eps, out); return NXS_OK;
    default: return NXS_EINVAL;
  }
}

template <class T> double sinc32(double t) {  // waveforms.ex:451-457
  const double tp = mul<T>(t, lit<T>(kPi));
  if ((T)tp == (T)0) return 1.0;
  return dvd<T>((double)rt<T>(sin(tp)), tp);
}

template <class T>
int firwin_impl(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta, int pass_zero,
                int scale, double sampling_rate, T* out) {
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
  const double m = lit<T>((double)(num_taps - 1) / 2.0);
  std::vector<double> alpha(num_taps), h(num_taps, 0.0);
  for (int64_t i = 0; i < num_taps; ++i) alpha[i] = sub<T>((double)i, m);
  std::vector<double> freqs;
  freqs.push_back(0.0);
  for (double c : cl) freqs.push_back(c);
  freqs.push_back(1.0);
  for (size_t i = 0; i + 1 < freqs.size(); ++i) {
    const bool take = pass_zero ? (i % 2 == 0) : (i % 2 == 1);
    if (!take) continue;
    const double a = lit<T>(freqs[i]), b = lit<T>(freqs[i + 1]);
    for (int64_t k = 0; k < num_taps; ++k) {
      const double ca = mul<T>(a, sinc32<T>(mul<T>(a, alpha[k])));
      const double cb = mul<T>(b, sinc32<T>(mul<T>(b, alpha[k])));
      h[k] = sub<T>(add<T>(h[k], cb), ca);  // filters.ex:223-227
    }
  }
  std::vector<T> w(num_taps);
  int rc = window_into<T>(window_kind, num_taps, 0, beta, 1.0e-7, w.data());
  if (rc) return rc;
  for (int64_t k = 0; k < num_taps; ++k) h[k] = mul<T>(h[k], (double)w[k]);
  if (scale) {  // filters.ex:229-252
    double sf;
    if (pass_zero) sf = 0.0;
    else if (ncut == 1) sf = 1.0;
    else sf = (cl[0] + cl[1]) / 2.0;
    double dot = 0.0;
    for (int64_t k = 0; k < num_taps; ++k) dot += h[k] * (double)rt<T>(cos(mul<T>(alpha[k], lit<T>(kPi * sf))));
    const double s = (double)rt<T>(fabs(dot));
    for (int64_t k = 0; k < num_taps; ++k) h[k] = dvd<T>(h[k], s);
  }
  for (int64_t k = 0; k < num_taps; ++k) out[k] = (T)h[k];
  return NXS_OK;
}

// lib/nx_signal.ex:154-166: Nx.linspace(0, step * fft_length, n: fft_length, type:, endpoint:)
template <class T>
int fft_frequencies_impl(double sampling_rate, int64_t fft_length, int endpoint, T* out) {
  if (fft_length < 1 || !out) return NXS_EINVAL;
  const double sr = lit<float>(sampling_rate);  // the sampling rate enters the defn as an f32 scalar
  const double step = dvd<float>(sr, (double)fft_length);
  linspace<T>(0.0, mul<float>(step, (double)fft_length), fft_length, endpoint != 0, out);
  return NXS_OK;
}

}  // namespace

extern "C" {

int nxs_window_f32(int kind, int64_t n, int periodic, double beta, double eps, float* out) {
  return window_into<float>(kind, n, periodic, beta, eps, out);
}

int nxs_window_f64(int kind, int64_t n, int periodic, double beta, double eps, double* out) {
  return window_into<double>(kind, n, periodic, beta, eps, out);
}

int nxs_firwin_f32(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta,
                   int pass_zero, int scale, double sampling_rate, float* out) {
  return firwin_impl<float>(num_taps, cutoffs, ncut, window_kind, beta, pass_zero, scale, sampling_rate, out);
}

int nxs_firwin_f64(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta,
                   int pass_zero, int scale, double sampling_rate, double* out) {
  return firwin_impl<double>(num_taps, cutoffs, ncut, window_kind, beta, pass_zero, scale, sampling_rate, out);
}

int nxs_fft_frequencies_f32(double sampling_rate, int64_t fft_length, float* out) {
  return fft_frequencies_impl<float>(sampling_rate, fft_length, 0, out);
}

int nxs_fft_frequencies_ex(double sampling_rate, int64_t fft_length, int endpoint, int is_f64, void* out) {
  return is_f64 ? fft_frequencies_impl<double>(sampling_rate, fft_length, endpoint, (double*)out)
                : fft_frequencies_impl<float>(sampling_rate, fft_length, endpoint, (float*)out);
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
  linspace<float>(0.0, lit<float>(max_mel / f_sp), mel_bins + 2, true, mels.data());  // :412
  const double min_log_hz = 1000.0, min_log_mel = lit<float>(min_log_hz / f_sp);
  const double logstep = dvd<float>((double)rt<float>(log(lit<float>(6.4))), 27.0);  // :419
  for (int64_t i = 0; i < mel_bins + 2; ++i) {
    const double m = (double)mels[i];
    if ((float)m >= (float)min_log_mel) {  // :421-426
      const double e = (double)rt<float>(exp(mul<float>(logstep, sub<float>(m, min_log_mel))));
      mel_f[i] = rt<float>(mul<float>(min_log_hz, e));
    } else {
      mel_f[i] = rt<float>(mul<float>(lit<float>(f_sp), m));
    }
  }
  for (int64_t j = 0; j < mel_bins; ++j) {
    const double fd0 = sub<float>((double)mel_f[j + 1], (double)mel_f[j]);
    const double fd1 = sub<float>((double)mel_f[j + 2], (double)mel_f[j + 1]);
    const double enorm = dvd<float>(2.0, sub<float>((double)mel_f[j + 2], (double)mel_f[j]));  // :436
    for (int64_t k = 0; k < fft_length; ++k) {
      const double r0 = sub<float>((double)mel_f[j], (double)fftfreqs[k]);
      const double r2 = sub<float>((double)mel_f[j + 2], (double)fftfreqs[k]);
      const float lower = rt<float>(dvd<float>(-r0, fd0));  // :431
      const float upper = rt<float>(dvd<float>(r2, fd1));   // :432
      float wgt = lower < upper ? lower : upper;
      if (lower != lower || upper != upper) wgt = NAN;  // Nx.min propagates NaN (0/0 when two edges coincide)
      if (!(wgt > 0.0f) && wgt == wgt) wgt = 0.0f;
      out[j * fft_length + k] = rt<float>(mul<float>((double)wgt, enorm));
    }
  }
  return NXS_OK;
}

int nxs_stft_times_f32(int64_t frame_length, double sampling_rate, int64_t num_frames, float* out) {
  if (num_frames < 0 || (!out && num_frames > 0)) return NXS_EINVAL;
  if (num_frames == 0) return NXS_OK;
  const double sr = lit<float>(sampling_rate);
  const double ts = dvd<float>((double)frame_length, mul<float>(2.0, sr));
  const double last = mul<float>(ts, (double)num_frames);
  linspace<float>(ts, last, num_frames, true, out);
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
