/*
 * nxsignal_b200.h -- C ABI of the B200-native backend for NxSignal's
 * STFT / ISTFT / windows / FIR hot path.
 *
 * Every entry point below replaces one public function head of the reference
 * (elixir-nx/nx_signal v0.3.0 @ dcf5b81, 100 % Elixir: it has no FFI of its
 * own, so the seam is the function heads themselves -- SURVEY.md 8b).  The
 * reference interface each entry replaces is cited as file:line relative to
 * /root/reference.  A NIF or any other FFI binds these
 * symbols 1:1; signatures use only plain pointers and sizes.
 *
 * Conventions
 *   - return value: 0 = NXS_OK, negative = error (nxs_strerror()).  No
 *     exceptions, no abort, no stdout/stderr output.
 *   - layouts are row-major, native-endian, exactly what Nx.to_binary/1
 *     yields: f32 = float, c64 = interleaved (re, im) float pairs.
 *   - "_dev" entries take DEVICE pointers, enqueue on `stream` (a cudaStream_t
 *     passed as void*, NULL = CUDA's default stream) and return without
 *     synchronising.  "_host" entries take HOST pointers, stage through pinned
 *     buffers owned by the context, and return when the result is in `out`.
 *   - the caller owns every input/output buffer; the library never keeps a
 *     caller pointer past return (_host) / past stream completion (_dev).
 *   - one nxs_ctx is single-threaded; distinct contexts may be used from
 *     distinct threads concurrently.
 *   - a context may be used from several CUDA streams in turn: its internal
 *     device state (prepared window, scratch) is ordered behind the previous
 *     call with an event, so a call on stream B issued after a call on stream A
 *     starts after A's call on the device (the host never blocks).  Two calls
 *     on one context therefore never overlap on the device; use one context
 *     per stream for concurrency.
 */
#ifndef NXSIGNAL_B200_H
#define NXSIGNAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NXS_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------- */
enum {
  NXS_OK = 0,
  NXS_EINVAL = -1,       /* bad option value   -> ArgumentError in the shim  */
  NXS_ESHAPE = -2,       /* incompatible shapes -> ArgumentError in the shim */
  NXS_EUNSUPPORTED = -3, /* valid in the reference, not built here           */
  NXS_ECUDA = -4,        /* CUDA runtime failure (nxs_last_error)            */
  NXS_ENCCL = -5,
  NXS_ENOMEM = -6,
  NXS_ENODEVICE = -7     /* no CUDA device: the product path has no CPU fallback */
};

/* ---- option enums (values of the reference's keyword options) ---------- */
/* window_padding / padding: lib/nx_signal.ex:76, 175-178, 303-331 */
enum { NXS_PAD_VALID = 0, NXS_PAD_SAME = 1, NXS_PAD_REFLECT = 2, NXS_PAD_EXPLICIT = 3 };
/* :scaling: lib/nx_signal.ex:113-127, 611-625 */
enum { NXS_SCALE_NONE = 0, NXS_SCALE_SPECTRUM = 1, NXS_SCALE_PSD = 2 };
/* NxSignal.Windows.*: lib/nx_signal/windows.ex:33,57,98,160,225,278,341 */
enum {
  NXS_WIN_RECTANGULAR = 0, NXS_WIN_BARTLETT = 1, NXS_WIN_TRIANGULAR = 2, NXS_WIN_BLACKMAN = 3,
  NXS_WIN_HAMMING = 4, NXS_WIN_HANN = 5, NXS_WIN_KAISER = 6
};
/* :mode of convolve/correlate/fftconvolve: lib/nx_signal/convolution.ex:39-44 */
enum { NXS_MODE_FULL = 0, NXS_MODE_SAME = 1, NXS_MODE_VALID = 2 };
/* PeakFinding comparator: &Nx.less/2 (argrelmin), &Nx.greater/2 (argrelmax), and the two non-strict forms */
enum { NXS_CMP_LESS = 0, NXS_CMP_GREATER = 1, NXS_CMP_LESS_EQUAL = 2, NXS_CMP_GREATER_EQUAL = 3 };

typedef struct nxs_ctx nxs_ctx;

/* ---- library / context -------------------------------------------------- */
int nxs_abi_version(void);
/* "src_sha256=<16 hex digits> built=<UTC time> nvcc=<version> arch=sm_100a": the hash is over the library's
 * source files in sorted order (csrc Makefile, STAMP_SRC), so a caller can tell whether the binary it loaded
 * was built from the sources beside it (bench.py prints both) */
const char* nxs_build_info(void);
const char* nxs_strerror(int code);
/* number of visible CUDA devices (0 when there is none; never an error) */
int nxs_device_count(void);
/* creates a context bound to CUDA device `device` (own stream, twiddle tables,
 * pinned staging).  NXS_ENODEVICE when there is no GPU. */
int nxs_ctx_create(int device, nxs_ctx** out);
int nxs_ctx_destroy(nxs_ctx* ctx);
/* text of the last CUDA/NCCL failure seen by this context ("" if none) */
const char* nxs_last_error(const nxs_ctx* ctx);
int nxs_ctx_synchronize(nxs_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t nxs_ctx_launch_count(const nxs_ctx* ctx);

/* kernel timing for the roofline figure: while enabled, the dominant kernel of every compute
 * call is bracketed by CUDA events on its launch stream.  nxs_ctx_profile_read waits for
 * them, returns the summed kernel time and launch count since the last read, and resets. */
int nxs_ctx_profile(nxs_ctx* ctx, int enable);
int nxs_ctx_profile_read(nxs_ctx* ctx, double* total_ms, int64_t* launches);

/* phases of the last nxs_stft_f32_host call on this context, seconds since the call began:
 * [0] all copies/kernels enqueued, [1] first result slab in host memory, [2] last slab in host
 * memory, [3] host mirror threads done (= result complete).  bench.py reports them. */
int nxs_ctx_host_timeline(const nxs_ctx* ctx, double out_seconds[4]);

/* nxs_stft_f32_host moves the result in one of five ways, reported by nxs_ctx_host_mode for the
 * last call: 0 = both spectrum halves over PCIe; 1 = bins 0 .. fft_length/2 over PCIe straight into
 * the caller's rows, host threads write the conjugate-mirror bins; 3 = mixed: three channel chunks of
 * four as in 1, the fourth as in 0 (balances PCIe against host memory bandwidth); 2 = the result
 * buffer is pageable (not cudaHostRegister'ed -- e.g. a BEAM binary): the lower half lands in the
 * context's pinned ring and host threads copy it out and write the mirror half in one pass.
 * +16: the input was pageable and went through the pinned input ring.  For pinned results the
 * context measures the cost of modes 0 and 1 on its own calls and uses the cheaper (which one
 * depends on whether the box is short of PCIe or of host memory bandwidth; mode 3 never won where it
 * was measured and can only be pinned); nxs_ctx_set_host_mode pins a mode: -1 auto (default), 0, 1
 * or 3.  4 = small call (input and result up to 4 MiB, no mode pinned): one stream, both halves
 * over PCIe, no host threads -- the latency path of the short calls NxSignal.stft usually gets.
 * All modes give bit-identical results. */
int nxs_ctx_set_host_mode(nxs_ctx* ctx, int mode);
int nxs_ctx_host_mode(const nxs_ctx* ctx, int* mode);

/* ---- host-side closed forms (O(n); no GPU needed) ----------------------- */
/* NxSignal.Windows.{rectangular,bartlett,triangular,blackman,hamming,hann,kaiser}(n, opts)
 * lib/nx_signal/windows.ex:33,57,98,160,225,278,341 -- f32 output, `periodic`
 * = :is_periodic (ignored by rectangular/bartlett/triangular), beta/eps = kaiser
 * options.  Bit-compatible with Nx.BinaryBackend's per-op f32 rounding. */
int nxs_window_f32(int kind, int64_t n, int periodic, double beta, double eps, float* out);
/* the same windows with `type: :f64`: the reference's graph evaluated in double (windows.ex:58,161,
 * 226,279,342); kaiser keeps the reference's f32 scalars (I0(beta), Nx.Constants.pi()) */
int nxs_window_f64(int kind, int64_t n, int periodic, double beta, double eps, double* out);

/* NxSignal.Filters.firwin(num_taps, cutoff, opts)  lib/nx_signal/filters.ex:147-252
 * cutoffs in the units of sampling_rate; window_kind one of NXS_WIN_{HAMMING,HANN,
 * BLACKMAN,BARTLETT,RECTANGULAR,KAISER}; errors: NXS_EINVAL with the reference's
 * three ArgumentError conditions (filters.ex:170-178, 189-193, 274-277). */
int nxs_firwin_f32(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta,
                   int pass_zero, int scale, double sampling_rate, float* out);
/* `type: :f64` (filters.ex:153): the same graph in double */
int nxs_firwin_f64(int64_t num_taps, const double* cutoffs, int ncut, int window_kind, double beta,
                   int pass_zero, int scale, double sampling_rate, double* out);

/* NxSignal.fft_frequencies(sampling_rate, fft_length: n)  lib/nx_signal.ex:154-166 */
int nxs_fft_frequencies_f32(double sampling_rate, int64_t fft_length, float* out);
/* with the head's other options: `endpoint:` (Nx.linspace divides by n - 1 instead of n) and
 * `type: :f64` (out is double* when is_f64) */
int nxs_fft_frequencies_ex(double sampling_rate, int64_t fft_length, int endpoint, int is_f64, void* out);

/* NxSignal.mel_filters(fft_length, mel_bins, sampling_rate, max_mel:, mel_frequency_spacing:)
 * lib/nx_signal.ex:397-445 -- out [mel_bins][fft_length] f32, bit-compatible with the
 * reference's per-op f32 rounding (defaults: max_mel 3016, mel_frequency_spacing 200/3). */
int nxs_mel_filters_f32(int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel,
                        double mel_frequency_spacing, float* out);

/* frame times of stft/3: linspace(N/(2 sr), N/(2 sr) * M, n: M)  lib/nx_signal.ex:108-111 */
int nxs_stft_times_f32(int64_t frame_length, double sampling_rate, int64_t num_frames, float* out);

/* frame count of as_windowed/2 (shape rule at lib/nx_signal.ex:289-298).
 * pad_mode NXS_PAD_*; pad_lo/pad_hi used only for NXS_PAD_EXPLICIT. */
int nxs_num_frames(int64_t length, int64_t window_length, int64_t stride, int pad_mode,
                   int64_t pad_lo, int64_t pad_hi, int64_t* num_frames);

/* ---- STFT: NxSignal.stft(data, window, opts)  lib/nx_signal.ex:68-130 ----
 * x      [channels][x_ld] f32, the first `length` samples of each row are used
 *        (channels = product of the Nx vectorised axes)
 * window [frame_length] f32
 * hop    = frame_length - overlap_length
 * z      [channels][num_frames][fft_length] c64 (full two-sided spectrum)
 * Scaling :spectrum divides by sum(w), :psd by sqrt(sr * sum(w^2)).
 * fft_length != frame_length zero-pads / truncates each windowed frame (Nx.fft).
 * times/frequencies are host-side: nxs_stft_times_f32 / nxs_fft_frequencies_f32. */
int nxs_stft_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                     int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                     float* z, void* stream);
int nxs_stft_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                      const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                      int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                      float* z);

/* The same head on COMPLEX data (the reference's graph takes any numeric tensor: Nx.multiply(window) ->
 * Nx.fft, lib/nx_signal.ex:101-102): x [channels][x_ld] c64 (interleaved), z as above. */
int nxs_stft_c64_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                     int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                     float* z, void* stream);
int nxs_stft_c64_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                      const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                      int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                      float* z);

/* One-sided variant (opt-in): only bins 0 .. fft_length/2 of every frame are
 * stored, z [channels][num_frames][z_ld] c64 with z_ld >= fft_length/2 + 1.  The remaining
 * bins of the reference's two-sided result (lib/nx_signal.ex:49) are conj(z[fft_length-k]);
 * nxs_stft_f32_host uses this form on the wire and mirrors on the host. */
int nxs_stft_onesided_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length,
                              int64_t x_ld, const float* window, int64_t frame_length, int64_t hop,
                              int64_t fft_length, int pad_mode, int64_t pad_lo, int64_t pad_hi,
                              int scaling, double sampling_rate, float* z, int64_t z_ld, void* stream);

/* ---- log-mel: NxSignal.stft_to_mel(z, sampling_rate, fft_length:, mel_bins:, ...)
 * lib/nx_signal.ex:486-513.
 * z   [channels][num_frames][z_ld] c64; only bins 0 .. fft_length/2 - 1 are read, so both the
 *     two-sided (z_ld = fft_length) and the one-sided (z_ld >= fft_length/2 + 1) STFT fit
 * out [channels][num_frames][mel_bins] f32 = (max(log10(clip(|z|^2 . filters, 1e-10)), max - 8) + 4) / 4,
 *     the maximum taken per channel (Nx.reduce_max on a vectorised tensor). */
int nxs_stft_to_mel_f32_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                            int64_t z_ld, int64_t fft_length, int64_t mel_bins, double sampling_rate,
                            double max_mel, double mel_frequency_spacing, float* out, void* stream);
int nxs_stft_to_mel_f32_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                             int64_t z_ld, int64_t fft_length, int64_t mel_bins, double sampling_rate,
                             double max_mel, double mel_frequency_spacing, float* out);

/* Fused form of NxSignal.stft/3 |> NxSignal.stft_to_mel/3: the same
 * values as nxs_stft_f32_dev followed by nxs_stft_to_mel_f32_dev, but the spectrum never leaves
 * the SM (8 fft_length bytes per frame not written and not re-read).  Served by the TMA-staged
 * kernels (power-of-two fft_length 512 .. 8192 == frame_length, hop % 4 == 0, 16-byte aligned
 * rows); NXS_EUNSUPPORTED otherwise -- the caller then chains the two entries.
 * out [channels][num_frames][mel_bins] f32 */
int nxs_stft_mel_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                         const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                         int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                         int64_t mel_bins, double max_mel, double mel_frequency_spacing, float* out,
                         void* stream);
/* host form: x / window / out are host pointers (H2D -> kernels -> D2H inside, synchronous).  0.13x the
 * PCIe bytes of the STFT host entry (4 mel_bins bytes per frame come back instead of 4 (fft_length + 2));
 * configurations the fused kernel does not serve are chained on the device inside the call. */
int nxs_stft_mel_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                         const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                         int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                         int64_t mel_bins, double max_mel, double mel_frequency_spacing, float* out);

/* ---- ISTFT: NxSignal.istft(data, window, opts)  lib/nx_signal.ex:582-638 ---
 * z      [channels][num_frames][z_len] c64; Nx.ifft(length: fft_length) pads /
 *        truncates the last axis to fft_length, which must equal frame_length
 *        (the reference's `frames * window` broadcast, :628)
 * y      [channels][num_frames*hop + frame_length - hop] c64 */
int nxs_istft_c64_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                      int64_t z_len, const float* window, int64_t frame_length, int64_t hop,
                      int64_t fft_length, int scaling, double sampling_rate, float* y, void* stream);
int nxs_istft_c64_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                       int64_t z_len, const float* window, int64_t frame_length, int64_t hop,
                       int64_t fft_length, int scaling, double sampling_rate, float* y);

/* ---- c2r ISTFT (opt-in extension, not a reference head) ---
 * The counterpart of nxs_stft_onesided_f32_dev: z [channels][num_frames][z_ld]
 * c64 holds bins 0 .. fft_length/2 (z_ld >= fft_length/2 + 1), y
 * [channels][num_frames*hop + frame_length - hop] is REAL f32 and equals
 * Re(NxSignal.istft(ext(z), window, opts)) (lib/nx_signal.ex:582-638) with
 * ext(z)[k] = conj(z[fft_length - k]) for k > fft_length/2.  fft_length must be
 * even and equal frame_length.  fft_length in {512, 1024, 2048, 4096} with
 * hop = N/2, N/4, N/8 runs the packed half-length kernel (asynchronous); other
 * shapes extend the spectrum on the device, run the c64 path and synchronise
 * the stream before returning. */
int nxs_istft_c2r_f32_dev(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                          int64_t z_ld, const float* window, int64_t frame_length, int64_t hop,
                          int64_t fft_length, int scaling, double sampling_rate, float* y, void* stream);
int nxs_istft_c2r_f32_host(nxs_ctx* ctx, const float* z, int64_t channels, int64_t num_frames,
                           int64_t z_ld, const float* window, int64_t frame_length, int64_t hop,
                           int64_t fft_length, int scaling, double sampling_rate, float* y);

/* ---- framing: NxSignal.as_windowed(tensor, opts)  lib/nx_signal.ex:249-364 -
 * x [channels][x_ld] (elem_size 4 or 8 bytes: f32/s32 or c64/s64/f64)
 * out [channels][num_frames][window_length] of the same element type */
int nxs_as_windowed_dev(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                        int64_t x_ld, int64_t window_length, int64_t stride, int pad_mode,
                        int64_t pad_lo, int64_t pad_hi, void* out, void* stream);
int nxs_as_windowed_host(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                         int64_t x_ld, int64_t window_length, int64_t stride, int pad_mode,
                         int64_t pad_lo, int64_t pad_hi, void* out);

/* ---- NxSignal.overlap_and_add(tensor, overlap_length:)  lib/nx_signal.ex:684-735
 * t [batch][num_frames][frame_length] f32 or c64 -> out [batch][num_frames*hop + overlap] */
int nxs_overlap_and_add_f32_dev(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                int64_t frame_length, int64_t overlap_length, float* out, void* stream);
int nxs_overlap_and_add_c64_dev(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                int64_t frame_length, int64_t overlap_length, float* out, void* stream);
int nxs_overlap_and_add_f32_host(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                 int64_t frame_length, int64_t overlap_length, float* out);
int nxs_overlap_and_add_c64_host(nxs_ctx* ctx, const float* t, int64_t batch, int64_t num_frames,
                                 int64_t frame_length, int64_t overlap_length, float* out);

/* ---- FIR apply: NxSignal.Convolution.convolve(x, taps, mode:, method:) for the
 * batched-FIR form x {C, L} * h {1, K}  (lib/nx_signal/convolution.ex:38-58,
 * 95-211, 252-329; the "broadcastable" test, test/nx_signal/convolutions_test.exs:95-143).
 * Overlap-save on the GPU reproduces the values of both :direct and :fft methods.
 * y [channels][out_len], out_len = L+K-1 (full) | L (same) | max(L,K)-min(L,K)+1 (valid) */
int nxs_fir_out_len(int64_t length, int64_t num_taps, int mode, int64_t* out_len);
int nxs_fir_f32_dev(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* taps, int64_t num_taps, int mode, float* y, int64_t y_ld, void* stream);
int nxs_fir_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                     const float* taps, int64_t num_taps, int mode, float* y, int64_t y_ld);

/* ---- general N-d convolution (rank <= 3 after the shim squeezes), real or complex:
 * NxSignal.Convolution.convolve / correlate / fftconvolve for arbitrary small
 * operands (lib/nx_signal/convolution.ex:38-93, 252-298).  Direct summation in
 * fp32 with fp64 accumulation disabled; shapes as int64[3] (padded with 1s).
 * is_complex: operands and result are c64 (interleaved) when 1, f32 when 0. */
int nxs_convolve_nd_dev(nxs_ctx* ctx, const float* a, const int64_t a_shape[3], const float* b,
                        const int64_t b_shape[3], int is_complex, int mode, float* out, void* stream);
int nxs_convolve_nd_host(nxs_ctx* ctx, const float* a, const int64_t a_shape[3], const float* b,
                         const int64_t b_shape[3], int is_complex, int mode, float* out);

/* ---- spectrogram-adjacent operators -------------------
 * Tensors are row-major, rank <= 3 for median / wiener (rank <= 8 for argrel*).
 *
 * NxSignal.Filters.median(t, kernel_shape: ks)  lib/nx_signal/filters.ex:17-56
 *   out[i] = Nx.median of the ks-window that STARTS at i, the start clamped so
 *   the window stays inside the tensor (Nx.slice); f32 out.  kernel_shape must
 *   have t's rank ("kernel shape must be of the same rank as the tensor", :36)
 *   and 1 <= ks[d] <= shape[d].
 * NxSignal.Filters.wiener(t, kernel_size:, noise:)  filters.ex:80-110, 281-303
 *   computed in f64 like the reference; t / out are f32 (is_f64 = 0) or f64;
 *   has_noise = 0 estimates the noise as mean(local variance).
 * NxSignal.PeakFinding.argrelextrema / argrelmin / argrelmax(data, axis:, order:)
 *   lib/nx_signal/peak_finding.ex:131-391.  comparator NXS_CMP_*; indices is
 *   s32 [prod(shape)][rank]: the multi-indices of the extrema in row-major order
 *   followed by rows of -1; *valid_count their number (device pointer in the
 *   _dev form).  Arbitrary comparator functions are not supported. */
int nxs_median_f32_dev(nxs_ctx* ctx, const float* t, int rank, const int64_t* shape,
                       const int64_t* kernel_shape, float* out, void* stream);
int nxs_median_f32_host(nxs_ctx* ctx, const float* t, int rank, const int64_t* shape,
                        const int64_t* kernel_shape, float* out);
int nxs_wiener_dev(nxs_ctx* ctx, const void* t, int is_f64, int rank, const int64_t* shape,
                   const int64_t* kernel_size, int has_noise, double noise, void* out, void* stream);
int nxs_wiener_host(nxs_ctx* ctx, const void* t, int is_f64, int rank, const int64_t* shape,
                    const int64_t* kernel_size, int has_noise, double noise, void* out);
int nxs_argrelextrema_f32_dev(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis,
                              int order, int comparator, int32_t* indices, int64_t* valid_count,
                              void* stream);
int nxs_argrelextrema_f32_host(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis,
                               int order, int comparator, int32_t* indices, int64_t* valid_count);

/* ---- multi-GPU setup: one broadcast of the coefficient block (window or FIR
 * taps) from rank 0 over NCCL; no other collective exists on this path
 * (SURVEY.md 8e).  `comm` is an ncclComm_t passed as void*; buf is a device
 * pointer on every rank. */
int nxs_bcast_coeffs_dev(nxs_ctx* ctx, void* comm, float* buf, int64_t count, int root, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NXSIGNAL_B200_H */
