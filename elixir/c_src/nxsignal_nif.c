/*
 * nxsignal_nif.c -- NIF shim over include/nxsignal_b200.h: one NIF per "_host" export plus the
 * host-side closed forms, so that every public head of the accelerated path
 * (NxSignal.stft/istft/as_windowed/overlap_and_add/fft_frequencies/mel_filters/stft_to_mel,
 * NxSignal.Windows.*, NxSignal.Filters.firwin/median/wiener,
 * NxSignal.Convolution.convolve/correlate/fftconvolve, NxSignal.PeakFinding.argrel*) can be served by
 * the B200 backend (elixir/lib/nx_signal_b200.ex holds the heads).
 *
 * There is no BEAM in this repository's build image, so this file is never linked here; it is
 * syntax- and type-checked against tests/stubs/erl_nif.h (tests/test_elixir_boundary.py), which also
 * checks that every NIF the Elixir module declares exists in funcs[] with the same arity.  Build
 * on a machine with Erlang/OTP 27 + CUDA:
 *   cc -O2 -fPIC -shared -I$ERL_ROOT/usr/include -I../../include nxsignal_nif.c \
 *      -L../../nx_signal_b200/lib -lnxsignal_b200 -o priv/nxsignal_nif.so
 *
 * Conventions
 *  - tensors cross as Nx.to_binary/1 binaries (row-major, native endian, c64 = interleaved f32);
 *    results are allocated with enif_make_new_binary so the VM owns them (pageable memory: the
 *    library stages it through its pinned rings, see nxs_stft_f32_host);
 *  - every size is validated BEFORE anything is allocated: dimensions must be positive, products
 *    must not overflow, and each inspected binary must hold exactly what the dimensions say --
 *    otherwise {:error, :argument_error, msg} (raised as ArgumentError by the Elixir side);
 *  - one nxs_ctx is single-threaded (nxsignal_b200.h), but dirty NIFs of many BEAM processes run
 *    concurrently on different scheduler threads: the context resource carries a mutex that is held
 *    around every library call;
 *  - every compute entry is a dirty IO-bound NIF: calls block on PCIe transfers for milliseconds.
 */
#include <erl_nif.h>
#include <stdint.h>
#include <string.h>

#include "nxsignal_b200.h"

static ErlNifResourceType* CTX_TYPE;

typedef struct {
  nxs_ctx* ctx;
  ErlNifMutex* mu;
} ctx_res;

static void ctx_dtor(ErlNifEnv* env, void* obj) {
  ctx_res* r = (ctx_res*)obj;
  (void)env;
  if (r->ctx) nxs_ctx_destroy(r->ctx);
  if (r->mu) enif_mutex_destroy(r->mu);
}

static ERL_NIF_TERM mk_error(ErlNifEnv* env, int rc) {
  /* NXS_EINVAL / NXS_ESHAPE -> :argument_error (raised as ArgumentError by the Elixir side) */
  const char* kind = (rc == NXS_EINVAL || rc == NXS_ESHAPE) ? "argument_error" : "runtime_error";
  return enif_make_tuple3(env, enif_make_atom(env, "error"), enif_make_atom(env, kind),
                          enif_make_string(env, nxs_strerror(rc), ERL_NIF_LATIN1));
}

static ERL_NIF_TERM mk_ok1(ErlNifEnv* env, ERL_NIF_TERM a) {
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), a);
}

/* ---- size arithmetic: everything that sizes a binary goes through these -------------------- */
#define NXS_NIF_MAX_BYTES ((int64_t)1 << 46) /* 64 TiB: anything larger is a caller bug, not a tensor */

/* *out = a * b for non-negative operands; 0 on overflow or when the product exceeds the cap */
static int mul_ok(int64_t a, int64_t b, int64_t* out) {
  if (a < 0 || b < 0) return 0;
  if (a != 0 && b > NXS_NIF_MAX_BYTES / a) return 0;
  *out = a * b;
  return 1;
}

/* bytes = d0 * d1 * d2 * elem for positive dimensions; 0 when any is < 1 or the product overflows */
static int bytes3(int64_t d0, int64_t d1, int64_t d2, int64_t elem, int64_t* bytes) {
  int64_t t;
  if (d0 < 1 || d1 < 1 || d2 < 1 || elem < 1) return 0;
  return mul_ok(d0, d1, &t) && mul_ok(t, d2, &t) && mul_ok(t, elem, bytes);
}

static int get_i64(ErlNifEnv* env, ERL_NIF_TERM t, int64_t* v) {
  ErlNifSInt64 x;
  if (!enif_get_int64(env, t, &x)) return 0;
  *v = (int64_t)x;
  return 1;
}

/* double from an Erlang float or integer */
static int get_f64(ErlNifEnv* env, ERL_NIF_TERM t, double* v) {
  ErlNifSInt64 x;
  if (enif_get_double(env, t, v)) return 1;
  if (enif_get_int64(env, t, &x)) {
    *v = (double)x;
    return 1;
  }
  return 0;
}

/* list of up to `cap` integers -> v[], *n */
static int get_i64_list(ErlNifEnv* env, ERL_NIF_TERM list, int64_t* v, unsigned cap, unsigned* n) {
  ERL_NIF_TERM head, tail = list;
  unsigned len = 0;
  if (!enif_get_list_length(env, list, &len) || len > cap) return 0;
  for (unsigned i = 0; i < len; ++i)
    if (!enif_get_list_cell(env, tail, &head, &tail) || !get_i64(env, head, &v[i])) return 0;
  *n = len;
  return 1;
}

static int get_ctx(ErlNifEnv* env, ERL_NIF_TERM t, ctx_res** r) {
  return enif_get_resource(env, t, CTX_TYPE, (void**)r) && (*r)->ctx != NULL;
}

#define LOCK(r) enif_mutex_lock((r)->mu)
#define UNLOCK(r) enif_mutex_unlock((r)->mu)

/* ---- context ------------------------------------------------------------------------------- */
static ERL_NIF_TERM ctx_create(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int dev;
  if (argc != 1 || !enif_get_int(env, argv[0], &dev)) return enif_make_badarg(env);
  nxs_ctx* c = NULL;
  int rc = nxs_ctx_create(dev, &c);
  if (rc) return mk_error(env, rc);
  ctx_res* r = (ctx_res*)enif_alloc_resource(CTX_TYPE, sizeof(ctx_res));
  r->ctx = c;
  r->mu = enif_mutex_create((char*)"nxs_ctx");
  if (!r->mu) {
    enif_release_resource(r); /* the destructor frees the context */
    return mk_error(env, NXS_ENOMEM);
  }
  ERL_NIF_TERM t = enif_make_resource(env, r);
  enif_release_resource(r);
  return mk_ok1(env, t);
}

/* ---- stft(ctx, x_bin, channels, length, window_bin, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling, sr)
 * NxSignal.stft/3, lib/nx_signal.ex:68-130 -> {:ok, z, times, frequencies, num_frames} */
static ERL_NIF_TERM stft_common(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[], int cplx) {
  ctx_res* r;
  ErlNifBinary x, w;
  int64_t ch, len, hop, nfft, lo, hi, frames = 0, xb, zb;
  int pad, scal;
  double sr;
  if (argc != 12 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &x) ||
      !get_i64(env, argv[2], &ch) || !get_i64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &w) ||
      !get_i64(env, argv[5], &hop) || !get_i64(env, argv[6], &nfft) || !enif_get_int(env, argv[7], &pad) ||
      !get_i64(env, argv[8], &lo) || !get_i64(env, argv[9], &hi) || !enif_get_int(env, argv[10], &scal) ||
      !get_f64(env, argv[11], &sr))
    return enif_make_badarg(env);
  const int64_t n = (int64_t)(w.size / sizeof(float));
  if (n < 1 || w.size != (size_t)n * sizeof(float) || hop < 1 || nfft < 1) return mk_error(env, NXS_ESHAPE);
  if (!bytes3(ch, len, 1, (cplx ? 2 : 1) * sizeof(float), &xb) || (int64_t)x.size != xb) return mk_error(env, NXS_ESHAPE);
  int rc = nxs_num_frames(len, n, hop, pad, lo, hi, &frames);
  if (rc) return mk_error(env, rc);
  if (!bytes3(ch, frames, nfft, 2 * sizeof(float), &zb)) return mk_error(env, NXS_ESHAPE); /* no frame fits, or too large */
  ERL_NIF_TERM zt, tt, ft;
  float* z = (float*)enif_make_new_binary(env, (size_t)zb, &zt);
  float* times = (float*)enif_make_new_binary(env, (size_t)frames * sizeof(float), &tt);
  float* freqs = (float*)enif_make_new_binary(env, (size_t)nfft * sizeof(float), &ft);
  LOCK(r);
  rc = cplx ? nxs_stft_c64_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)w.data, n, hop, nfft, pad,
                                lo, hi, scal, sr, z)
            : nxs_stft_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)w.data, n, hop, nfft, pad,
                                lo, hi, scal, sr, z);
  UNLOCK(r);
  if (!rc) rc = nxs_stft_times_f32(n, sr, frames, times);
  if (!rc) rc = nxs_fft_frequencies_f32(sr, nfft, freqs);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple5(env, enif_make_atom(env, "ok"), zt, tt, ft, enif_make_int64(env, frames));
}

static ERL_NIF_TERM stft(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  return stft_common(env, argc, argv, 0);
}
/* the same head on complex data (x_bin holds c64): Nx.multiply(window) -> Nx.fft, lib/nx_signal.ex:101-102 */
static ERL_NIF_TERM stft_c64(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  return stft_common(env, argc, argv, 1);
}

/* shared argument block of istft / istft_c2r:
 * (ctx, z_bin, channels, frames, z_len, window_bin, hop, fft_length, scaling, sr) */
static ERL_NIF_TERM istft_common(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[], int c2r) {
  ctx_res* r;
  ErlNifBinary z, w;
  int64_t ch, frames, zlen, hop, nfft, zb, yb, out_len, t;
  int scal;
  double sr;
  if (argc != 10 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &z) ||
      !get_i64(env, argv[2], &ch) || !get_i64(env, argv[3], &frames) || !get_i64(env, argv[4], &zlen) ||
      !enif_inspect_binary(env, argv[5], &w) || !get_i64(env, argv[6], &hop) || !get_i64(env, argv[7], &nfft) ||
      !enif_get_int(env, argv[8], &scal) || !get_f64(env, argv[9], &sr))
    return enif_make_badarg(env);
  const int64_t n = (int64_t)(w.size / sizeof(float));
  if (n < 1 || w.size != (size_t)n * sizeof(float)) return mk_error(env, NXS_ESHAPE);
  if (hop < 1 || hop > n) return mk_error(env, NXS_EINVAL); /* overlap_length in [0, window) */
  if (!bytes3(ch, frames, zlen, 2 * sizeof(float), &zb) || (int64_t)z.size != zb) return mk_error(env, NXS_ESHAPE);
  if (!mul_ok(frames, hop, &t)) return mk_error(env, NXS_ESHAPE);
  out_len = t + (n - hop);
  if (!bytes3(ch, out_len, 1, (c2r ? 1 : 2) * sizeof(float), &yb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM yt;
  float* y = (float*)enif_make_new_binary(env, (size_t)yb, &yt);
  LOCK(r);
  int rc = c2r ? nxs_istft_c2r_f32_host(r->ctx, (const float*)z.data, ch, frames, zlen, (const float*)w.data, n, hop,
                                        nfft, scal, sr, y)
               : nxs_istft_c64_host(r->ctx, (const float*)z.data, ch, frames, zlen, (const float*)w.data, n, hop, nfft,
                                    scal, sr, y);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), yt, enif_make_int64(env, out_len));
}

/* NxSignal.istft/3, lib/nx_signal.ex:582-638 -> {:ok, y (c64), out_len} */
static ERL_NIF_TERM istft(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  return istft_common(env, argc, argv, 0);
}
/* opt-in: one-sided spectrum in (z_len >= fft_length/2 + 1), REAL signal out -> {:ok, y (f32), out_len} */
static ERL_NIF_TERM istft_c2r(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  return istft_common(env, argc, argv, 1);
}

/* ---- fir(ctx, x_bin, channels, length, taps_bin, mode): the batched FIR form of
 * NxSignal.Convolution.convolve/3 (x {C, L} * h {1, K}), lib/nx_signal/convolution.ex:38-58 */
static ERL_NIF_TERM fir(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x, h;
  int64_t ch, len, xb, yb, out_len = 0;
  int mode;
  if (argc != 6 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &x) ||
      !get_i64(env, argv[2], &ch) || !get_i64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &h) ||
      !enif_get_int(env, argv[5], &mode))
    return enif_make_badarg(env);
  const int64_t k = (int64_t)(h.size / sizeof(float));
  if (k < 1 || h.size != (size_t)k * sizeof(float)) return mk_error(env, NXS_ESHAPE);
  if (!bytes3(ch, len, 1, sizeof(float), &xb) || (int64_t)x.size != xb) return mk_error(env, NXS_ESHAPE);
  int rc = nxs_fir_out_len(len, k, mode, &out_len);
  if (rc) return mk_error(env, rc);
  if (!bytes3(ch, out_len, 1, sizeof(float), &yb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM yt;
  float* y = (float*)enif_make_new_binary(env, (size_t)yb, &yt);
  LOCK(r);
  rc = nxs_fir_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)h.data, k, mode, y, out_len);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), yt, enif_make_int64(env, out_len));
}

/* ---- convolve_nd(ctx, a_bin, a_shape, b_bin, b_shape, is_complex, mode): general small N-d operands
 * (rank <= 3 after the shim squeezes), convolution.ex:38-93, 252-298 -> {:ok, out, out_shape} */
static ERL_NIF_TERM convolve_nd(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary a, b;
  int64_t as[3], bs[3], os[3], ab, bb, ob;
  unsigned ra = 0, rb = 0;
  int cplx, mode;
  if (argc != 7 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &a) ||
      !get_i64_list(env, argv[2], as, 3, &ra) || !enif_inspect_binary(env, argv[3], &b) ||
      !get_i64_list(env, argv[4], bs, 3, &rb) || !enif_get_int(env, argv[5], &cplx) ||
      !enif_get_int(env, argv[6], &mode) || ra != 3 || rb != 3)
    return enif_make_badarg(env);
  const int64_t es = (cplx ? 2 : 1) * (int64_t)sizeof(float);
  if (!bytes3(as[0], as[1], as[2], es, &ab) || (int64_t)a.size != ab) return mk_error(env, NXS_ESHAPE);
  if (!bytes3(bs[0], bs[1], bs[2], es, &bb) || (int64_t)b.size != bb) return mk_error(env, NXS_ESHAPE);
  int ok1 = 1, ok2 = 1;
  for (int d = 0; d < 3; ++d) {
    ok1 = ok1 && as[d] >= bs[d];
    ok2 = ok2 && as[d] <= bs[d];
  }
  for (int d = 0; d < 3; ++d) {
    if (mode == NXS_MODE_FULL) os[d] = as[d] + bs[d] - 1;
    else if (mode == NXS_MODE_SAME) os[d] = as[d];
    else if (mode == NXS_MODE_VALID) {
      if (!ok1 && !ok2) return mk_error(env, NXS_ESHAPE); /* convolution.ex:131-134 */
      os[d] = (ok1 ? as[d] - bs[d] : bs[d] - as[d]) + 1;
    } else return mk_error(env, NXS_EINVAL);
  }
  if (!bytes3(os[0], os[1], os[2], es, &ob)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM ot;
  float* out = (float*)enif_make_new_binary(env, (size_t)ob, &ot);
  LOCK(r);
  int rc = nxs_convolve_nd_host(r->ctx, (const float*)a.data, as, (const float*)b.data, bs, cplx, mode, out);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), ot,
                          enif_make_list3(env, enif_make_int64(env, os[0]), enif_make_int64(env, os[1]),
                                          enif_make_int64(env, os[2])));
}

/* ---- as_windowed(ctx, x_bin, elem_size, channels, length, window_length, stride, pad_mode, lo, hi)
 * NxSignal.as_windowed/2, lib/nx_signal.ex:249-364 -> {:ok, frames_bin, num_frames} */
static ERL_NIF_TERM as_windowed(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x;
  int64_t ch, len, wl, stride, lo, hi, xb, ob, frames = 0;
  int es, pad;
  if (argc != 10 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &x) ||
      !enif_get_int(env, argv[2], &es) || !get_i64(env, argv[3], &ch) || !get_i64(env, argv[4], &len) ||
      !get_i64(env, argv[5], &wl) || !get_i64(env, argv[6], &stride) || !enif_get_int(env, argv[7], &pad) ||
      !get_i64(env, argv[8], &lo) || !get_i64(env, argv[9], &hi))
    return enif_make_badarg(env);
  if (es != 4 && es != 8) return mk_error(env, NXS_EUNSUPPORTED);
  if (stride < 1) return mk_error(env, NXS_EINVAL); /* "expected an integer >= 1", lib/nx_signal.ex:282-284 */
  if (!bytes3(ch, len, 1, es, &xb) || (int64_t)x.size != xb) return mk_error(env, NXS_ESHAPE);
  int rc = nxs_num_frames(len, wl, stride, pad, lo, hi, &frames);
  if (rc) return mk_error(env, rc);
  if (!bytes3(ch, frames, wl, es, &ob)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM ot;
  void* out = enif_make_new_binary(env, (size_t)ob, &ot);
  LOCK(r);
  rc = nxs_as_windowed_host(r->ctx, x.data, es, ch, len, len, wl, stride, pad, lo, hi, out);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), ot, enif_make_int64(env, frames));
}

/* ---- overlap_and_add(ctx, t_bin, is_complex, batch, num_frames, frame_length, overlap_length)
 * NxSignal.overlap_and_add/2, lib/nx_signal.ex:684-735 -> {:ok, out_bin, out_len} */
static ERL_NIF_TERM overlap_and_add(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary t;
  int64_t batch, frames, flen, ov, tb, ob, out_len, m;
  int cplx;
  if (argc != 7 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &t) ||
      !enif_get_int(env, argv[2], &cplx) || !get_i64(env, argv[3], &batch) || !get_i64(env, argv[4], &frames) ||
      !get_i64(env, argv[5], &flen) || !get_i64(env, argv[6], &ov))
    return enif_make_badarg(env);
  const int64_t es = (cplx ? 2 : 1) * (int64_t)sizeof(float);
  if (ov < 0 || ov >= flen) return mk_error(env, NXS_EINVAL); /* lib/nx_signal.ex:692-695 */
  if (!bytes3(batch, frames, flen, es, &tb) || (int64_t)t.size != tb) return mk_error(env, NXS_ESHAPE);
  if (!mul_ok(frames, flen - ov, &m)) return mk_error(env, NXS_ESHAPE);
  out_len = m + ov;
  if (!bytes3(batch, out_len, 1, es, &ob)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM ot;
  float* out = (float*)enif_make_new_binary(env, (size_t)ob, &ot);
  LOCK(r);
  int rc = cplx ? nxs_overlap_and_add_c64_host(r->ctx, (const float*)t.data, batch, frames, flen, ov, out)
                : nxs_overlap_and_add_f32_host(r->ctx, (const float*)t.data, batch, frames, flen, ov, out);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), ot, enif_make_int64(env, out_len));
}

/* ---- stft_to_mel(ctx, z_bin, channels, frames, z_len, fft_length, mel_bins, sr, max_mel, f_sp)
 * NxSignal.stft_to_mel/3, lib/nx_signal.ex:486-513 */
static ERL_NIF_TERM stft_to_mel(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary z;
  int64_t ch, frames, zlen, nfft, mels, zb, mb;
  double sr, max_mel, f_sp;
  if (argc != 10 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &z) ||
      !get_i64(env, argv[2], &ch) || !get_i64(env, argv[3], &frames) || !get_i64(env, argv[4], &zlen) ||
      !get_i64(env, argv[5], &nfft) || !get_i64(env, argv[6], &mels) || !get_f64(env, argv[7], &sr) ||
      !get_f64(env, argv[8], &max_mel) || !get_f64(env, argv[9], &f_sp))
    return enif_make_badarg(env);
  if (!bytes3(ch, frames, zlen, 2 * sizeof(float), &zb) || (int64_t)z.size != zb) return mk_error(env, NXS_ESHAPE);
  if (nfft < 2 || !bytes3(ch, frames, mels, sizeof(float), &mb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM mt;
  float* mel = (float*)enif_make_new_binary(env, (size_t)mb, &mt);
  LOCK(r);
  int rc = nxs_stft_to_mel_f32_host(r->ctx, (const float*)z.data, ch, frames, zlen, nfft, mels, sr, max_mel, f_sp, mel);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, mt);
}

/* ---- stft_mel(ctx, x_bin, channels, length, window_bin, hop, fft_length, pad_mode, lo, hi, scaling, sr,
 *          mel_bins, max_mel, f_sp) -- NxSignal.stft/3 |> NxSignal.stft_to_mel/3 in one call
 * (lib/nx_signal.ex:68-130, 486-513): only the [frames][mel_bins] tensor comes back over PCIe */
static ERL_NIF_TERM stft_mel(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x, w;
  int64_t ch, len, hop, nfft, lo, hi, mels, xb, mb, frames = 0;
  int pad, scal;
  double sr, max_mel, f_sp;
  if (argc != 15 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &x) ||
      !get_i64(env, argv[2], &ch) || !get_i64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &w) ||
      !get_i64(env, argv[5], &hop) || !get_i64(env, argv[6], &nfft) || !enif_get_int(env, argv[7], &pad) ||
      !get_i64(env, argv[8], &lo) || !get_i64(env, argv[9], &hi) || !enif_get_int(env, argv[10], &scal) ||
      !get_f64(env, argv[11], &sr) || !get_i64(env, argv[12], &mels) || !get_f64(env, argv[13], &max_mel) ||
      !get_f64(env, argv[14], &f_sp))
    return enif_make_badarg(env);
  const int64_t n = (int64_t)(w.size / sizeof(float));
  if (n < 1 || w.size != (size_t)n * sizeof(float) || hop < 1 || nfft < 2) return mk_error(env, NXS_ESHAPE);
  if (!bytes3(ch, len, 1, sizeof(float), &xb) || (int64_t)x.size != xb) return mk_error(env, NXS_ESHAPE);
  int rc = nxs_num_frames(len, n, hop, pad, lo, hi, &frames);
  if (rc) return mk_error(env, rc);
  if (!bytes3(ch, frames, mels, sizeof(float), &mb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM mt;
  float* mel = (float*)enif_make_new_binary(env, (size_t)mb, &mt);
  LOCK(r);
  rc = nxs_stft_mel_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)w.data, n, hop, nfft, pad, lo,
                             hi, scal, sr, mels, max_mel, f_sp, mel);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), mt, enif_make_int64(env, frames));
}

/* shape / kernel lists of rank <= 3 and the element count they imply */
static int get_shape3(ErlNifEnv* env, ERL_NIF_TERM shape_t, ERL_NIF_TERM kernel_t, int64_t* shape, int64_t* ks,
                      unsigned* rank, int64_t* total, int* same_rank) {
  unsigned krank = 0;
  if (!get_i64_list(env, shape_t, shape, 3, rank) || !get_i64_list(env, kernel_t, ks, 3, &krank)) return 0;
  *same_rank = (*rank == krank);
  *total = 1;
  for (unsigned i = 0; i < *rank; ++i) {
    if (shape[i] < 1 || !mul_ok(*total, shape[i], total)) return 0;
  }
  return 1;
}

/* ---- median(ctx, t_bin, shape, kernel_shape) -- NxSignal.Filters.median/2, lib/nx_signal/filters.ex:17-56 */
static ERL_NIF_TERM median(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary t;
  int64_t shape[3], ks[3], total, tb;
  unsigned rank = 0;
  int same;
  if (argc != 4 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &t) ||
      !get_shape3(env, argv[2], argv[3], shape, ks, &rank, &total, &same))
    return enif_make_badarg(env);
  if (!same) return mk_error(env, NXS_ESHAPE); /* "kernel shape must be of the same rank as the tensor" */
  if (!mul_ok(total, sizeof(float), &tb) || (int64_t)t.size != tb) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM ot;
  float* out = (float*)enif_make_new_binary(env, (size_t)tb, &ot);
  LOCK(r);
  int rc = nxs_median_f32_host(r->ctx, (const float*)t.data, (int)rank, shape, ks, out);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, ot);
}

/* ---- wiener(ctx, t_bin, is_f64, shape, kernel_size, has_noise, noise)
 * NxSignal.Filters.wiener/2, lib/nx_signal/filters.ex:80-110, 281-303 */
static ERL_NIF_TERM wiener(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary t;
  int64_t shape[3], ks[3], total, tb;
  unsigned rank = 0;
  int same, is_f64, has_noise;
  double noise;
  if (argc != 7 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &t) ||
      !enif_get_int(env, argv[2], &is_f64) || !get_shape3(env, argv[3], argv[4], shape, ks, &rank, &total, &same) ||
      !enif_get_int(env, argv[5], &has_noise) || !get_f64(env, argv[6], &noise))
    return enif_make_badarg(env);
  if (!same) return mk_error(env, NXS_ESHAPE);
  if (!mul_ok(total, is_f64 ? sizeof(double) : sizeof(float), &tb) || (int64_t)t.size != tb)
    return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM ot;
  void* out = enif_make_new_binary(env, (size_t)tb, &ot);
  LOCK(r);
  int rc = nxs_wiener_host(r->ctx, t.data, is_f64, (int)rank, shape, ks, has_noise, noise, out);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, ot);
}

/* ---- argrelextrema(ctx, data_bin, shape, axis, order, comparator)
 * NxSignal.PeakFinding.argrelmin/argrelmax/argrelextrema, lib/nx_signal/peak_finding.ex:131-391
 * -> {:ok, indices_bin (s32 [prod(shape)][rank], -1 padded), valid_count} */
static ERL_NIF_TERM argrelextrema(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary d;
  int64_t shape[8], total = 1, db, ib, valid = 0;
  unsigned rank = 0;
  int axis, order, cmp;
  if (argc != 6 || !get_ctx(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &d) ||
      !get_i64_list(env, argv[2], shape, 8, &rank) || !enif_get_int(env, argv[3], &axis) ||
      !enif_get_int(env, argv[4], &order) || !enif_get_int(env, argv[5], &cmp))
    return enif_make_badarg(env);
  if (rank < 1) return mk_error(env, NXS_ESHAPE);
  for (unsigned i = 0; i < rank; ++i)
    if (shape[i] < 1 || !mul_ok(total, shape[i], &total)) return mk_error(env, NXS_ESHAPE);
  if (!mul_ok(total, sizeof(float), &db) || (int64_t)d.size != db) return mk_error(env, NXS_ESHAPE);
  if (!mul_ok(total, (int64_t)rank * (int64_t)sizeof(int32_t), &ib)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM it;
  int32_t* idx = (int32_t*)enif_make_new_binary(env, (size_t)ib, &it);
  LOCK(r);
  int rc = nxs_argrelextrema_f32_host(r->ctx, (const float*)d.data, (int)rank, shape, axis, order, cmp, idx, &valid);
  UNLOCK(r);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), it, enif_make_int64(env, valid));
}

/* ---- host-side closed forms (no context, no GPU) --------------------------------------------- */
/* window(kind, n, periodic, beta, eps) -- NxSignal.Windows.*, lib/nx_signal/windows.ex:33-341 (f32) */
static ERL_NIF_TERM window(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int kind, periodic;
  int64_t n, nb;
  double beta, eps;
  if (argc != 5 || !enif_get_int(env, argv[0], &kind) || !get_i64(env, argv[1], &n) ||
      !enif_get_int(env, argv[2], &periodic) || !get_f64(env, argv[3], &beta) || !get_f64(env, argv[4], &eps))
    return enif_make_badarg(env);
  if (!bytes3(n, 1, 1, sizeof(float), &nb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM t;
  float* out = (float*)enif_make_new_binary(env, (size_t)nb, &t);
  int rc = nxs_window_f32(kind, n, periodic, beta, eps, out);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, t);
}

/* firwin(num_taps, cutoffs, window_kind, beta, pass_zero, scale, sampling_rate)
 * NxSignal.Filters.firwin/3, lib/nx_signal/filters.ex:147-279 */
static ERL_NIF_TERM firwin(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int64_t taps, nb;
  double cut[64], beta, sr;
  int kind, pass_zero, scale;
  unsigned ncut = 0;
  ERL_NIF_TERM head, tail;
  if (argc != 7 || !get_i64(env, argv[0], &taps) || !enif_get_list_length(env, argv[1], &ncut) || ncut < 1 ||
      ncut > 64 || !enif_get_int(env, argv[2], &kind) || !get_f64(env, argv[3], &beta) ||
      !enif_get_int(env, argv[4], &pass_zero) || !enif_get_int(env, argv[5], &scale) || !get_f64(env, argv[6], &sr))
    return enif_make_badarg(env);
  tail = argv[1];
  for (unsigned i = 0; i < ncut; ++i)
    if (!enif_get_list_cell(env, tail, &head, &tail) || !get_f64(env, head, &cut[i])) return enif_make_badarg(env);
  if (!bytes3(taps, 1, 1, sizeof(float), &nb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM t;
  float* out = (float*)enif_make_new_binary(env, (size_t)nb, &t);
  int rc = nxs_firwin_f32(taps, cut, (int)ncut, kind, beta, pass_zero, scale, sr, out);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, t);
}

/* fft_frequencies(sampling_rate, fft_length) -- NxSignal.fft_frequencies/2, lib/nx_signal.ex:154-166 */
static ERL_NIF_TERM fft_frequencies(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  double sr;
  int64_t n, nb;
  if (argc != 2 || !get_f64(env, argv[0], &sr) || !get_i64(env, argv[1], &n)) return enif_make_badarg(env);
  if (!bytes3(n, 1, 1, sizeof(float), &nb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM t;
  float* out = (float*)enif_make_new_binary(env, (size_t)nb, &t);
  int rc = nxs_fft_frequencies_f32(sr, n, out);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, t);
}

/* mel_filters(fft_length, mel_bins, sampling_rate, max_mel, mel_frequency_spacing)
 * NxSignal.mel_filters/4, lib/nx_signal.ex:397-445 -> [mel_bins][fft_length] f32 */
static ERL_NIF_TERM mel_filters(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int64_t nfft, mels, nb;
  double sr, max_mel, f_sp;
  if (argc != 5 || !get_i64(env, argv[0], &nfft) || !get_i64(env, argv[1], &mels) || !get_f64(env, argv[2], &sr) ||
      !get_f64(env, argv[3], &max_mel) || !get_f64(env, argv[4], &f_sp))
    return enif_make_badarg(env);
  if (!bytes3(mels, nfft, 1, sizeof(float), &nb)) return mk_error(env, NXS_ESHAPE);
  ERL_NIF_TERM t;
  float* out = (float*)enif_make_new_binary(env, (size_t)nb, &t);
  int rc = nxs_mel_filters_f32(nfft, mels, sr, max_mel, f_sp, out);
  if (rc) return mk_error(env, rc);
  return mk_ok1(env, t);
}

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
  (void)priv;
  (void)info;
  CTX_TYPE = enif_open_resource_type(env, NULL, "nxs_ctx", ctx_dtor, ERL_NIF_RT_CREATE, NULL);
  return CTX_TYPE ? 0 : 1;
}

static ErlNifFunc funcs[] = {
    {"ctx_create", 1, ctx_create, 0},
    {"stft", 12, stft, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"stft_c64", 12, stft_c64, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"istft", 10, istft, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"istft_c2r", 10, istft_c2r, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"fir", 6, fir, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"convolve_nd", 7, convolve_nd, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"as_windowed", 10, as_windowed, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"overlap_and_add", 7, overlap_and_add, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"stft_to_mel", 10, stft_to_mel, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"stft_mel", 15, stft_mel, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"median", 4, median, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"wiener", 7, wiener, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"argrelextrema", 6, argrelextrema, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"window", 5, window, 0},
    {"firwin", 7, firwin, 0},
    {"fft_frequencies", 2, fft_frequencies, 0},
    {"mel_filters", 5, mel_filters, 0},
};

ERL_NIF_INIT(Elixir.NxSignalB200.NIF, funcs, load, NULL, NULL, NULL)
