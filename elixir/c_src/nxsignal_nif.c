/*
 * nxsignal_nif.c -- source-only NIF shim over include/nxsignal_b200.h.
 *
 * NOT compiled in this repository's environment (no erl_nif.h, no BEAM); it is the binding a
 * maintainer of elixir-nx/nx_signal would add so that NxSignal.stft/3 (lib/nx_signal.ex:68),
 * NxSignal.istft/3 (:582) and Convolution.convolve/3 for the FIR form (lib/nx_signal/convolution.ex:38)
 * run on the B200 backend.  Build (on a machine with Erlang/OTP 27 + CUDA):
 *   cc -O2 -fPIC -shared -I$ERL_ROOT/usr/include -I../../include nxsignal_nif.c \
 *      -L../../nx_signal_b200/lib -lnxsignal_b200 -o priv/nxsignal_nif.so
 *
 * Tensors cross as Nx.to_binary/1 binaries (row-major, native endian, c64 = interleaved f32).
 * Results are allocated with enif_make_new_binary so the VM owns them.  Every entry is a dirty
 * IO-bound NIF: calls block on PCIe transfers for milliseconds.
 */
#include <erl_nif.h>
#include <string.h>

#include "nxsignal_b200.h"

static ErlNifResourceType* CTX_TYPE;

typedef struct { nxs_ctx* ctx; } ctx_res;

static void ctx_dtor(ErlNifEnv* env, void* obj) { (void)env; nxs_ctx_destroy(((ctx_res*)obj)->ctx); }

static ERL_NIF_TERM mk_error(ErlNifEnv* env, int rc) {
  /* NXS_EINVAL / NXS_ESHAPE -> :argument_error (raised as ArgumentError by the Elixir side) */
  const char* kind = (rc == NXS_EINVAL || rc == NXS_ESHAPE) ? "argument_error" : "runtime_error";
  return enif_make_tuple3(env, enif_make_atom(env, "error"), enif_make_atom(env, kind),
                          enif_make_string(env, nxs_strerror(rc), ERL_NIF_LATIN1));
}

static ERL_NIF_TERM ctx_create(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int dev;
  if (argc != 1 || !enif_get_int(env, argv[0], &dev)) return enif_make_badarg(env);
  nxs_ctx* c = NULL;
  int rc = nxs_ctx_create(dev, &c);
  if (rc) return mk_error(env, rc);
  ctx_res* r = enif_alloc_resource(CTX_TYPE, sizeof(ctx_res));
  r->ctx = c;
  ERL_NIF_TERM t = enif_make_resource(env, r);
  enif_release_resource(r);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), t);
}

/* stft(ctx, x_bin, channels, length, window_bin, hop, fft_length, pad_mode, pad_lo, pad_hi, scaling, sr) */
static ERL_NIF_TERM stft(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x, w;
  ErlNifSInt64 ch, len, hop, nfft, lo, hi;
  int pad, scal;
  double sr;
  if (argc != 12 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) ||
      !enif_inspect_binary(env, argv[1], &x) || !enif_get_int64(env, argv[2], &ch) ||
      !enif_get_int64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &w) ||
      !enif_get_int64(env, argv[5], &hop) || !enif_get_int64(env, argv[6], &nfft) ||
      !enif_get_int(env, argv[7], &pad) || !enif_get_int64(env, argv[8], &lo) ||
      !enif_get_int64(env, argv[9], &hi) || !enif_get_int(env, argv[10], &scal) ||
      !enif_get_double(env, argv[11], &sr))
    return enif_make_badarg(env);
  const ErlNifSInt64 n = (ErlNifSInt64)(w.size / sizeof(float));
  if ((ErlNifSInt64)x.size != ch * len * (ErlNifSInt64)sizeof(float)) return mk_error(env, NXS_ESHAPE);
  int64_t frames = 0;
  int rc = nxs_num_frames(len, n, hop, pad, lo, hi, &frames);
  if (rc) return mk_error(env, rc);
  ERL_NIF_TERM zt, tt, ft;
  float* z = (float*)enif_make_new_binary(env, (size_t)(ch * frames * nfft) * 2 * sizeof(float), &zt);
  float* times = (float*)enif_make_new_binary(env, (size_t)frames * sizeof(float), &tt);
  float* freqs = (float*)enif_make_new_binary(env, (size_t)nfft * sizeof(float), &ft);
  rc = nxs_stft_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)w.data, n, hop, nfft, pad, lo,
                         hi, scal, sr, z);
  if (!rc) rc = nxs_stft_times_f32(n, sr, frames, times);
  if (!rc) rc = nxs_fft_frequencies_f32(sr, nfft, freqs);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple5(env, enif_make_atom(env, "ok"), zt, tt, ft, enif_make_int64(env, frames));
}

/* istft(ctx, z_bin, channels, frames, z_len, window_bin, hop, fft_length, scaling, sr) */
static ERL_NIF_TERM istft(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary z, w;
  ErlNifSInt64 ch, frames, zlen, hop, nfft;
  int scal;
  double sr;
  if (argc != 10 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) ||
      !enif_inspect_binary(env, argv[1], &z) || !enif_get_int64(env, argv[2], &ch) ||
      !enif_get_int64(env, argv[3], &frames) || !enif_get_int64(env, argv[4], &zlen) ||
      !enif_inspect_binary(env, argv[5], &w) || !enif_get_int64(env, argv[6], &hop) ||
      !enif_get_int64(env, argv[7], &nfft) || !enif_get_int(env, argv[8], &scal) ||
      !enif_get_double(env, argv[9], &sr))
    return enif_make_badarg(env);
  const ErlNifSInt64 n = (ErlNifSInt64)(w.size / sizeof(float));
  const ErlNifSInt64 out_len = frames * hop + (n - hop);
  ERL_NIF_TERM yt;
  float* y = (float*)enif_make_new_binary(env, (size_t)(ch * out_len) * 2 * sizeof(float), &yt);
  int rc = nxs_istft_c64_host(r->ctx, (const float*)z.data, ch, frames, zlen, (const float*)w.data, n, hop, nfft,
                              scal, sr, y);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), yt);
}

/* fir(ctx, x_bin, channels, length, taps_bin, mode) */
static ERL_NIF_TERM fir(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x, h;
  ErlNifSInt64 ch, len;
  int mode;
  if (argc != 6 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) ||
      !enif_inspect_binary(env, argv[1], &x) || !enif_get_int64(env, argv[2], &ch) ||
      !enif_get_int64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &h) ||
      !enif_get_int(env, argv[5], &mode))
    return enif_make_badarg(env);
  const ErlNifSInt64 k = (ErlNifSInt64)(h.size / sizeof(float));
  int64_t out_len = 0;
  int rc = nxs_fir_out_len(len, k, mode, &out_len);
  if (rc) return mk_error(env, rc);
  ERL_NIF_TERM yt;
  float* y = (float*)enif_make_new_binary(env, (size_t)(ch * out_len) * sizeof(float), &yt);
  rc = nxs_fir_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)h.data, k, mode, y, out_len);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), yt, enif_make_int64(env, out_len));
}

/* stft_to_mel(ctx, z_bin, channels, frames, z_len, fft_length, mel_bins, sr, max_mel, f_sp)
 * -- NxSignal.stft_to_mel/3, lib/nx_signal.ex:486-513 */
static ERL_NIF_TERM stft_to_mel(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary z;
  ErlNifSInt64 ch, frames, zlen, nfft, mels;
  double sr, max_mel, f_sp;
  if (argc != 10 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) ||
      !enif_inspect_binary(env, argv[1], &z) || !enif_get_int64(env, argv[2], &ch) ||
      !enif_get_int64(env, argv[3], &frames) || !enif_get_int64(env, argv[4], &zlen) ||
      !enif_get_int64(env, argv[5], &nfft) || !enif_get_int64(env, argv[6], &mels) ||
      !enif_get_double(env, argv[7], &sr) || !enif_get_double(env, argv[8], &max_mel) ||
      !enif_get_double(env, argv[9], &f_sp))
    return enif_make_badarg(env);
  if (z.size < (size_t)(ch * frames * zlen) * 2 * sizeof(float)) return enif_make_badarg(env);
  ERL_NIF_TERM mt;
  float* mel = (float*)enif_make_new_binary(env, (size_t)(ch * frames * mels) * sizeof(float), &mt);
  int rc = nxs_stft_to_mel_f32_host(r->ctx, (const float*)z.data, ch, frames, zlen, nfft, mels, sr, max_mel, f_sp, mel);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), mt);
}

/* stft_mel(ctx, x_bin, channels, length, window_bin, hop, fft_length, pad_mode, lo, hi, scaling, sr,
 *          mel_bins, max_mel, f_sp) -- NxSignal.stft/3 |> NxSignal.stft_to_mel/3 in one call
 * (lib/nx_signal.ex:68-130, 486-513): only the [frames][mel_bins] tensor comes back over PCIe */
static ERL_NIF_TERM stft_mel(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary x, w;
  ErlNifSInt64 ch, len, hop, nfft, lo, hi, mels;
  int pad, scal;
  double sr, max_mel, f_sp;
  if (argc != 15 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) ||
      !enif_inspect_binary(env, argv[1], &x) || !enif_get_int64(env, argv[2], &ch) ||
      !enif_get_int64(env, argv[3], &len) || !enif_inspect_binary(env, argv[4], &w) ||
      !enif_get_int64(env, argv[5], &hop) || !enif_get_int64(env, argv[6], &nfft) ||
      !enif_get_int(env, argv[7], &pad) || !enif_get_int64(env, argv[8], &lo) ||
      !enif_get_int64(env, argv[9], &hi) || !enif_get_int(env, argv[10], &scal) ||
      !enif_get_double(env, argv[11], &sr) || !enif_get_int64(env, argv[12], &mels) ||
      !enif_get_double(env, argv[13], &max_mel) || !enif_get_double(env, argv[14], &f_sp))
    return enif_make_badarg(env);
  const ErlNifSInt64 n = (ErlNifSInt64)(w.size / sizeof(float));
  int64_t frames = 0;
  int rc = nxs_num_frames(len, n, hop, pad, lo, hi, &frames);
  if (rc) return mk_error(env, rc);
  ERL_NIF_TERM mt;
  float* mel = (float*)enif_make_new_binary(env, (size_t)(ch * frames * mels) * sizeof(float), &mt);
  rc = nxs_stft_mel_f32_host(r->ctx, (const float*)x.data, ch, len, len, (const float*)w.data, n, hop, nfft, pad, lo,
                             hi, scal, sr, mels, max_mel, f_sp, mel);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple3(env, enif_make_atom(env, "ok"), mt, enif_make_int64(env, frames));
}

/* median(ctx, t_bin, shape_list, kernel_shape_list) -- NxSignal.Filters.median/2, lib/nx_signal/filters.ex:17-56 */
static ERL_NIF_TERM median(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  ctx_res* r;
  ErlNifBinary t;
  int64_t shape[3], ks[3];
  unsigned rank = 0, krank = 0;
  ERL_NIF_TERM head, tail;
  if (argc != 4 || !enif_get_resource(env, argv[0], CTX_TYPE, (void**)&r) || !enif_inspect_binary(env, argv[1], &t) ||
      !enif_get_list_length(env, argv[2], &rank) || !enif_get_list_length(env, argv[3], &krank) || rank > 3)
    return enif_make_badarg(env);
  if (rank != krank) return mk_error(env, NXS_ESHAPE); /* "kernel shape must be of the same rank as the tensor" */
  size_t total = 1;
  tail = argv[2];
  for (unsigned i = 0; i < rank; ++i) {
    ErlNifSInt64 v;
    if (!enif_get_list_cell(env, tail, &head, &tail) || !enif_get_int64(env, head, &v)) return enif_make_badarg(env);
    shape[i] = v;
    total *= (size_t)v;
  }
  tail = argv[3];
  for (unsigned i = 0; i < rank; ++i) {
    ErlNifSInt64 v;
    if (!enif_get_list_cell(env, tail, &head, &tail) || !enif_get_int64(env, head, &v)) return enif_make_badarg(env);
    ks[i] = v;
  }
  if (t.size < total * sizeof(float)) return enif_make_badarg(env);
  ERL_NIF_TERM ot;
  float* out = (float*)enif_make_new_binary(env, total * sizeof(float), &ot);
  int rc = nxs_median_f32_host(r->ctx, (const float*)t.data, (int)rank, shape, ks, out);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), ot);
}

/* window(kind, n, periodic, beta, eps) -- host-side, bit-compatible with Nx.BinaryBackend */
static ERL_NIF_TERM window(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
  int kind, periodic;
  ErlNifSInt64 n;
  double beta, eps;
  if (argc != 5 || !enif_get_int(env, argv[0], &kind) || !enif_get_int64(env, argv[1], &n) ||
      !enif_get_int(env, argv[2], &periodic) || !enif_get_double(env, argv[3], &beta) ||
      !enif_get_double(env, argv[4], &eps))
    return enif_make_badarg(env);
  ERL_NIF_TERM t;
  float* out = (float*)enif_make_new_binary(env, (size_t)n * sizeof(float), &t);
  int rc = nxs_window_f32(kind, n, periodic, beta, eps, out);
  if (rc) return mk_error(env, rc);
  return enif_make_tuple2(env, enif_make_atom(env, "ok"), t);
}

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
  (void)priv; (void)info;
  CTX_TYPE = enif_open_resource_type(env, NULL, "nxs_ctx", ctx_dtor, ERL_NIF_RT_CREATE, NULL);
  return CTX_TYPE ? 0 : 1;
}

static ErlNifFunc funcs[] = {
    {"ctx_create", 1, ctx_create, 0},
    {"stft", 12, stft, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"istft", 10, istft, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"fir", 6, fir, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"stft_to_mel", 10, stft_to_mel, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"stft_mel", 15, stft_mel, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"median", 4, median, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"window", 5, window, 0},
};

ERL_NIF_INIT(Elixir.NxSignalB200.NIF, funcs, load, NULL, NULL, NULL)
