// nxs_hostio.cu -- nxs_stft_f32_host: NxSignal.stft/3 (lib/nx_signal.ex:68-130) on HOST buffers,
// the call a NIF makes and what bench.py's `e2e` measures.
//
// The result is 8x the input (nfft/hop x 2 floats per sample), so the call is bound by the
// device->host leg.  Three things shorten it:
//   1. a channel-chunk pipeline on three streams (H2D | kernels | D2H in slabs of a few MiB);
//   2. bins nfft/2+1 .. nfft-1 of a real signal's spectrum are the exact conjugate mirror of bins
//      nfft/2-1 .. 1, so only bins 0 .. nfft/2 cross PCIe and host threads write the other half
//      (a bit copy with one sign flip) behind a gate that rises as slabs land;
//   3. caller memory the DMA engines cannot address (pageable -- a BEAM binary) goes through pinned
//      ring slots owned by the context: input chunks are staged by the host threads, result slabs
//      land in the ring and the same pass that copies them out writes the mirror half.
// Whether (2) pays depends on what the box is short of: PCIe (one GPU per host: mirror wins) or
// host memory bandwidth (many ranks per host: the mirror's extra read + write loses).  The context
// therefore keeps a per-mode cost estimate (seconds per result byte, measured on its own calls)
// and picks the cheaper mode, re-probing the other one now and then; nxs_ctx_set_host_mode pins it.
#include <stdlib.h>

#include "nxs_hostio.cuh"

namespace nxs {

namespace {

enum { kIn = 0, kOut = 1 };

struct Item {
  int kind;
  int group;       // kIn: chunk index; kOut: slab index
  int64_t a, b;    // kIn: piece [a, b) in bytes of the chunk; kOut: rows [a, b)
};

struct StftHostJob {
  // result
  float* z = nullptr;
  int64_t nfft = 0, nout = 0;
  bool mirror = false, unstage = false;
  const float* ring = nullptr;       // pinned result ring: nslots slots of slot_rows rows, ring_pitch complex apart
  int64_t ring_pitch = 0, slot_rows = 0;
  int nslots = 0;
  std::vector<int64_t> slab_r0;      // first row of every slab
  // input staging
  const char* x = nullptr;           // caller's rows (bytes)
  char* hx = nullptr;                // pinned input ring: 2 slots of in_slot bytes
  size_t in_slot = 0;
  std::vector<size_t> chunk_off;     // byte offset of every chunk in x
  // progress
  std::vector<Item> items;
  std::vector<std::atomic<int>> in_left, slab_left;
  std::atomic<int64_t> landed{0};    // slabs whose D2H has completed
  std::atomic<int64_t> h2d_done{0};  // chunks whose H2D has completed (their pinned input slot is free again)
  const std::atomic<bool>* abort = nullptr;

  StftHostJob(size_t nchunks, size_t nslabs) : in_left(nchunks), slab_left(nslabs) {}

  bool wait(const std::atomic<int64_t>& v, int64_t above) const {
    int spins = 0;
    while (v.load(std::memory_order_acquire) <= above) {
      if (abort->load(std::memory_order_relaxed)) return false;
      cpu_relax();
      if (++spins == 4096) {
        spins = 0;
        sched_yield();
      }
    }
    return true;
  }

  static void run(void* p, int64_t i) {
    StftHostJob& j = *static_cast<StftHostJob*>(p);
    const Item& it = j.items[(size_t)i];
    if (it.kind == kIn) {
      // the slot was last used by chunk group - 2
      if (it.group >= 2 && !j.wait(j.h2d_done, it.group - 2)) return;
      memcpy(j.hx + (it.group & 1) * j.in_slot + it.a, j.x + j.chunk_off[it.group] + it.a, size_t(it.b - it.a));
      j.in_left[it.group].fetch_sub(1, std::memory_order_release);
    } else {
      if (!j.wait(j.landed, it.group)) return;
      if (j.unstage) {
        const int64_t r0 = j.slab_r0[it.group];
        const float* src = j.ring + 2 * (int64_t(it.group % j.nslots) * j.slot_rows + (it.a - r0)) * j.ring_pitch;
        unstage_rows_c64(j.z, j.nfft, j.nout, src, j.ring_pitch, it.a, it.b, j.mirror);
      } else {
        mirror_rows_c64(j.z, j.nfft, it.a, it.b);
      }
      j.slab_left[it.group].fetch_sub(1, std::memory_order_release);
    }
  }
};

// Transfer modes of a pinned result: 0 = both spectrum halves over PCIe, 1 = lower half + host mirror,
// 3 = mixed (three chunks of four as in 1, the fourth as in 0).  Mode 1 halves the PCIe bytes but adds a read
// and a write of host memory per mirrored row, mode 0 the reverse.  The automatic choice is between 0 and 1
// (seconds per result byte of each, exponentially averaged over the context's own calls); the mix can only be
// pinned: where it was measured (one B200 per 16-vCPU guest: 103 ms mirrored, 112 ms mixed, 139 ms full) the
// copy engine, slowed by the host threads' memory traffic, stays the bottleneck, so moving more bytes over it
// to spare host traffic loses.
constexpr size_t kSmallCallBytes = size_t(4) << 20;  // results and inputs up to here take the single-stream path

int choose_mode(nxs_ctx* ctx, bool can_mirror, size_t result_bytes, int64_t nchunks) {
  if (!can_mirror) return 0;
  if (const char* e = getenv("NXS_HOST_NO_MIRROR")) {
    if (e[0] && e[0] != '0') return 0;
  }
  if (ctx->host_mode_forced >= 0) return ctx->host_mode_forced == 3 && nchunks < 4 ? 1 : ctx->host_mode_forced;
  if (result_bytes < (size_t(32) << 20)) return 1;  // too small to measure: one-sided transfer
  if (ctx->host_cost[1] <= 0.0) return 1;           // unknown costs first
  if (ctx->host_cost[0] <= 0.0) return 0;
  const int best = ctx->host_cost[1] <= ctx->host_cost[0] ? 1 : 0;
  // re-probe the other mode every 16th call: what the box is short of changes with its load
  return (++ctx->host_calls % 16 == 0) ? 1 - best : best;
}

}  // namespace
}  // namespace nxs

using namespace nxs;

extern "C" {

int nxs_ctx_set_host_mode(nxs_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 3 || mode == 2) return NXS_EINVAL;
  ctx->host_mode_forced = mode;
  return NXS_OK;
}

int nxs_ctx_host_mode(const nxs_ctx* ctx, int* mode) {
  if (!ctx || !mode) return NXS_EINVAL;
  *mode = ctx->host_mode_last;
  return NXS_OK;
}

int nxs_stft_f32_host(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                      const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                      int pad_mode, int64_t pad_lo, int64_t pad_hi, int scaling, double sampling_rate,
                      float* z) {
  if (!ctx || !x || !window || !z) return NXS_EINVAL;
  PadGeom g;
  int64_t M = 0;
  int rc = stft_check(channels, length, x_ld, frame_length, hop, fft_length, pad_mode, pad_lo, pad_hi,
                      scaling, sampling_rate, &g, &M);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (channels == 0 || M == 0) return NXS_OK;
  StreamOrder stream_order(ctx, ctx->stream);
  const double t_start = wall_seconds();
  for (double& t : ctx->host_t) t = 0.0;

  const int64_t rows = channels * M;
  const size_t result_bytes = size_t(rows) * size_t(fft_length) * sizeof(float2);
  // Small calls (what most NxSignal.stft calls are: BASELINE configs[0] is 184 frames): one stream, both
  // spectrum halves over PCIe, no events, no host threads -- the call costs its four enqueues and one wait.
  // A pageable input goes through the driver's own bounce buffer; a pageable RESULT keeps the ring path below
  // (lower half into the pinned ring, host threads copy it out and mirror it: 142 us against 159 us for the
  // driver's pageable D2H of the 1.5 MB of configs[0]).
  const size_t span_bytes = size_t((channels - 1) * x_ld + length) * sizeof(float);
  const bool z_pinned = host_is_pinned(z), x_pinned = host_is_pinned(x);
  if (result_bytes <= kSmallCallBytes && span_bytes <= kSmallCallBytes && ctx->host_mode_forced < 0 && z_pinned) {
    rc = grow_buf(ctx, &ctx->d_stage_in, &ctx->d_stage_in_bytes, span_bytes + size_t(frame_length) * sizeof(float) + 512, false);
    if (rc) return rc;
    rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, result_bytes + 256, false);
    if (rc) return rc;
    float* d_x = (float*)ctx->d_stage_in;
    float* d_w = (float*)((char*)ctx->d_stage_in + (span_bytes + 255) / 256 * 256);
    float2* d_z = (float2*)ctx->d_stage_out;
    ctx->host_mode_last = 4;
    cudaError_t e = cudaMemcpyAsync(d_x, x, span_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, window, size_t(frame_length) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      rc = launch_stft(ctx, d_x, channels, length, x_ld, d_w, frame_length, hop, fft_length, g, M, scaling, sampling_rate,
                       d_z, fft_length, 0, ctx->stream);
      if (rc == NXS_OK) e = cudaMemcpyAsync(z, d_z, result_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    }
    const cudaError_t es = cudaStreamSynchronize(ctx->stream);  // nothing of ours touches the caller's buffers after return
    if (rc) return rc;
    NXS_CUDA(ctx, e);
    NXS_CUDA(ctx, es);
    ctx->host_t[0] = ctx->host_t[1] = ctx->host_t[2] = ctx->host_t[3] = wall_seconds() - t_start;
    return NXS_OK;
  }
  const bool can_mirror = stft_has_exact_mirror(fft_length);
  const bool unstage = !z_pinned;
  const bool stage_in = !x_pinned;
  // geometry of the two row forms on the device: bins 0 .. nfft/2 (pitch a multiple of 32 bytes) or all bins
  const int64_t nout_m = fft_length / 2 + 1, zld_m = (nout_m + 3) / 4 * 4;
  // channel chunks (~128 MiB of one-sided device result each)
  const size_t chunk_unit = size_t(M) * size_t(can_mirror ? zld_m : fft_length) * sizeof(float2);
  int64_t cc = int64_t((size_t(128) << 20) / (chunk_unit ? chunk_unit : 1));
  if (cc < 1) cc = 1;
  if (cc > channels) cc = channels;
  const int64_t nchunks = (channels + cc - 1) / cc;
  // pageable result: always the one-sided transfer -- the pass that copies a slab out of the ring
  // writes the mirror half too, so it costs no extra read
  const int level = z_pinned ? choose_mode(ctx, can_mirror, result_bytes, nchunks) : (can_mirror ? 1 : 0);
  ctx->host_mode_last = (unstage ? 2 : level) | (stage_in ? 16 : 0);
  std::vector<char> cm((size_t)nchunks);           // chunk i moves the lower half only (and is mirrored on the host)
  std::vector<size_t> dz_off((size_t)nchunks + 1);  // its first row in the device result, in complex elements
  bool mirror = false;                              // any chunk mirrored
  dz_off[0] = 0;
  for (int64_t i = 0; i < nchunks; ++i) {
    cm[(size_t)i] = level == 1 || (level == 3 && (i & 3) != 3);
    mirror = mirror || cm[(size_t)i];
    const int64_t n = channels - i * cc < cc ? channels - i * cc : cc;
    dz_off[(size_t)i + 1] = dz_off[(size_t)i] + size_t(n) * size_t(M) * size_t(cm[(size_t)i] ? zld_m : fft_length);
  }
  const int64_t nout = can_mirror && (unstage || mirror) ? nout_m : fft_length;  // bins per staged row (host items)
  const size_t in_bytes = size_t((channels - 1) * x_ld + length) * sizeof(float);
  rc = grow_buf(ctx, &ctx->d_stage_in, &ctx->d_stage_in_bytes, in_bytes + size_t(frame_length) * sizeof(float) + 512, false);
  if (rc) return rc;
  rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, dz_off[(size_t)nchunks] * sizeof(float2) + 256, false);
  if (rc) return rc;
  if (!ctx->out_stream) NXS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
  float* d_x = (float*)ctx->d_stage_in;
  float* d_w = (float*)((char*)ctx->d_stage_in + (in_bytes + 255) / 256 * 256);
  float2* d_z = (float2*)ctx->d_stage_out;
  float2* hz = reinterpret_cast<float2*>(z);

  // D2H slabs (whole 64-frame work items, a few MiB on the wire: small enough that the host threads read a
  // slab while it is still cache-resident, large enough to keep the copy engine busy)
  size_t slab_bytes = size_t(8) << 20;
  if (const char* e = getenv("NXS_HOST_SLAB_KB")) {
    if (atol(e) > 0) slab_bytes = size_t(atol(e)) << 10;
  }
  auto rows_per_slab = [&](int64_t pitch) {
    const int64_t r = int64_t(slab_bytes / (size_t(pitch) * sizeof(float2)));
    return r < 64 ? int64_t(64) : r / 64 * 64;
  };
  const int64_t slab_rows = rows_per_slab(zld_m);  // the larger of the two: sizes the pinned ring
  std::vector<int64_t> slab_r0, slab_r1;
  std::vector<int> chunk_first_slab((size_t)nchunks + 1, 0);
  for (int64_t i = 0; i < nchunks; ++i) {
    const int64_t c0 = i * cc, n = channels - c0 < cc ? channels - c0 : cc;
    const int64_t per = rows_per_slab(cm[(size_t)i] ? zld_m : fft_length);
    chunk_first_slab[(size_t)i] = (int)slab_r0.size();
    for (int64_t r0 = c0 * M, r_end = (c0 + n) * M; r0 < r_end;) {
      int64_t r1 = (r0 / 64 + per / 64) * 64;  // slabs end on work-item boundaries except at the chunk's end
      if (r1 > r_end) r1 = r_end;
      slab_r0.push_back(r0);
      slab_r1.push_back(r1);
      r0 = r1;
    }
  }
  const size_t nslabs = slab_r0.size();
  chunk_first_slab[(size_t)nchunks] = (int)nslabs;
  rc = ensure_events(ctx, nslabs + 2 * (size_t)nchunks);
  if (rc) return rc;
  cudaEvent_t* ev_slab = ctx->slab_events.data();
  cudaEvent_t* ev_h2d = ev_slab + nslabs;
  cudaEvent_t* ev_krn = ev_h2d + nchunks;
  const int64_t z_ld = unstage && can_mirror ? zld_m : fft_length;  // pitch of the pinned result ring's rows

  // pinned rings for pageable caller memory
  const int nslots = 6;
  const size_t ring_slot = size_t(slab_rows + 64) * size_t(z_ld) * sizeof(float2);
  const size_t in_slot = (size_t((cc - 1) * x_ld + length) * sizeof(float) + 255) / 256 * 256;
  const size_t pin_need = (unstage ? nslots * ring_slot : 0) + (stage_in ? 2 * in_slot : 0);
  if (pin_need) {
    rc = grow_buf(ctx, &ctx->h_pinned, &ctx->h_pinned_bytes, pin_need, true);
    if (rc) return rc;
  }
  char* const h_ring = (char*)ctx->h_pinned;
  char* const h_in = h_ring + (unstage ? nslots * ring_slot : 0);

  // host work: input staging pieces and result items, interleaved in the order the data flows
  // (IN 0, IN 1, OUT 0, IN 2, OUT 1, ...); items are handed out in order and wait for their own data
  const bool host_work = mirror || unstage || stage_in;
  StftHostJob job((size_t)nchunks, nslabs);
  if (host_work) {
    job.z = z;
    job.nfft = fft_length;
    job.nout = nout;
    job.mirror = can_mirror;  // staged rows hold the lower half (unstage), or the item mirrors in place
    job.unstage = unstage;
    job.ring = (const float*)h_ring;
    job.ring_pitch = z_ld;
    job.slot_rows = slab_rows + 64;
    job.nslots = nslots;
    job.slab_r0 = slab_r0;
    job.x = (const char*)x;
    job.hx = h_in;
    job.in_slot = in_slot;
    job.abort = ensure_pool(ctx)->abort_flag();
    auto add_in = [&](int64_t i) {
      const int64_t c0 = i * cc, n = channels - c0 < cc ? channels - c0 : cc;
      const size_t bytes = size_t((n - 1) * x_ld + length) * sizeof(float), piece = size_t(4) << 20;
      int cnt = 0;
      for (size_t o = 0; o < bytes; o += piece, ++cnt)
        job.items.push_back({kIn, (int)i, (int64_t)o, (int64_t)(bytes - o < piece ? bytes : o + piece)});
      job.in_left[(size_t)i].store(cnt, std::memory_order_relaxed);
    };
    auto add_out = [&](int64_t i) {
      for (int s = chunk_first_slab[(size_t)i]; s < chunk_first_slab[(size_t)i + 1]; ++s) {
        int cnt = 0;
        for (int64_t a = slab_r0[(size_t)s]; a < slab_r1[(size_t)s]; a += 64, ++cnt)
          job.items.push_back({kOut, s, a, a + 64 < slab_r1[(size_t)s] ? a + 64 : slab_r1[(size_t)s]});
        job.slab_left[(size_t)s].store(cnt, std::memory_order_relaxed);
      }
    };
    job.chunk_off.resize((size_t)nchunks);
    for (int64_t i = 0; i < nchunks; ++i) job.chunk_off[(size_t)i] = size_t(i * cc * x_ld) * sizeof(float);
    for (int64_t i = 0; i < nchunks + 2; ++i) {
      if (stage_in && i < nchunks) add_in(i);
      if (i >= 2 && (unstage || cm[(size_t)(i - 2)])) add_out(i - 2);
    }
    if (!(mirror || unstage)) job.items.shrink_to_fit();
    ctx->pool->begin((int64_t)job.items.size(), &StftHostJob::run, &job, nullptr);
  }

  // from here on every exit must join the workers
  size_t next_land = 0, enq_slabs = 0;
  int64_t next_h2d = 0, enq_chunks = 0;
  cudaError_t async_err = cudaSuccess;  // a failed copy / kernel surfaces in the event queries: stop waiting for it
  auto pump = [&]() {  // publish completed copies to the workers
    while (next_land < enq_slabs) {
      const cudaError_t q = cudaEventQuery(ev_slab[next_land]);
      if (q != cudaSuccess) {
        if (q != cudaErrorNotReady) async_err = q;
        break;
      }
      if (next_land == 0) ctx->host_t[1] = wall_seconds() - t_start;
      job.landed.store((int64_t)++next_land, std::memory_order_release);
    }
    while (next_h2d < enq_chunks) {
      const cudaError_t q = cudaEventQuery(ev_h2d[next_h2d]);
      if (q != cudaSuccess) {
        if (q != cudaErrorNotReady) async_err = q;
        break;
      }
      job.h2d_done.store(++next_h2d, std::memory_order_release);
    }
  };
  auto spin_until = [&](const std::atomic<int>& left) {
    int spins = 0;
    while (left.load(std::memory_order_acquire) > 0 && async_err == cudaSuccess) {
      pump();
      cpu_relax();
      if (++spins == 256) {
        spins = 0;
        sched_yield();
      }
    }
    return async_err == cudaSuccess;
  };
  auto enqueue = [&]() -> int {
    NXS_CUDA(ctx, cudaMemcpyAsync(d_w, window, size_t(frame_length) * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    for (int64_t i = 0; i < nchunks; ++i) {
      const int64_t c0 = i * cc, n = channels - c0 < cc ? channels - c0 : cc;
      const size_t xb = size_t((n - 1) * x_ld + length) * sizeof(float);
      const float* src = x + c0 * x_ld;
      if (stage_in) {
        if (!spin_until(job.in_left[(size_t)i])) return set_cuda_error(ctx, async_err, "nxs_stft_f32_host (input staging)");
        src = (const float*)(h_in + (i & 1) * in_slot);
      }
      NXS_CUDA(ctx, cudaMemcpyAsync(d_x + c0 * x_ld, src, xb, cudaMemcpyHostToDevice, ctx->copy_stream));
      NXS_CUDA(ctx, cudaEventRecord(ev_h2d[i], ctx->copy_stream));
      enq_chunks = i + 1;
      NXS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[i], 0));
      const bool one = cm[(size_t)i];
      const int64_t pitch = one ? zld_m : fft_length;
      float2* const dz = d_z + dz_off[(size_t)i];
      int rcl = launch_stft(ctx, d_x + c0 * x_ld, n, length, x_ld, d_w, frame_length, hop, fft_length, g, M, scaling,
                            sampling_rate, dz, pitch, one ? 1 : 0, ctx->stream);
      if (rcl) return rcl;
      NXS_CUDA(ctx, cudaEventRecord(ev_krn[i], ctx->stream));
      NXS_CUDA(ctx, cudaStreamWaitEvent(ctx->out_stream, ev_krn[i], 0));
      for (int s = chunk_first_slab[(size_t)i]; s < chunk_first_slab[(size_t)i + 1]; ++s) {
        const int64_t r0 = slab_r0[(size_t)s], r1 = slab_r1[(size_t)s];
        if (unstage) {
          if (s >= nslots && !spin_until(job.slab_left[(size_t)(s - nslots)]))  // the ring slot has been copied out
            return set_cuda_error(ctx, async_err, "nxs_stft_f32_host (result ring)");
          NXS_CUDA(ctx, cudaMemcpyAsync(h_ring + size_t(s % nslots) * ring_slot, dz + size_t(r0 - c0 * M) * pitch,
                                        size_t(r1 - r0) * pitch * sizeof(float2), cudaMemcpyDeviceToHost, ctx->out_stream));
        } else if (one) {
          NXS_CUDA(ctx, cudaMemcpy2DAsync(hz + size_t(r0) * fft_length, size_t(fft_length) * sizeof(float2),
                                          dz + size_t(r0 - c0 * M) * pitch, size_t(pitch) * sizeof(float2),
                                          size_t(nout_m) * sizeof(float2), size_t(r1 - r0), cudaMemcpyDeviceToHost,
                                          ctx->out_stream));
        } else {
          NXS_CUDA(ctx, cudaMemcpyAsync(hz + size_t(r0) * fft_length, dz + size_t(r0 - c0 * M) * fft_length,
                                        size_t(r1 - r0) * fft_length * sizeof(float2), cudaMemcpyDeviceToHost,
                                        ctx->out_stream));
        }
        NXS_CUDA(ctx, cudaEventRecord(ev_slab[s], ctx->out_stream));
        enq_slabs = (size_t)s + 1;
        if (host_work) pump();
      }
    }
    ctx->host_t[0] = wall_seconds() - t_start;  // everything enqueued
    if (host_work) {
      for (; next_land < nslabs;) {
        NXS_CUDA(ctx, cudaEventSynchronize(ev_slab[next_land]));
        pump();
        if (async_err != cudaSuccess) return set_cuda_error(ctx, async_err, "nxs_stft_f32_host (device to host)");
      }
    }
    return NXS_OK;
  };
  rc = enqueue();
  ctx->host_t[2] = wall_seconds() - t_start;  // last slab landed
  if (host_work) ctx->pool->finish(rc != NXS_OK);
  ctx->host_t[3] = wall_seconds() - t_start;  // host threads done
  // nothing of ours may still read or write the caller's buffers when the call returns
  cudaError_t e1 = cudaStreamSynchronize(ctx->out_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->copy_stream);
  if (rc) return rc;
  NXS_CUDA(ctx, e1);
  NXS_CUDA(ctx, e2);
  NXS_CUDA(ctx, e3);
  if (ctx->host_t[1] == 0.0) ctx->host_t[1] = ctx->host_t[2] = wall_seconds() - t_start;
  // feed the mode's cost estimate (pinned results only; large calls only)
  if (z_pinned && can_mirror && result_bytes >= (size_t(32) << 20)) {
    const double cost = (wall_seconds() - t_start) / double(result_bytes);
    double& c = ctx->host_cost[level];
    c = c <= 0.0 ? cost : 0.5 * c + 0.5 * cost;
  }
  return NXS_OK;
}

}  // extern "C"
