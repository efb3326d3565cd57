// nxs_hostio.cuh -- plumbing of the "_host" entry points (what a NIF calls: host pointers in, host
// pointers out, synchronous).  Every _host entry is a software pipeline over independent rows
// (channels): H2D of chunk k+1 | kernels of chunk k | D2H of chunk k-1 on three streams through a
// ring of device slots, so a call costs max(H2D, kernels, D2H) instead of their sum.  Caller
// memory that the DMA engines cannot address (pageable: what enif_make_new_binary or a plain
// malloc gives) is staged through pinned ring slots by the context's host threads instead of by
// the driver's single-threaded bounce buffer.
#pragma once
#include <string.h>

#include <atomic>
#include <chrono>
#include <vector>

#include "nxs_common.cuh"

namespace nxs {

int grow_buf(nxs_ctx* ctx, void** p, size_t* have, size_t need, bool host);

inline double wall_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// can the copy engines address this host pointer directly (cudaMallocHost / cudaHostRegister memory)?
inline bool host_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

inline int ensure_events(nxs_ctx* ctx, size_t n) {
  while (ctx->slab_events.size() < n) {
    cudaEvent_t e = nullptr;
    NXS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->slab_events.push_back(e);
  }
  return NXS_OK;
}

inline HostPool* ensure_pool(nxs_ctx* ctx) {
  // at least one worker besides the calling thread: the staged host calls wait for items only workers run
  if (!ctx->pool) ctx->pool = new HostPool(HostPool::default_threads() < 2 ? 2 : HostPool::default_threads());
  return ctx->pool;
}

// multi-threaded memcpy of a list of pieces (pageable <-> pinned staging)
struct CopyPiece {
  void* dst;
  const void* src;
  size_t bytes;
};
inline void parallel_copy(nxs_ctx* ctx, std::vector<CopyPiece>& pieces) {
  if (pieces.empty()) return;
  ensure_pool(ctx)->parallel_for((int64_t)pieces.size(), [](void* p, int64_t i) {
    const CopyPiece& c = (*static_cast<std::vector<CopyPiece>*>(p))[(size_t)i];
    memcpy(c.dst, c.src, c.bytes);
  }, &pieces);
}
// rows [0, n) of `row_bytes` at pitches sp / dp, cut into pieces of at most 2 MiB
inline void add_row_pieces(std::vector<CopyPiece>& v, char* dst, size_t dp, const char* src, size_t sp,
                           size_t row_bytes, int64_t n) {
  const size_t kPiece = size_t(2) << 20;
  if (dp == row_bytes && sp == row_bytes) {  // contiguous block
    const size_t total = row_bytes * (size_t)n;
    for (size_t o = 0; o < total; o += kPiece) v.push_back({dst + o, src + o, total - o < kPiece ? total - o : kPiece});
    return;
  }
  for (int64_t r = 0; r < n; ++r)
    for (size_t o = 0; o < row_bytes; o += kPiece)
      v.push_back({dst + r * dp + o, src + r * sp + o, row_bytes - o < kPiece ? row_bytes - o : kPiece});
}

// one _host call over `rows` independent rows
struct PipeSpec {
  const void* in = nullptr;   // [rows] rows of in_row_bytes at stride in_pitch (bytes)
  size_t in_row_bytes = 0, in_pitch = 0;
  void* out = nullptr;        // [rows] rows of out_row_bytes at stride out_pitch
  size_t out_row_bytes = 0, out_pitch = 0;
  int64_t rows = 0;
  const void* aux = nullptr;  // small second operand (window / taps): copied once, before the pipeline
  size_t aux_bytes = 0;
  size_t chunk_target = size_t(48) << 20;  // bytes of the larger side per chunk
};

// launch(row0, nrows, d_in, d_aux, d_out) enqueues the kernels of rows [row0, row0 + nrows) on
// ctx->stream; device rows keep the host pitches.  Returns NXS_*.
template <class F>
int host_pipeline(nxs_ctx* ctx, const PipeSpec& s, F&& launch) {
  if (s.rows <= 0) return NXS_OK;
  constexpr int NS = 3;  // device ring slots
  const size_t big = s.in_pitch > s.out_pitch ? s.in_pitch : s.out_pitch;
  int64_t per = (int64_t)(s.chunk_target / (big ? big : 1));
  if (per < 1) per = 1;
  const bool small_call = size_t(s.rows) * big <= (size_t(4) << 20);  // not worth a pipeline: one chunk, one stream
  if (s.rows >= 4 && per > (s.rows + 3) / 4 && !small_call) per = (s.rows + 3) / 4;  // at least four chunks when there are four rows
  if (per > s.rows) per = s.rows;
  const int64_t nchunks = (s.rows + per - 1) / per;
  auto span = [](int64_t n, size_t pitch, size_t row) { return n > 0 ? size_t(n - 1) * pitch + row : size_t(0); };
  const size_t in_slot = (span(per, s.in_pitch, s.in_row_bytes) + 255) / 256 * 256;
  const size_t out_slot = (span(per, s.out_pitch, s.out_row_bytes) + 255) / 256 * 256;
  const size_t aux_pad = (s.aux_bytes + 255) / 256 * 256;
  // a single small chunk: everything on the compute stream (no cross-stream hand-offs), and pageable memory
  // through the driver's own bounce buffer, which beats waking the host threads below about a MiB
  const bool single = nchunks == 1;
  const size_t kDriverStaged = size_t(1) << 20;
  cudaStream_t const s_in = single ? ctx->stream : ctx->copy_stream;
  const bool in_pg = s.in_row_bytes && !(single && in_slot <= kDriverStaged) && !host_is_pinned(s.in);
  const bool out_pg = s.out_row_bytes && !(single && out_slot <= kDriverStaged) && !host_is_pinned(s.out);

  int rc = grow_buf(ctx, &ctx->d_stage_in, &ctx->d_stage_in_bytes, NS * in_slot + aux_pad + 256, false);
  if (rc) return rc;
  rc = grow_buf(ctx, &ctx->d_stage_out, &ctx->d_stage_out_bytes, NS * out_slot + 256, false);
  if (rc) return rc;
  const size_t pin_need = (in_pg ? 2 * in_slot : 0) + (out_pg ? 2 * out_slot : 0);
  if (pin_need) {
    rc = grow_buf(ctx, &ctx->h_pinned, &ctx->h_pinned_bytes, pin_need, true);
    if (rc) return rc;
  }
  if (!ctx->out_stream) NXS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
  cudaStream_t const s_out = single ? ctx->stream : ctx->out_stream;
  rc = ensure_events(ctx, 3 * NS);
  if (rc) return rc;
  char* const d_in = (char*)ctx->d_stage_in;
  char* const d_aux = d_in + NS * in_slot;
  char* const d_out = (char*)ctx->d_stage_out;
  char* const h_in = (char*)ctx->h_pinned;
  char* const h_out = h_in + (in_pg ? 2 * in_slot : 0);
  cudaEvent_t* ev_h2d = ctx->slab_events.data();
  cudaEvent_t* ev_krn = ev_h2d + NS;
  cudaEvent_t* ev_d2h = ev_krn + NS;

  auto fail = [&](int code) {  // nothing of ours may touch the caller's buffers after return
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->out_stream);
    return code;
  };
  auto cuda_fail = [&](cudaError_t e, const char* what) { return fail(set_cuda_error(ctx, e, what)); };
#define NXS_PIPE_CUDA(call)                                   \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
  } while (0)

  if (s.aux_bytes) NXS_PIPE_CUDA(cudaMemcpyAsync(d_aux, s.aux, s.aux_bytes, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<CopyPiece> pieces;
  auto unstage = [&](int64_t k) -> cudaError_t {  // chunk k: pinned out slot -> the caller's pageable rows
    const int64_t r0 = k * per, n = s.rows - r0 < per ? s.rows - r0 : per;
    cudaError_t e = cudaEventSynchronize(ev_d2h[k % NS]);
    if (e != cudaSuccess) return e;
    pieces.clear();
    add_row_pieces(pieces, (char*)s.out + r0 * s.out_pitch, s.out_pitch, h_out + (k & 1) * out_slot, s.out_pitch,
                   s.out_row_bytes, n);
    parallel_copy(ctx, pieces);
    return cudaSuccess;
  };

  for (int64_t k = 0; k < nchunks; ++k) {
    const int slot = (int)(k % NS);
    const int64_t r0 = k * per, n = s.rows - r0 < per ? s.rows - r0 : per;
    char* di = d_in + slot * in_slot;
    char* dout = d_out + slot * out_slot;
    // H2D of chunk k (its device slot is free once the kernels of chunk k - NS have run)
    if (s.in_row_bytes) {
      if (k >= NS) NXS_PIPE_CUDA(cudaStreamWaitEvent(s_in, ev_krn[slot], 0));
      const char* src = (const char*)s.in + r0 * s.in_pitch;
      if (in_pg) {
        char* hs = h_in + (k & 1) * in_slot;
        if (k >= 2) NXS_PIPE_CUDA(cudaEventSynchronize(ev_h2d[(k - 2) % NS]));  // the pinned slot has been read
        pieces.clear();
        add_row_pieces(pieces, hs, s.in_pitch, src, s.in_pitch, s.in_row_bytes, n);
        parallel_copy(ctx, pieces);
        src = hs;
      }
      NXS_PIPE_CUDA(cudaMemcpyAsync(di, src, span(n, s.in_pitch, s.in_row_bytes), cudaMemcpyHostToDevice, s_in));
      if (!single) {
        NXS_PIPE_CUDA(cudaEventRecord(ev_h2d[slot], s_in));
        NXS_PIPE_CUDA(cudaStreamWaitEvent(ctx->stream, ev_h2d[slot], 0));
      }
    }
    // kernels (the out slot is free once the D2H of chunk k - NS has drained it)
    if (k >= NS && s.out_row_bytes) NXS_PIPE_CUDA(cudaStreamWaitEvent(ctx->stream, ev_d2h[slot], 0));
    rc = launch(r0, n, (void*)di, (void*)d_aux, (void*)dout);
    if (rc) return fail(rc);
    if (!single) NXS_PIPE_CUDA(cudaEventRecord(ev_krn[slot], ctx->stream));
    // D2H
    if (s.out_row_bytes) {
      if (!single) NXS_PIPE_CUDA(cudaStreamWaitEvent(s_out, ev_krn[slot], 0));
      char* dst = out_pg ? h_out + (k & 1) * out_slot : (char*)s.out + r0 * s.out_pitch;
      if (s.out_pitch == s.out_row_bytes || out_pg)
        NXS_PIPE_CUDA(cudaMemcpyAsync(dst, dout, span(n, s.out_pitch, s.out_row_bytes), cudaMemcpyDeviceToHost, s_out));
      else  // never write the caller's bytes between rows
        NXS_PIPE_CUDA(cudaMemcpy2DAsync(dst, s.out_pitch, dout, s.out_pitch, s.out_row_bytes, (size_t)n,
                                        cudaMemcpyDeviceToHost, s_out));
      if (!single || out_pg) NXS_PIPE_CUDA(cudaEventRecord(ev_d2h[slot], s_out));
      if (out_pg && k >= 1) NXS_PIPE_CUDA(unstage(k - 1));
    }
  }
  if (out_pg) NXS_PIPE_CUDA(unstage(nchunks - 1));
  if (!single) NXS_PIPE_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  NXS_PIPE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!single) NXS_PIPE_CUDA(cudaStreamSynchronize(ctx->out_stream));
#undef NXS_PIPE_CUDA
  return NXS_OK;
}

}  // namespace nxs
