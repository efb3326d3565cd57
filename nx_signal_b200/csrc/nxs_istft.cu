// nxs_istft.cu -- fused inverse STFT for sm_100a.
//
// Replaces NxSignal.istft/3 (lib/nx_signal.ex:582-638): Nx.ifft(length:) (:609) -> rescale
// (:611-625) -> * window (:628) -> overlap_and_add (:627-628, :684-735) -> divide by the
// overlap-added |window|^2 with the `> 1e-10 else 1` guard (:630-637), returning c64.
//
// Kernels, fastest first (launch_istft picks):
//   istft_rola_kernel      hop = N/2, N/4, N/8: a group of T threads walks consecutive frames
//                          and keeps the running overlap-add in registers (no CTA barrier, no
//                          atomics, TMA-staged input) -- the hot path
//   istft_ring_kernel      any hop <= N (at most 64 covering frames): the same per-group,
//                          TMA-staged structure with the running overlap-add in a per-group
//                          ring of N complex values in shared memory
//   istft_kernel           z_len != N / unaligned rows: a CTA owns a segment of frames, every
//                          group inverse-transforms one frame into shared memory and the CTA
//                          gathers the overlap-add (ascending frame order + a carry: deterministic)
//   ifft_frames_kernel /   nfft >= 4096 with an unsupported hop, and generic lengths: frames
//   istft_dft_frames_kernel  to a scratch tensor, then istft_ola_norm_kernel gathers
//   istft_edge_f64_kernel  always last: recomputes the ill-conditioned head / tail samples in
//                          double, as the reference's f64 backend effectively does
//   istft_rola_c2r_kernel  opt-in (nxs_istft_c2r_f32_*): one-sided Hermitian spectrum in, real
//                          signal out; a half-length packed transform per frame (end of file)
// Segments after the first recompute the few frames that overlap their start (warm-up)
// instead of exchanging partial sums.  The normaliser is accumulated the same way from
// |w|^2, which reproduces the reference's edge behaviour (fewer covering frames at both
// ends) exactly.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "nxs_common.cuh"
#include "nxs_fft.cuh"
#include "nxs_tma.cuh"

namespace nxs {

int launch_prep_window(nxs_ctx* ctx, const float* window, int64_t n, int64_t nfft, int scaling,
                       double sampling_rate, float prescale, int invert, float* out, cudaStream_t st);
int get_dft_table(nxs_ctx* ctx, int64_t n, int sign, float2** out);

struct IstftArgs {
  const float2* z;  // [C][M][z_len]
  int64_t M, z_len;
  const float* wprep;  // [nfft] = w * S / nfft
  const float* w;      // [nfft] raw window (for |w|^2)
  float2* y;           // [C][M*hop + nfft - hop]
  int64_t out_len;
  int hop;
  int seg_frames;  // frames per segment (multiple of G)
  int segs_per_channel;
  int total_segs;
  int warm_batches;  // batches recomputed before a segment that does not start at frame 0
  const float2* tw;
};

template <class PL, int THREADS>
struct IstftCfg {
  static constexpr int G = THREADS / PL::T, NFFT = PL::N;
  static constexpr size_t BUF_BYTES = size_t(G) * 2 * PL::BUF * sizeof(cpx);
  static constexpr size_t WIN_OFF = BUF_BYTES;                              // w' [NFFT]
  static constexpr size_t W2_OFF = WIN_OFF + size_t(NFFT) * sizeof(float);  // |w|^2 [NFFT]
  static constexpr size_t CARRY_OFF = W2_OFF + size_t(NFFT) * sizeof(float);
  // carry: 2 x (NFFT complex + NFFT float) (OV <= NFFT - 1 entries used)
  static constexpr size_t CARRY_ONE = size_t(NFFT) * (sizeof(cpx) + sizeof(float));
  static constexpr size_t TW_OFF = CARRY_OFF + 2 * CARRY_ONE;
  static constexpr size_t SMEM = TW_OFF + size_t(PL::TW_TOTAL) * sizeof(cpx) + 16;
};

template <class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) istft_kernel(const IstftArgs a) {
  using CF = IstftCfg<PL, THREADS>;
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const buf_base = reinterpret_cast<cpx*>(smem_raw);
  cpx* const bufA = buf_base + (size_t)(2 * g) * PL::BUF;
  cpx* const bufB = bufA + PL::BUF;
  float* wsm = reinterpret_cast<float*>(smem_raw + CF::WIN_OFF);
  float* w2sm = reinterpret_cast<float*>(smem_raw + CF::W2_OFF);
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);

  for (int i = tid; i < N; i += THREADS) {
    wsm[i] = a.wprep[i];
    const float w = a.w[i];
    w2sm[i] = (float)((double)fabsf(w) * (double)fabsf(w));  // Nx.abs(window) ** 2, f32
  }
  for (int i = tid; i < PL::TW_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  __syncthreads();
  TwTable<PL> tw;
  tw.init(twsm, t);
  const SyncBlock sync;

  const int hop = a.hop;
  const int OV = N - hop;
  const int span = (G - 1) * hop + N;

  for (int seg = blockIdx.x; seg < a.total_segs; seg += gridDim.x) {
    const int c = seg / a.segs_per_channel;
    const int si = seg - c * a.segs_per_channel;
    const int64_t ms = (int64_t)si * a.seg_frames;
    int64_t me = ms + a.seg_frames;
    if (me > a.M) me = a.M;
    int64_t mb = ms - (int64_t)a.warm_batches * G;
    if (mb < 0) mb = 0;
    const float2* __restrict__ zc = a.z + (int64_t)c * a.M * a.z_len;
    float2* __restrict__ yc = a.y + (int64_t)c * a.out_len;

    int cur = 0;  // carry buffer in use
    {
      cpx* cc = reinterpret_cast<cpx*>(smem_raw + CF::CARRY_OFF);
      float* cn = reinterpret_cast<float*>(smem_raw + CF::CARRY_OFF + size_t(N) * sizeof(cpx));
      for (int i = tid; i < N; i += THREADS) {
        cc[i] = make_float2(0.f, 0.f);
        cn[i] = 0.f;
      }
    }
    __syncthreads();

    for (int64_t m0 = mb; m0 < me; m0 += G) {
      const int64_t m = m0 + g;
      const bool active = m < me;  // frames beyond this segment belong to the next one
      cpx v[P];
      if (active) {
        const float2* __restrict__ zf = zc + m * a.z_len;
#pragma unroll
        for (int b = 0; b < B0; ++b)
#pragma unroll
          for (int q = 0; q < R0; ++q) {
            const int i = fft_in_index<PL>(t, b, q);
            float2 val = make_float2(0.f, 0.f);
            if (i < a.z_len) val = __ldg(zf + i);
            v[b * R0 + q] = make_float2(val.y, val.x);  // swap: ifft(x) = swap(fft(swap(x))) / n
          }
      } else {
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = make_float2(0.f, 0.f);
      }
      block_fft<PL>(v, t, bufA, bufB, tw, sync);
      // windowed time-domain frame -> this group's frame buffer (unpadded)
      cpx* fb = ((PL::NP - 1) & 1) ? bufB : bufA;
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int n = fft_out_index<PL>(t, b, q);
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float w = wsm[n];
          fb[n] = make_float2(r.y * w, r.x * w);
        }
      __syncthreads();

      // gather overlap-add of this batch
      int64_t gl = me - m0;  // active frames in this batch
      const int gact = gl < G ? (int)gl : G;
      const bool last = (m0 + G >= me) && (me == a.M);  // flush the tail of the channel
      const bool emit = m0 >= ms;                       // warm-up batches only build the carry
      const cpx* cold = reinterpret_cast<const cpx*>(smem_raw + CF::CARRY_OFF + size_t(cur) * CF::CARRY_ONE);
      const float* nold = reinterpret_cast<const float*>(smem_raw + CF::CARRY_OFF + size_t(cur) * CF::CARRY_ONE +
                                                         size_t(N) * sizeof(cpx));
      cpx* cnew = reinterpret_cast<cpx*>(smem_raw + CF::CARRY_OFF + size_t(cur ^ 1) * CF::CARRY_ONE);
      float* nnew = reinterpret_cast<float*>(smem_raw + CF::CARRY_OFF + size_t(cur ^ 1) * CF::CARRY_ONE +
                                             size_t(N) * sizeof(cpx));
      const int done = gact * hop;  // positions completed by this batch
      const int64_t pos0 = m0 * hop;
      const int fb_rel = ((PL::NP - 1) & 1) ? PL::BUF : 0;  // every group keeps its frame at the same offset
      const int emit_end = !emit ? 0 : (last ? done + OV : done);
      for (int j = tid; j < span; j += THREADS) {
        float re = 0.f, im = 0.f, nr = 0.f;
        if (j < OV) {
          const cpx cv = cold[j];
          re = cv.x;
          im = cv.y;
          nr = nold[j];
        }
        const int g_lo = j - N + 1 <= 0 ? 0 : (j - N + hop) / hop;
        int g_hi = j / hop;
        if (g_hi > gact - 1) g_hi = gact - 1;
        for (int gg = g_lo; gg <= g_hi; ++gg) {
          const int n = j - gg * hop;
          const cpx fv = buf_base[(size_t)(2 * gg) * PL::BUF + fb_rel + n];
          re += fv.x;
          im += fv.y;
          nr += w2sm[n];
        }
        if (j < emit_end) {
          const float d = nr > 1.0e-10f ? nr : 1.0f;  // select(norm > 1e-10, norm, 1.0)
          yc[pos0 + j] = make_float2(re / d, im / d);
        }
        if (j >= done && j < done + OV) {
          cnew[j - done] = make_float2(re, im);
          nnew[j - done] = nr;
        }
      }
      cur ^= 1;
      __syncthreads();  // frame buffers and the old carry are free again
    }
  }
}

// ------------------------------------------------------------------------------------------
// Register overlap-add variant (the hot path): hop = N / HOPDIV with hop a multiple of the
// group width T.  A group of T threads walks a segment of consecutive frames of one channel on
// its own -- no CTA-wide barrier anywhere.  After the last FFT pass thread t holds the frame's
// samples n = t + j*T (j < P), so the running overlap-add lives in P complex registers per
// thread: acc[j] += frame[t + j*T] * w'; the first hop samples are then final (emit acc[j],
// j < hop/T, divided by the overlap-added |w|^2) and the accumulators shift down by hop/T --
// a compile-time register renaming.  Each frame's spectrum (N complex, contiguous) arrives by
// one cp.async.bulk (TMA 1-D) into the group's stage buffer, issued as soon as the previous
// frame has been read out of it, so the load of frame m+1 overlaps the FFT of frame m.
// Segments after the first recompute the HOPDIV-1 frames that overlap their start.
// ------------------------------------------------------------------------------------------
// XD: two exchange buffers per group (no barrier between a middle pass's reads and writes)
template <class PL, int THREADS, bool XD = false>
struct RolaCfg {
  static constexpr int G = THREADS / PL::T, N = PL::N;
  static constexpr size_t GROUP_BYTES = (size_t(N) + size_t(XD ? 2 : 1) * size_t(PL::BUF)) * sizeof(cpx);  // stage + exchange
  static constexpr size_t WIN_OFF = size_t(G) * GROUP_BYTES;
  static constexpr size_t TW_OFF = WIN_OFF + size_t(N) * sizeof(float);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(PL::TW_TOTAL) * sizeof(cpx);
  static constexpr size_t SMEM = BAR_OFF + 8 * size_t(G) + 8;
};

template <class PL, int THREADS, int MINB, int HOPDIV, bool XD = false>
__global__ void __launch_bounds__(THREADS, MINB) istft_rola_kernel(const IstftArgs a) {
  using CF = RolaCfg<PL, THREADS, XD>;
  static_assert(!XD || PL::NP == 3, "the two-buffer form relies on the 3-pass buffer rotation");
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  constexpr int HOP = N / HOPDIV, S = HOP / T;  // S accumulators complete per frame
  static_assert(HOP % T == 0 && S >= 1 && N % HOPDIV == 0, "hop must be a multiple of the group width");
  static_assert((N / RL) % T == 0, "last pass must leave n = t (mod T) in every thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const stage = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::GROUP_BYTES);
  cpx* const xbuf = stage + N;
  float* wsm = reinterpret_cast<float*>(smem_raw + CF::WIN_OFF);
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
  const uint32_t mybar = smem_u32(smem_raw + CF::BAR_OFF) + 8 * g;

  for (int i = tid; i < N; i += THREADS) wsm[i] = a.wprep[i];
  // Nx.abs(window) ** 2 in f32; only the per-thread constants and the channel edges need it
  auto w2 = [&](int n) {
    const float w = fabsf(__ldg(a.w + n));
    return (float)((double)w * (double)w);
  };
  for (int i = tid; i < PL::TW_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  if (tid == 0) {
    for (int i = 0; i < G; ++i) mbar_init(smem_u32(smem_raw + CF::BAR_OFF) + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TwDerive<PL> tw;
  tw.init(twsm, t);
  const GroupSync<T> sync{1 + g};

  // interior normaliser of the S samples a frame completes (kept as its reciprocal): all HOPDIV
  // covering frames present, summed in ascending frame order (= descending window offset)
  float normc[S];
#pragma unroll
  for (int j = 0; j < S; ++j) {
    float nr = 0.f;
#pragma unroll
    for (int k = HOPDIV - 1; k >= 0; --k) nr += w2(t + j * T + k * HOP);
    normc[j] = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);  // reciprocal of select(norm > 1e-10, norm, 1.0): one multiply per sample
  }
  // exact normaliser at output position p of a channel (edges: fewer covering frames)
  auto norm_at = [&](int64_t p) {
    int64_t m_lo = p - N + 1 <= 0 ? 0 : (p - N + HOP) / HOP;
    int64_t m_hi = p / HOP;
    if (m_hi > a.M - 1) m_hi = a.M - 1;
    float nr = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) nr += w2((int)(p - m * HOP));
    return nr;
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  // frame iterator over this group's segments
  auto seg_bounds = [&](int seg, int& c, int64_t& mb, int64_t& ms, int64_t& me) {
    c = seg / a.segs_per_channel;
    const int si = seg - c * a.segs_per_channel;
    ms = (int64_t)si * a.seg_frames;
    me = ms + a.seg_frames;
    if (me > a.M) me = a.M;
    mb = ms - (HOPDIV - 1);
    if (mb < 0) mb = 0;
  };
  auto issue = [&](int c, int64_t m) {
    mbar_expect_tx(mybar, (uint32_t)(N * sizeof(cpx)));
    tma_load_1d(smem_u32(stage), a.z + ((int64_t)c * a.M + m) * N, (uint32_t)(N * sizeof(cpx)), mybar);
  };

  uint32_t parity = 0;
  int seg = gid;
  int c = 0;
  int64_t mb = 0, ms = 0, me = 0;
  if (seg < a.total_segs) {
    seg_bounds(seg, c, mb, ms, me);
    if (t == 0) issue(c, mb);
  }
  while (seg < a.total_segs) {
    float2* __restrict__ yc = a.y + (int64_t)c * a.out_len;
    cpx acc[P];
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
    // the segment after this one (for the prefetch across the segment boundary)
    const int nseg = seg + ngroups;
    int nc = 0;
    int64_t nmb = 0, nms = 0, nme = 0;
    if (nseg < a.total_segs) seg_bounds(nseg, nc, nmb, nms, nme);

    for (int64_t m = mb; m < me; ++m) {
      cpx v[P];
      mbar_wait(mybar, parity);
      parity ^= 1;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const cpx val = stage[fft_in_index<PL>(t, b, q)];
          v[b * R0 + q] = make_float2(val.y, val.x);  // swap: ifft(x) = swap(fft(swap(x))) / n
        }
      sync();  // stage read out (and the previous frame's last exchange reads are done)
      if (t == 0) {
        if (m + 1 < me) issue(c, m + 1);
        else if (nseg < a.total_segs) issue(nc, nmb);
      }
      if constexpr (XD) block_fft<PL>(v, t, xbuf, xbuf + PL::BUF, tw, sync);
      else block_fft_single<PL>(v, t, xbuf, tw, sync);
      // window, accumulate: thread t holds n = t + j*T at v[fft_out_reg(b, q)], j = b + q*BL
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int j = b + q * BL;
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float w = wsm[t + j * T];
          if constexpr (PL::PK) {
            acc[j] = __ffma2_rn(make_float2(r.y, r.x), make_float2(w, w), acc[j]);
          } else {
            acc[j].x += r.y * w;
            acc[j].y += r.x * w;
          }
        }
      if (m >= ms) {
        const int64_t pos = m * HOP + t;
        const bool interior = m >= HOPDIV - 1;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          float rd = normc[j];
          if (!interior) {
            const float nr = norm_at(pos + j * T);
            rd = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);  // select(norm > 1e-10, norm, 1.0)
          }
          __stcs(yc + pos + j * T, cscale_<PL::PK>(acc[j], rd));
        }
      }
#pragma unroll
      for (int j = 0; j < P - S; ++j) acc[j] = acc[j + S];
#pragma unroll
      for (int j = P - S; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
    }
    if (me == a.M) {  // tail of the channel: the N - hop samples no further frame completes
      const int64_t pos = a.M * HOP + t;
#pragma unroll
      for (int j = 0; j < P - S; ++j) {
        const float nr = norm_at(pos + j * T);
        const float rd = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);
        __stcs(yc + pos + j * T, make_float2(acc[j].x * rd, acc[j].y * rd));
      }
    }
    seg = nseg;
    c = nc;
    mb = nmb;
    ms = nms;
    me = nme;
  }
}

// ------------------------------------------------------------------------------------------
// One warp per frame (Plan<N, 32, 32, 32>: 32 points per lane, radices 32 x 32).  Two passes mean ONE
// exchange per transform (the 64-thread plans above need two), the group barrier is __syncwarp, and the
// twiddled pass is one instead of two -- 7 % fewer instructions and a third less shared-memory traffic per
// frame.  What it costs is registers: 32 data points plus the running overlap-add do not fit 168 registers,
// so the carry of the overlap-add (the P - S samples per lane a frame leaves unfinished) lives in a private
// shared-memory column per lane (each lane reads slot j and writes slot j - S of its own column, in
// ascending j: no cross-lane traffic, no barrier).  The exchange buffer doubles as the TMA stage; the next
// frame's copy is issued from the FFT's hook, right behind the last pass's reads.
// ------------------------------------------------------------------------------------------
template <class PL, int THREADS, int HOPDIV>
struct WarpRolaCfg {
  static constexpr int G = THREADS / 32, N = PL::N, P = PL::P, S = (N / HOPDIV) / 32;
  static_assert(PL::T == 32 && PL::NP == 2, "one warp per frame, two passes");
  static_assert(size_t(PL::BUF) >= size_t(N), "the exchange buffer must hold a staged frame");
  static constexpr size_t GROUP_BYTES = (size_t(PL::BUF) + size_t(P - S) * 32) * sizeof(cpx);  // stage/exchange + carry
  static constexpr size_t WIN_OFF = size_t(G) * GROUP_BYTES;
  static constexpr size_t TW_OFF = WIN_OFF + size_t(N) * sizeof(float);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(PL::TW_TOTAL) * sizeof(cpx);
  static constexpr size_t SMEM = BAR_OFF + 8 * size_t(G) + 8;
};

template <class PL, int THREADS, int HOPDIV>
__global__ void __launch_bounds__(THREADS, 1) istft_warp_kernel(const IstftArgs a) {
  using CF = WarpRolaCfg<PL, THREADS, HOPDIV>;
  constexpr int N = PL::N, T = 32, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  constexpr int HOP = N / HOPDIV, S = CF::S;
  static_assert(HOP % T == 0 && S >= 1 && (N / RL) % T == 0, "hop must be a multiple of the warp width");
  static_assert(BL == 1, "the carry column is walked in ascending sample order");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const xbuf = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::GROUP_BYTES);
  cpx* const stage = xbuf;                       // the frame's spectrum lands where the exchange happens later
  cpx* const carry = xbuf + PL::BUF + t;          // this lane's column: carry[j * 32], j < P - S
  float* wsm = reinterpret_cast<float*>(smem_raw + CF::WIN_OFF);
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
  const uint32_t mybar = smem_u32(smem_raw + CF::BAR_OFF) + 8 * g;

  for (int i = tid; i < N; i += THREADS) wsm[i] = a.wprep[i];
  auto w2 = [&](int n) {
    const float w = fabsf(__ldg(a.w + n));
    return (float)((double)w * (double)w);
  };
  for (int i = tid; i < PL::TW_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  if (tid == 0) {
    for (int i = 0; i < G; ++i) mbar_init(smem_u32(smem_raw + CF::BAR_OFF) + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TwDerive<PL> tw;
  tw.init(twsm, t);
  const GroupSync<T> sync{1 + g};

  float normc[S];  // reciprocal of the interior normaliser of the S samples a frame completes
#pragma unroll
  for (int j = 0; j < S; ++j) {
    float nr = 0.f;
#pragma unroll
    for (int k = HOPDIV - 1; k >= 0; --k) nr += w2(t + j * T + k * HOP);
    normc[j] = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);
  }
  auto norm_at = [&](int64_t p) {  // exact normaliser at output position p of a channel (edges)
    int64_t m_lo = p - N + 1 <= 0 ? 0 : (p - N + HOP) / HOP;
    int64_t m_hi = p / HOP;
    if (m_hi > a.M - 1) m_hi = a.M - 1;
    float nr = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) nr += w2((int)(p - m * HOP));
    return nr;
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  auto seg_bounds = [&](int seg, int& c, int64_t& mb, int64_t& ms, int64_t& me) {
    c = seg / a.segs_per_channel;
    const int si = seg - c * a.segs_per_channel;
    ms = (int64_t)si * a.seg_frames;
    me = ms + a.seg_frames;
    if (me > a.M) me = a.M;
    mb = ms - (HOPDIV - 1);
    if (mb < 0) mb = 0;
  };
  auto issue = [&](int c, int64_t m) {
    // the buffer was last written by this warp's generic-proxy stores (the exchange): order them before the
    // bulk copy's async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(mybar, (uint32_t)(N * sizeof(cpx)));
    tma_load_1d(smem_u32(stage), a.z + ((int64_t)c * a.M + m) * N, (uint32_t)(N * sizeof(cpx)), mybar);
  };

  uint32_t parity = 0;
  int seg = gid;
  int c = 0;
  int64_t mb = 0, ms = 0, me = 0;
  if (seg < a.total_segs) {
    seg_bounds(seg, c, mb, ms, me);
    if (t == 0) issue(c, mb);
  }
  while (seg < a.total_segs) {
    float2* __restrict__ yc = a.y + (int64_t)c * a.out_len;
#pragma unroll
    for (int j = 0; j < P - S; ++j) carry[j * T] = make_float2(0.f, 0.f);
    const int nseg = seg + ngroups;
    int nc = 0;
    int64_t nmb = 0, nms = 0, nme = 0;
    if (nseg < a.total_segs) seg_bounds(nseg, nc, nmb, nms, nme);

    for (int64_t m = mb; m < me; ++m) {
      cpx v[P];
      mbar_wait(mybar, parity);
      parity ^= 1;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const cpx val = stage[fft_in_index<PL>(t, b, q)];
          v[b * R0 + q] = make_float2(val.y, val.x);  // swap: ifft(x) = swap(fft(swap(x))) / n
        }
      sync();  // the staged frame is in registers: the buffer may take the exchange
      const bool more = m + 1 < me, next_seg = nseg < a.total_segs;
      auto rearm = [&]() {  // behind the last pass's reads: the buffer is free again, stage the next frame
        sync();
        if (t == 0) {
          if (more) issue(c, m + 1);
          else if (next_seg) issue(nc, nmb);
        }
      };
      block_fft_single_hook<PL>(v, t, xbuf, tw, sync, rearm);
      // window, overlap-add through the lane's carry column, emit the S finished samples
      const bool emit = m >= ms, interior = m >= HOPDIV - 1;
      const int64_t pos = m * HOP + t;
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int j = b + q * BL;
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float w = wsm[t + j * T];
          cpx acc;
          if constexpr (PL::PK) {
            if (j < P - S) acc = __ffma2_rn(make_float2(r.y, r.x), make_float2(w, w), carry[j * T]);
            else acc = __fmul2_rn(make_float2(r.y, r.x), make_float2(w, w));
          } else {
            acc = make_float2(r.y * w, r.x * w);
            if (j < P - S) {
              const cpx old = carry[j * T];
              acc.x += old.x;
              acc.y += old.y;
            }
          }
          if (j < S) {
            if (emit) {
              float rd = normc[j];
              if (!interior) {
                const float nr = norm_at(pos + j * T);
                rd = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);
              }
              __stcs(yc + pos + j * T, cscale_<PL::PK>(acc, rd));
            }
          } else {
            carry[(j - S) * T] = acc;
          }
        }
    }
    if (me == a.M) {  // tail of the channel: the N - hop samples no further frame completes
      const int64_t pos = a.M * HOP + t;
#pragma unroll
      for (int j = 0; j < P - S; ++j) {
        const cpx acc = carry[j * T];
        const float nr = norm_at(pos + j * T);
        const float rd = 1.0f / (nr > 1.0e-10f ? nr : 1.0f);
        __stcs(yc + pos + j * T, make_float2(acc.x * rd, acc.y * rd));
      }
    }
    seg = nseg;
    c = nc;
    mb = nmb;
    ms = nms;
    me = nme;
  }
}

// ------------------------------------------------------------------------------------------
// large fft_length (shared memory cannot hold frames + carry): inverse-transform frames with the
// same engine into a scratch frame tensor; istft_ola_norm_kernel finishes.
// ------------------------------------------------------------------------------------------
template <class PL, int THREADS>
__global__ void __launch_bounds__(THREADS) ifft_frames_kernel(const float2* __restrict__ z, int64_t total_frames,
                                                              int64_t z_len, const float* __restrict__ wprep,
                                                              const float2* __restrict__ twg,
                                                              float2* __restrict__ frames) {
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = THREADS / T;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const bufA = reinterpret_cast<cpx*>(smem_raw) + (size_t)(2 * g) * PL::BUF;
  cpx* const bufB = bufA + PL::BUF;
  TwTable<PL> tw;
  tw.init(twg, t);
  const SyncBlock sync;
  const int64_t tiles = (total_frames + G - 1) / G;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t f = tile * G + g;
    const bool active = f < total_frames;
    cpx v[P];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int i = fft_in_index<PL>(t, b, q);
        float2 val = make_float2(0.f, 0.f);
        if (active && i < z_len) val = __ldg(z + f * z_len + i);
        v[b * R0 + q] = make_float2(val.y, val.x);
      }
    __syncthreads();  // previous tile's exchange reads are complete
    block_fft<PL>(v, t, bufA, bufB, tw, sync);
    if (active) {
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int n = fft_out_index<PL>(t, b, q);
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float w = __ldg(wprep + n);
          frames[f * N + n] = make_float2(r.y * w, r.x * w);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// generic path (any fft_length): direct inverse DFT of every frame into a scratch frame
// tensor, then a gather overlap-add + normalise kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) istft_dft_frames_kernel(const float2* __restrict__ z, int64_t total_frames,
                                                               int z_len, int nfft, const float* __restrict__ wprep,
                                                               const float2* __restrict__ tab,
                                                               float2* __restrict__ frames) {
  extern __shared__ float2 zs[];
  const int nin = z_len < nfft ? z_len : nfft;
  for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < nin; i += blockDim.x) zs[i] = z[f * z_len + i];
    __syncthreads();
    for (int n = threadIdx.x; n < nfft; n += blockDim.x) {
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int k = 0; k < nin; ++k) {
        const float2 w = __ldg(tab + idx);  // exp(+2 pi i k n / nfft)
        const float2 x = zs[k];
        re += x.x * w.x - x.y * w.y;
        im += x.x * w.y + x.y * w.x;
        idx += n;
        if (idx >= nfft) idx -= nfft;
      }
      const float w = wprep[n];
      frames[f * nfft + n] = make_float2(re * w, im * w);
    }
  }
}

__global__ void __launch_bounds__(256) istft_ola_norm_kernel(const float2* __restrict__ frames, int64_t channels,
                                                             int64_t M, int nfft, int hop, int64_t out_len,
                                                             const float* __restrict__ w, float2* __restrict__ y) {
  const int64_t total = channels * out_len;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i % out_len, c = i / out_len;
    int64_t m_hi = n / hop;
    if (m_hi > M - 1) m_hi = M - 1;
    const int64_t m_lo = n - nfft + 1 <= 0 ? 0 : (n - nfft + hop) / hop;
    float re = 0.f, im = 0.f, nr = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) {
      const int k = (int)(n - m * hop);
      const float2 v = frames[(c * M + m) * nfft + k];
      re += v.x;
      im += v.y;
      const float a = fabsf(w[k]);
      nr += (float)((double)a * (double)a);
    }
    const float d = nr > 1.0e-10f ? nr : 1.0f;
    y[i] = make_float2(re / d, im / d);
  }
}

template <class PL>
static int get_tw_table(nxs_ctx* ctx, float2** out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP) | (uint64_t(1) << 62);
  auto it = ctx->tables.find(key);
  if (it != ctx->tables.end()) {
    *out = it->second.tw;
    return NXS_OK;
  }
  std::vector<float2> tw(PL::TW_TOTAL > 0 ? PL::TW_TOTAL : 1);
  for (int p = 1; p < PL::NP; ++p) {
    const int R = PL::R(p), NS = PL::NS(p);
    for (int q = 1; q < R; ++q)
      for (int k = 0; k < NS; ++k) {
        const double ang = -2.0 * M_PI * double(q) * double(k) / double(NS * R);
        tw[PL::twOffset(p) + (q - 1) * NS + k] = make_float2((float)cos(ang), (float)sin(ang));
      }
  }
  PlanTables t;
  NXS_CUDA(ctx, cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
  NXS_CUDA(ctx, cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->tables[key] = t;
  *out = t.tw;
  return NXS_OK;
}

template <class PL, int THREADS, int MINB>
static int run_istft(nxs_ctx* ctx, IstftArgs a, int64_t channels, cudaStream_t st) {
  using CF = IstftCfg<PL, THREADS>;
  float2* tw = nullptr;
  int rc = get_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  const int G = CF::G;
  const int OV = PL::N - a.hop;
  const int warm_frames = (OV + a.hop - 1) / a.hop;
  a.warm_batches = (warm_frames + G - 1) / G;
  int64_t seg = 256;
  if (seg < int64_t(8) * a.warm_batches * G) seg = int64_t(8) * a.warm_batches * G;
  seg = (seg + G - 1) / G * G;
  if (seg > a.M) seg = (a.M + G - 1) / G * G;
  a.seg_frames = (int)seg;
  a.segs_per_channel = (int)((a.M + seg - 1) / seg);
  const int64_t total = int64_t(a.segs_per_channel) * channels;
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.total_segs = (int)total;
  auto kern = istft_kernel<PL, THREADS, MINB>;
  if (CF::SMEM > 231424) return NXS_EUNSUPPORTED;
  static LaunchCache cache;
  int occ = 1;
  rc = cache.get(ctx, kern, THREADS, CF::SMEM, &occ);
  if (rc) return rc;
  int64_t grid = int64_t(ctx->sm_count) * occ;
  if (grid > total) grid = total;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, CF::SMEM, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// ------------------------------------------------------------------------------------------
// Ring overlap-add variant: any hop <= N (z_len == N).  Same per-group walk as the register
// kernel, but the running overlap-add lives in a per-group ring of N complex values in shared
// memory: after the inverse FFT every thread adds its P windowed samples at (base + n) mod N,
// the hop samples at the ring's base are then final (stored, normalised, zeroed) and the base
// advances by hop -- no data moves, no CTA-wide barrier, deterministic.  Costs ~2 P extra
// shared-memory accesses per thread and frame over the register form, and serves the hops
// the register form cannot (hop not a multiple of the group width).
// ------------------------------------------------------------------------------------------
template <class PL, int THREADS>
struct RingCfg {
  static constexpr int G = THREADS / PL::T, N = PL::N;
  static constexpr size_t GROUP_BYTES = (2 * size_t(N) + size_t(PL::BUF)) * sizeof(cpx);  // stage + ring + exchange
  static constexpr size_t WIN_OFF = size_t(G) * GROUP_BYTES;
  static constexpr size_t NORM_OFF = WIN_OFF + size_t(N) * sizeof(float);  // interior normaliser, hop <= N floats
  static constexpr size_t TW_OFF = NORM_OFF + size_t(N) * sizeof(float);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(PL::TWC_TOTAL) * sizeof(cpx);
  static constexpr size_t SMEM = BAR_OFF + 8 * size_t(G) + 8;
};

template <class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) istft_ring_kernel(const IstftArgs a) {
  using CF = RingCfg<PL, THREADS>;
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  static_assert((N / RL) % T == 0, "last pass must leave n = t (mod T) in every thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const stage = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::GROUP_BYTES);
  cpx* const ring = stage + N;
  cpx* const xbuf = ring + N;
  float* wsm = reinterpret_cast<float*>(smem_raw + CF::WIN_OFF);
  float* normsm = reinterpret_cast<float*>(smem_raw + CF::NORM_OFF);
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
  const uint32_t mybar = smem_u32(smem_raw + CF::BAR_OFF) + 8 * g;
  const int hop = a.hop;
  const int kcov = (N + hop - 1) / hop;  // frames covering an interior sample (at most)

  auto w2 = [&](int n) {
    const float w = fabsf(__ldg(a.w + n));
    return (float)((double)w * (double)w);  // Nx.abs(window) ** 2, f32
  };
  for (int i = tid; i < N; i += THREADS) wsm[i] = a.wprep[i];
  // interior normaliser of the sample at offset p < hop of a frame that has all its predecessors:
  // frames m, m-1, ... contribute w2[p], w2[p + hop], ...; ascending frame order = descending offset
  for (int p = tid; p < hop; p += THREADS) {
    float nr = 0.f;
    for (int k = (N - 1 - p) / hop; k >= 0; --k) nr += w2(p + k * hop);
    normsm[p] = nr;
  }
  for (int i = tid; i < PL::TWC_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  if (tid == 0) {
    for (int i = 0; i < G; ++i) mbar_init(smem_u32(smem_raw + CF::BAR_OFF) + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TwDeriveC<PL> tw;
  tw.init(twsm, t);
  const GroupSync<T> sync{1 + g};

  auto norm_at = [&](int64_t p) {  // exact normaliser at output position p (edges: fewer covering frames)
    int64_t m_lo = p - N + 1 <= 0 ? 0 : (p - N + hop) / hop;
    int64_t m_hi = p / hop;
    if (m_hi > a.M - 1) m_hi = a.M - 1;
    float nr = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) nr += w2((int)(p - m * hop));
    return nr;
  };
  auto seg_bounds = [&](int seg, int& c, int64_t& mb, int64_t& ms, int64_t& me) {
    c = seg / a.segs_per_channel;
    const int si = seg - c * a.segs_per_channel;
    ms = (int64_t)si * a.seg_frames;
    me = ms + a.seg_frames;
    if (me > a.M) me = a.M;
    mb = ms - (kcov - 1);
    if (mb < 0) mb = 0;
  };
  auto issue = [&](int c, int64_t m) {
    mbar_expect_tx(mybar, (uint32_t)(N * sizeof(cpx)));
    tma_load_1d(smem_u32(stage), a.z + ((int64_t)c * a.M + m) * N, (uint32_t)(N * sizeof(cpx)), mybar);
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  uint32_t parity = 0;
  int seg = gid;
  int c = 0;
  int64_t mb = 0, ms = 0, me = 0;
  if (seg < a.total_segs) {
    seg_bounds(seg, c, mb, ms, me);
    if (t == 0) issue(c, mb);
  }
  while (seg < a.total_segs) {
    float2* __restrict__ yc = a.y + (int64_t)c * a.out_len;
    for (int i = t; i < N; i += T) ring[i] = make_float2(0.f, 0.f);
    int base = 0;  // ring index of the current frame's sample 0
    const int nseg = seg + ngroups;
    int nc = 0;
    int64_t nmb = 0, nms = 0, nme = 0;
    if (nseg < a.total_segs) seg_bounds(nseg, nc, nmb, nms, nme);

    for (int64_t m = mb; m < me; ++m) {
      cpx v[P];
      mbar_wait(mybar, parity);
      parity ^= 1;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const cpx val = stage[fft_in_index<PL>(t, b, q)];
          v[b * R0 + q] = make_float2(val.y, val.x);  // swap: ifft(x) = swap(fft(swap(x))) / n
        }
      sync();  // stage read out; ring zeroing / the previous frame's emission are complete
      if (t == 0) {
        if (m + 1 < me) issue(c, m + 1);
        else if (nseg < a.total_segs) issue(nc, nmb);
      }
      block_fft_single<PL>(v, t, xbuf, tw, sync);
      // ring[(base + n) mod N] += frame[n] * w'[n], n = t + j*T: each thread owns its own slots
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int n = t + (b + q * BL) * T;
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float w = wsm[n];
          int idx = base + n;
          if (idx >= N) idx -= N;
          cpx acc = ring[idx];
          acc.x += r.y * w;
          acc.y += r.x * w;
          ring[idx] = acc;
        }
      sync();  // the frame is added: its first hop samples are final
      const bool emit = m >= ms;
      const bool interior = m >= kcov - 1;
      const int64_t pos = m * (int64_t)hop;
      for (int p = t; p < hop; p += T) {
        int idx = base + p;
        if (idx >= N) idx -= N;
        if (emit) {
          const cpx acc = ring[idx];
          const float nr = interior ? normsm[p] : norm_at(pos + p);
          const float d = nr > 1.0e-10f ? nr : 1.0f;  // select(norm > 1e-10, norm, 1.0)
          __stcs(yc + pos + p, make_float2(acc.x / d, acc.y / d));
        }
        ring[idx] = make_float2(0.f, 0.f);  // becomes the tail of the next frames
      }
      base += hop;
      if (base >= N) base -= N;
    }
    if (me == a.M) {  // tail of the channel: the N - hop samples no further frame completes
      sync();
      const int64_t pos = a.M * (int64_t)hop;
      for (int p = t; p < N - hop; p += T) {
        int idx = base + p;
        if (idx >= N) idx -= N;
        const cpx acc = ring[idx];
        const float nr = norm_at(pos + p);
        const float d = nr > 1.0e-10f ? nr : 1.0f;
        __stcs(yc + pos + p, make_float2(acc.x / d, acc.y / d));
      }
    }
    sync();  // emission done before the next segment zeroes the ring
    seg = nseg;
    c = nc;
    mb = nmb;
    ms = nms;
    me = nme;
  }
}

// Frames per segment (a segment = consecutive frames of one channel that one group walks, after `warm`
// recomputed frames that rebuild the overlap-add carry).  Large calls: a few segments per group so the tail is
// balanced, never so short that the recomputed frames exceed ~6 %.  Calls too small to give every group such a
// segment are latency-bound (BASELINE configs[0] is 184 frames): segments shrink until every group has one,
// down to 8 frames -- more recomputation on a machine that is mostly idle anyway.  The sums are formed in
// the same order whatever the segmentation, so the result does not depend on it.
static inline int64_t segment_frames(int64_t total_frames, int64_t groups, int64_t warm, int64_t frames_per_channel) {
  int64_t seg = (total_frames + groups * 4 - 1) / (groups * 4);
  const int64_t seg_min = 16 * warm > 32 ? 16 * warm : 32;
  if (seg < seg_min) {
    seg = (total_frames + groups - 1) / groups;
    if (seg > seg_min) seg = seg_min;
    const int64_t lat_min = 2 * warm > 8 ? 2 * warm : 8;
    if (seg < lat_min) seg = lat_min;
  }
  if (seg > frames_per_channel) seg = frames_per_channel;
  return seg < 1 ? 1 : seg;
}

// register overlap-add kernel: requires z_len == N, hop * HOPDIV == N, hop % T == 0 and 16-byte aligned rows
template <class PL, int THREADS, int MINB, int HOPDIV, bool XD = false>
static int run_istft_rola(nxs_ctx* ctx, IstftArgs a, int64_t channels, cudaStream_t st) {
  using CF = RolaCfg<PL, THREADS, XD>;
  float2* tw = nullptr;
  int rc = get_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  auto kern = istft_rola_kernel<PL, THREADS, MINB, HOPDIV, XD>;
  static LaunchCache cache;
  int occ = 1;
  rc = cache.get(ctx, kern, THREADS, CF::SMEM, &occ);
  if (rc) return rc;
  // segments: a few per group so the tail is balanced; each costs HOPDIV-1 recomputed frames
  const int64_t groups = int64_t(ctx->sm_count) * occ * CF::G;
  const int64_t total_frames = channels * a.M;
  const int64_t seg = segment_frames(total_frames, groups, HOPDIV - 1, a.M);
  a.seg_frames = (int)seg;
  a.segs_per_channel = (int)((a.M + seg - 1) / seg);
  const int64_t total = int64_t(a.segs_per_channel) * channels;
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.total_segs = (int)total;
  a.warm_batches = 0;
  int64_t grid = (total + CF::G - 1) / CF::G;
  if (grid > int64_t(ctx->sm_count) * occ) grid = int64_t(ctx->sm_count) * occ;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, CF::SMEM, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS, int HOPDIV>
static int run_istft_warp(nxs_ctx* ctx, IstftArgs a, int64_t channels, cudaStream_t st) {
  using CF = WarpRolaCfg<PL, THREADS, HOPDIV>;
  float2* tw = nullptr;
  int rc = get_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  auto kern = istft_warp_kernel<PL, THREADS, HOPDIV>;
  static_assert(CF::SMEM <= 232448, "istft_warp_kernel: shared memory");
  static int attr_done[16] = {0};
  if (!attr_done[ctx->device & 15]) {
    NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
    attr_done[ctx->device & 15] = 1;
  }
  // segments: a few per warp so the tail is balanced; each costs HOPDIV - 1 recomputed frames
  const int64_t groups = int64_t(ctx->sm_count) * CF::G;
  const int64_t total_frames = channels * a.M;
  const int64_t seg = segment_frames(total_frames, groups, HOPDIV - 1, a.M);
  a.seg_frames = (int)seg;
  a.segs_per_channel = (int)((a.M + seg - 1) / seg);
  const int64_t total = int64_t(a.segs_per_channel) * channels;
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.total_segs = (int)total;
  a.warm_batches = 0;
  int64_t grid = (total + CF::G - 1) / CF::G;
  if (grid > int64_t(ctx->sm_count)) grid = int64_t(ctx->sm_count);
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, CF::SMEM, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL>
static int get_twc_table(nxs_ctx* ctx, float2** out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP) | (uint64_t(1) << 62) |
                       (uint64_t(1) << 61);
  auto it = ctx->tables.find(key);
  if (it != ctx->tables.end()) {
    *out = it->second.tw;
    return NXS_OK;
  }
  std::vector<float2> tw(PL::TWC_TOTAL > 0 ? PL::TWC_TOTAL : 1);
  build_compact_twiddles<PL>(tw.data());
  PlanTables t;
  NXS_CUDA(ctx, cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
  NXS_CUDA(ctx, cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->tables[key] = t;
  *out = t.tw;
  return NXS_OK;
}

// ring overlap-add kernel: requires z_len == N, 16-byte aligned rows, hop <= N
template <class PL, int THREADS, int MINB>
static int run_istft_ring(nxs_ctx* ctx, IstftArgs a, int64_t channels, cudaStream_t st) {
  using CF = RingCfg<PL, THREADS>;
  float2* tw = nullptr;
  int rc = get_twc_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  auto kern = istft_ring_kernel<PL, THREADS, MINB>;
  if (CF::SMEM > 232448) return NXS_EUNSUPPORTED;
  static LaunchCache cache;
  int occ = 1;
  rc = cache.get(ctx, kern, THREADS, CF::SMEM, &occ);
  if (rc) return rc;
  const int64_t kcov = (PL::N + a.hop - 1) / a.hop;
  const int64_t groups = int64_t(ctx->sm_count) * occ * CF::G;
  const int64_t total_frames = channels * a.M;
  const int64_t seg = segment_frames(total_frames, groups, kcov - 1, a.M);
  a.seg_frames = (int)seg;
  a.segs_per_channel = (int)((a.M + seg - 1) / seg);
  const int64_t total = int64_t(a.segs_per_channel) * channels;
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.total_segs = (int)total;
  a.warm_batches = 0;
  int64_t grid = (total + CF::G - 1) / CF::G;
  if (grid > int64_t(ctx->sm_count) * occ) grid = int64_t(ctx->sm_count) * occ;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, CF::SMEM, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS, int MINB>
static int try_istft_rola(nxs_ctx* ctx, const IstftArgs& a, int64_t channels, cudaStream_t st, bool* done) {
  *done = false;
  if (getenv("NXS_ISTFT_NO_ROLA")) return NXS_OK;
  if (a.z_len != PL::N || (reinterpret_cast<uintptr_t>(a.z) & 15) != 0 || a.hop % PL::T != 0) return NXS_OK;
  if (a.M >= (int64_t(1) << 40)) return NXS_OK;
  *done = true;
  if (a.hop * 2 == PL::N) return run_istft_rola<PL, THREADS, MINB, 2>(ctx, a, channels, st);
  if (a.hop * 4 == PL::N) {
    if constexpr (PL::N == 1024) {  // tuning variant (tests/test_istft_gpu.py)
      const char* var = getenv("NXS_ISTFT_VARIANT");
      const int v = var ? atoi(var) : 0;
      if (v == 1) return run_istft_rola<PL, THREADS, MINB, 4>(ctx, a, channels, st);
      if (v == 2) return run_istft_rola<PL, THREADS, MINB, 4, true>(ctx, a, channels, st);  // T = 64, two exchange buffers, 2 CTAs/SM
      if (v == 4) return run_istft_rola<Plan<1024, 32, 32, 32, 1, 1, PL::PK>, 320, 1, 4>(ctx, a, channels, st);
      // the warp-per-frame plan with the overlap-add carry in a private shared-memory column per lane
      // (istft_warp_kernel): more warps per SM, but the carry traffic costs more than the occupancy buys
      if (v == 5) return run_istft_warp<Plan<1024, 32, 32, 32, 1, 1, PL::PK>, 384, 4>(ctx, a, channels, st);
      if (v == 6) return run_istft_warp<Plan<1024, 32, 32, 32, 1, 1, PL::PK>, 320, 4>(ctx, a, channels, st);
      // default: one warp per frame, 32 points per lane, radices 32 x 32 -- ONE exchange per transform and no
      // group barrier (0.713 ms at cfg5 against 0.735 for v == 2 and 0.758 for v == 5, profiles/r02v_istft_variants.txt)
      // with the butterflies on the packed fp32x2 instructions (Plan::PK): 0.720 -> 0.668 ms
      return run_istft_rola<Plan<1024, 32, 32, 32, 1, 1, PL::PK>, 256, 1, 4>(ctx, a, channels, st);
    }
    return run_istft_rola<PL, THREADS, MINB, 4>(ctx, a, channels, st);
  }
  if constexpr (PL::P >= 8) {
    if (a.hop * 8 == PL::N) return run_istft_rola<PL, THREADS, MINB, 8>(ctx, a, channels, st);
  }
  *done = false;
  return NXS_OK;
}

static int run_ola_norm(nxs_ctx* ctx, const float2* frames, int64_t channels, int64_t M, int64_t nfft, int64_t hop,
                        int64_t out_len, const float* window, float2* y, cudaStream_t st) {
  const int64_t total = channels * out_len;
  int64_t grid = (total + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  istft_ola_norm_kernel<<<(unsigned)grid, 256, 0, st>>>(frames, channels, M, (int)nfft, (int)hop, out_len, window, y);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS>
static int run_istft_two_kernels(nxs_ctx* ctx, const IstftArgs& a, int64_t channels, cudaStream_t st) {
  float2* tw = nullptr;
  int rc = get_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  const int64_t total_frames = channels * a.M;
  rc = ensure_scratch(ctx, size_t(total_frames) * PL::N * sizeof(float2));
  if (rc) return rc;
  constexpr int G = THREADS / PL::T;
  const size_t smem = size_t(G) * 2 * PL::BUF * sizeof(cpx);
  auto kern = ifft_frames_kernel<PL, THREADS>;
  NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (total_frames + G - 1) / G;
  int64_t grid = tiles < int64_t(ctx->sm_count) ? tiles : int64_t(ctx->sm_count);
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a.z, total_frames, a.z_len, a.wprep, tw, (float2*)ctx->d_scratch);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return run_ola_norm(ctx, (const float2*)ctx->d_scratch, channels, a.M, PL::N, a.hop, a.out_len, a.w, a.y, st);
}

// ------------------------------------------------------------------------------------------
// Edge fix-up in double.  Near both ends of a channel the reference divides by an overlap-added
// window energy D[p] that tends to zero (lib/nx_signal.ex:630-637, e.g. the first / last ~0.1 N
// samples under a Hann window): y = (frame sample ~ x w) / w^2.  Nx.BinaryBackend computes the
// inverse FFT in f64 and rounds each sample *relatively*, so it stays accurate there; an fp32
// FFT has an error proportional to the frame's largest sample, which the division amplifies by
// 1 / w.  This kernel recomputes exactly those samples -- positions in the partially covered
// head / tail of a channel with D[p] < 1 % of the full-coverage maximum -- as the reference
// does: an f64 inverse DFT of the covering frames, rounded to c64 after the ifft, the rescale,
// the window multiply and the overlap-add, then divided by D[p].  One warp per sample.
// ------------------------------------------------------------------------------------------
// C2R: z holds bins 0 .. nfft/2 of a Hermitian spectrum (row stride z_len), y is real -- the real
// part of the same computation on the conjugate-extended spectrum.
template <bool C2R>
__global__ void __launch_bounds__(256) istft_edge_f64_kernel(const float2* __restrict__ z, int64_t M, int64_t z_len,
                                                             int nfft, int hop, const float* __restrict__ w,
                                                             int scaling, float sr, int64_t out_len,
                                                             const double2* __restrict__ tab,
                                                             void* __restrict__ y_out) {
  __shared__ double red[256];
  __shared__ float s_dmax, s_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, edge = blockIdx.y;
  // Dmax: largest full-coverage normaliser; S: the :spectrum / :psd factor (lib/nx_signal.ex:611-625)
  float dm = 0.f;
  for (int r = tid; r < hop && r < nfft; r += blockDim.x) {
    float d = 0.f;
    for (int n = r; n < nfft; n += hop) {
      const float a = fabsf(w[n]);
      d += (float)((double)a * (double)a);
    }
    dm = fmaxf(dm, d);
  }
  double acc = 0.0;
  if (scaling != NXS_SCALE_NONE)
    for (int i = tid; i < nfft; i += blockDim.x) {
      const float v = w[i];
      acc += (scaling == NXS_SCALE_SPECTRUM) ? (double)v : (double)(float)((double)v * (double)v);
    }
  red[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const double total = red[0];
  __syncthreads();
  red[tid] = (double)dm;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmax(red[tid], red[tid + s]);
    __syncthreads();
  }
  if (tid == 0) {
    s_dmax = (float)red[0];
    float S = 1.f;
    if (scaling == NXS_SCALE_SPECTRUM) S = (float)total;
    else if (scaling == NXS_SCALE_PSD) S = (float)sqrt((double)(float)((double)sr * (double)(float)total));
    s_scale = S;
  }
  __syncthreads();
  const float thresh = 0.01f * s_dmax;
  const double S = (double)s_scale;
  const int OV = nfft - hop;
  // head: [0, OV), tail: [M hop, out_len) (= the last OV samples); never the same sample twice
  int64_t p_lo, p_hi;
  const int64_t head_end = OV < out_len ? OV : out_len;
  if (edge == 0) {
    p_lo = 0;
    p_hi = head_end;
  } else {
    p_lo = M * hop > head_end ? M * hop : head_end;
    p_hi = out_len;
  }
  const float2* __restrict__ zc = z + (int64_t)c * M * z_len;
  const int nin = C2R ? nfft / 2 + 1 : (int)(z_len < nfft ? z_len : nfft);
  for (int64_t p = p_lo + warp; p < p_hi; p += blockDim.x / 32) {
    const int64_t m_lo = p - nfft + 1 <= 0 ? 0 : (p - nfft + hop) / hop;
    int64_t m_hi = p / hop;
    if (m_hi > M - 1) m_hi = M - 1;
    float D = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) {
      const float a = fabsf(w[p - m * hop]);
      D += (float)((double)a * (double)a);
    }
    if (!(D < thresh) || !(D > 1.0e-10f)) continue;  // well conditioned, or the reference's guard case
    double ore = 0.0, oim = 0.0;  // overlap-add of c64 terms, accumulated in f64
    for (int64_t m = m_lo; m <= m_hi; ++m) {
      const int n = (int)(p - m * hop);
      const float2* __restrict__ zf = zc + m * z_len;
      // s = sum_k Z[k] exp(+2 pi i k n / nfft), k split over the lanes
      double sre = 0.0, sim = 0.0;
      int idx = (int)(((int64_t)lane * n) % nfft);
      const int step = (int)(((int64_t)32 * n) % nfft);
      for (int k = lane; k < nin; k += 32) {
        float2 v = zf[k];
        const double2 e = tab[idx];
        if constexpr (C2R) {
          // Re(Z[k] e + conj(Z[k]) conj(e)) = 2 Re(Z[k] e); DC and Nyquist count once, real parts only
          const bool self = k == 0 || 2 * k == nfft;
          if (self) v.y = 0.f;
          const double term = (double)v.x * e.x - (double)v.y * e.y;
          sre += self ? term : 2.0 * term;
        } else {
          sre += (double)v.x * e.x - (double)v.y * e.y;
          sim += (double)v.x * e.y + (double)v.y * e.x;
        }
        idx += step;
        if (idx >= nfft) idx -= nfft;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sre += __shfl_xor_sync(0xffffffffu, sre, o);
        sim += __shfl_xor_sync(0xffffffffu, sim, o);
      }
      float fre = (float)(sre / nfft), fim = (float)(sim / nfft);  // Nx.ifft -> c64
      if (scaling != NXS_SCALE_NONE) {
        fre = (float)((double)fre * S);
        fim = (float)((double)fim * S);
      }
      const double wn = (double)w[n];
      ore += (double)(float)((double)fre * wn);  // frames * window -> c64
      oim += (double)(float)((double)fim * wn);
    }
    if (lane == 0) {
      const float rre = (float)ore, rim = (float)oim;  // overlap_and_add -> c64
      if constexpr (C2R)
        reinterpret_cast<float*>(y_out)[(int64_t)c * out_len + p] = (float)((double)rre / (double)D);
      else
        reinterpret_cast<float2*>(y_out)[(int64_t)c * out_len + p] =
            make_float2((float)((double)rre / (double)D), (float)((double)rim / (double)D));
    }
  }
}

int get_dft_table_f64(nxs_ctx* ctx, int64_t n, double2** out) {
  const uint64_t key = (uint64_t(5) << 32) | uint64_t(n);
  auto it = ctx->dft_tables.find(key);
  if (it != ctx->dft_tables.end()) {
    *out = reinterpret_cast<double2*>(it->second);
    return NXS_OK;
  }
  std::vector<double2> tab(n);
  for (int64_t m = 0; m < n; ++m) {
    // exact at the multiples of pi/2, as cospi / sinpi would be
    const int64_t m8 = (8 * m) % (8 * n);
    double cs, sn;
    if (m8 == 0) { cs = 1; sn = 0; }
    else if (m8 == 2 * n) { cs = 0; sn = 1; }
    else if (m8 == 4 * n) { cs = -1; sn = 0; }
    else if (m8 == 6 * n) { cs = 0; sn = -1; }
    else {
      const double ang = 2.0 * M_PI * double(m) / double(n);
      cs = cos(ang);
      sn = sin(ang);
    }
    tab[m] = make_double2(cs, sn);
  }
  double2* d = nullptr;
  NXS_CUDA(ctx, cudaMalloc(&d, n * sizeof(double2)));
  NXS_CUDA(ctx, cudaMemcpy(d, tab.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  ctx->dft_tables[key] = reinterpret_cast<float2*>(d);
  *out = d;
  return NXS_OK;
}

static int launch_istft_main(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_len,
                             const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                             double sampling_rate, float2* y, cudaStream_t st);

int launch_istft(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_len,
                 const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                 double sampling_rate, float2* y, cudaStream_t st) {
  if (channels <= 0) return NXS_OK;
  int rc = launch_istft_main(ctx, z, channels, num_frames, z_len, window, frame_length, hop, fft_length, scaling,
                             sampling_rate, y, st);
  if (rc) return rc;
  // no overlap: nothing is partially covered; beyond 2^16 points the f64 table is not worth its memory
  if (hop >= fft_length || fft_length > 65536 || getenv("NXS_ISTFT_NO_EDGE_F64")) return NXS_OK;
  double2* tab = nullptr;
  rc = get_dft_table_f64(ctx, fft_length, &tab);
  if (rc) return rc;
  const int64_t out_len = num_frames * hop + (fft_length - hop);
  int64_t done = 0;
  while (done < channels) {  // gridDim.x limit is 2^31 - 1; channels beyond that come in slices
    const int64_t n = channels - done < (int64_t(1) << 30) ? channels - done : (int64_t(1) << 30);
    istft_edge_f64_kernel<false><<<dim3((unsigned)n, 2), 256, 0, st>>>(z + done * num_frames * z_len, num_frames, z_len,
                                                                (int)fft_length, (int)hop, window, scaling,
                                                                (float)sampling_rate, out_len, tab, y + done * out_len);
    ctx->launches++;
    NXS_CUDA(ctx, cudaGetLastError());
    done += n;
  }
  return NXS_OK;
}

static int launch_istft_main(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_len,
                             const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                             double sampling_rate, float2* y, cudaStream_t st) {
  if (channels <= 0) return NXS_OK;
  const int64_t nfft = fft_length;
  if (nfft > (int64_t(1) << 24)) return NXS_EUNSUPPORTED;
  int rc = ensure_coef(ctx, size_t(nfft) * sizeof(float));
  if (rc) return rc;
  // w' = w * S / nfft (istft multiplies by S = sum w | sqrt(sr * sum w^2), lib/nx_signal.ex:611-625)
  rc = launch_prep_window(ctx, window, frame_length, nfft, scaling, sampling_rate, (float)(1.0 / double(nfft)), 1,
                          ctx->d_coef, st);
  if (rc) return rc;

  IstftArgs a;
  a.z = z;
  a.M = num_frames;
  a.z_len = z_len;
  a.wprep = ctx->d_coef;
  a.w = window;
  a.y = y;
  a.out_len = num_frames * hop + (nfft - hop);
  a.hop = (int)hop;
  a.tw = nullptr;
  a.seg_frames = a.segs_per_channel = a.total_segs = a.warm_batches = 0;

  const bool pow2 = (nfft & (nfft - 1)) == 0;
  if (pow2 && nfft >= 128 && nfft <= 4096) {  // register overlap-add fast path (hop = N/2, N/4, N/8)
    bool done = false;
    const bool scalar = getenv("NXS_ISTFT_SCALAR") != nullptr;
    switch (nfft) {
      // FFT engine, window and overlap-add on the packed fp32x2 instructions (Plan::PK; nxs_fft.cuh): cfg5 0.720 ->
      // 0.637 ms, nfft 2048 0.847 -> 0.769 ms (profiles/r02x_packed_fp32x2.txt); NXS_ISTFT_SCALAR=1 runs the scalar plans
#define NXS_ROLA(TH, MB, ...)                                                                          \
  rc = scalar ? try_istft_rola<Plan<__VA_ARGS__>, TH, MB>(ctx, a, channels, st, &done)                 \
              : try_istft_rola<Plan<__VA_ARGS__, 1, true>, TH, MB>(ctx, a, channels, st, &done)
      case 128:  // half-warp groups (16 threads per frame); NXS_ISTFT_NO_ROLA128 keeps the gather kernel (A/B)
        if (!getenv("NXS_ISTFT_NO_ROLA128")) NXS_ROLA(256, 2, 128, 16, 8, 8, 2);
        break;
      case 256: NXS_ROLA(256, 2, 256, 32, 8, 8, 4); break;
      case 512: NXS_ROLA(256, 2, 512, 64, 8, 8, 8); break;
      case 1024: NXS_ROLA(256, 2, 1024, 64, 16, 8, 8); break;
      case 2048: NXS_ROLA(256, 2, 2048, 128, 16, 16, 8); break;
      case 4096: NXS_ROLA(512, 1, 4096, 256, 16, 16, 16); break;
#undef NXS_ROLA
      default: break;
    }
    if (rc || done) return rc;
    // any other hop (<= N, at most 64 covering frames): ring overlap-add
    const bool ring_ok = a.z_len == nfft && (reinterpret_cast<uintptr_t>(a.z) & 15) == 0 && a.hop * 64 >= nfft &&
                         !getenv("NXS_ISTFT_NO_ROLA") && !getenv("NXS_ISTFT_NO_RING");
    if (ring_ok) {
      // FFT engine on packed fp32x2 (Plan::PK): hop 250 1.112 -> 1.048 ms, hop 441 0.732 -> 0.689 ms at nfft 1024,
      // 2048 / 700 1.210 -> 1.178 ms, 512 / 160 1.687 -> 1.638 ms; NXS_ISTFT_SCALAR=1 runs the scalar plans
      const bool ring_pk = !scalar;
      switch (nfft) {
#define NXS_RING(TH, MB, ...)                                                                  \
  return ring_pk ? run_istft_ring<Plan<__VA_ARGS__, 1, true>, TH, MB>(ctx, a, channels, st)    \
                 : run_istft_ring<Plan<__VA_ARGS__>, TH, MB>(ctx, a, channels, st)
        case 256: NXS_RING(256, 2, 256, 32, 8, 8, 4);
        case 512: NXS_RING(256, 2, 512, 64, 8, 8, 8);
        case 1024: NXS_RING(256, 2, 1024, 64, 16, 8, 8);
        case 2048: NXS_RING(256, 1, 2048, 128, 16, 16, 8);
#undef NXS_RING
        default: break;  // 4096: stage + ring + exchange of two groups exceed shared memory -> scratch path below
      }
    }
  }
  if (pow2 && nfft >= 32 && nfft <= 8192) {
    switch (nfft) {
      case 32: return run_istft<Plan<32, 4, 8, 4>, 128, 1>(ctx, a, channels, st);
      case 64: return run_istft<Plan<64, 8, 8, 8>, 128, 1>(ctx, a, channels, st);
      case 128: return run_istft<Plan<128, 16, 8, 8, 2>, 128, 1>(ctx, a, channels, st);
      case 256: return run_istft<Plan<256, 32, 8, 8, 4>, 256, 2>(ctx, a, channels, st);
      case 512: return run_istft<Plan<512, 64, 8, 8, 8>, 256, 2>(ctx, a, channels, st);
      case 1024: return run_istft<Plan<1024, 64, 16, 8, 8>, 256, 2>(ctx, a, channels, st);
      case 2048: return run_istft<Plan<2048, 128, 16, 16, 8>, 512, 1>(ctx, a, channels, st);
      case 4096: return run_istft_two_kernels<Plan<4096, 256, 16, 16, 16>, 512>(ctx, a, channels, st);
      case 8192: return run_istft_two_kernels<Plan<8192, 512, 16, 16, 16, 2>, 512>(ctx, a, channels, st);
      default: break;
    }
  }
  // generic: frames = ifft(z) * w' in scratch, then overlap-add + normalise
  float2* tab = nullptr;
  rc = get_dft_table(ctx, nfft, +1, &tab);
  if (rc) return rc;
  const int64_t total_frames = channels * num_frames;
  rc = ensure_scratch(ctx, size_t(total_frames) * nfft * sizeof(float2));
  if (rc) return rc;
  const int nin = (int)(z_len < nfft ? z_len : nfft);
  const size_t smem = size_t(nin) * sizeof(float2);
  if (smem > 200 * 1024) return NXS_EUNSUPPORTED;
  NXS_CUDA(ctx, cudaFuncSetAttribute(istft_dft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = total_frames < int64_t(ctx->sm_count) * 8 ? total_frames : int64_t(ctx->sm_count) * 8;
  istft_dft_frames_kernel<<<(unsigned)grid, 256, smem, st>>>(z, total_frames, (int)z_len, (int)nfft, ctx->d_coef, tab,
                                                            (float2*)ctx->d_scratch);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return run_ola_norm(ctx, (const float2*)ctx->d_scratch, channels, num_frames, nfft, hop, a.out_len, window, y, st);
}


// ==========================================================================================
// c2r ISTFT (opt-in): z holds only bins 0 .. nfft/2 of each frame -- what
// nxs_stft_onesided_f32_dev writes -- and y is REAL.  Defined as Re(istft(ext(z))) with ext the
// conjugate-mirror extension (the imaginary parts of the DC and Nyquist bins do not reach the
// real part).  For a spectrum that came from a real signal this is the reference's result with
// its (rounding-noise) imaginary part dropped, at half the input bytes, a quarter of the
// output bytes and half the FFT work:
//   2 Z[k] = (X[k] + conj(X[Nh-k])) + i e^{+2 pi i k / nfft} (X[k] - conj(X[Nh-k])),  k < Nh = nfft/2
//   sum_k 2 Z[k] e^{+2 pi i j k / Nh} = nfft (x[2j] + i x[2j+1])
// so one Nh-point complex transform per frame yields the frame's nfft real samples packed in
// pairs, and the register overlap-add of istft_rola_kernel runs on float2 = (even, odd) sample.
// ==========================================================================================
template <class PL, int THREADS>
struct RolaC2rCfg {
  static constexpr int G = THREADS / PL::T, NH = PL::N;
  static constexpr size_t STAGE = size_t(NH) + 2;  // bins 0 .. Nh-1 behind 0 / 1 alignment slots
  static constexpr size_t GROUP_BYTES = (STAGE + size_t(PL::BUF)) * sizeof(cpx);
  static constexpr size_t WIN_OFF = size_t(G) * GROUP_BYTES;
  static constexpr size_t TW_OFF = WIN_OFF + 2 * size_t(NH) * sizeof(float);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(PL::TW_TOTAL) * sizeof(cpx);
  static constexpr size_t NORM_OFF = (BAR_OFF + 8 * size_t(G) + 15) / 16 * 16;  // interior normaliser, large plans
  static constexpr size_t SMEM = NORM_OFF + (PL::P <= 8 ? 0 : size_t(NH) / 2 * sizeof(float2));
};

struct IstftC2rArgs {
  const float2* z;  // [C][M][z_ld], bins 0 .. nfft/2
  int64_t M, z_ld;
  const float* wprep;  // [nfft] = w * S / nfft
  const float* w;      // raw window
  float* y;            // [C][out_len] real
  int64_t out_len;
  int seg_frames, segs_per_channel, total_segs;
  const float2* tw;
  const float2* pre;  // [Nh] i e^{+2 pi i k / nfft}
};

template <class PL, int THREADS, int MINB, int HOPDIV>
__global__ void __launch_bounds__(THREADS, MINB) istft_rola_c2r_kernel(const IstftC2rArgs a) {
  using CF = RolaC2rCfg<PL, THREADS>;
  constexpr int NH = PL::N, NFFT = 2 * NH, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  constexpr int HOP = NFFT / HOPDIV, HOP2 = HOP / 2, S = HOP2 / T;  // S packed accumulators complete per frame
  static_assert(HOP2 % T == 0 && S >= 1 && NFFT % HOPDIV == 0, "hop / 2 must be a multiple of the group width");
  static_assert((NH / RL) % T == 0, "last pass must leave j = t (mod T) in every thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const stage = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::GROUP_BYTES);
  cpx* const xbuf = stage + CF::STAGE;
  float2* wsm = reinterpret_cast<float2*>(smem_raw + CF::WIN_OFF);  // (w'[2j], w'[2j+1])
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
  const uint32_t mybar = smem_u32(smem_raw + CF::BAR_OFF) + 8 * g;

  for (int i = tid; i < NFFT; i += THREADS) reinterpret_cast<float*>(wsm)[i] = a.wprep[i];
  auto w2 = [&](int n) {
    const float w = fabsf(__ldg(a.w + n));
    return (float)((double)w * (double)w);
  };
  for (int i = tid; i < PL::TW_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  if (tid == 0) {
    for (int i = 0; i < G; ++i) mbar_init(smem_u32(smem_raw + CF::BAR_OFF) + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TwDerive<PL> tw;
  tw.init(twsm, t);
  const GroupSync<T> sync{1 + g};

  // pre-pass twiddles of this thread's input points (constant across frames): registers when the
  // plan leaves room (P <= 8), else re-read through L1 every frame
  constexpr bool PRE_REGS = P <= 8;
  cpx pre[PRE_REGS ? P : 1];
  if constexpr (PRE_REGS) {
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) pre[b * R0 + q] = __ldg(a.pre + fft_in_index<PL>(t, b, q));
  }

  auto guard = [](float nr) { return nr > 1.0e-10f ? nr : 1.0f; };  // select(norm > 1e-10, norm, 1.0)
  // reciprocal interior normaliser of the 2 S samples a frame completes (ascending frame order):
  // registers for the small plans, shared memory when P = 16 leaves no room
  constexpr bool NORM_REGS = P <= 8;
  float2 normc[NORM_REGS ? S : 1];
  float2* const nsm = reinterpret_cast<float2*>(smem_raw + CF::NORM_OFF);
  auto norm_interior = [&](int i2) {  // packed index i2 < HOP2
    float n0 = 0.f, n1 = 0.f;
#pragma unroll
    for (int k = HOPDIV - 1; k >= 0; --k) {
      n0 += w2(2 * i2 + k * HOP);
      n1 += w2(2 * i2 + 1 + k * HOP);
    }
    return make_float2(1.0f / guard(n0), 1.0f / guard(n1));  // reciprocals: one multiply per sample
  };
  if constexpr (NORM_REGS) {
#pragma unroll
    for (int j = 0; j < S; ++j) normc[j] = norm_interior(t + j * T);
  } else {
    for (int i = tid; i < HOP2; i += THREADS) nsm[i] = norm_interior(i);
    __syncthreads();
  }
  auto norm_at = [&](int64_t p) {  // exact normaliser at output sample p (edges: fewer covering frames)
    int64_t m_lo = p - NFFT + 1 <= 0 ? 0 : (p - NFFT + HOP) / HOP;
    int64_t m_hi = p / HOP;
    if (m_hi > a.M - 1) m_hi = a.M - 1;
    float nr = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) nr += w2((int)(p - m * HOP));
    return nr;
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  auto seg_bounds = [&](int seg, int& c, int64_t& mb, int64_t& ms, int64_t& me) {
    c = seg / a.segs_per_channel;
    const int si = seg - c * a.segs_per_channel;
    ms = (int64_t)si * a.seg_frames;
    me = ms + a.seg_frames;
    if (me > a.M) me = a.M;
    mb = ms - (HOPDIV - 1);
    if (mb < 0) mb = 0;
  };
  // the copy starts at the 16-byte boundary at or below the row and covers bins 0 .. Nh-1
  auto row_of = [&](int c, int64_t m) { return a.z + ((int64_t)c * a.M + m) * a.z_ld; };
  auto issue = [&](int c, int64_t m) {
    const float2* row = row_of(c, m);
    const uint32_t off = (uint32_t)((reinterpret_cast<uintptr_t>(row) >> 3) & 1);
    const uint32_t bytes = (uint32_t)(NH * sizeof(cpx)) + 16 * off;
    mbar_expect_tx(mybar, bytes);
    tma_load_1d(smem_u32(stage), row - off, bytes, mybar);
  };

  uint32_t parity = 0;
  int seg = gid;
  int c = 0;
  int64_t mb = 0, ms = 0, me = 0;
  if (seg < a.total_segs) {
    seg_bounds(seg, c, mb, ms, me);
    if (t == 0) issue(c, mb);
  }
  while (seg < a.total_segs) {
    float2* __restrict__ yc = reinterpret_cast<float2*>(a.y + (int64_t)c * a.out_len);  // out_len is even
    cpx acc[P];
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
    const int nseg = seg + ngroups;
    int nc = 0;
    int64_t nmb = 0, nms = 0, nme = 0;
    if (nseg < a.total_segs) seg_bounds(nseg, nc, nmb, nms, nme);

    for (int64_t m = mb; m < me; ++m) {
      cpx v[P];
      const float2* row = row_of(c, m);
      const int off = (int)((reinterpret_cast<uintptr_t>(row) >> 3) & 1);
      float nyq = 0.f;
      if (t == 0) nyq = __ldg(&row[NH].x);  // Re X[Nh]: the one bin the copy does not bring
      mbar_wait(mybar, parity);
      parity ^= 1;
      const cpx* __restrict__ sx = stage + off;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int k = fft_in_index<PL>(t, b, q);
          cpx A = sx[k];
          cpx Bc = cconj(sx[NH - k]);
          if (b == 0 && q == 0 && t == 0) {  // k = 0 pairs DC with Nyquist, real parts only
            A.y = 0.f;
            Bc = make_float2(nyq, 0.f);
          }
          const cpx pw = PRE_REGS ? pre[PRE_REGS ? b * R0 + q : 0] : __ldg(a.pre + k);
          // scalar on purpose: the packed forms of this pre-pass and of the overlap-add below cost 4 % at cfg5's
          // shape (0.489 -> 0.508 ms; the operands are assembled from scalars) -- only the FFT engine follows PL::PK
          const cpx Z = cadd(cadd(A, Bc), cmul(pw, csub(A, Bc)));
          v[b * R0 + q] = make_float2(Z.y, Z.x);  // swap: ifft(x) = swap(fft(swap(x))) / n
          // large plans: keep the compiler from hoisting all 2 P stage loads (register pressure)
          if constexpr (P > 8) {
            if ((q & 3) == 3) asm volatile("" ::: "memory");
          }
        }
      sync();  // stage read out (and the previous frame's last exchange reads are done)
      if (t == 0) {
        if (m + 1 < me) issue(c, m + 1);
        else if (nseg < a.total_segs) issue(nc, nmb);
      }
      block_fft_single<PL>(v, t, xbuf, tw, sync);
      // thread t holds packed outputs j' = t + j*T at v[fft_out_reg(b, q)], j = b + q*BL:
      // (swapped) re = x[2j'+1], im = x[2j']
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const int j = b + q * BL;
          const cpx r = v[fft_out_reg<PL>(b, q)];
          const float2 w = wsm[t + j * T];
          acc[j].x += r.y * w.x;
          acc[j].y += r.x * w.y;
        }
      if (m >= ms) {
        const int64_t pos2 = m * HOP2 + t;  // packed output index: samples 2 pos2, 2 pos2 + 1
        const bool interior = m >= HOPDIV - 1;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          const int64_t p2 = pos2 + j * T;
          float2 rn = NORM_REGS ? normc[NORM_REGS ? j : 0] : nsm[t + j * T];
          if (!interior) rn = make_float2(1.0f / guard(norm_at(2 * p2)), 1.0f / guard(norm_at(2 * p2 + 1)));
          __stcs(yc + p2, make_float2(acc[j].x * rn.x, acc[j].y * rn.y));
        }
      }
#pragma unroll
      for (int j = 0; j < P - S; ++j) acc[j] = acc[j + S];
#pragma unroll
      for (int j = P - S; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
    }
    if (me == a.M) {  // tail of the channel: the nfft - hop samples no further frame completes
      const int64_t pos2 = a.M * HOP2 + t;
#pragma unroll
      for (int j = 0; j < P - S; ++j) {
        const int64_t p2 = pos2 + j * T;
        __stcs(yc + p2, make_float2(acc[j].x / guard(norm_at(2 * p2)), acc[j].y / guard(norm_at(2 * p2 + 1))));
      }
    }
    seg = nseg;
    c = nc;
    mb = nmb;
    ms = nms;
    me = nme;
  }
}

// i e^{+2 pi i k / nfft}, k < nfft / 2 (double-computed), cached per context
static int get_c2r_pre_table(nxs_ctx* ctx, int64_t nfft, float2** out) {
  const uint64_t key = (uint64_t(6) << 32) | uint64_t(nfft);
  auto it = ctx->dft_tables.find(key);
  if (it != ctx->dft_tables.end()) {
    *out = it->second;
    return NXS_OK;
  }
  std::vector<float2> tab(nfft / 2);
  for (int64_t k = 0; k < nfft / 2; ++k) {
    const double ang = 2.0 * M_PI * double(k) / double(nfft);
    tab[k] = make_float2((float)-sin(ang), (float)cos(ang));
  }
  float2* d = nullptr;
  NXS_CUDA(ctx, cudaMalloc(&d, tab.size() * sizeof(float2)));
  NXS_CUDA(ctx, cudaMemcpy(d, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->dft_tables[key] = d;
  *out = d;
  return NXS_OK;
}

template <class PL, int THREADS, int MINB, int HOPDIV>
static int run_istft_rola_c2r(nxs_ctx* ctx, IstftC2rArgs a, int64_t channels, cudaStream_t st) {
  using CF = RolaC2rCfg<PL, THREADS>;
  float2* tw = nullptr;
  int rc = get_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  float2* pre = nullptr;
  rc = get_c2r_pre_table(ctx, 2 * PL::N, &pre);
  if (rc) return rc;
  a.pre = pre;
  auto kern = istft_rola_c2r_kernel<PL, THREADS, MINB, HOPDIV>;
  static LaunchCache cache;
  int occ = 1;
  rc = cache.get(ctx, kern, THREADS, CF::SMEM, &occ);
  if (rc) return rc;
  const int64_t groups = int64_t(ctx->sm_count) * occ * CF::G;
  const int64_t total_frames = channels * a.M;
  const int64_t seg = segment_frames(total_frames, groups, HOPDIV - 1, a.M);
  a.seg_frames = (int)seg;
  a.segs_per_channel = (int)((a.M + seg - 1) / seg);
  const int64_t total = int64_t(a.segs_per_channel) * channels;
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.total_segs = (int)total;
  int64_t grid = (total + CF::G - 1) / CF::G;
  if (grid > int64_t(ctx->sm_count) * occ) grid = int64_t(ctx->sm_count) * occ;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, CF::SMEM, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS, int MINB>
static int try_istft_rola_c2r(nxs_ctx* ctx, const IstftC2rArgs& a, int64_t hop, int64_t channels, cudaStream_t st,
                              bool* done) {
  constexpr int NFFT = 2 * PL::N;
  *done = true;
  if (hop * 2 == NFFT) return run_istft_rola_c2r<PL, THREADS, MINB, 2>(ctx, a, channels, st);
  if (hop * 4 == NFFT) return run_istft_rola_c2r<PL, THREADS, MINB, 4>(ctx, a, channels, st);
  if constexpr (NFFT / 16 >= PL::T) {
    if (hop * 8 == NFFT) return run_istft_rola_c2r<PL, THREADS, MINB, 8>(ctx, a, channels, st);
  }
  *done = false;
  return NXS_OK;
}

// other shapes: extend to the two-sided spectrum, run the c64 path, keep the real part
__global__ void __launch_bounds__(256) hermitian_extend_kernel(const float2* __restrict__ z, int64_t frames,
                                                               int64_t z_ld, int nfft, float2* __restrict__ full) {
  const int64_t total = frames * nfft;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / nfft;
    const int k = (int)(i - f * nfft);
    const float2* row = z + f * z_ld;
    float2 v;
    if (2 * k <= nfft) {
      v = row[k];
      if (k == 0 || 2 * k == nfft) v.y = 0.f;
    } else {
      v = cconj(row[nfft - k]);
    }
    full[i] = v;
  }
}
__global__ void __launch_bounds__(256) real_part_kernel(const float2* __restrict__ y, int64_t total,
                                                        float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = y[i].x;
}

int launch_istft_c2r(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                     const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, int scaling,
                     double sampling_rate, float* y, cudaStream_t st) {
  if (channels <= 0) return NXS_OK;
  const int64_t nfft = fft_length;
  if (nfft > (int64_t(1) << 24) || (nfft & 1)) return NXS_EUNSUPPORTED;
  const int64_t out_len = num_frames * hop + (nfft - hop);
  const bool fast_shape = (nfft == 512 || nfft == 1024 || nfft == 2048 || nfft == 4096) &&
                          (hop * 2 == nfft || hop * 4 == nfft || hop * 8 == nfft) &&
                          (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0 &&
                          num_frames < (int64_t(1) << 40) && !getenv("NXS_ISTFT_NO_C2R");
  if (fast_shape) {
    int rc = ensure_coef(ctx, size_t(nfft) * sizeof(float));
    if (rc) return rc;
    rc = launch_prep_window(ctx, window, frame_length, nfft, scaling, sampling_rate, (float)(1.0 / double(nfft)), 1,
                            ctx->d_coef, st);
    if (rc) return rc;
    IstftC2rArgs a;
    a.z = z;
    a.M = num_frames;
    a.z_ld = z_ld;
    a.wprep = ctx->d_coef;
    a.w = window;
    a.y = y;
    a.out_len = out_len;
    a.seg_frames = a.segs_per_channel = a.total_segs = 0;
    a.tw = a.pre = nullptr;
    bool done = false;
    const bool scalar = getenv("NXS_ISTFT_SCALAR") != nullptr;
    switch (nfft) {
#define NXS_ROLA_C2R(TH, MB, ...)                                                                      \
  rc = scalar ? try_istft_rola_c2r<Plan<__VA_ARGS__>, TH, MB>(ctx, a, hop, channels, st, &done)         \
              : try_istft_rola_c2r<Plan<__VA_ARGS__, 1, true>, TH, MB>(ctx, a, hop, channels, st, &done)
      case 512: NXS_ROLA_C2R(256, 2, 256, 32, 8, 8, 4); break;
      case 1024: NXS_ROLA_C2R(256, 2, 512, 64, 8, 8, 8); break;  // packed engine: 0.514 -> 0.488 ms at cfg5's shape
      case 2048: NXS_ROLA_C2R(256, 2, 1024, 64, 16, 8, 8); break;
      case 4096: NXS_ROLA_C2R(256, 2, 2048, 128, 16, 16, 8); break;
#undef NXS_ROLA_C2R
      default: break;
    }
    if (rc) return rc;
    if (done) {
      if (hop >= nfft || getenv("NXS_ISTFT_NO_EDGE_F64")) return NXS_OK;
      double2* tab = nullptr;
      rc = get_dft_table_f64(ctx, nfft, &tab);
      if (rc) return rc;
      int64_t cdone = 0;
      while (cdone < channels) {
        const int64_t n = channels - cdone < (int64_t(1) << 30) ? channels - cdone : (int64_t(1) << 30);
        istft_edge_f64_kernel<true><<<dim3((unsigned)n, 2), 256, 0, st>>>(
            z + cdone * num_frames * z_ld, num_frames, z_ld, (int)nfft, (int)hop, window, scaling,
            (float)sampling_rate, out_len, tab, y + cdone * out_len);
        ctx->launches++;
        NXS_CUDA(ctx, cudaGetLastError());
        cdone += n;
      }
      return NXS_OK;
    }
  }
  // general shapes: temporaries for the two-sided spectrum and the c64 result (freed after the stream drains)
  const int64_t frames = channels * num_frames;
  float2 *full = nullptr, *yc = nullptr;
  NXS_CUDA(ctx, cudaMalloc(&full, size_t(frames) * nfft * sizeof(float2)));
  cudaError_t e = cudaMalloc(&yc, size_t(channels) * out_len * sizeof(float2));
  if (e != cudaSuccess) {
    cudaFree(full);
    return set_cuda_error(ctx, e, "cudaMalloc(c2r result)");
  }
  int64_t grid = (frames * nfft + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  hermitian_extend_kernel<<<(unsigned)grid, 256, 0, st>>>(z, frames, z_ld, (int)nfft, full);
  ctx->launches++;
  int rc = launch_istft(ctx, full, channels, num_frames, nfft, window, frame_length, hop, fft_length, scaling,
                        sampling_rate, yc, st);
  if (rc == NXS_OK) {
    const int64_t total = channels * out_len;
    int64_t g2 = (total + 255) / 256;
    if (g2 > int64_t(ctx->sm_count) * 16) g2 = int64_t(ctx->sm_count) * 16;
    real_part_kernel<<<(unsigned)g2, 256, 0, st>>>(yc, total, y);
    ctx->launches++;
  }
  cudaError_t es = cudaStreamSynchronize(st);
  cudaFree(full);
  cudaFree(yc);
  if (rc) return rc;
  NXS_CUDA(ctx, es);
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

}  // namespace nxs
