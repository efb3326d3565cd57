// nxs_stft.cu -- fused forward STFT for sm_100a.
//
// Replaces NxSignal.stft/3 (lib/nx_signal.ex:68-130): as_windowed (:94-100, :249-364) ->
// Nx.multiply(window) (:101) -> Nx.fft(length: fft_length) (:102) -> optional scaling
// (:113-127), without ever materialising the {M, N} frame tensor.  One group of T threads
// transforms one frame: the real frame is packed as nfft/2 complex points, run through
// the register/shared-memory Stockham FFT (nxs_fft.cuh), and a split post-pass produces
// the full two-sided spectrum (the reference returns all nfft bins, lib/nx_signal.ex:49)
// with coalesced 8-byte stores.  The window is pre-multiplied by 1/2 (the split
// post-pass' factor) and by the :spectrum / :psd scale, so scaling costs nothing.
//
// Any fft_length that is not a supported power of two falls back to a direct O(n^2)
// DFT kernel on the GPU (the reference does the same for odd lengths).
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "nxs_common.cuh"
#include "nxs_fft.cuh"
#include "nxs_tma.cuh"

namespace nxs {

// ------------------------------------------------------------------------------------------
// window preparation: out[s] = w[s] * prescale / S for s < min(N, nfft), 0 up to nfft
//   :spectrum  S = f32(sum w)                       (lib/nx_signal.ex:116)
//   :psd       S = f32(sqrt(f32(sr * f32(sum w^2))))  (lib/nx_signal.ex:119)
// ------------------------------------------------------------------------------------------
__global__ void prep_window_kernel(const float* __restrict__ w, int n, int nfft, int scaling, float sr,
                                   float prescale, int invert, float* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  if (scaling != NXS_SCALE_NONE) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float v = w[i];
      acc += (scaling == NXS_SCALE_SPECTRUM) ? (double)v : (double)(float)((double)v * (double)v);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  double S = 1.0;
  if (scaling == NXS_SCALE_SPECTRUM) S = (double)(float)red[0];
  else if (scaling == NXS_SCALE_PSD) S = (double)(float)sqrt((double)(float)((double)sr * (double)(float)red[0]));
  // stft divides by S, istft multiplies by S
  const float f = (float)(invert ? (double)prescale * S : (double)prescale / S);
  const int nload = n < nfft ? n : nfft;
  for (int i = threadIdx.x; i < nfft; i += blockDim.x) out[i] = (i < nload) ? w[i] * f : 0.f;
}

struct StftArgs {
  const float* x;
  int64_t x_ld, L;
  const float* wprep;  // [nfft], scaled, zero-extended (prep_window_kernel) -- or, with winline, the caller's raw window
  int winline = 0;     // 1: no :scaling, the staged kernel forms w[i] * wmul itself (one launch less per call)
  float wmul = 1.f;
  // what prep_window_kernel needs when a path does not take the window inline
  const float* wraw = nullptr;
  int wn = 0, scaling = 0;
  float sr = 0.f;
  float* wdst = nullptr;
  float2* z;           // [total_frames][z_ld]; nfft bins per frame, or nfft/2 + 1 when onesided
  int64_t z_ld;
  int onesided;        // 1: only bins 0 .. nfft/2 are stored (the rest is their conjugate mirror)
  int64_t M, total_frames, hop, pad_lo;
  int nload;           // min(frame_length, nfft)
  int reflect;
  const float2* tw;
  const float2* post;
  // fused log-mel epilogue (MODE 2): bin-major filterbank (MelLayout, nxs_common.cuh) and outputs
  const float2* mel_w2 = nullptr;  // [nfft / 2] per-bin weights for filters jl, jl + 1
  const int* mel_desc = nullptr;   // [T] piece base | boundary mask << 12 | skip << 31
  const int* mel_ps = nullptr;     // [mel_bins + 3] first piece of each segment
  int mel_bins = 0;
  MelBank* mel_bank = nullptr;     // host side only: the layout is resolved per plan in run_r2c_staged
  float* mel_out = nullptr;      // [total_frames][mel_bins] log10 mel power (before the clamp)
  int* mel_chmax = nullptr;      // [channels] ordered-int maximum
};

// MODE of the r2c kernels' epilogue
enum { kTwoSided = 0, kOneSided = 1, kMel = 2 };

__device__ __forceinline__ int mel_float_key(float v) {
  const int b = __float_as_int(v);
  return b >= 0 ? b : b ^ 0x7fffffff;
}

__device__ __forceinline__ float load_padded(const float* __restrict__ xrow, int64_t src, int64_t L,
                                             int reflect) {
  if (src >= 0 && src < L) return __ldg(xrow + src);
  if (reflect) return __ldg(xrow + reflect_index(src, L));
  return 0.f;
}

template <class PL, class TW, int THREADS>
__global__ void __launch_bounds__(THREADS) stft_r2c_kernel(const StftArgs a) {
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = THREADS / T, NFFT = 2 * N;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  static_assert(THREADS % T == 0 && P >= 2, "bad plan");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* bufA = reinterpret_cast<cpx*>(smem_raw) + (size_t)(2 * g) * PL::BUF;
  cpx* bufB = bufA + PL::BUF;

  TW tw;
  tw.init(a.tw, t);
  cpx wpost[P / 2];
#pragma unroll
  for (int i = 0; i < P / 2; ++i) wpost[i] = __ldg(a.post + t + i * T);
  const SyncBlock sync;

  const int64_t num_tiles = (a.total_frames + G - 1) / G;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t f = tile * G + g;
    const bool active = f < a.total_frames;
    cpx v[P];
    if (active) {
      const int64_t c = f / a.M, m = f - c * a.M;
      const float* __restrict__ xrow = a.x + c * a.x_ld;
      const int64_t src0 = m * a.hop - a.pad_lo;
      const bool fast = (a.nload == NFFT) && src0 >= 0 && src0 + NFFT <= a.L &&
                        ((reinterpret_cast<uintptr_t>(xrow + src0) & 7) == 0);
      if (fast) {
        const float2* __restrict__ xp = reinterpret_cast<const float2*>(xrow + src0);
        const float2* __restrict__ wp = reinterpret_cast<const float2*>(a.wprep);
#pragma unroll
        for (int b = 0; b < B0; ++b)
#pragma unroll
          for (int q = 0; q < R0; ++q) {
            const int i = fft_in_index<PL>(t, b, q);
            const float2 xx = __ldg(xp + i), ww = __ldg(wp + i);
            v[b * R0 + q] = make_float2(xx.x * ww.x, xx.y * ww.y);
          }
      } else {
#pragma unroll
        for (int b = 0; b < B0; ++b)
#pragma unroll
          for (int q = 0; q < R0; ++q) {
            const int s = 2 * fft_in_index<PL>(t, b, q);
            float re = 0.f, im = 0.f;
            if (s < a.nload) re = load_padded(xrow, src0 + s, a.L, a.reflect) * __ldg(a.wprep + s);
            if (s + 1 < a.nload) im = load_padded(xrow, src0 + s + 1, a.L, a.reflect) * __ldg(a.wprep + s + 1);
            v[b * R0 + q] = make_float2(re, im);
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = make_float2(0.f, 0.f);
    }

    block_fft<PL>(v, t, bufA, bufB, tw, sync);

    // split post-pass through the next exchange buffer (unpadded: both Z[k] and Z[N-k]
    // reads are stride +-1 across threads)
    cpx* pb = ((PL::NP - 1) & 1) ? bufB : bufA;
#pragma unroll
    for (int b = 0; b < BL; ++b)
#pragma unroll
      for (int q = 0; q < RL; ++q) pb[fft_out_index<PL>(t, b, q)] = v[fft_out_reg<PL>(b, q)];
    sync();
    if (active) {
      float2* __restrict__ zf = a.z + f * a.z_ld;
      const bool two = !a.onesided;
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int kk = t + i * T;
        const cpx A = pb[kk];
        const cpx Bc = cconj(pb[(N - kk) & (N - 1)]);
        const cpx E = cadd(A, Bc), O = csub(A, Bc);
        const cpx Tm = cmul(wpost[i], O);
        const cpx X0 = cadd(E, Tm), X1 = csub(E, Tm);
        zf[kk] = X0;
        if (two || kk == 0) zf[N + kk] = X1;
        if (kk > 0) {
          zf[N - kk] = cconj(X1);
          if (two) zf[NFFT - kk] = cconj(X0);
        } else {
          const cpx Zh = pb[N / 2];
          zf[N / 2] = make_float2(2.f * Zh.x, -2.f * Zh.y);
          if (two) zf[N + N / 2] = make_float2(2.f * Zh.x, 2.f * Zh.y);
        }
      }
    }
    if constexpr (PL::NP & 1) {
      cpx* tmp = bufA;
      bufA = bufB;
      bufB = tmp;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Staged variant (the hot path).  Input arrives by cp.async.bulk (TMA, 1-D) signalled through
// mbarriers, so no thread ever waits on a global load, and groups synchronise on their own
// named barrier, so the frames in flight on an SM drift apart and overlap their FP, shared-
// memory and global phases.  Two staging layouts:
//   per tile  (PERGROUP = false, the first version): a CTA walks tiles of G consecutive frames
//             of one channel and ONE bulk copy brings the tile's contiguous span,
//             (G-1)*hop + nfft samples, double-buffered; a CTA-wide barrier per tile.
//   per group (PERGROUP = true, what ships): every group stages its own frame and free-runs;
//             the nfft/hop-fold re-read of the input is served by L2, and no CTA-wide barrier
//             exists in the steady state.
// Tiles that touch the padding region (or are not 16-byte aligned) take the per-thread load
// path instead.  MODE selects the epilogue: the reference's two-sided spectrum, the one-sided
// half (bins 0 .. nfft/2), or the fused log-mel reduction (the spectrum is never stored).
// ------------------------------------------------------------------------------------------
// HOPDIV: the stage holds spans for hop <= nfft / HOPDIV; TWREG: twiddles in registers, else a
// shared-memory copy of the per-pass table.
// PERGROUP: every frame group stages its own frame (nfft floats, its own mbarrier pair) and
// free-runs with no CTA-wide barrier; the nfft/hop-fold re-read of the input is served by L2.
// LEAN (PERGROUP only): one stage buffer and one exchange buffer per group -- the next frame's
// TMA is issued as soon as the group has read the stage into registers, so it still overlaps the
// whole FFT; halves the shared memory of the large plans (2 CTAs/SM instead of 1) for one
// extra group barrier per middle pass.
// WINREG: the thread's P window values (constant across frames) live in registers.
// PAIRED: conjugate-paired last pass (nxs_fft.cuh): thread t ends up holding Z[k] and Z[N - k] for each
// of its bins, so the split pass runs out of registers -- no exchange, and no misaligned descending
// shared-memory reads (ncu: they were 10 % excess wavefronts on the nfft = 4096 plan).
// XD (LEAN + PAIRED only): two exchange buffers per group and the compact twiddle table (power-of-two rows only),
// so no barrier separates a middle pass's reads from its writes: three group barriers per frame instead of
// four, paid for with shared memory (one 512-thread CTA per SM instead of two of 256).
template <class PL_, int THREADS_, int HOPDIV_, bool TWREG_, bool PERGROUP_ = false, bool LEAN_ = false,
          bool WINREG_ = false, bool PAIRED_ = false, bool XD_ = false, bool TWC_ = XD_, int KPK_ = 0>
struct StagedCfg {
  using PL = PL_;
  // packed fp32x2 arithmetic outside the FFT engine (which follows Plan::PK): 1 = the window multiply,
  // 2 = the split pass, 4 = the conjugate stores formed from the split pass's operands
  static constexpr int KPK = KPK_;
  static constexpr int THREADS = THREADS_, HOPDIV = HOPDIV_;
  static constexpr bool TWREG = TWREG_, PERGROUP = PERGROUP_, LEAN = LEAN_, WINREG = WINREG_, PAIRED = PAIRED_, XD = XD_;
  static constexpr bool TWC = TWC_ || XD_;  // compact twiddle table (power-of-two rows; TwDeriveC)
  static_assert(!LEAN || PERGROUP, "LEAN needs per-group staging");
  static_assert(!PAIRED || !TWREG, "the paired last pass reads its twiddles from the table");
  static_assert(!XD || (LEAN && PAIRED && !TWREG), "XD is the LEAN + PAIRED layout with a second exchange buffer");
  static constexpr int G = THREADS / PL::T, NFFT = 2 * PL::N;
  static constexpr int NSTAGE = LEAN ? 1 : 2, NXBUF = (LEAN && !XD) ? 1 : 2;
  static constexpr int TW_ENTRIES = TWREG ? 0 : (TWC ? PL::TWC_TOTAL : PL::TW_TOTAL);
  // floats per stage: one tile span, or G private frames
  // (per-group frames carry 4 floats of slack: the copy starts at the 16-byte boundary below the frame)
  static constexpr int FRAME_STAGE = NFFT + 4;
  static constexpr int STAGE = PERGROUP ? G * FRAME_STAGE : NFFT + (G - 1) * (NFFT / HOPDIV);
  static constexpr size_t BUF_BYTES = size_t(G) * NXBUF * PL::BUF * sizeof(cpx);
  static constexpr size_t WIN_OFF = BUF_BYTES;
  static constexpr size_t STAGE_OFF = WIN_OFF + size_t(NFFT) * sizeof(float);
  static constexpr size_t TW_OFF = STAGE_OFF + NSTAGE * size_t(STAGE) * sizeof(float);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(TW_ENTRIES) * sizeof(cpx);
  static constexpr size_t SMEM = BAR_OFF + 16 * (PERGROUP ? G : 1);
};

template <class CF, int MINB, int MODE>
__global__ void __launch_bounds__(CF::THREADS, MINB) stft_r2c_staged_kernel(const StftArgs a, const int tpc,
                                                                            const int total_tiles) {
  using PL = typename CF::PL;
  using TW = typename std::conditional<CF::TWREG, TwRegs<PL>,
                                       typename std::conditional<CF::TWC, TwDeriveC<PL>, TwDerive<PL>>::type>::type;
  constexpr int THREADS = CF::THREADS;
  constexpr int N = PL::N, T = PL::T, P = PL::P, G = CF::G, NFFT = 2 * N;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1), BL = P / RL;
  static_assert(THREADS % T == 0 && P >= 2 && G <= (T > 32 ? 15 : 64), "bad plan");  // named barriers 1 .. 15 (groups of a warp or less use __syncwarp)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* bufA = reinterpret_cast<cpx*>(smem_raw) + (size_t)(CF::NXBUF * g) * PL::BUF;
  cpx* bufB = CF::NXBUF == 1 ? bufA : bufA + PL::BUF;
  float* wsm = reinterpret_cast<float*>(smem_raw + CF::WIN_OFF);
  float* stage0 = reinterpret_cast<float*>(smem_raw + CF::STAGE_OFF);
  const uint32_t bar0 = smem_u32(smem_raw + CF::BAR_OFF);

  constexpr bool ONESIDED = MODE == kOneSided;
  // mel epilogue tables live behind the configuration's own shared memory
  float2* const mel_w = reinterpret_cast<float2*>(smem_raw + ((CF::SMEM + 15) / 16) * 16);  // [N] per-bin weights
  int* const mel_ps = reinterpret_cast<int*>(mel_w + N);                                   // [mel_bins + 3]
  int mel_d = 0;  // this thread's piece descriptor (constant across frames)
  if constexpr (MODE == kMel) {
    for (int i = tid; i < N; i += THREADS) mel_w[i] = a.mel_w2[i];
    for (int i = tid; i < a.mel_bins + 3; i += THREADS) mel_ps[i] = a.mel_ps[i];
    mel_d = __ldg(a.mel_desc + t);
  }
  int mel_c = -1;            // channel the running maximum belongs to
  float mel_max = -INFINITY;  // running maximum of this thread's log-mel values
  if (a.winline) {
    for (int i = tid; i < NFFT; i += THREADS) wsm[i] = i < a.nload ? a.wprep[i] * a.wmul : 0.f;
  } else {
    for (int i = tid; i < NFFT; i += THREADS) wsm[i] = a.wprep[i];
  }
  if constexpr (!CF::TWREG) {
    cpx* twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
    for (int i = tid; i < CF::TW_ENTRIES; i += THREADS) twsm[i] = a.tw[i];
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * (CF::PERGROUP ? G : 1); ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  TW tw;
  if constexpr (CF::TWREG) tw.init(a.tw, t);
  else tw.init(reinterpret_cast<const cpx*>(smem_raw + CF::TW_OFF), t);
  constexpr int JL = N / RL;  // butterflies of the last pass
  // split-pass twiddles of this thread's P / 2 bin pairs: bins t + i T, or (PAIRED) slot b's bins j_b + q JL
  cpx wpost[P / 2];
#pragma unroll
  for (int i = 0; i < P / 2; ++i) {
    if constexpr (CF::PAIRED) {
      int kk = t + (i / RL) * T + (i % RL) * JL;
      if (t == 0 && i >= RL / 2 && i < RL) kk = JL / 2 + (i - RL / 2) * JL;  // thread 0's self-paired butterfly JL / 2
      wpost[i] = __ldg(a.post + kk);
    } else {
      wpost[i] = __ldg(a.post + t + i * T);
    }
  }
  // the pair (Z[kk], conj Z[N - kk]) number i of this thread, its twiddle and (kk == 0 only) Z[N / 2].
  // PAIRED: slot b < B/2 holds butterfly j = t + b T (bins j + q JL in v[b RL + bitrev q]) and slot b + B/2 its
  // conjugate partner butterfly JL - j, so pair q is (slot b, q) with (slot b + B/2, RL - 1 - q).  Thread 0's first
  // slot pair is special -- butterflies 0 and JL / 2 pair with THEMSELVES (q with RL - q, resp. q with RL - 1 - q) --
  // and is brought into the general form by a register permutation done with selects (no divergent branch: a
  // divergent special case here sits on the critical path of the whole group's next barrier):
  //   A' = [A_0 .. A_{R/2-1}, B_0 .. B_{R/2-1}],  B' = [B_{R/2} .. B_{R-1}, A_{R/2+1} .. A_{R-1}, A_0],  Zh = A_{R/2}
  // with bins kk_i = i JL (i < R/2), JL / 2 + (i - R/2) JL (i >= R/2); wpost[] holds the matching twiddles.
  const bool z0 = CF::PAIRED && t == 0;
  auto load_pair = [&](const cpx (&v)[P], const cpx* pb, int i, int& kk, cpx& A, cpx& Bc, cpx& w, cpx& Zh) {
    if constexpr (CF::PAIRED) {
      constexpr int LB = ilog2(RL);
      const int b = i / RL, q = i % RL;
      w = wpost[i];
      if (b == 0) {
        const cpx a_q = v[bitrev(q, LB)];
        const cpx b_p = v[(BL / 2) * RL + bitrev(RL - 1 - q, LB)];  // general partner
        cpx a0, b0;  // what thread 0 uses instead
        if (q < RL / 2) {
          a0 = a_q;
          b0 = v[bitrev((RL - q) % RL, LB)];  // A_{R-q}; q = 0: A_0 itself (DC)
        } else {
          a0 = v[(BL / 2) * RL + bitrev(q - RL / 2, LB)];               // B_{q - R/2}
          b0 = v[(BL / 2) * RL + bitrev(RL - 1 - (q - RL / 2), LB)];    // B_{R-1-(q-R/2)}
        }
        A = make_float2(z0 ? a0.x : a_q.x, z0 ? a0.y : a_q.y);
        Bc = make_float2(z0 ? b0.x : b_p.x, z0 ? -b0.y : -b_p.y);
        kk = (z0 && q >= RL / 2) ? JL / 2 + (q - RL / 2) * JL : t + q * JL;
        if (q == 0) Zh = v[bitrev(RL / 2, LB)];  // only read when kk == 0 (thread 0)
      } else {
        kk = t + b * T + q * JL;
        A = v[b * RL + bitrev(q, LB)];
        Bc = cconj(v[(b + BL / 2) * RL + bitrev(RL - 1 - q, LB)]);
      }
    } else {
      kk = t + i * T;
      A = pb[kk];
      Bc = cconj(pb[(N - kk) & (N - 1)]);
      w = wpost[i];
      if (kk == 0) Zh = pb[N / 2];
    }
  };
  const GroupSync<T> sync{1 + g};
  const int hop = (int)a.hop;
  float2 wreg[CF::WINREG ? P : 1];
  if constexpr (CF::WINREG) {
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q)
        wreg[b * R0 + q] = reinterpret_cast<const float2*>(wsm)[fft_in_index<PL>(t, b, q)];
  }

  // is this tile (PERGROUP: this group's frame) fully interior (no padding) -> staged through TMA
  // tiles are walked with a stride of gridDim.x: (channel, tile-in-channel) advance incrementally,
  // one division per thread at kernel start instead of one per tile (ncu: 4-6 % of all instructions)
  const int step_c = (int)gridDim.x / tpc, step_r = (int)gridDim.x - step_c * tpc;
  auto advance = [&](int& cc, int& rr) {
    cc += step_c;
    rr += step_r;
    if (rr >= tpc) {
      rr -= tpc;
      ++cc;
    }
  };
  auto tile_geom = [&](int c, int r, int& m0, int& gact, int64_t& src0) {
    m0 = r * G;
    const int64_t left = a.M - m0;
    gact = left < G ? (int)left : G;
    src0 = (int64_t)m0 * hop - a.pad_lo;
    if constexpr (CF::PERGROUP) {
      // any hop: the copy covers [s & ~3, ...) in whole 16-byte units and must stay inside the row
      const int64_t s = src0 + (int64_t)g * hop;
      return g < gact && s >= 0 && (s & ~(int64_t)3) + (((s & 3) + NFFT + 3) & ~3) <= a.L;
    } else {
      return src0 >= 0 && src0 + (int64_t)(gact - 1) * hop + NFFT <= a.L;
    }
  };
  // mbarrier / stage buffer of (stage) for this thread's group
  const uint32_t mybar = CF::PERGROUP ? bar0 + 16 * g : bar0;
  float* const mystage = CF::PERGROUP ? stage0 + (size_t)g * CF::FRAME_STAGE : stage0;
  auto issue = [&](int c, int r, int stage) {
    int m0, gact;
    int64_t src0;
    if (tile_geom(c, r, m0, gact, src0)) {
      const uint32_t bar = mybar + 8 * stage;
      int64_t s = CF::PERGROUP ? src0 + (int64_t)g * hop : src0;
      uint32_t bytes = (uint32_t)((CF::PERGROUP ? NFFT : (gact - 1) * hop + NFFT) * sizeof(float));
      if constexpr (CF::PERGROUP) {
        bytes = (uint32_t)((((s & 3) + NFFT + 3) & ~3) * sizeof(float));
        s &= ~(int64_t)3;
      }
      mbar_expect_tx(bar, bytes);
      tma_load_1d(smem_u32(mystage + (size_t)stage * CF::STAGE), a.x + (int64_t)c * a.x_ld + s, bytes, bar);
    }
  };
  const bool issuer = CF::PERGROUP ? (t == 0) : (tid == 0);

  uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s
  int tile = blockIdx.x;
  int c = tile / tpc, rt = tile - c * tpc;  // this tile's channel and index inside the channel
  if (tile < total_tiles && issuer) issue(c, rt, 0);
  for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
    const int stage = CF::LEAN ? 0 : (it & 1);
    int cn = c, rn = rt;  // the tile this CTA takes next
    advance(cn, rn);
    // stage^1 was last read in iteration it-1: PERGROUP -- this group's threads all passed that
    // iteration's barriers before the issuer gets here; else a CTA-wide barrier says so
    if constexpr (!CF::PERGROUP) __syncthreads();
    if constexpr (!CF::LEAN) {
      if (issuer && tile + (int)gridDim.x < total_tiles) issue(cn, rn, stage ^ 1);
    }

    int m0, gact;
    int64_t src0;
    const bool staged = tile_geom(c, rt, m0, gact, src0);
    const int m = m0 + g;
    const bool active = g < gact;
    const int64_t f = (int64_t)c * a.M + m;
    cpx v[P];
    if (staged) {
      mbar_wait(mybar + 8 * stage, (phase_bits >> stage) & 1);
      phase_bits ^= (1u << stage);
      if (active) {
        // offset of the frame inside the staged copy (per group: 0..3 samples past the 16-byte boundary)
        const int off = CF::PERGROUP ? (int)((src0 + (int64_t)g * hop) & 3) : g * hop;
        const float* __restrict__ xs = mystage + (size_t)stage * CF::STAGE + off;
        const float2* __restrict__ wp = reinterpret_cast<const float2*>(wsm);
        if (!CF::PERGROUP || !(off & 1)) {
          const float2* __restrict__ xp = reinterpret_cast<const float2*>(xs);
#pragma unroll
          for (int b = 0; b < B0; ++b)
#pragma unroll
            for (int q = 0; q < R0; ++q) {
              const int i = fft_in_index<PL>(t, b, q);
              const float2 xx = xp[i];
              float2 ww;
              if constexpr (CF::WINREG) ww = wreg[b * R0 + q];
              else ww = wp[i];
              if constexpr ((CF::KPK & 1) != 0) v[b * R0 + q] = __fmul2_rn(xx, ww);
              else v[b * R0 + q] = make_float2(xx.x * ww.x, xx.y * ww.y);
            }
        } else {  // odd sample offset: the pairs are not 8-byte aligned in shared memory
#pragma unroll
          for (int b = 0; b < B0; ++b)
#pragma unroll
            for (int q = 0; q < R0; ++q) {
              const int i = fft_in_index<PL>(t, b, q);
              float2 ww;
              if constexpr (CF::WINREG) ww = wreg[b * R0 + q];
              else ww = wp[i];
              v[b * R0 + q] = make_float2(xs[2 * i] * ww.x, xs[2 * i + 1] * ww.y);
            }
        }
      } else {
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = make_float2(0.f, 0.f);
      }
    } else if (active) {
      const float* __restrict__ xrow = a.x + (int64_t)c * a.x_ld;
      const int64_t s0 = src0 + (int64_t)g * hop;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int s = 2 * fft_in_index<PL>(t, b, q);
          const float re = load_padded(xrow, s0 + s, a.L, a.reflect) * wsm[s];
          const float im = load_padded(xrow, s0 + s + 1, a.L, a.reflect) * wsm[s + 1];
          v[b * R0 + q] = make_float2(re, im);
        }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = make_float2(0.f, 0.f);
    }

    if constexpr (CF::LEAN) {
      // the group has read its stage (and finished the previous frame's post-pass reads of the
      // exchange buffer): refill the stage with the next frame while this one is transformed
      sync();
      if (issuer && tile + (int)gridDim.x < total_tiles) issue(cn, rn, 0);
      if constexpr (CF::XD) block_fft_paired<PL>(v, t, bufA, bufB, tw, sync);
      else block_fft_single<PL, TW, GroupSync<T>, CF::PAIRED>(v, t, bufA, tw, sync);
      if constexpr (!CF::PAIRED) sync();  // last pass's reads done before the post-pass reuses the buffer
    } else if constexpr (CF::PAIRED) {
      block_fft_paired<PL>(v, t, bufA, bufB, tw, sync);
    } else {
      block_fft<PL>(v, t, bufA, bufB, tw, sync);
    }

    cpx* pb = ((PL::NP - 1) & 1) ? bufB : bufA;
    if constexpr (!CF::PAIRED) {
#pragma unroll
      for (int b = 0; b < BL; ++b)
#pragma unroll
        for (int q = 0; q < RL; ++q) pb[fft_out_index<PL>(t, b, q)] = v[fft_out_reg<PL>(b, q)];
      sync();
    }
    if constexpr (MODE == kMel) {
      // power of bins 0 .. N-1 (the lower half-spectrum, lib/nx_signal.ex:496) -> shared memory in bin order
      float p0[P / 2], p1[P / 2], ph = 0.f;
      int kks[P / 2];
      if (active) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i) {
          int kk;
          cpx A, Bc, w, Zh = make_float2(0.f, 0.f);
          load_pair(v, pb, i, kk, A, Bc, w, Zh);
          kks[i] = kk;
          const cpx E = cadd(A, Bc), O = csub(A, Bc);
          const cpx Tm = cmul(w, O);
          const cpx X0 = cadd(E, Tm), X1 = csub(E, Tm);
          p0[i] = X0.x * X0.x + X0.y * X0.y;  // bin kk
          p1[i] = X1.x * X1.x + X1.y * X1.y;  // bin N - kk (|conj| = |.|)
          if (kk == 0) ph = 4.f * (Zh.x * Zh.x + Zh.y * Zh.y);  // bin N / 2
        }
      }
      // the power spectrum goes to the exchange buffer the last pass did not use (its reads ended
      // before the barrier above), so no barrier is needed here; with a single buffer (LEAN) or
      // without the split pass's barrier (PAIRED) every read of it must finish first
      if constexpr (CF::LEAN || CF::PAIRED) sync();
      float* pw = reinterpret_cast<float*>(CF::LEAN ? pb : (pb == bufA ? bufB : bufA));
      if (active) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i) {
          const int kk = kks[i];
          pw[kk] = p0[i];
          if (kk > 0) pw[N - kk] = p1[i];
          else pw[N / 2] = ph;
        }
      }
      sync();
      // bin-major reduction (MelLayout): this thread owns bins [P t, P t + P) and emits one partial
      // sum per filter segment it touches (pieces, numbered in bin order, in the buffer's upper half)
      // (two buffers: pieces overwrite pb, whose pair reads all precede the barrier above)
      float2* const piece = CF::LEAN ? reinterpret_cast<float2*>(pw + N) : reinterpret_cast<float2*>(pb);
      if (active && mel_d >= 0) {
        int r = mel_d & 0xfff;
        const unsigned mask = ((unsigned)mel_d >> 12) & 0xffffu;
        const float4* __restrict__ p4 = reinterpret_cast<const float4*>(pw + P * t);
        // weights are stored load-major ([load i][thread t], MelLayout): conflict-free 128-bit reads
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(mel_w) + t;
        float U = 0.f, D = 0.f;
#pragma unroll
        for (int h = 0; h < P / 4; ++h) {
          const float4 pv = p4[h];
          const float4 wa = w4[(2 * h) * T], wb = w4[(2 * h + 1) * T];
          const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
          const float wl[4] = {wa.x, wa.z, wb.x, wb.z}, wh[4] = {wa.y, wa.w, wb.y, wb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int b = 4 * h + e;
            if (b > 0 && ((mask >> b) & 1u)) {
              piece[r] = make_float2(U, D);
              ++r;
              U = 0.f;
              D = 0.f;
            }
            U = fmaf(wl[e], pp[e], U);
            D = fmaf(wh[e], pp[e], D);
          }
        }
        piece[r] = make_float2(U, D);
      }
      sync();
      if (active) {
        if (c != mel_c) {  // uniform over the group
          if (mel_c >= 0) {
            float m = mel_max;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((tid & 31) == 0) atomicMax(a.mel_chmax + mel_c, mel_float_key(m));
          }
          mel_c = c;
          mel_max = -INFINITY;
        }
        float* __restrict__ orow = a.mel_out + f * a.mel_bins;
        for (int j = t; j < a.mel_bins; j += T) {
          // filter j = rising-slope pieces of segment j (.y) + falling-slope pieces of segment j + 1 (.x)
          const int q0 = mel_ps[j], q1 = mel_ps[j + 1], q2 = mel_ps[j + 2];
          float acc = 0.f;
          // a handful of pieces per filter: a plain counted loop (unrolling only adds prologue code)
#pragma unroll 1
          for (int q = q0; q < q2; ++q) {
            const float2 pc = piece[q];
            acc += q < q1 ? pc.y : pc.x;
          }
          const float v = __log2f(fmaxf(acc, 1.0e-10f)) * 0.30102999566f;  // log10(clip(mel, 1e-10)), :511
          orow[j] = v;
          mel_max = fmaxf(mel_max, v);
        }
      }
    } else if (active) {
      float2* __restrict__ zf = a.z + f * (ONESIDED ? a.z_ld : (int64_t)NFFT);
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        int kk;
        cpx A, Bc, w, Zh = make_float2(0.f, 0.f);
        load_pair(v, pb, i, kk, A, Bc, w, Zh);
        constexpr bool PK = (CF::KPK & 2) != 0, PKC = (CF::KPK & 4) != 0;
        const cpx E = cadd_<PK>(A, Bc), O = csub_<PK>(A, Bc);
        const cpx Tm = cmul_<PK>(O, w);
        const cpx X0 = cadd_<PK>(E, Tm), X1 = csub_<PK>(E, Tm);
        __stcs(zf + kk, X0);
        if (!ONESIDED || kk == 0) __stcs(zf + N + kk, X1);
        if (kk > 0) {
          // conj(X1), conj(X0) formed from E and Tm (packed: the half negations are operand modifiers)
          __stcs(zf + N - kk, PKC ? cadd_<PKC>(make_float2(E.x, -E.y), make_float2(-Tm.x, Tm.y)) : cconj(X1));
          if constexpr (!ONESIDED)
            __stcs(zf + NFFT - kk, PKC ? cadd_<PKC>(make_float2(E.x, -E.y), make_float2(Tm.x, -Tm.y)) : cconj(X0));
        } else {
          __stcs(zf + N / 2, make_float2(2.f * Zh.x, -2.f * Zh.y));
          if constexpr (!ONESIDED) __stcs(zf + N + N / 2, make_float2(2.f * Zh.x, 2.f * Zh.y));
        }
      }
    }
    if constexpr ((PL::NP & 1) && !CF::LEAN) {
      cpx* tmp = bufA;
      bufA = bufB;
      bufB = tmp;
    }
    c = cn;
    rt = rn;
  }
  if constexpr (MODE == kMel) {
    if (mel_c >= 0) {  // uniform over a warp: its threads belong to one group
      float m = mel_max;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((tid & 31) == 0) atomicMax(a.mel_chmax + mel_c, mel_float_key(m));
    }
  }
}

// ------------------------------------------------------------------------------------------
// generic path: direct DFT, any fft_length >= 1.  One CTA per frame.
//   tab[m] = exp(-2 pi i m / nfft) (double-computed), X[k] = sum_s xw[s] * tab[(k*s) mod nfft]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_dft_kernel(const StftArgs a, int nfft, const float2* __restrict__ tab) {
  extern __shared__ float xw[];
  for (int64_t f = blockIdx.x; f < a.total_frames; f += gridDim.x) {
    const int64_t c = f / a.M, m = f - c * a.M;
    const float* __restrict__ xrow = a.x + c * a.x_ld;
    const int64_t src0 = m * a.hop - a.pad_lo;
    __syncthreads();
    for (int s = threadIdx.x; s < a.nload; s += blockDim.x)
      xw[s] = load_padded(xrow, src0 + s, a.L, a.reflect) * a.wprep[s];
    __syncthreads();
    const int nout = a.onesided ? nfft / 2 + 1 : nfft;
    for (int k = threadIdx.x; k < nout; k += blockDim.x) {
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int s = 0; s < a.nload; ++s) {
        const float2 w = __ldg(tab + idx);
        const float xv = xw[s];
        re = fmaf(xv, w.x, re);
        im = fmaf(xv, w.y, im);
        idx += k;
        if (idx >= nfft) idx -= nfft;
      }
      // Nx.fft snaps |re|, |im| <= eps (1e-10) to zero
      a.z[f * a.z_ld + k] = make_float2(fabsf(re) <= 1e-10f ? 0.f : re, fabsf(im) <= 1e-10f ? 0.f : im);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <class PL>
static int get_tables(nxs_ctx* ctx, PlanTables* out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP);
  auto it = ctx->tables.find(key);
  if (it != ctx->tables.end()) {
    *out = it->second;
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
  std::vector<float2> post(PL::N + 1);  // the paired split pass indexes bins up to N - 1
  for (int k = 0; k <= PL::N; ++k) {
    const double th = M_PI * double(k) / double(PL::N);
    post[k] = make_float2((float)(-sin(th)), (float)(-cos(th)));
  }
  PlanTables t;
  NXS_CUDA(ctx, cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
  NXS_CUDA(ctx, cudaMalloc(&t.post, post.size() * sizeof(float2)));
  NXS_CUDA(ctx, cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(t.post, post.data(), post.size() * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->tables[key] = t;
  *out = t;
  return NXS_OK;
}

int get_dft_table(nxs_ctx* ctx, int64_t n, int sign, float2** out) {
  const uint64_t key = (uint64_t(sign < 0 ? 2 : 3) << 32) | uint64_t(n);
  auto it = ctx->dft_tables.find(key);
  if (it != ctx->dft_tables.end()) {
    *out = it->second;
    return NXS_OK;
  }
  std::vector<float2> tab(n);
  for (int64_t m = 0; m < n; ++m) {
    const double ang = (sign < 0 ? -2.0 : 2.0) * M_PI * double(m) / double(n);
    double c = cos(ang), sn = sin(ang);
    if (fabs(c) < 1e-15) c = 0.0;  // exact zeros at multiples of pi/2
    if (fabs(sn) < 1e-15) sn = 0.0;
    tab[m] = make_float2((float)c, (float)sn);
  }
  float2* d = nullptr;
  NXS_CUDA(ctx, cudaMalloc(&d, n * sizeof(float2)));
  NXS_CUDA(ctx, cudaMemcpy(d, tab.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
  ctx->dft_tables[key] = d;
  *out = d;
  return NXS_OK;
}

int launch_prep_window(nxs_ctx* ctx, const float* window, int64_t n, int64_t nfft, int scaling,
                       double sampling_rate, float prescale, int invert, float* out, cudaStream_t st);

// scaled, zero-extended window into the context's coefficient block (paths that do not take it inline)
static int prep_for(nxs_ctx* ctx, StftArgs& a, int nfft, cudaStream_t st) {
  int rc = launch_prep_window(ctx, a.wraw, a.wn, nfft, a.scaling, a.sr, a.wmul, 0, a.wdst, st);
  a.wprep = a.wdst;
  a.winline = 0;
  return rc;
}

template <class PL, class TW, int THREADS>
static int run_r2c(nxs_ctx* ctx, StftArgs a, cudaStream_t st) {
  if (a.mel_out) return NXS_EUNSUPPORTED;  // the caller chains stft -> stft_to_mel instead
  PlanTables tabs;
  int rc = get_tables<PL>(ctx, &tabs);
  if (rc) return rc;
  rc = prep_for(ctx, a, 2 * PL::N, st);
  if (rc) return rc;
  a.tw = tabs.tw;
  a.post = tabs.post;
  constexpr int G = THREADS / PL::T;
  const size_t smem = size_t(G) * 2 * PL::BUF * sizeof(cpx);
  auto kern = stft_r2c_kernel<PL, TW, THREADS>;
  static LaunchCache cache;
  int occ = 1;
  rc = cache.get(ctx, kern, THREADS, smem, &occ);
  if (rc) return rc;
  const int64_t tiles = (a.total_frames + G - 1) / G;
  int64_t grid = int64_t(ctx->sm_count) * occ;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// compact twiddle table (TwDeriveC: only the power-of-two rows of every pass), cached per context
template <class PL>
static int get_compact_tw(nxs_ctx* ctx, float2** out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP) | (uint64_t(1) << 61);
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

template <class CF, int MINB>
static int run_r2c_staged(nxs_ctx* ctx, StftArgs a, int64_t channels, cudaStream_t st) {
  using PL = typename CF::PL;
  constexpr int THREADS = CF::THREADS;
  PlanTables tabs;
  int rc = get_tables<PL>(ctx, &tabs);
  if (rc) return rc;
  a.tw = tabs.tw;
  a.post = tabs.post;
  if constexpr (CF::TWC) {
    float2* twc = nullptr;
    rc = get_compact_tw<PL>(ctx, &twc);
    if (rc) return rc;
    a.tw = twc;
  }
  if constexpr (CF::XD) {
    if (a.mel_out) return NXS_EUNSUPPORTED;  // the fused log-mel epilogue keeps the single-buffer layout
  }
  const int64_t tpc = (a.M + CF::G - 1) / CF::G;
  const int64_t tiles = tpc * channels;
  auto kern = a.mel_out ? stft_r2c_staged_kernel<CF, MINB, kMel>
              : a.onesided ? stft_r2c_staged_kernel<CF, MINB, kOneSided>
                           : stft_r2c_staged_kernel<CF, MINB, kTwoSided>;
  size_t smem = CF::SMEM;
  if (a.mel_out) {
    const MelLayout* lay = nullptr;
    rc = get_mel_layout(ctx, a.mel_bank, PL::P, &lay);
    if (rc) return rc;
    if (!lay->ok) return NXS_EUNSUPPORTED;  // not a triangular bank: the caller chains stft -> stft_to_mel
    a.mel_w2 = lay->d_w2;
    a.mel_desc = lay->d_desc;
    a.mel_ps = lay->d_ps;
    smem = (CF::SMEM + 15) / 16 * 16 + size_t(PL::N) * sizeof(float2) + size_t(a.mel_bins + 3) * sizeof(int);
  }
  if (smem > 232448) return NXS_EUNSUPPORTED;
  static LaunchCache cache[3];  // one per epilogue mode
  int occ = 1;
  rc = cache[a.mel_out ? 2 : a.onesided ? 1 : 0].get(ctx, kern, THREADS, smem, &occ);
  if (rc) return rc;
  if (a.scaling == NXS_SCALE_NONE) {  // the kernel scales the raw window while it loads it
    a.wprep = a.wraw;
    a.winline = 1;
  } else {
    rc = prep_for(ctx, a, CF::NFFT, st);
    if (rc) return rc;
  }
  int64_t grid = int64_t(ctx->sm_count) * occ;
  if (grid > tiles) grid = tiles;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a, (int)tpc, (int)tiles);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// can the staged (TMA) kernel serve this call?  (16-byte alignment of every tile span, hop <= nfft/2)
static int variant_env() {
  const char* var = getenv("NXS_STFT_VARIANT");
  return var ? atoi(var) : 0;
}

template <class CF>
static bool staged_ok(const StftArgs& a, int64_t channels) {
  const int64_t tiles = ((a.M + CF::G - 1) / CF::G) * channels;
  // per-group staging takes any hop / padding (each copy starts at the 16-byte boundary below its frame);
  // the per-tile layout needs aligned spans.  Rows must start 16-byte aligned either way.
  const bool aligned = CF::PERGROUP || (a.hop % 4 == 0 && a.pad_lo % 4 == 0 && a.hop <= CF::NFFT / CF::HOPDIV);
  return a.nload == CF::NFFT && aligned && a.x_ld % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && tiles < (int64_t(1) << 30) &&
         CF::SMEM <= 231424;
}

int launch_prep_window(nxs_ctx* ctx, const float* window, int64_t n, int64_t nfft, int scaling,
                       double sampling_rate, float prescale, int invert, float* out, cudaStream_t st) {
  prep_window_kernel<<<1, 256, 0, st>>>(window, (int)n, (int)nfft, scaling, (float)sampling_rate, prescale,
                                        invert, out);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

bool stft_has_exact_mirror(int64_t nfft) {
  return (nfft & (nfft - 1)) == 0 && nfft >= 64 && nfft <= 16384;
}

int launch_stft(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                const float* window, int64_t frame_length, int64_t hop, int64_t fft_length,
                const PadGeom& g, int64_t num_frames, int scaling, double sampling_rate, float2* z,
                int64_t z_ld, int onesided, cudaStream_t st, const MelEpilogue* mel) {
  if (num_frames <= 0 || channels <= 0) return NXS_OK;
  if (fft_length > (int64_t(1) << 24) || frame_length > (int64_t(1) << 24)) return NXS_EUNSUPPORTED;
  const int64_t nfft = fft_length;
  // One nfft-4096 launch over tens of GB runs below the rate of shard-sized launches over the same tensors (the achieved
  // bandwidth falls with the footprint of a launch: 1024 ch x 60 s, 106 GB: 22.3 ms as one launch against 19.6 ms as
  // seven; 256 ch: 4.78 against 4.42 ms; profiles/r03i_footprint.txt): such inputs are walked in channel blocks of
  // ~16 GB.  Measured for this kernel only -- cfg2's nfft-1024 kernel LOSES 5 % when split (66 GB: 10.8 -> 11.4 ms) --
  // so the other sizes keep the single launch.
  if (!mel && channels > 1 && nfft == 4096) {
    const char* e = getenv("NXS_STFT_SPLIT_BYTES");
    const double split_bytes = e && atof(e) > 0 ? atof(e) : 16e9;
    const int64_t z_row = onesided ? z_ld : nfft;
    const double per_ch = 4.0 * double(length) + 8.0 * double(num_frames) * double(z_row);
    int64_t blk = (int64_t)(split_bytes / per_ch);
    if (blk < 1) blk = 1;
    if (channels > blk + blk / 2) {
      const int64_t nblk = (channels + blk - 1) / blk;
      blk = (channels + nblk - 1) / nblk;
      for (int64_t c0 = 0; c0 < channels; c0 += blk) {
        const int64_t n = channels - c0 < blk ? channels - c0 : blk;
        const int rcb = launch_stft(ctx, x + c0 * x_ld, n, length, x_ld, window, frame_length, hop, fft_length, g, num_frames,
                                    scaling, sampling_rate, z + c0 * num_frames * z_row, z_ld, onesided, st, nullptr);
        if (rcb) return rcb;
      }
      return NXS_OK;
    }
  }
  int rc = ensure_coef(ctx, size_t(nfft) * sizeof(float));
  if (rc) return rc;

  StftArgs a;
  a.x = x;
  a.x_ld = x_ld;
  a.L = length;
  a.wprep = ctx->d_coef;
  a.wdst = ctx->d_coef;
  a.wraw = window;
  a.wn = (int)frame_length;
  a.scaling = scaling;
  a.sr = (float)sampling_rate;
  a.z = z;
  a.z_ld = z_ld;
  a.onesided = onesided;
  a.M = num_frames;
  a.total_frames = num_frames * channels;
  a.hop = hop;
  a.pad_lo = g.lo;
  a.nload = (int)(frame_length < nfft ? frame_length : nfft);
  a.reflect = g.reflect;
  a.tw = nullptr;
  a.post = nullptr;
  if (mel) {  // fused log-mel epilogue: served by the staged kernels only
    a.mel_bank = mel->bank;
    a.mel_bins = (int)mel->bank->mel_bins;
    a.mel_out = mel->out;
    a.mel_chmax = mel->chmax;
  }

  const bool fast = stft_has_exact_mirror(nfft);
  a.wmul = fast ? 0.5f : 1.0f;  // the split pass' 1/2 rides on the window

  if (fast) {
    // staged (TMA) configurations first; each falls back to the general kernel when the call
    // does not meet its alignment / hop conditions
#define NXS_TRY_STAGED(CF, MINB) \
  if (staged_ok<CF>(a, channels)) return run_r2c_staged<CF, MINB>(ctx, a, channels, st)
    switch (nfft) {
      case 64:  // the staged form does not pay at 256-byte frames (3.26 against 3.22 ms): tuning variant
        if (variant_env() == 16 && !a.mel_out) { using CF = StagedCfg<Plan<32, 4, 8, 4, 1, 1, true>, 128, 2, true, true>; NXS_TRY_STAGED(CF, 4); }
        return run_r2c<Plan<32, 4, 8, 4>, TwTable<Plan<32, 4, 8, 4>>, 256>(ctx, a, st);
      // nfft 128 / 256: packed fp32x2 engine by default (2.24 -> 2.20 ms, 1.90 -> 1.86 ms on 8 ch x 600 s); variant 9 = scalar
      case 128:
        if (variant_env() == 9) return run_r2c<Plan<64, 8, 8, 8>, TwTable<Plan<64, 8, 8, 8>>, 256>(ctx, a, st);
        // default: the TMA-staged per-group kernel, 8 threads per frame: 2.20 -> 1.86 ms at hop 32, 1.24 -> 0.95 ms at hop 64
        // (8 ch x 600 s); variant 18 = the general kernel
        if (variant_env() != 18 && !a.mel_out) { using CF = StagedCfg<Plan<64, 8, 8, 8, 1, 1, true>, 128, 2, true, true>; NXS_TRY_STAGED(CF, 4); }
        return run_r2c<Plan<64, 8, 8, 8, 1, 1, true>, TwTable<Plan<64, 8, 8, 8, 1, 1, true>>, 256>(ctx, a, st);
      case 256:
        if (variant_env() == 9) return run_r2c<Plan<128, 16, 8, 8, 2>, TwTable<Plan<128, 16, 8, 8, 2>>, 256>(ctx, a, st);
        // default: the TMA-staged per-group kernel with half-warp groups (16 threads per frame, __syncwarp as the group
        // barrier, 16 groups per 256-thread CTA): 1.86 -> 1.38 ms on 8 ch x 600 s at hop 64 (0.68 -> 0.92 of the HBM peak), any hop / padding offset;
        // variant 18 = the general kernel, 17 = window in registers as well (1.52 ms)
        if (variant_env() == 17 && !a.mel_out) { using CF = StagedCfg<Plan<128, 16, 8, 8, 2, 1, true>, 128, 2, true, true, false, true>; NXS_TRY_STAGED(CF, 4); }
        if (variant_env() == 16 && !a.mel_out) { using CF = StagedCfg<Plan<128, 16, 8, 8, 2, 1, true>, 128, 2, true, true>; NXS_TRY_STAGED(CF, 4); }  // 1.40 ms
        if (variant_env() != 18 && !a.mel_out) { using CF = StagedCfg<Plan<128, 16, 8, 8, 2, 1, true>, 256, 2, true, true>; NXS_TRY_STAGED(CF, 2); }
        return run_r2c<Plan<128, 16, 8, 8, 2, 1, true>, TwTable<Plan<128, 16, 8, 8, 2, 1, true>>, 256>(ctx, a, st);
      case 512: {
        using PL = Plan<256, 32, 8, 8, 4>;
        if (variant_env() == 9) { using CF = StagedCfg<PL, 256, 2, true, true>; NXS_TRY_STAGED(CF, 2); }  // scalar arithmetic
        // default: FFT engine on packed fp32x2 (Plan::PK): 1.420 -> 1.354 ms on 8 ch x 600 s (0.89 -> 0.94 of HBM peak)
        { using CF = StagedCfg<Plan<256, 32, 8, 8, 4, 1, true>, 256, 2, true, true>; NXS_TRY_STAGED(CF, 2); }
        return run_r2c<PL, TwTable<PL>, 256>(ctx, a, st);
      }
      case 1024: {
        using PL = Plan<512, 64, 8, 8, 8>;
        // tuning variants (tests/test_stft_variants_gpu.py); default = per-group staging, 256 x 2,
        // twiddles and window in registers
        const char* var = getenv("NXS_STFT_VARIANT");
        const int variant = var ? atoi(var) : 0;
        if (variant == 1) { using CF = StagedCfg<PL, 256, 2, true, false>; NXS_TRY_STAGED(CF, 2); }
        if (variant == 2) { using CF = StagedCfg<PL, 512, 2, true, true>; NXS_TRY_STAGED(CF, 1); }
        if (variant == 3) { using CF = StagedCfg<PL, 128, 2, true, true>; NXS_TRY_STAGED(CF, 4); }
        if (variant == 4) { using CF = StagedCfg<PL, 256, 2, true, true, true>; NXS_TRY_STAGED(CF, 2); }
        if (variant == 5) { using CF = StagedCfg<PL, 256, 2, true, true>; NXS_TRY_STAGED(CF, 2); }
        // one warp per frame (16 points per lane, radices 16 16 2): group barriers are warp barriers.
        // Measured slower than the default on cfg2 (1.41 - 1.51 ms vs 1.34 ms): kept as tuning variants.
        if (variant == 6) { using CF = StagedCfg<Plan<512, 32, 16, 16, 2>, 256, 2, false, true, true>; NXS_TRY_STAGED(CF, 2); }
        if (variant == 7) { using CF = StagedCfg<Plan<512, 32, 16, 16, 2>, 128, 2, false, true, true>; NXS_TRY_STAGED(CF, 4); }
        if (variant == 8) { using CF = StagedCfg<Plan<512, 32, 16, 16, 2>, 256, 2, false, true, false>; NXS_TRY_STAGED(CF, 2); }
        // FFT engine on packed fp32x2 (Plan::PK): 1.354 ms against 1.345 (with the window multiply too: 1.396) -- the kernel
        // is HBM-bound and the re-pairing moves cost more than the packed butterflies save
        if (variant == 14) { using CF = StagedCfg<Plan<512, 64, 8, 8, 8, 1, true>, 256, 2, true, true, false, true>; NXS_TRY_STAGED(CF, 2); }
        { using CF = StagedCfg<PL, 256, 2, true, true, false, true>; NXS_TRY_STAGED(CF, 2); }
        return run_r2c<PL, TwRegs<PL>, 512>(ctx, a, st);
      }
      case 2048: {
        using PL = Plan<1024, 64, 16, 8, 8>;
        if (variant_env() == 1) { using CF = StagedCfg<PL, 512, 2, false, false>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() == 2) { using CF = StagedCfg<PL, 256, 2, false, true>; NXS_TRY_STAGED(CF, 2); }
        // the paired split pass loses here (1.58 vs 1.45 ms on 8 ch x 600 s): with T = 64 the thread-0 special
        // case of the pairing diverges in every second warp; kept as a tuning variant
        if (variant_env() == 3) { using CF = StagedCfg<PL, 256, 2, false, true, true, false, true>; NXS_TRY_STAGED(CF, 2); }
        if (variant_env() == 9) { using CF = StagedCfg<PL, 256, 2, false, true, true>; NXS_TRY_STAGED(CF, 2); }  // scalar arithmetic
        // default: the FFT engine on packed fp32x2 (Plan::PK): 1.161 -> 1.111 ms on 64 ch x 60 s (with the window too: 1.114)
        { using CF = StagedCfg<Plan<1024, 64, 16, 8, 8, 1, true>, 256, 2, false, true, true>; NXS_TRY_STAGED(CF, 2); }
        return run_r2c<PL, TwTable<PL>, 512>(ctx, a, st);
      }
      case 4096: {
        using PL = Plan<2048, 128, 16, 16, 8>;
        if (variant_env() == 1) { using CF = StagedCfg<PL, 512, 4, false, false>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() == 2) { using CF = StagedCfg<PL, 256, 4, false, true>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() == 3) { using CF = StagedCfg<PL, 512, 4, false, true, true>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() == 4) { using CF = StagedCfg<PL, 256, 4, false, true, true>; NXS_TRY_STAGED(CF, 2); }
        // two warps per frame, 32 points per thread (more independent butterflies per thread, barriers between two warps only)
        if (variant_env() == 6 && !a.mel_out) { using CF = StagedCfg<Plan<2048, 64, 16, 16, 8>, 384, 4, false, true, true, false, true, false, true>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() == 7 && !a.mel_out) { using CF = StagedCfg<Plan<2048, 64, 16, 16, 8>, 256, 4, false, true, true, false, true, false, true>; NXS_TRY_STAGED(CF, 1); }
        // default for the spectrum outputs: paired split pass, two exchange buffers, one 512-thread CTA per SM
        // (cfg3 shard: 2.48 ms; paired 256 x 2 (variant 8): 2.52 - 2.64 ms; unpaired (4): 2.53 ms; P = 32 (6 / 7): 2.61 / 2.91 ms);
        // the fused log-mel epilogue keeps the single-buffer paired layout
        // scalar arithmetic = variant 9; FFT engine on the packed fp32x2 instructions (Plan::PK) = variant 10
        if (variant_env() == 9 && !a.mel_out) { using CF = StagedCfg<PL, 512, 4, false, true, true, false, true, true>; NXS_TRY_STAGED(CF, 1); }
        using PLK = Plan<2048, 128, 16, 16, 8, 1, true>;
        if (variant_env() == 10 && !a.mel_out) { using CF = StagedCfg<PLK, 512, 4, false, true, true, false, true, true, true, 0>; NXS_TRY_STAGED(CF, 1); }
        // default: engine + window multiply packed.  cfg3 shard, ms: scalar 2.39, engine 2.27, engine + window 2.21,
        // engine + split pass 2.44, all three 2.37 (the split pass's select-built operands have to be moved into pairs)
        if (variant_env() != 8 && !a.mel_out) { using CF = StagedCfg<PLK, 512, 4, false, true, true, false, true, true, true, 1>; NXS_TRY_STAGED(CF, 1); }
        { using CF = StagedCfg<PL, 256, 4, false, true, true, false, true>; NXS_TRY_STAGED(CF, 2); }
        return run_r2c<PL, TwTable<PL>, 512>(ctx, a, st);
      }
      case 8192: {
        using PL = Plan<4096, 256, 16, 16, 16>;
        // packed fp32x2: engine only 0.751 ms, engine + window (variant 15) 0.700 ms against 0.706 ms scalar: no gain, tuning only
        if (variant_env() == 15) { using CF = StagedCfg<Plan<4096, 256, 16, 16, 16, 1, true>, 512, 4, false, true, true, false, false, false, false, 1>; NXS_TRY_STAGED(CF, 1); }
        if (variant_env() != 1) { using CF = StagedCfg<PL, 512, 4, false, true, true>; NXS_TRY_STAGED(CF, 1); }
        return run_r2c<PL, TwTable<PL>, 512>(ctx, a, st);
      }
      case 16384:
        return run_r2c<Plan<8192, 512, 16, 16, 16, 2>, TwTable<Plan<8192, 512, 16, 16, 16, 2>>, 512>(ctx, a, st);
      default: break;
    }
  }
  if (a.mel_out) return NXS_EUNSUPPORTED;
  a.wmul = 1.0f;
  rc = prep_for(ctx, a, (int)nfft, st);
  if (rc) return rc;
  float2* tab = nullptr;
  rc = get_dft_table(ctx, nfft, -1, &tab);
  if (rc) return rc;
  const size_t smem = size_t(a.nload) * sizeof(float);
  if (smem > 200 * 1024) return NXS_EUNSUPPORTED;
  NXS_CUDA(ctx, cudaFuncSetAttribute(stft_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = a.total_frames < int64_t(ctx->sm_count) * 8 ? a.total_frames : int64_t(ctx->sm_count) * 8;
  stft_dft_kernel<<<(unsigned)grid, 256, smem, st>>>(a, (int)nfft, tab);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

}  // namespace nxs
