// nxs_fir.cu -- FIR filtering by overlap-save FFT convolution for sm_100a.
//
// Replaces NxSignal.Convolution.convolve(x, taps, mode:, method:) for the batched-FIR form
// x {C, L} * h {1, K} (lib/nx_signal/convolution.ex:38-58; direct :95-211, fft :252-329).  The
// reference's :fft method transforms the whole row at length L + K - 1 (:260-284); here the
// row is cut into blocks of V = F - K + 1 outputs and each block is circularly convolved at a
// fixed power-of-two F (overlap-save), which yields the same linear-convolution values.
//
// Two consecutive blocks of a channel ride one complex FFT: z = blk0 + i*blk1, Y = FFT(z) * H,
// y = IFFT(Y); because h is real, Re(y) is blk0's result and Im(y) blk1's -- no split pass.
// The plans are palindromic (first radix == last radix), so the forward transform's output
// registers are exactly the inverse transform's input registers: forward FFT, multiply by
// H / F, inverse FFT (re/im swap identity) and the store all happen without leaving the
// register file except for the in-FFT exchanges.  `mode` only shifts the output window:
// full 0, same (K-1)/2, valid min(L,K)-1 (convolution.ex:300-329).
#include <math.h>
#include <stdlib.h>

#include "nxs_common.cuh"
#include "nxs_fft.cuh"
#include "nxs_tma.cuh"

namespace nxs {

struct FirArgs {
  const float* x;
  int64_t L, x_ld;
  float* y;
  int64_t out_len, y_ld;
  int64_t start;  // y[o] = full[o + start]
  int K;
  int V;  // valid outputs per block = F - K + 1
  int64_t b_lo;
  int pairs_per_channel;
  FastDiv div_ppc;  // tile / pairs_per_channel without the emulated division (twice per block pair per thread)
  int total_tiles;
  const float2* H;  // [F], already divided by F
  const float2* PQ = nullptr;  // fir_ols_r2c_kernel: P[N] then Q[N] (fir_pq_kernel), already divided by N
  const float2* tw;
  int accumulate = 0;  // 1: y += result (later partitions of a long filter)
};

// H[k] = (1/F) sum_j h[j] exp(-2 pi i j k / F), accumulated in double.  One warp per bin k (the
// taps are split over the lanes); tab[m] = exp(+2 pi i m / F) in double (nxs_common: get_dft_table_f64).
__global__ void __launch_bounds__(256) fir_spectrum_kernel(const float* __restrict__ taps, int K, int F,
                                                           const double2* __restrict__ tab, float2* __restrict__ H) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= F) return;
  double re = 0.0, im = 0.0;
  int idx = (int)(((int64_t)lane * k) % F);
  const int step = (int)(((int64_t)32 * k) % F);
  for (int j = lane; j < K; j += 32) {
    const double2 e = tab[idx];
    const double h = (double)taps[j];
    re += h * e.x;
    im -= h * e.y;
    idx += step;
    if (idx >= F) idx -= F;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if (lane == 0) H[k] = make_float2((float)(re / F), (float)(im / F));
}

static int launch_fir_spectrum(nxs_ctx* ctx, const float* taps, int K, int F, float2* H, cudaStream_t st) {
  double2* tab = nullptr;
  int rc = get_dft_table_f64(ctx, F, &tab);
  if (rc) return rc;
  fir_spectrum_kernel<<<(F + 7) / 8, 256, 0, st>>>(taps, K, F, tab, H);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fir_ols_kernel(const FirArgs a) {
  constexpr int F = PL::N, T = PL::T, P = PL::P, G = THREADS / T;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1);
  static_assert(R0 == RL, "FIR plans must be palindromic (first radix == last radix)");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const bufA = reinterpret_cast<cpx*>(smem_raw) + (size_t)(2 * g) * PL::BUF;
  cpx* const bufB = bufA + PL::BUF;
  cpx* twsm = reinterpret_cast<cpx*>(smem_raw + size_t(G) * 2 * PL::BUF * sizeof(cpx));
  for (int i = tid; i < PL::TW_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  __syncthreads();
  TwTable<PL> tw;
  tw.init(twsm, t);
  const SyncBlock sync;
  const int K1 = a.K - 1;

  const int iters = (a.total_tiles + G - 1) / G;
  for (int it = blockIdx.x; it < iters; it += gridDim.x) {
    const int tile = it * G + g;
    const bool active = tile < a.total_tiles;
    int c = 0, pi = 0;
    if (active) {
      c = a.div_ppc.div(tile);
      pi = tile - c * a.pairs_per_channel;
    }
    const float* __restrict__ xrow = a.x + (int64_t)c * a.x_ld;
    const int64_t n0 = (a.b_lo + 2 * (int64_t)pi) * a.V;  // first full-convolution index of block 0
    const int64_t s0 = n0 - K1, s1 = s0 + a.V;           // input segment starts of the two blocks
    cpx v[P];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int i = fft_in_index<PL>(t, b, q);
        float re = 0.f, im = 0.f;
        if (active) {
          const int64_t i0 = s0 + i, i1 = s1 + i;
          if (i0 >= 0 && i0 < a.L) re = __ldg(xrow + i0);
          if (i1 >= 0 && i1 < a.L) im = __ldg(xrow + i1);
        }
        v[b * R0 + q] = make_float2(re, im);
      }
    __syncthreads();  // the previous iteration's exchange reads are complete
    block_fft<PL>(v, t, bufA, bufB, tw, sync);
    // pointwise multiply by H / F; palindromic plan: output register (b, q) is input register (b, q)
    cpx u[P];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int k = fft_out_index<PL>(t, b, q);
        const cpx yk = cmul(v[fft_out_reg<PL>(b, q)], __ldg(a.H + k));
        u[b * R0 + q] = make_float2(yk.y, yk.x);  // swap for the inverse transform
      }
    __syncthreads();  // forward exchanges fully consumed before the inverse reuses the buffers
    block_fft<PL>(u, t, bufA, bufB, tw, sync);
    if (active) {
      float* __restrict__ yrow = a.y + (int64_t)c * a.y_ld;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          if (i >= K1) {
            const cpx r = u[fft_out_reg<PL>(b, q)];
            const int64_t o0 = n0 + (i - K1) - a.start, o1 = o0 + a.V;
            if (o0 >= 0 && o0 < a.out_len) __stcs(yrow + o0, r.y);  // Re(y): block 0
            if (o1 >= 0 && o1 < a.out_len) __stcs(yrow + o1, r.x);  // Im(y): block 1
          }
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Per-group variant (the hot path for K > 129): every group of T threads walks its own block
// pairs with no CTA-wide barrier.  The pair's input span (V + F samples, contiguous) arrives by
// one cp.async.bulk (TMA 1-D, from the 16-byte-aligned address below the span) into the group's
// stage buffer, re-issued for the next pair as soon as this pair has been read into registers.
// One exchange buffer per group, H / F and the compact twiddle table (power-of-two rows, the
// other twiddles are products) in shared memory: the kernel is bounded by the shared-memory
// pipe, so every table load saved is time saved.
// ------------------------------------------------------------------------------------------
template <class PL, int THREADS>
struct FirPgCfg {
  static constexpr int F = PL::N, G = THREADS / PL::T;
  // the stage holds one pair's span: V + F samples plus alignment slack (a multiple of 4 floats)
  static __host__ __device__ int stage_floats(int V) { return (V + F + 3 + 3) / 4 * 4 + 4; }
  static __host__ __device__ size_t group_bytes(int V) {
    return size_t(stage_floats(V)) * sizeof(float) + size_t(PL::BUF) * sizeof(cpx);
  }
  static __host__ __device__ size_t h_off(int V) { return size_t(G) * group_bytes(V); }
  static __host__ __device__ size_t tw_off(int V) { return h_off(V) + size_t(F) * sizeof(cpx); }
  static __host__ __device__ size_t bar_off(int V) { return tw_off(V) + size_t(PL::TWC_TOTAL) * sizeof(cpx); }
  static __host__ __device__ size_t smem(int V) { return bar_off(V) + 8 * size_t(G) + 8; }
};

template <class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fir_ols_pg_kernel(const FirArgs a, const int aligned_rows) {
  using CF = FirPgCfg<PL, THREADS>;
  constexpr int F = PL::N, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1);
  static_assert(R0 == RL, "FIR plans must be palindromic (first radix == last radix)");
  static_assert(T >= 32, "per-group barriers need whole warps");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  float* const stage = reinterpret_cast<float*>(smem_raw + size_t(g) * CF::group_bytes(a.V));
  cpx* const xbuf = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::group_bytes(a.V) +
                                           size_t(CF::stage_floats(a.V)) * sizeof(float));
  cpx* const Hs = reinterpret_cast<cpx*>(smem_raw + CF::h_off(a.V));
  cpx* const twsm = reinterpret_cast<cpx*>(smem_raw + CF::tw_off(a.V));
  const uint32_t mybar = smem_u32(smem_raw + CF::bar_off(a.V)) + 8 * g;
  for (int i = tid; i < F; i += THREADS) Hs[i] = a.H[i];
  for (int i = tid; i < PL::TWC_TOTAL; i += THREADS) twsm[i] = a.tw[i];
  if (tid == 0) {
    for (int i = 0; i < G; ++i) mbar_init(smem_u32(smem_raw + CF::bar_off(a.V)) + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TwDeriveC<PL> tw;
  tw.init(twsm, t);
  const GroupSync<T> sync{1 + g};
  const int K1 = a.K - 1;

  // geometry of a block pair; returns whether its span can be staged by TMA
  auto geom = [&](int tile, int& c, int64_t& n0, int64_t& s0a, int& off, uint32_t& bytes) {
    c = a.div_ppc.div(tile);
    const int pi = tile - c * a.pairs_per_channel;
    n0 = (a.b_lo + 2 * (int64_t)pi) * a.V;
    const int64_t s0 = n0 - K1;
    s0a = s0 & ~(int64_t)3;  // floor to a multiple of 4 samples (also for negative s0)
    off = (int)(s0 - s0a);
    const int len4 = (off + a.V + F + 3) & ~3;
    bytes = (uint32_t)len4 * (uint32_t)sizeof(float);
    return aligned_rows && s0a >= 0 && s0a + len4 <= a.L;
  };
  auto issue = [&](int tile) {
    int c, off;
    int64_t n0, s0a;
    uint32_t bytes;
    if (geom(tile, c, n0, s0a, off, bytes)) {
      mbar_expect_tx(mybar, bytes);
      tma_load_1d(smem_u32(stage), a.x + (int64_t)c * a.x_ld + s0a, bytes, mybar);
    }
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  uint32_t parity = 0;
  int tile = gid;
  if (tile < a.total_tiles && t == 0) issue(tile);
  for (; tile < a.total_tiles; tile += ngroups) {
    int c, off;
    int64_t n0, s0a;
    uint32_t bytes;
    const bool staged = geom(tile, c, n0, s0a, off, bytes);
    cpx v[P];
    if (staged) {
      mbar_wait(mybar, parity);
      parity ^= 1;
      const float* __restrict__ p0 = stage + off;
      const float* __restrict__ p1 = p0 + a.V;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_in_index<PL>(t, b, q);
          v[b * R0 + q] = make_float2(p0[i], p1[i]);
        }
    } else {
      const float* __restrict__ xrow = a.x + (int64_t)c * a.x_ld;
      const int64_t s0 = n0 - K1, s1 = s0 + a.V;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_in_index<PL>(t, b, q);
          const int64_t i0 = s0 + i, i1 = s1 + i;
          float re = 0.f, im = 0.f;
          if (i0 >= 0 && i0 < a.L) re = __ldg(xrow + i0);
          if (i1 >= 0 && i1 < a.L) im = __ldg(xrow + i1);
          v[b * R0 + q] = make_float2(re, im);
        }
    }
    sync();  // stage read out; the previous pair's last exchange reads are done
    if (t == 0 && tile + ngroups < a.total_tiles) issue(tile + ngroups);
    block_fft_single<PL>(v, t, xbuf, tw, sync);
    // pointwise multiply by H / F; palindromic plan: output register (b, q) is input register (b, q)
    cpx u[P];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int k = fft_out_index<PL>(t, b, q);
        const cpx yk = cmul(v[fft_out_reg<PL>(b, q)], Hs[k]);
        u[b * R0 + q] = make_float2(yk.y, yk.x);  // swap for the inverse transform
      }
    sync();  // forward transform's last exchange fully consumed before the inverse reuses the buffer
    block_fft_single<PL>(u, t, xbuf, tw, sync);
    float* __restrict__ yrow = a.y + (int64_t)c * a.y_ld;
    const int64_t obase = n0 - K1 - a.start;  // output index of FFT sample i is obase + i (block 0)
    if (obase + K1 >= 0 && obase + F + a.V <= a.out_len) {
      // interior pair: every kept sample of both blocks lands inside the row
      float* __restrict__ y0 = yrow + obase;
      float* __restrict__ y1 = y0 + a.V;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          if (i >= K1) {
            cpx r = u[fft_out_reg<PL>(b, q)];
            if (a.accumulate) {
              r.y += y0[i];
              r.x += y1[i];
            }
            __stcs(y0 + i, r.y);  // Re(y): block 0
            __stcs(y1 + i, r.x);  // Im(y): block 1
          }
        }
    } else {
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          if (i >= K1) {
            const cpx r = u[fft_out_reg<PL>(b, q)];
            const int64_t o0 = obase + i, o1 = o0 + a.V;
            if (o0 >= 0 && o0 < a.out_len) __stcs(yrow + o0, a.accumulate ? r.y + yrow[o0] : r.y);
            if (o1 >= 0 && o1 < a.out_len) __stcs(yrow + o1, a.accumulate ? r.x + yrow[o1] : r.x);
          }
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Real-packed overlap-save (the hot path for K > 513).  The two kernels above spend one complex
// FFT of F points on two real blocks and keep F - K + 1 outputs of each: at F = 4096, K = 2049 half
// of every inverse transform is thrown away.  Here ONE real block of 2N samples is packed as N
// complex points z[i] = x[2i] + i x[2i+1] (the same N = 4096-point engine, the same registers), so a
// block keeps 2N - K + 1 = 6144 outputs instead of 2 x 2048 for the same two transforms.  The
// spectral multiply becomes "widely linear": with He / Ho the N-point DFTs of the even / odd taps,
//     Z'[k] = P[k] Z[k] + Q[k] conj(Z[N - k]),
//     P = He + (i/2)(1 - e^{-2 pi i k / N}) Ho,   Q = (i/2)(1 + e^{-2 pi i k / N}) Ho
// (the polyphase form ye = he * xe + z^-1 ho * xo, yo = ho * xe + he * xo written on the packed
// spectra), and the inverse transform of Z' is the circular convolution, packed the same way.  No
// split passes: two complex multiply-adds per bin and one exchange for conj(Z[N - k]).
// Per block 2 FFTs + 8 flops per bin for 6144 outputs: ~28 % fewer instructions per output sample
// than the pair kernel (cfg4: 64 ch x 600 s, K = 2049).
// The group's exchange buffer doubles as its TMA stage: the kernel is issue-bound with three groups
// per SM, so a group waiting ~1 us for its own block costs nothing while the others compute.
// ------------------------------------------------------------------------------------------
// P / Q tables in double: one warp per bin k < N; tab[m] = exp(+2 pi i m / N) (get_dft_table_f64)
__global__ void __launch_bounds__(256) fir_pq_kernel(const float* __restrict__ taps, int K, int N,
                                                     const double2* __restrict__ tab, float2* __restrict__ PQ) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= N) return;
  double er = 0.0, ei = 0.0, orr = 0.0, oi = 0.0;
  // tap j = 2 m (+1) contributes h[j] e^{-2 pi i m k / N}
  const int M2 = (K + 1) / 2;
  int idx = (int)(((int64_t)lane * k) % N);
  const int step = (int)(((int64_t)32 * k) % N);
  for (int m = lane; m < M2; m += 32) {
    const double2 e = tab[idx];
    const double h0 = (double)taps[2 * m];
    const double h1 = 2 * m + 1 < K ? (double)taps[2 * m + 1] : 0.0;
    er += h0 * e.x;
    ei -= h0 * e.y;
    orr += h1 * e.x;
    oi -= h1 * e.y;
    idx += step;
    if (idx >= N) idx -= N;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    er += __shfl_xor_sync(0xffffffffu, er, o);
    ei += __shfl_xor_sync(0xffffffffu, ei, o);
    orr += __shfl_xor_sync(0xffffffffu, orr, o);
    oi += __shfl_xor_sync(0xffffffffu, oi, o);
  }
  if (lane == 0) {
    const double2 w = tab[k];  // e^{+i theta}; e^{-i theta} = (w.x, -w.y)
    // a = (i/2)(1 - e^{-i theta}) = (i/2)(1 - w.x + i w.y) = (-w.y/2, (1 - w.x)/2);  b = (i/2)(1 + e^{-i theta}) = (w.y/2, (1 + w.x)/2)
    const double ar = -0.5 * w.y, ai = 0.5 * (1.0 - w.x), br = 0.5 * w.y, bi = 0.5 * (1.0 + w.x);
    const double pr = er + ar * orr - ai * oi, pi = ei + ar * oi + ai * orr;
    const double qr = br * orr - bi * oi, qi = br * oi + bi * orr;
    PQ[k] = make_float2((float)(pr / N), (float)(pi / N));
    PQ[N + k] = make_float2((float)(qr / N), (float)(qi / N));
  }
}

template <class PL, int THREADS>
struct FirR2cCfg {
  static constexpr int N = PL::N, G = THREADS / PL::T;
  static_assert(size_t(PL::BUF) * sizeof(cpx) >= size_t(2 * N + 8) * sizeof(float), "the exchange buffer must hold a staged block");
  static constexpr size_t GROUP_BYTES = size_t(PL::BUF) * sizeof(cpx);
  static constexpr size_t PQ_OFF = size_t(G) * GROUP_BYTES;
  static constexpr size_t TW_OFF = PQ_OFF + size_t(2 * N) * sizeof(cpx);
  static constexpr size_t BAR_OFF = TW_OFF + size_t(PL::TWC_TOTAL) * sizeof(cpx);
  static constexpr size_t SMEM = BAR_OFF + 8 * size_t(G) + 8;
};

template <class PL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fir_ols_r2c_kernel(const FirArgs a, const int aligned_rows) {
  using CF = FirR2cCfg<PL, THREADS>;
  constexpr int N = PL::N, F2 = 2 * N, T = PL::T, P = PL::P, G = CF::G;
  constexpr int R0 = PL::R(0), B0 = P / R0;
  constexpr int RL = PL::R(PL::NP - 1);
  static_assert(R0 == RL, "FIR plans must be palindromic (first radix == last radix)");
  static_assert(T >= 32, "per-group barriers need whole warps");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  cpx* const xbuf = reinterpret_cast<cpx*>(smem_raw + size_t(g) * CF::GROUP_BYTES);
  float* const stage = reinterpret_cast<float*>(xbuf);  // the block's 2N samples land where the exchanges happen later
  const cpx* const Ps = reinterpret_cast<const cpx*>(smem_raw + CF::PQ_OFF);
  const cpx* const Qs = Ps + N;
  cpx* const twsm = reinterpret_cast<cpx*>(smem_raw + CF::TW_OFF);
  const uint32_t mybar = smem_u32(smem_raw + CF::BAR_OFF) + 8 * g;
  {
    cpx* pq = reinterpret_cast<cpx*>(smem_raw + CF::PQ_OFF);
    for (int i = tid; i < 2 * N; i += THREADS) pq[i] = a.PQ[i];
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
  const int K1 = a.K - 1;

  // geometry of a block; returns whether its 2N-sample span can be staged by TMA (copied from the
  // 16-byte boundary at or below it; the frame then sits at a 0..3 sample offset inside the stage)
  auto geom = [&](int tile, int& c, int64_t& n0, int64_t& s0a, int& off, uint32_t& bytes) {
    c = a.div_ppc.div(tile);
    const int bi = tile - c * a.pairs_per_channel;
    n0 = (a.b_lo + (int64_t)bi) * a.V;  // first full-convolution index the block produces
    const int64_t s0 = n0 - K1;
    s0a = s0 & ~(int64_t)3;
    off = (int)(s0 - s0a);
    const int len4 = (off + F2 + 3) & ~3;
    bytes = (uint32_t)len4 * (uint32_t)sizeof(float);
    return aligned_rows && s0a >= 0 && s0a + len4 <= a.L;
  };
  auto issue = [&](int tile) {
    int c, off;
    int64_t n0, s0a;
    uint32_t bytes;
    if (geom(tile, c, n0, s0a, off, bytes)) {
      // the buffer was last written by this group's generic-proxy stores (the inverse transform's
      // exchanges): order them before the bulk copy's async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(mybar, bytes);
      tma_load_1d(smem_u32(stage), a.x + (int64_t)c * a.x_ld + s0a, bytes, mybar);
    }
  };

  const int gid = blockIdx.x * G + g, ngroups = gridDim.x * G;
  uint32_t parity = 0;
  int tile = gid;
  if (tile < a.total_tiles && t == 0) issue(tile);
  for (; tile < a.total_tiles; tile += ngroups) {
    int c, off;
    int64_t n0, s0a;
    uint32_t bytes;
    const bool staged = geom(tile, c, n0, s0a, off, bytes);
    cpx v[P];
    if (staged) {
      mbar_wait(mybar, parity);
      parity ^= 1;
      if (!(off & 1)) {
        const float2* __restrict__ p2 = reinterpret_cast<const float2*>(stage + off);
#pragma unroll
        for (int b = 0; b < B0; ++b)
#pragma unroll
          for (int q = 0; q < R0; ++q) v[b * R0 + q] = p2[fft_in_index<PL>(t, b, q)];
      } else {  // odd sample offset: the (even, odd) pairs are not 8-byte aligned in the stage
        const float* __restrict__ p1 = stage + off;
#pragma unroll
        for (int b = 0; b < B0; ++b)
#pragma unroll
          for (int q = 0; q < R0; ++q) {
            const int i = fft_in_index<PL>(t, b, q);
            v[b * R0 + q] = make_float2(p1[2 * i], p1[2 * i + 1]);
          }
      }
    } else {
      const float* __restrict__ xrow = a.x + (int64_t)c * a.x_ld;
      const int64_t s0 = n0 - K1;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int64_t i0 = s0 + 2 * fft_in_index<PL>(t, b, q), i1 = i0 + 1;
          float re = 0.f, im = 0.f;
          if (i0 >= 0 && i0 < a.L) re = __ldg(xrow + i0);
          if (i1 >= 0 && i1 < a.L) im = __ldg(xrow + i1);
          v[b * R0 + q] = make_float2(re, im);
        }
    }
    sync();  // the staged block is in registers
    block_fft_single<PL>(v, t, xbuf, tw, sync);
    sync();  // last pass's reads done before Z overwrites the buffer
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) xbuf[fft_out_index<PL>(t, b, q)] = v[fft_out_reg<PL>(b, q)];
    sync();
    // Z'[k] = P[k] Z[k] + Q[k] conj(Z[N - k]); palindromic plan: output register (b, q) is input register (b, q)
    cpx u[P];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int k = fft_out_index<PL>(t, b, q);
        const cpx A = v[fft_out_reg<PL>(b, q)];
        const cpx Bc = cconj(xbuf[(N - k) & (N - 1)]);
        const cpx pk = Ps[k], qk = Qs[k];
        if constexpr (PL::PK) {  // (yi, yr) = B.y (-qk.x, qk.y) + B.x (qk.y, qk.x) + A.y (pk.x, -pk.y) + A.x (pk.y, pk.x), B = conj(Bc)
          cpx s = __fmul2_rn(make_float2(-qk.x, qk.y), make_float2(-Bc.y, -Bc.y));
          s = __ffma2_rn(make_float2(qk.y, qk.x), make_float2(Bc.x, Bc.x), s);
          s = __ffma2_rn(make_float2(pk.x, -pk.y), make_float2(A.y, A.y), s);
          u[b * R0 + q] = __ffma2_rn(make_float2(pk.y, pk.x), make_float2(A.x, A.x), s);
        } else {
          const float yr = fmaf(pk.x, A.x, fmaf(-pk.y, A.y, fmaf(qk.x, Bc.x, -qk.y * Bc.y)));
          const float yi = fmaf(pk.x, A.y, fmaf(pk.y, A.x, fmaf(qk.x, Bc.y, qk.y * Bc.x)));
          u[b * R0 + q] = make_float2(yi, yr);  // swap for the inverse transform
        }
      }
    sync();  // every partner read done before the inverse transform's exchanges
    block_fft_single<PL>(u, t, xbuf, tw, sync);
    sync();  // the buffer is free: stage the group's next block while this one is stored
    if (t == 0 && tile + ngroups < a.total_tiles) issue(tile + ngroups);
    // sample n of the block sits in u at pair i = n / 2: (swap identity) .y = even sample, .x = odd sample
    float* __restrict__ yrow = a.y + (int64_t)c * a.y_ld;
    const int64_t obase = n0 - K1 - a.start;  // output index of block sample n is obase + n
    const bool interior = obase + K1 >= 0 && obase + F2 <= a.out_len;
    if (interior && !(K1 & 1) && ((reinterpret_cast<uintptr_t>(yrow + obase) & 7) == 0)) {
      float2* __restrict__ y2 = reinterpret_cast<float2*>(yrow + obase);
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          if (2 * i >= K1) {
            const cpx r = u[fft_out_reg<PL>(b, q)];
            float2 o = make_float2(r.y, r.x);
            if (a.accumulate) {
              const float2 old = y2[i];
              o.x += old.x;
              o.y += old.y;
            }
            __stcs(y2 + i, o);
          }
        }
    } else if (interior) {  // odd output offset (or odd K - 1): element stores, no bounds to check
      float* __restrict__ y0 = yrow + obase;
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          const cpx r = u[fft_out_reg<PL>(b, q)];
          if (2 * i >= K1) __stcs(y0 + 2 * i, a.accumulate ? r.y + y0[2 * i] : r.y);
          if (2 * i + 1 >= K1) __stcs(y0 + 2 * i + 1, a.accumulate ? r.x + y0[2 * i + 1] : r.x);
        }
    } else {
#pragma unroll
      for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int q = 0; q < R0; ++q) {
          const int i = fft_out_index<PL>(t, b, q);
          const cpx r = u[fft_out_reg<PL>(b, q)];
          const int64_t o0 = obase + 2 * i, o1 = o0 + 1;
          if (2 * i >= K1 && o0 >= 0 && o0 < a.out_len) __stcs(yrow + o0, a.accumulate ? r.y + yrow[o0] : r.y);
          if (2 * i + 1 >= K1 && o1 >= 0 && o1 < a.out_len) __stcs(yrow + o1, a.accumulate ? r.x + yrow[o1] : r.x);
        }
    }
  }
}

template <class PL>
static int fir_tw_table(nxs_ctx* ctx, float2** out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP) | (uint64_t(2) << 62) |
                       (uint64_t(PL::R(0)) << 20);
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
static int run_fir(nxs_ctx* ctx, FirArgs a, int64_t channels, const float* taps, cudaStream_t st) {
  constexpr int F = PL::N, G = THREADS / PL::T;
  float2* tw = nullptr;
  int rc = fir_tw_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  rc = ensure_scratch(ctx, size_t(F) * sizeof(float2));
  if (rc) return rc;
  rc = launch_fir_spectrum(ctx, taps, a.K, F, (float2*)ctx->d_scratch, st);
  if (rc) return rc;
  a.H = (const float2*)ctx->d_scratch;
  a.V = F - a.K + 1;
  a.b_lo = a.start / a.V;
  const int64_t b_hi = (a.start + a.out_len - 1) / a.V;
  const int64_t nblocks = b_hi - a.b_lo + 1;
  const int64_t pairs = (nblocks + 1) / 2;
  const int64_t tiles = pairs * channels;
  if (tiles >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.pairs_per_channel = (int)pairs;
  a.div_ppc = FastDiv((int)pairs);
  a.total_tiles = (int)tiles;
  const size_t smem = size_t(G) * 2 * PL::BUF * sizeof(cpx) + size_t(PL::TW_TOTAL) * sizeof(cpx);
  auto kern = fir_ols_kernel<PL, THREADS, MINB>;
  NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  NXS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (occ < 1) occ = 1;
  const int64_t iters = (tiles + G - 1) / G;
  int64_t grid = int64_t(ctx->sm_count) * occ;
  if (grid > iters) grid = iters;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL>
static int fir_twc_table(nxs_ctx* ctx, float2** out) {
  const uint64_t key = (uint64_t(PL::N) << 32) | (uint64_t(PL::T) << 8) | uint64_t(PL::NP) | (uint64_t(3) << 62) |
                       (uint64_t(PL::R(0)) << 20);
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

template <class PL, int THREADS, int MINB>
static int run_fir_pg(nxs_ctx* ctx, FirArgs a, int64_t channels, const float* taps, cudaStream_t st) {
  using CF = FirPgCfg<PL, THREADS>;
  constexpr int F = PL::N;
  float2* tw = nullptr;
  int rc = fir_twc_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  rc = ensure_scratch(ctx, size_t(F) * sizeof(float2));
  if (rc) return rc;
  rc = launch_fir_spectrum(ctx, taps, a.K, F, (float2*)ctx->d_scratch, st);
  if (rc) return rc;
  a.H = (const float2*)ctx->d_scratch;
  a.V = F - a.K + 1;
  // blocks whose outputs fall inside [0, out_len): full-convolution indices [start, start + out_len),
  // clipped at 0 (a later partition of a long filter has start < 0: its first outputs do not exist)
  if (a.start + a.out_len <= 0) return NXS_OK;
  a.b_lo = (a.start > 0 ? a.start : 0) / a.V;
  const int64_t b_hi = (a.start + a.out_len - 1) / a.V;
  const int64_t pairs = (b_hi - a.b_lo + 2) / 2;
  const int64_t tiles = pairs * channels;
  if (tiles >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.pairs_per_channel = (int)pairs;
  a.div_ppc = FastDiv((int)pairs);
  a.total_tiles = (int)tiles;
  auto kern = fir_ols_pg_kernel<PL, THREADS, MINB>;
  const size_t smem = CF::smem(a.V);
  if (smem > 232448) return NXS_EUNSUPPORTED;
  NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  NXS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (occ < 1) occ = 1;
  int64_t grid = int64_t(ctx->sm_count) * occ;
  const int64_t need = (tiles + CF::G - 1) / CF::G;
  if (grid > need) grid = need;
  // TMA needs 16-byte aligned spans: row starts aligned (the kernel floors each span's start itself)
  const int aligned_rows = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.x_ld % 4 == 0;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a, aligned_rows);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

template <class PL, int THREADS, int MINB>
static int run_fir_r2c(nxs_ctx* ctx, FirArgs a, int64_t channels, const float* taps, cudaStream_t st) {
  using CF = FirR2cCfg<PL, THREADS>;
  constexpr int N = PL::N;
  float2* tw = nullptr;
  int rc = fir_twc_table<PL>(ctx, &tw);
  if (rc) return rc;
  a.tw = tw;
  rc = ensure_scratch(ctx, size_t(2 * N) * sizeof(float2));
  if (rc) return rc;
  double2* tab = nullptr;
  rc = get_dft_table_f64(ctx, N, &tab);
  if (rc) return rc;
  fir_pq_kernel<<<(N + 7) / 8, 256, 0, st>>>(taps, a.K, N, tab, (float2*)ctx->d_scratch);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  a.PQ = (const float2*)ctx->d_scratch;
  a.V = 2 * N - a.K + 1;
  if (a.start + a.out_len <= 0) return NXS_OK;
  a.b_lo = (a.start > 0 ? a.start : 0) / a.V;
  const int64_t b_hi = (a.start + a.out_len - 1) / a.V;
  const int64_t blocks = b_hi - a.b_lo + 1;
  const int64_t tiles = blocks * channels;
  if (tiles >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  a.pairs_per_channel = (int)blocks;  // one block per tile here
  a.div_ppc = FastDiv((int)blocks);
  a.total_tiles = (int)tiles;
  auto kern = fir_ols_r2c_kernel<PL, THREADS, MINB>;
  const size_t smem = CF::SMEM;
  static_assert(CF::SMEM <= 232448, "fir_ols_r2c_kernel: shared memory");
  static int attr_done[16] = {0};
  if (!attr_done[ctx->device & 15]) {
    NXS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[ctx->device & 15] = 1;
  }
  int64_t grid = int64_t(ctx->sm_count) * MINB;
  const int64_t need = (tiles + CF::G - 1) / CF::G;
  if (grid > need) grid = need;
  const int aligned_rows = (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.x_ld % 4 == 0;
  prof_begin(ctx, st);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(a, aligned_rows);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// short filters / very long filters: direct summation, one thread per output (fp32 FMA chain in
// ascending tap order)
__global__ void __launch_bounds__(256) fir_direct_kernel(const float* __restrict__ x, int64_t channels, int64_t L,
                                                         int64_t x_ld, const float* __restrict__ taps, int K,
                                                         int64_t start, int64_t out_len, int64_t y_ld,
                                                         float* __restrict__ y) {
  const int64_t total = channels * out_len;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i % out_len, c = i / out_len;
    const int64_t n = o + start;
    const float* __restrict__ xr = x + c * x_ld;
    int64_t k_lo = n - (L - 1);
    if (k_lo < 0) k_lo = 0;
    int64_t k_hi = n < K - 1 ? n : K - 1;
    float acc = 0.f;
    for (int64_t k = k_lo; k <= k_hi; ++k) acc = fmaf(__ldg(taps + k), __ldg(xr + (n - k)), acc);
    y[c * y_ld + o] = acc;
  }
}

int launch_fir(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld, const float* taps,
               int64_t num_taps, int mode, float* y, int64_t y_ld, cudaStream_t st) {
  if (channels <= 0) return NXS_OK;
  int64_t out_len = 0;
  int rc = nxs_fir_out_len(length, num_taps, mode, &out_len);
  if (rc) return rc;
  FirArgs a;
  a.x = x;
  a.L = length;
  a.x_ld = x_ld;
  a.y = y;
  a.out_len = out_len;
  a.y_ld = y_ld;
  a.K = (int)num_taps;
  a.start = mode == NXS_MODE_FULL ? 0
            : mode == NXS_MODE_SAME ? (num_taps - 1) / 2
                                    : (length < num_taps ? length : num_taps) - 1;
  a.V = 0;
  a.b_lo = 0;
  a.pairs_per_channel = a.total_tiles = 0;
  a.H = nullptr;
  a.tw = nullptr;
  const int64_t K = num_taps;
  const bool pg = !getenv("NXS_FIR_NO_PG");
  const char* var = getenv("NXS_FIR_VARIANT");  // tuning variants (tests/test_fir_conv_gpu.py)
  const int variant = var ? atoi(var) : 0;
  // short filters: F = 256 blocks keep only 256 - K + 1 of every 256 samples; on long rows the
  // per-group F = 1024 kernel (V = 1025 - K, 88 - 98 % kept) is the better trade
  const bool short_on_long_rows = K >= 16 && K <= 129 && pg && variant != 2 && a.out_len >= 8192;
  if (K >= 16 && K <= 129 && !short_on_long_rows) return run_fir<Plan<256, 16, 16, 16>, 256, 2>(ctx, a, channels, taps, st);
  // mid-size filters on long rows take the real-packed N = 4096 kernel below from K = 320 on (94 - 96 % of every
  // 8192-sample block kept).  Measured on 64 ch x 600 s, pair kernel (F = 1024) / real-packed N = 4096 / real-packed
  // N = 1024 (variant 6: small transforms, little work per barrier): K = 193 4.56 / 4.83 / - ms, K = 255 4.95 / 5.47 /
  // 5.77, K = 385 5.84 / 4.93 / 5.14, K = 513 7.23 / 5.01 / 5.56 (profiles/r02_fir_timings.txt)
  const bool mid_r2c = K >= 320 && K <= 513 && a.out_len >= 32768;
  if (mid_r2c && pg && variant == 6) return run_fir_r2c<Plan<1024, 64, 8, 16, 8>, 384, 2>(ctx, a, channels, taps, st);
  const bool r2c_mid = mid_r2c && pg && variant != 1 && variant != 3;  // served by the real-packed block below
  if (!r2c_mid && (short_on_long_rows || (K > 129 && K <= 513))) {
    if (pg && variant == 1) return run_fir_pg<Plan<1024, 64, 8, 16, 8>, 256, 2>(ctx, a, channels, taps, st);
    // packed fp32x2 butterflies (Plan::PK): 4.94 -> 4.53 ms at K = 255 on 64 ch x 600 s (profiles/r02x_packed_fp32x2.txt)
    if (pg) return run_fir_pg<Plan<1024, 64, 8, 16, 8, 1, true>, 384, 2>(ctx, a, channels, taps, st);
    return run_fir<Plan<1024, 64, 8, 16, 8>, 256, 2>(ctx, a, channels, taps, st);
  }
  if ((K > 513 || mid_r2c) && pg && variant != 1 && variant != 3) {
    // real-packed overlap-save: one 8192-sample real block per 4096-point transform pair, V = 8193 - KP
    // outputs kept; long filters are cut into n equal runs of KP taps whose partial convolutions are
    // accumulated, y[n] += (x * h_p)[n - p KP] (work per output ~ n / (8193 - K / n); later passes
    // read-modify-write y).  Packed fp32x2 butterflies (Plan::PK): cfg4 6.15 -> 5.82 ms.
    using PL = Plan<4096, 256, 16, 16, 16, 1, true>;
    int64_t n_best = 1;
    double c_best = 1e300;
    for (int64_t n = 1; n <= (K + 1023) / 1024; ++n) {
      const int64_t kp = (K + n - 1) / n;
      if (kp > 7169) continue;
      const double c = (double(n) + 0.15 * double(n - 1)) / double(8193 - kp);
      if (c < c_best) {
        c_best = c;
        n_best = n;
      }
    }
    const int64_t KP = (K + n_best - 1) / n_best;
    for (int64_t p = 0; p * KP < K; ++p) {
      FirArgs ap = a;
      ap.K = (int)(K - p * KP < KP ? K - p * KP : KP);
      ap.start = a.start - p * KP;
      ap.accumulate = p > 0;
      // variant 7: two warps per block pair, 64 points per thread, radices 64 x 64 -- ONE exchange per transform
      // variant 8: the default layout on scalar arithmetic (A/B against the packed fp32x2 plan)
      rc = variant == 8 ? run_fir_r2c<Plan<4096, 256, 16, 16, 16>, 768, 1>(ctx, ap, channels, taps + p * KP, st)
           : variant == 7 ? run_fir_r2c<Plan<4096, 64, 64, 64, 1, 1, true>, 256, 1>(ctx, ap, channels, taps + p * KP, st)
           : variant == 4 ? run_fir_r2c<PL, 512, 1>(ctx, ap, channels, taps + p * KP, st)
                          : run_fir_r2c<PL, 768, 1>(ctx, ap, channels, taps + p * KP, st);
      if (rc) return rc;
    }
    return NXS_OK;
  }
  if (K > 513 && (K <= 3585 || pg)) {
    using PL = Plan<4096, 256, 16, 16, 16>;
    if (!pg) return run_fir<PL, 256, 2>(ctx, a, channels, taps, st);
    // Long filters are cut into n equal runs of KP taps and the partial convolutions accumulated,
    // y[n] += (x * h_p)[n - p KP]: a block of F = 4096 yields 4097 - KP outputs, so the work per
    // output is ~ n / (4097 - K / n); n = 1 up to K ~ 2900, then 2, ...  (K = 3585 in one pass would
    // keep only 512 of every 4096 samples.)
    int64_t n_best = 1;
    double c_best = 1e300;
    for (int64_t n = 1; n <= (K + 511) / 512; ++n) {
      const int64_t kp = (K + n - 1) / n;
      if (kp > 3585) continue;
      const double c = (double(n) + 0.15 * double(n - 1)) / double(4097 - kp);  // later passes read-modify-write y
      if (c < c_best) {
        c_best = c;
        n_best = n;
      }
    }
    const int64_t KP = (K + n_best - 1) / n_best;
    for (int64_t p = 0; p * KP < K; ++p) {
      FirArgs ap = a;
      ap.K = (int)(K - p * KP < KP ? K - p * KP : KP);
      ap.start = a.start - p * KP;
      ap.accumulate = p > 0;
      // three groups per SM when the stage (V + F samples) is small enough, else two
      const bool three = variant != 1 && FirPgCfg<PL, 768>::smem(int(4096 - ap.K + 1)) <= 232448;  // variant 3: the pair kernel, three groups
      rc = three ? run_fir_pg<PL, 768, 1>(ctx, ap, channels, taps + p * KP, st)
                 : run_fir_pg<PL, 512, 1>(ctx, ap, channels, taps + p * KP, st);
      if (rc) return rc;
    }
    return NXS_OK;
  }
  // K < 16 (or the per-group kernels switched off): direct
  if (K > (int64_t(1) << 30)) return NXS_EUNSUPPORTED;
  const int64_t total = channels * out_len;
  int64_t grid = (total + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  prof_begin(ctx, st);
  fir_direct_kernel<<<(unsigned)grid, 256, 0, st>>>(x, channels, length, x_ld, taps, (int)K, a.start, out_len, y_ld, y);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

}  // namespace nxs
