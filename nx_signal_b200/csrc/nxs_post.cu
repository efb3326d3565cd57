// nxs_post.cu -- spectrogram-adjacent operators for sm_100a: the typical steps
// after an STFT, kept on the device so the spectrogram never crosses PCIe between them.
//
//   median   NxSignal.Filters.median(t, kernel_shape:)         lib/nx_signal/filters.ex:17-56
//   wiener   NxSignal.Filters.wiener(t, kernel_size:, noise:)  lib/nx_signal/filters.ex:80-110, 281-303
//   argrel*  NxSignal.PeakFinding.argrelextrema / argrelmin / argrelmax
//                                                              lib/nx_signal/peak_finding.ex:131-391
//
// These are stencil / selection / compaction kernels: one thread per output element, window reads
// served by L1/L2 (neighbouring threads share all but one window column), every result
// deterministic.  Tensors are rank <= 3 here (leading dimensions of size 1 for lower ranks).
#include <math.h>
#include <stdlib.h>

#include "nxs_common.cuh"

namespace nxs {

// ------------------------------------------------------------------------------------------
// median: out[i] = Nx.median of the window that STARTS at i, the start clamped per axis so the
// window stays inside the tensor (Nx.slice semantics, filters.ex:25-27); f32 out.
// Selection by rank counting: candidate a is the r-th smallest iff #(w < a) <= r < #(w <= a).
// Small windows are staged once per thread in shared memory ([element][thread]: conflict-free).
// ------------------------------------------------------------------------------------------
struct MedianArgs {
  const float* t;
  float* out;
  int d0, d1, d2;  // tensor shape
  int k0, k1, k2;  // window shape
  int64_t total;
};

constexpr int kMedianStage = 64;    // windows up to this many elements are staged in shared memory
constexpr int kMedianThreads = 128;

template <bool STAGED>
__global__ void __launch_bounds__(kMedianThreads) median_kernel(const MedianArgs a) {
  extern __shared__ float win_sm[];  // [n][kMedianThreads]
  const int n = a.k0 * a.k1 * a.k2;
  const int r_hi = n / 2, r_lo = (n & 1) ? r_hi : r_hi - 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
    const int i2 = (int)(i % a.d2);
    const int64_t q = i / a.d2;
    const int i1 = (int)(q % a.d1), i0 = (int)(q / a.d1);
    const int s0 = min(i0, a.d0 - a.k0), s1 = min(i1, a.d1 - a.k1), s2 = min(i2, a.d2 - a.k2);
    auto at = [&](int e) {  // e-th window element (row-major over the window)
      const int e2 = e % a.k2, e1 = (e / a.k2) % a.k1, e0 = e / (a.k2 * a.k1);
      return __ldg(a.t + ((int64_t)(s0 + e0) * a.d1 + (s1 + e1)) * a.d2 + (s2 + e2));
    };
    if constexpr (STAGED) {
      for (int e = 0; e < n; ++e) win_sm[e * kMedianThreads + threadIdx.x] = at(e);
    }
    auto get = [&](int e) { return STAGED ? win_sm[e * kMedianThreads + threadIdx.x] : at(e); };
    float lo = 0.f, hi = 0.f;
    bool have_lo = false, have_hi = false;
    for (int c = 0; c < n && !(have_lo && have_hi); ++c) {
      const float v = get(c);
      int less = 0, leq = 0;
      for (int e = 0; e < n; ++e) {
        const float w = get(e);
        less += w < v;
        leq += w <= v;
      }
      if (!have_lo && less <= r_lo && r_lo < leq) {
        lo = v;
        have_lo = true;
      }
      if (!have_hi && less <= r_hi && r_hi < leq) {
        hi = v;
        have_hi = true;
      }
    }
    // odd count: the middle element; even: Nx.median averages the two middle elements (in double, one rounding)
    a.out[i] = (n & 1) ? hi : (float)(((double)lo + (double)hi) / 2.0);
  }
}

// Windows of up to 64 elements (every usual spectrogram smoothing window: 1 x 17, 31 x 1, 5 x 5 ...):
// the window is gathered into KB registers (KB = 4 .. 64, the tail padded with +inf, which sorts last
// and leaves the lower ranks unchanged) and sorted by a bitonic network of min/max pairs -- no
// branches, no shared memory; neighbouring threads re-read the same lines from L1.  The window's
// element offsets are precomputed on the host (kernel parameters = constant bank, indexed at compile
// time after unrolling).
template <int KB>
struct MedianNetArgs {
  const float* t;
  float* out;
  int d0, d1, d2, k0, k1, k2, n, total;
  FastDiv div_d2, div_d1;
  int off[KB];  // offset of window element e from the window's first element
};

template <int KB>
__device__ __forceinline__ void bitonic_sort(float (&v)[KB]) {
#pragma unroll
  for (int k = 2; k <= KB; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float a = v[i], b = v[l];
          const bool asc = (i & k) == 0;
          v[i] = asc ? fminf(a, b) : fmaxf(a, b);
          v[l] = asc ? fmaxf(a, b) : fminf(a, b);
        }
      }
}

template <int KB>
__global__ void __launch_bounds__(256) median_net_kernel(const MedianNetArgs<KB> a) {
  const int r_hi = a.n / 2, r_lo = (a.n & 1) ? r_hi : r_hi - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += gridDim.x * blockDim.x) {
    const int q = a.div_d2.div(i), i2 = i - q * a.d2;
    const int i0 = a.div_d1.div(q), i1 = q - i0 * a.d1;
    const int s0 = min(i0, a.d0 - a.k0), s1 = min(i1, a.d1 - a.k1), s2 = min(i2, a.d2 - a.k2);
    const float* __restrict__ w = a.t + ((int64_t)s0 * a.d1 + s1) * a.d2 + s2;
    float v[KB];
#pragma unroll
    for (int e = 0; e < KB; ++e) v[e] = e < a.n ? __ldg(w + a.off[e]) : INFINITY;
    bitonic_sort<KB>(v);
    float lo = 0.f, hi = 0.f;
#pragma unroll
    for (int e = 0; e < KB; ++e) {
      if (e == r_lo) lo = v[e];
      if (e == r_hi) hi = v[e];
    }
    // odd count: the middle element; even: Nx.median averages the two middle elements (in double, one rounding)
    a.out[i] = (a.n & 1) ? hi : (float)(((double)lo + (double)hi) / 2.0);
  }
}

template <int KB>
static int run_median_net(nxs_ctx* ctx, const float* t, const int64_t shape[3], const int64_t kernel[3], float* out,
                          cudaStream_t st) {
  MedianNetArgs<KB> a;
  a.t = t;
  a.out = out;
  a.d0 = (int)shape[0];
  a.d1 = (int)shape[1];
  a.d2 = (int)shape[2];
  a.k0 = (int)kernel[0];
  a.k1 = (int)kernel[1];
  a.k2 = (int)kernel[2];
  a.n = a.k0 * a.k1 * a.k2;
  a.total = (int)(shape[0] * shape[1] * shape[2]);
  a.div_d2 = FastDiv(a.d2);
  a.div_d1 = FastDiv(a.d1);
  for (int e = 0; e < KB; ++e) {
    const int ee = e < a.n ? e : 0;
    const int e2 = ee % a.k2, e1 = (ee / a.k2) % a.k1, e0 = ee / (a.k2 * a.k1);
    a.off[e] = (e0 * a.d1 + e1) * a.d2 + e2;
  }
  int64_t grid = (int64_t(a.total) + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  prof_begin(ctx, st);
  median_net_kernel<KB><<<(unsigned)grid, 256, 0, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// One-axis windows (1 x k, k x 1 ...: harmonic / percussive smoothing): two neighbouring outputs
// along the axis share k - 1 of their k elements.  A thread sorts that shared core once (bitonic
// network over KB >= k - 1 registers) and finishes each of its two outputs with a clamp: the rank-r
// element of (sorted core c) + {x} is min(max(x, c[r-1]), c[r]) (c[-1] = -inf, c[k-1] = +inf).
// The tensor is viewed as [outer][n][inner] around the axis; `inner` runs fastest over the threads so
// the loads stay coalesced whichever axis carries the window.  Outputs whose window start is clamped
// (the last k - 1 positions of the axis, Nx.slice semantics) take the plain sort of their window.
struct MedianAxisArgs {
  const float* t;
  float* out;
  int n, inner, k, pairs;  // axis length, inner size, window length, ceil(n / 2)
  int total_threads;       // outer * pairs * inner
  FastDiv div_inner, div_pairs;
};

template <int KB>
__device__ __forceinline__ float rank_with_extra(const float (&c)[KB], int r, float x) {
  // rank r of the union of the sorted core and x
  float below = -INFINITY, at = INFINITY;
#pragma unroll
  for (int e = 0; e < KB; ++e) {
    if (e == r - 1) below = c[e];
    if (e == r) at = c[e];
  }
  return fminf(fmaxf(x, below), at);
}

template <int KB>
__global__ void __launch_bounds__(256) median_axis_kernel(const MedianAxisArgs a) {
  const int k = a.k, core = k - 1;
  const int r_hi = k / 2, r_lo = (k & 1) ? r_hi : r_hi - 1;
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < a.total_threads; id += gridDim.x * blockDim.x) {
    const int q = a.div_inner.div(id), in = id - q * a.inner;
    const int o = a.div_pairs.div(q), p = q - o * a.pairs;
    const int j0 = 2 * p;  // outputs j0 and j0 + 1 along the axis
    const int64_t base = ((int64_t)o * a.n) * a.inner + in;
    const float* __restrict__ col = a.t + base;
    float* __restrict__ ocol = a.out + base;
    const int last = a.n - k;  // last unclamped window start
    if (j0 + 1 <= last) {
      float c[KB];
#pragma unroll
      for (int e = 0; e < KB; ++e) c[e] = e < core ? __ldg(col + (int64_t)(j0 + 1 + e) * a.inner) : INFINITY;
      bitonic_sort<KB>(c);
      const float x0 = __ldg(col + (int64_t)j0 * a.inner), x1 = __ldg(col + (int64_t)(j0 + k) * a.inner);
      float m0 = rank_with_extra<KB>(c, r_hi, x0), m1 = rank_with_extra<KB>(c, r_hi, x1);
      if (!(k & 1)) {  // even window: Nx.median averages the two middle elements (in double, one rounding)
        m0 = (float)(((double)rank_with_extra<KB>(c, r_lo, x0) + (double)m0) / 2.0);
        m1 = (float)(((double)rank_with_extra<KB>(c, r_lo, x1) + (double)m1) / 2.0);
      }
      ocol[(int64_t)j0 * a.inner] = m0;
      ocol[(int64_t)(j0 + 1) * a.inner] = m1;
    } else {
      // clamped starts: every output from `last` on is the median of the window [last, last + k)
      for (int j = j0; j < j0 + 2 && j < a.n; ++j) {
        const int s = j < last ? j : last;
        float c[KB];
#pragma unroll
        for (int e = 0; e < KB; ++e) c[e] = e < core ? __ldg(col + (int64_t)(s + 1 + e) * a.inner) : INFINITY;
        bitonic_sort<KB>(c);
        const float x = __ldg(col + (int64_t)s * a.inner);
        float m = rank_with_extra<KB>(c, r_hi, x);
        if (!(k & 1)) m = (float)(((double)rank_with_extra<KB>(c, r_lo, x) + (double)m) / 2.0);
        ocol[(int64_t)j * a.inner] = m;
      }
    }
  }
}

template <int KB>
static int run_median_axis(nxs_ctx* ctx, const MedianAxisArgs& a, cudaStream_t st) {
  int64_t grid = (int64_t(a.total_threads) + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  prof_begin(ctx, st);
  median_axis_kernel<KB><<<(unsigned)grid, 256, 0, st>>>(a);
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

int launch_median(nxs_ctx* ctx, const float* t, const int64_t shape[3], const int64_t kernel[3], float* out,
                  cudaStream_t st) {
  {
    // exactly one axis carries the window (2 <= k <= 65): shared-core kernel
    const int64_t total = shape[0] * shape[1] * shape[2];
    int axis = -1, nonunit = 0;
    for (int i = 0; i < 3; ++i)
      if (kernel[i] > 1) {
        axis = i;
        ++nonunit;
      }
    if (nonunit == 1 && kernel[axis] <= 65 && total > 0 && total < (int64_t(1) << 31) - (int64_t(1) << 24) &&
        !getenv("NXS_MEDIAN_NO_AXIS") && !getenv("NXS_MEDIAN_NO_NET")) {
      MedianAxisArgs a;
      a.t = t;
      a.out = out;
      a.n = (int)shape[axis];
      int64_t inner = 1, outer = 1;
      for (int i = axis + 1; i < 3; ++i) inner *= shape[i];
      for (int i = 0; i < axis; ++i) outer *= shape[i];
      a.inner = (int)inner;
      a.k = (int)kernel[axis];
      a.pairs = (a.n + 1) / 2;
      a.total_threads = (int)(outer * a.pairs * inner);
      a.div_inner = FastDiv(a.inner);
      a.div_pairs = FastDiv(a.pairs);
      const int core = a.k - 1;
      if (core <= 2) return run_median_axis<2>(ctx, a, st);
      if (core <= 4) return run_median_axis<4>(ctx, a, st);
      if (core <= 8) return run_median_axis<8>(ctx, a, st);
      if (core <= 16) return run_median_axis<16>(ctx, a, st);
      if (core <= 32) return run_median_axis<32>(ctx, a, st);
      return run_median_axis<64>(ctx, a, st);
    }
  }
  {
    const int64_t total = shape[0] * shape[1] * shape[2], n = kernel[0] * kernel[1] * kernel[2];
    if (total > 0 && total < (int64_t(1) << 31) - (int64_t(1) << 24) && n <= 64 && !getenv("NXS_MEDIAN_NO_NET")) {
      if (n <= 4) return run_median_net<4>(ctx, t, shape, kernel, out, st);
      if (n <= 8) return run_median_net<8>(ctx, t, shape, kernel, out, st);
      if (n <= 16) return run_median_net<16>(ctx, t, shape, kernel, out, st);
      if (n <= 32) return run_median_net<32>(ctx, t, shape, kernel, out, st);
      return run_median_net<64>(ctx, t, shape, kernel, out, st);
    }
  }
  MedianArgs a;
  a.t = t;
  a.out = out;
  a.d0 = (int)shape[0];
  a.d1 = (int)shape[1];
  a.d2 = (int)shape[2];
  a.k0 = (int)kernel[0];
  a.k1 = (int)kernel[1];
  a.k2 = (int)kernel[2];
  a.total = shape[0] * shape[1] * shape[2];
  if (a.total <= 0) return NXS_OK;
  const int64_t n = kernel[0] * kernel[1] * kernel[2];
  if (n > 8192) return NXS_EUNSUPPORTED;
  int64_t grid = (a.total + kMedianThreads - 1) / kMedianThreads;
  if (grid > int64_t(ctx->sm_count) * 16) grid = int64_t(ctx->sm_count) * 16;
  prof_begin(ctx, st);
  if (n <= kMedianStage) {
    const size_t smem = size_t(n) * kMedianThreads * sizeof(float);
    median_kernel<true><<<(unsigned)grid, kMedianThreads, smem, st>>>(a);
  } else {
    median_kernel<false><<<(unsigned)grid, kMedianThreads, 0, st>>>(a);
  }
  prof_end(ctx, st);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// ------------------------------------------------------------------------------------------
// wiener (filters.ex:281-303), computed in double like the reference (Nx.as_type(:f64), :103):
//   l_mean = correlate(t, ones, :same) / size,  l_var = correlate(t^2, ones, :same) / size - l_mean^2
//   noise  = given, or mean(l_var)
//   out    = l_var < noise ? l_mean : (t - l_mean) (1 - noise / l_var) + l_mean
// The :same window of output i covers [i - (k-1) + (k-1)/2, i + (k-1)/2] per axis, zeros outside
// (convolution.ex:95-211).  Pass 1 writes (l_mean, l_var) and per-block sums of l_var; a one-thread
// pass adds the block sums in a fixed order (deterministic); pass 2 applies the formula.  (Recomputing
// the statistics in pass 2 instead of storing them was measured slower: the f64 window sums, not the
// 32 bytes per element of scratch traffic, are the cost -- 5.5 vs 3.6 ms on a 360k x 513 spectrogram.)
// ------------------------------------------------------------------------------------------
struct WienerArgs {
  const void* t;
  int is_f64;
  int d0, d1, d2, k0, k1, k2;
  int64_t total;
  double2* mv;        // [total] (l_mean, l_var)
  double* block_sum;  // [grid]
};

__device__ __forceinline__ double wiener_load(const WienerArgs& a, int64_t i) {
  return a.is_f64 ? reinterpret_cast<const double*>(a.t)[i] : (double)reinterpret_cast<const float*>(a.t)[i];
}

__global__ void __launch_bounds__(256) wiener_stats_kernel(const WienerArgs a) {
  __shared__ double red[256];
  const double size = (double)a.k0 * (double)a.k1 * (double)a.k2;
  double acc = 0.0;
  // a block owns a contiguous run of elements so that the block sums add up in element order
  const int64_t per_block = (a.total + gridDim.x - 1) / gridDim.x;
  const int64_t begin = blockIdx.x * per_block, end = begin + per_block < a.total ? begin + per_block : a.total;
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int i2 = (int)(i % a.d2);
    const int64_t q = i / a.d2;
    const int i1 = (int)(q % a.d1), i0 = (int)(q / a.d1);
    const int h0 = i0 + (a.k0 - 1) / 2, h1 = i1 + (a.k1 - 1) / 2, h2 = i2 + (a.k2 - 1) / 2;
    double s = 0.0, s2 = 0.0;
    for (int e0 = 0; e0 < a.k0; ++e0) {  // kernel index ascending = source index descending (the oracle's order)
      const int p0 = h0 - e0;
      if (p0 < 0 || p0 >= a.d0) continue;
      for (int e1 = 0; e1 < a.k1; ++e1) {
        const int p1 = h1 - e1;
        if (p1 < 0 || p1 >= a.d1) continue;
        for (int e2 = 0; e2 < a.k2; ++e2) {
          const int p2 = h2 - e2;
          if (p2 < 0 || p2 >= a.d2) continue;
          const double v = wiener_load(a, ((int64_t)p0 * a.d1 + p1) * a.d2 + p2);
          s += v;
          s2 += v * v;
        }
      }
    }
    const double mean = s / size;
    const double var = s2 / size - mean * mean;
    a.mv[i] = make_double2(mean, var);
    acc += var;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.block_sum[blockIdx.x] = red[0];
}

__global__ void wiener_noise_kernel(const double* __restrict__ block_sum, int nblocks, int64_t total,
                                    double* __restrict__ noise) {
  double s = 0.0;
  for (int i = 0; i < nblocks; ++i) s += block_sum[i];
  *noise = s / (double)total;
}

__global__ void __launch_bounds__(256) wiener_apply_kernel(const WienerArgs a, const double* __restrict__ noise_dev,
                                                           double noise_given, int use_given, void* __restrict__ out) {
  const double noise = use_given ? noise_given : *noise_dev;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 mv = a.mv[i];
    const double t = wiener_load(a, i);
    const double res = (t - mv.x) * (1.0 - noise / mv.y);
    const double r = mv.y < noise ? mv.x : res + mv.x;
    if (a.is_f64) reinterpret_cast<double*>(out)[i] = r;
    else reinterpret_cast<float*>(out)[i] = (float)r;
  }
}

int launch_wiener(nxs_ctx* ctx, const void* t, int is_f64, const int64_t shape[3], const int64_t kernel[3],
                  int has_noise, double noise, void* out, cudaStream_t st) {
  WienerArgs a;
  a.t = t;
  a.is_f64 = is_f64;
  a.d0 = (int)shape[0];
  a.d1 = (int)shape[1];
  a.d2 = (int)shape[2];
  a.k0 = (int)kernel[0];
  a.k1 = (int)kernel[1];
  a.k2 = (int)kernel[2];
  a.total = shape[0] * shape[1] * shape[2];
  if (a.total <= 0) return NXS_OK;
  int64_t grid = (a.total + 255) / 256;
  if (grid > int64_t(ctx->sm_count) * 8) grid = int64_t(ctx->sm_count) * 8;
  const size_t mv_bytes = (size_t(a.total) * sizeof(double2) + 255) / 256 * 256;
  int rc = ensure_scratch(ctx, mv_bytes + (size_t(grid) + 1) * sizeof(double));
  if (rc) return rc;
  a.mv = reinterpret_cast<double2*>(ctx->d_scratch);
  a.block_sum = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->d_scratch) + mv_bytes);
  double* noise_dev = a.block_sum + grid;
  prof_begin(ctx, st);
  wiener_stats_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
  if (!has_noise) wiener_noise_kernel<<<1, 1, 0, st>>>(a.block_sum, (int)grid, a.total, noise_dev);
  wiener_apply_kernel<<<(unsigned)grid, 256, 0, st>>>(a, noise_dev, noise, has_noise, out);
  prof_end(ctx, st);
  ctx->launches += has_noise ? 2 : 3;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// ------------------------------------------------------------------------------------------
// argrelextrema (peak_finding.ex:339-391): element e is an extremum iff cmp(e, e -+ s) along the
// axis for every shift s = 1 .. order, neighbour indices clipped to the axis (so an edge element is
// compared with itself).  The tensor is viewed as [outer][n][inner] around the axis.  The result is
// the reference's `nonzero`: the multi-indices of the true elements in row-major order, then rows
// of -1, plus the count -- an order-preserving compaction (block counts -> one-block scan ->
// scatter), deterministic.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool relcmp(int cmp, float x, float y) {
  switch (cmp) {
    case NXS_CMP_LESS: return x < y;
    case NXS_CMP_GREATER: return x > y;
    case NXS_CMP_LESS_EQUAL: return x <= y;
    default: return x >= y;
  }
}

constexpr int kScanChunk = 4096;  // elements per compaction block (16 per thread)

// mask bytes + per-chunk counts.  Element k * 256 + tid of a chunk belongs to thread tid in iteration k
// (coalesced loads and byte stores); the mask buffer is padded to whole chunks and the pad is zeroed.
__global__ void __launch_bounds__(256) relextrema_mask_kernel(const float* __restrict__ d, int n, int inner,
                                                              const FastDiv div_inner, const FastDiv div_n,
                                                              int total, int order, int cmp,
                                                              unsigned char* __restrict__ mask,
                                                              int* __restrict__ block_count) {
  __shared__ int warp_cnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (element counts are below 2^31 -- checked by the launcher -- so index arithmetic is 32-bit)
  for (int blk = blockIdx.x; (int64_t)blk * kScanChunk < total; blk += gridDim.x) {
    int mine = 0;
    const int64_t base = (int64_t)blk * kScanChunk;
#pragma unroll 4
    for (int k = 0; k < kScanChunk / 256; ++k) {
      const int64_t i64 = base + k * 256 + threadIdx.x;
      bool ok = false;
      if (i64 < total) {
        const int i = (int)i64;
        const int row = div_inner.div(i);          // i / inner
        const int pos = row - div_n.div(row) * n;  // (i / inner) % n
        const float x = d[i];
        ok = true;
        for (int s = 1; s <= order && ok; ++s) {
          const int up = pos + s < n ? s : n - 1 - pos, dn = pos - s >= 0 ? s : pos;
          ok = relcmp(cmp, x, d[i + (int64_t)up * inner]) && relcmp(cmp, x, d[i - (int64_t)dn * inner]);
        }
      }
      mask[i64] = ok ? 1 : 0;
      mine += ok;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) warp_cnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      int c = 0;
      for (int w = 0; w < 8; ++w) c += warp_cnt[w];
      block_count[blk] = c;
    }
    __syncthreads();
  }
}

// exclusive scan of the chunk counts by one block (1024 at a time with a running carry)
__global__ void __launch_bounds__(1024) block_scan_kernel(int* __restrict__ counts, int64_t nblocks,
                                                          int64_t* __restrict__ total_out) {
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nblocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      const int64_t add = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += add;
      __syncthreads();
    }
    const int64_t excl = carry + buf[threadIdx.x] - v;
    if (i < nblocks) reinterpret_cast<int*>(counts)[i] = (int)excl;  // offsets fit: total < 2^31 (checked by the launcher)
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

struct ShapeN {
  int rank;
  int dim[8];
  FastDiv fd[8];
};

// A thread owns 16 consecutive mask bytes of the chunk (one 128-bit load); one block-wide exclusive
// scan of the 256 per-thread counts places every thread's rows, which it then writes in element order.
__global__ void __launch_bounds__(256) nonzero_scatter_kernel(const unsigned char* __restrict__ mask, int total,
                                                              const int* __restrict__ block_offset, const ShapeN shp,
                                                              int* __restrict__ indices) {
  __shared__ int warp_tot[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int blk = blockIdx.x; (int64_t)blk * kScanChunk < total; blk += gridDim.x) {
    const int64_t first = (int64_t)blk * kScanChunk + threadIdx.x * 16;
    const uint4 m = *reinterpret_cast<const uint4*>(mask + first);  // padded buffer: always in bounds
    const unsigned words[4] = {m.x, m.y, m.z, m.w};
    const int mine = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);  // bytes are 0 / 1
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) before += warp_tot[w];
    int64_t row = (int64_t)block_offset[blk] + before + incl - mine;
    if (mine) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((words[e >> 2] >> (8 * (e & 3))) & 1u) {
          int rem = (int)(first + e);
#pragma unroll
          for (int ax = 7; ax >= 0; --ax) {
            if (ax < shp.rank) {
              const int dd = shp.dim[ax], qq = shp.fd[ax].div(rem);
              indices[row * shp.rank + ax] = rem - qq * dd;
              rem = qq;
            }
          }
          ++row;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) nonzero_fill_kernel(int* __restrict__ indices, int64_t total, int rank,
                                                           const int64_t* __restrict__ valid) {
  const int64_t first = *valid * rank, end = total * rank;
  // scalar head up to the next 16-byte boundary, 128-bit stores over the body, scalar tail
  const int64_t body0 = (first + 3) & ~(int64_t)3, body1 = end & ~(int64_t)3;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  if (body0 >= body1) {
    for (int64_t i = first + tid; i < end; i += nthr) indices[i] = -1;
    return;
  }
  for (int64_t i = first + tid; i < body0; i += nthr) indices[i] = -1;
  int4* __restrict__ v = reinterpret_cast<int4*>(indices);
  for (int64_t i = body0 / 4 + tid; i < body1 / 4; i += nthr) v[i] = make_int4(-1, -1, -1, -1);
  for (int64_t i = body1 + tid; i < end; i += nthr) indices[i] = -1;
}

int launch_argrelextrema(nxs_ctx* ctx, const float* data, int rank, const int64_t* shape, int axis, int order, int cmp,
                         int* indices, int64_t* valid_dev, cudaStream_t st) {
  ShapeN shp;
  shp.rank = rank;
  int64_t total = 1, inner = 1;
  for (int i = 0; i < rank; ++i) {
    shp.dim[i] = (int)shape[i];
    shp.fd[i] = FastDiv((int)shape[i]);
    total *= shape[i];
    if (i > axis) inner *= shape[i];
  }
  if (total <= 0) {
    NXS_CUDA(ctx, cudaMemsetAsync(valid_dev, 0, sizeof(int64_t), st));
    return NXS_OK;
  }
  if (total >= (int64_t(1) << 31)) return NXS_EUNSUPPORTED;
  const int64_t n = shape[axis];
  const int64_t nblocks = (total + kScanChunk - 1) / kScanChunk;
  const size_t mask_bytes = size_t(nblocks) * kScanChunk;  // whole chunks: the kernels read / write the pad
  int rc = ensure_scratch(ctx, mask_bytes + size_t(nblocks) * sizeof(int));
  if (rc) return rc;
  unsigned char* mask = reinterpret_cast<unsigned char*>(ctx->d_scratch);
  int* counts = reinterpret_cast<int*>(mask + mask_bytes);
  int64_t grid = nblocks < int64_t(ctx->sm_count) * 8 ? nblocks : int64_t(ctx->sm_count) * 8;
  prof_begin(ctx, st);
  relextrema_mask_kernel<<<(unsigned)grid, 256, 0, st>>>(data, (int)n, (int)inner, FastDiv((int)inner), FastDiv((int)n),
                                                         (int)total, order, cmp, mask, counts);
  block_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, valid_dev);
  nonzero_scatter_kernel<<<(unsigned)grid, 256, 0, st>>>(mask, (int)total, counts, shp, indices);
  int64_t g2 = (total * rank + 255) / 256;
  if (g2 > int64_t(ctx->sm_count) * 8) g2 = int64_t(ctx->sm_count) * 8;
  nonzero_fill_kernel<<<(unsigned)g2, 256, 0, st>>>(indices, total, rank, valid_dev);
  prof_end(ctx, st);
  ctx->launches += 4;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

}  // namespace nxs
