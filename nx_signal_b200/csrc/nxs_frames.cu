// nxs_frames.cu -- the data-movement ops either side of the FFT kernels, as standalone
// entry points: as_windowed (gather of overlapping frames), overlap_and_add (deterministic
// gather formulation of the reference's Nx.indexed_add scatter) and the small general
// N-d direct convolution.
//   NxSignal.as_windowed/2      lib/nx_signal.ex:249-364
//   NxSignal.overlap_and_add/2  lib/nx_signal.ex:684-735
//   Convolution.convolve/3      lib/nx_signal/convolution.ex:38-58, 95-223, 300-329
#include "nxs_common.cuh"

namespace nxs {

template <typename E>
__global__ void __launch_bounds__(256) as_windowed_kernel(const E* __restrict__ x, int64_t channels, int64_t L,
                                                          int64_t x_ld, int64_t N, int64_t stride, int64_t lo,
                                                          int reflect, int64_t M, E* __restrict__ out) {
  const int64_t total = channels * M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i % N;
    const int64_t fm = i / N;
    const int64_t m = fm % M, c = fm / M;
    const int64_t src = m * stride + n - lo;
    E v = E();
    if (src >= 0 && src < L) v = x[c * x_ld + src];
    else if (reflect) v = x[c * x_ld + reflect_index(src, L)];
    out[i] = v;
  }
}

// the same gather with 32-bit indices and multiply-high divisions (outputs of fewer than 2^31
// elements -- every BASELINE shape): the emulated 64-bit divisions above cost ~200 instructions per
// copied element
template <typename E>
__global__ void __launch_bounds__(256) as_windowed_kernel32(const E* __restrict__ x, int L, int64_t x_ld, int N,
                                                            int stride, int lo, int reflect, int M,
                                                            const FastDiv div_n, const FastDiv div_m, int total,
                                                            E* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int fm = div_n.div(i), n = i - fm * N;
    const int c = div_m.div(fm), m = fm - c * M;
    const int64_t src = (int64_t)m * stride + n - lo;
    E v = E();
    if (src >= 0 && src < L) v = x[c * x_ld + src];
    else if (reflect) v = x[c * x_ld + reflect_index(src, L)];
    out[i] = v;
  }
}

// f32 frames whose every group of 4 samples is 16-byte aligned in both tensors (N, stride, lo, x_ld
// multiples of 4): one 128-bit load and store per thread; groups that touch the padding fall back to
// the per-sample rule
__global__ void __launch_bounds__(256) as_windowed_kernel_v4(const float* __restrict__ x, int L, int64_t x_ld, int N4,
                                                             int stride, int lo, int reflect, int M,
                                                             const FastDiv div_n4, const FastDiv div_m, int total4,
                                                             float4* __restrict__ out) {
  // four independent 128-bit loads in flight per thread, then four streaming stores (the frames are written once
  // and never re-read here; the 4x re-read of the input stays in L2)
  auto fetch = [&](int i) -> float4 {
    const int fm = div_n4.div(i), n = (i - fm * N4) * 4;
    const int c = div_m.div(fm), m = fm - c * M;
    const int64_t src = (int64_t)m * stride + n - lo;
    const float* __restrict__ row = x + c * x_ld;
    if (src >= 0 && src + 3 < L) return *reinterpret_cast<const float4*>(row + src);
    float e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t sk = src + k;
      e[k] = (sk >= 0 && sk < L) ? row[sk] : (reflect ? row[reflect_index(sk, L)] : 0.f);
    }
    return make_float4(e[0], e[1], e[2], e[3]);
  };
  const int step = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total4 - 3 * (int64_t)step; i += 4 * step) {
    const float4 v0 = fetch(i), v1 = fetch(i + step), v2 = fetch(i + 2 * step), v3 = fetch(i + 3 * step);
    __stcs(out + i, v0);
    __stcs(out + i + step, v1);
    __stcs(out + i + 2 * step, v2);
    __stcs(out + i + 3 * step, v3);
  }
  for (; i < total4; i += step) __stcs(out + i, fetch(i));
}

int launch_as_windowed(nxs_ctx* ctx, const void* x, int elem_size, int64_t channels, int64_t length,
                       int64_t x_ld, int64_t window_length, int64_t stride, const PadGeom& g,
                       int64_t num_frames, void* out, cudaStream_t st) {
  const int64_t total = channels * num_frames * window_length;
  if (total <= 0) return NXS_OK;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (grid > cap) grid = cap;
  const bool small = total < (int64_t(1) << 31) - (int64_t(1) << 24) && length < (int64_t(1) << 31) &&
                     stride < (int64_t(1) << 31) && g.lo < (int64_t(1) << 31);
  const bool vec4 = small && elem_size == 4 && window_length % 4 == 0 && stride % 4 == 0 && g.lo % 4 == 0 &&
                    x_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec4) {
    const int64_t total4 = total / 4;
    int64_t g4 = (total4 + 255) / 256;
    if (g4 > cap) g4 = cap;
    as_windowed_kernel_v4<<<(unsigned)g4, 256, 0, st>>>((const float*)x, (int)length, x_ld, (int)(window_length / 4),
                                                        (int)stride, (int)g.lo, g.reflect, (int)num_frames,
                                                        FastDiv((int)(window_length / 4)), FastDiv((int)num_frames),
                                                        (int)total4, (float4*)out);
  } else if (small && elem_size == 4)
    as_windowed_kernel32<float><<<(unsigned)grid, 256, 0, st>>>(
        (const float*)x, (int)length, x_ld, (int)window_length, (int)stride, (int)g.lo, g.reflect, (int)num_frames,
        FastDiv((int)window_length), FastDiv((int)num_frames), (int)total, (float*)out);
  else if (small)
    as_windowed_kernel32<double><<<(unsigned)grid, 256, 0, st>>>(
        (const double*)x, (int)length, x_ld, (int)window_length, (int)stride, (int)g.lo, g.reflect, (int)num_frames,
        FastDiv((int)window_length), FastDiv((int)num_frames), (int)total, (double*)out);
  else if (elem_size == 4)
    as_windowed_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)x, channels, length, x_ld,
                                                              window_length, stride, g.lo, g.reflect,
                                                              num_frames, (float*)out);
  else
    as_windowed_kernel<double><<<(unsigned)grid, 256, 0, st>>>((const double*)x, channels, length, x_ld,
                                                               window_length, stride, g.lo, g.reflect,
                                                               num_frames, (double*)out);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// out[b][n] = sum over frames m covering n of t[b][m][n - m*hop]   (ascending m, fp32)
template <int CPLX>
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ t, int64_t batch, int64_t M,
                                                          int64_t N, int64_t hop, int64_t out_len,
                                                          float* __restrict__ out) {
  const int64_t total = batch * out_len;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i % out_len, b = i / out_len;
    int64_t m_hi = n / hop;
    if (m_hi > M - 1) m_hi = M - 1;
    int64_t m_lo = n - N + 1 <= 0 ? 0 : (n - N + 1 + hop - 1) / hop;
    float re = 0.f, im = 0.f;
    for (int64_t m = m_lo; m <= m_hi; ++m) {
      const int64_t idx = (b * M + m) * N + (n - m * hop);
      if (CPLX) {
        const float2 v = reinterpret_cast<const float2*>(t)[idx];
        re += v.x;
        im += v.y;
      } else {
        re += t[idx];
      }
    }
    if (CPLX) reinterpret_cast<float2*>(out)[i] = make_float2(re, im);
    else out[i] = re;
  }
}

// 32-bit / multiply-high form for inputs and outputs of fewer than 2^31 elements
template <int CPLX>
__global__ void __launch_bounds__(256) overlap_add_kernel32(const float* __restrict__ t, int M, int N, int hop,
                                                            int out_len, const FastDiv div_out, const FastDiv div_hop,
                                                            int total, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = div_out.div(i), n = i - b * out_len;
    int m_hi = div_hop.div(n);
    if (m_hi > M - 1) m_hi = M - 1;
    const int m_lo = n - N + 1 <= 0 ? 0 : div_hop.div(n - N + hop);
    float re = 0.f, im = 0.f;
    int64_t idx = ((int64_t)b * M + m_lo) * N + (n - m_lo * hop);
    for (int m = m_lo; m <= m_hi; ++m, idx += N - hop) {  // ascending m, as the reference's indexed_add visits them
      if (CPLX) {
        const float2 v = reinterpret_cast<const float2*>(t)[idx];
        re += v.x;
        im += v.y;
      } else {
        re += t[idx];
      }
    }
    if (CPLX) reinterpret_cast<float2*>(out)[i] = make_float2(re, im);
    else out[i] = re;
  }
}

int launch_overlap_and_add(nxs_ctx* ctx, const float* t, int complex_, int64_t batch, int64_t num_frames,
                           int64_t frame_length, int64_t overlap, float* out, cudaStream_t st) {
  const int64_t hop = frame_length - overlap;
  const int64_t out_len = num_frames * hop + overlap;
  const int64_t total = batch * out_len;
  if (total <= 0) return NXS_OK;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (grid > cap) grid = cap;
  const int64_t lim = (int64_t(1) << 31) - (int64_t(1) << 24);
  if (total < lim && out_len + frame_length < lim && complex_)
    overlap_add_kernel32<1><<<(unsigned)grid, 256, 0, st>>>(t, (int)num_frames, (int)frame_length, (int)hop, (int)out_len,
                                                            FastDiv((int)out_len), FastDiv((int)hop), (int)total, out);
  else if (total < lim && out_len + frame_length < lim)
    overlap_add_kernel32<0><<<(unsigned)grid, 256, 0, st>>>(t, (int)num_frames, (int)frame_length, (int)hop, (int)out_len,
                                                            FastDiv((int)out_len), FastDiv((int)hop), (int)total, out);
  else if (complex_)
    overlap_add_kernel<1><<<(unsigned)grid, 256, 0, st>>>(t, batch, num_frames, frame_length, hop, out_len, out);
  else
    overlap_add_kernel<0><<<(unsigned)grid, 256, 0, st>>>(t, batch, num_frames, frame_length, hop, out_len, out);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// general direct convolution, rank <= 3, real or complex; double accumulation like the
// reference's Nx.conv on BinaryBackend (one rounding at the end)
struct ConvArgs {
  int64_t as[3], bs[3], os[3], start[3];
};

template <int CPLX>
__global__ void __launch_bounds__(128) convolve_nd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          ConvArgs g, float* __restrict__ out) {
  const int64_t total = g.os[0] * g.os[1] * g.os[2];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o2 = i % g.os[2], o1 = (i / g.os[2]) % g.os[1], o0 = i / (g.os[2] * g.os[1]);
    const int64_t f0 = o0 + g.start[0], f1 = o1 + g.start[1], f2 = o2 + g.start[2];
    double re = 0.0, im = 0.0;
    for (int64_t j0 = 0; j0 < g.bs[0]; ++j0) {
      const int64_t i0 = f0 - j0;
      if (i0 < 0 || i0 >= g.as[0]) continue;
      for (int64_t j1 = 0; j1 < g.bs[1]; ++j1) {
        const int64_t i1 = f1 - j1;
        if (i1 < 0 || i1 >= g.as[1]) continue;
        for (int64_t j2 = 0; j2 < g.bs[2]; ++j2) {
          const int64_t i2 = f2 - j2;
          if (i2 < 0 || i2 >= g.as[2]) continue;
          const int64_t ia = (i0 * g.as[1] + i1) * g.as[2] + i2;
          const int64_t ib = (j0 * g.bs[1] + j1) * g.bs[2] + j2;
          if (CPLX) {
            const float2 av = reinterpret_cast<const float2*>(a)[ia];
            const float2 bv = reinterpret_cast<const float2*>(b)[ib];
            re += (double)av.x * bv.x - (double)av.y * bv.y;
            im += (double)av.x * bv.y + (double)av.y * bv.x;
          } else {
            re += (double)a[ia] * (double)b[ib];
          }
        }
      }
    }
    if (CPLX) reinterpret_cast<float2*>(out)[i] = make_float2((float)re, (float)im);
    else out[i] = (float)re;
  }
}

int launch_convolve_nd(nxs_ctx* ctx, const float* a, const int64_t* as, const float* b, const int64_t* bs,
                       int is_complex, int mode, float* out, cudaStream_t st) {
  ConvArgs g;
  bool ok1 = true;
  for (int d = 0; d < 3; ++d) ok1 = ok1 && as[d] >= bs[d];
  for (int d = 0; d < 3; ++d) {
    g.as[d] = as[d];
    g.bs[d] = bs[d];
    if (mode == NXS_MODE_FULL) {
      g.os[d] = as[d] + bs[d] - 1;
      g.start[d] = 0;
    } else if (mode == NXS_MODE_SAME) {
      g.os[d] = as[d];
      g.start[d] = (bs[d] - 1) / 2;
    } else {
      const int64_t big = ok1 ? as[d] : bs[d], small = ok1 ? bs[d] : as[d];
      g.os[d] = big - small + 1;
      g.start[d] = small - 1;
    }
  }
  const int64_t total = g.os[0] * g.os[1] * g.os[2];
  if (total <= 0) return NXS_OK;
  int64_t grid = (total + 127) / 128;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (grid > cap) grid = cap;
  if (is_complex) convolve_nd_kernel<1><<<(unsigned)grid, 128, 0, st>>>(a, b, g, out);
  else convolve_nd_kernel<0><<<(unsigned)grid, 128, 0, st>>>(a, b, g, out);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

}  // namespace nxs
