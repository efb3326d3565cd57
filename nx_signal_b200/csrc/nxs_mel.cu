// nxs_mel.cu -- log-mel epilogue of the STFT for sm_100a.
//
// Replaces NxSignal.stft_to_mel/3 (lib/nx_signal.ex:486-513): |z|^2 over the first
// fft_length/2 bins (:490, :496-499) -> dot with the Slaney mel filterbank of mel_filters/4
// (:397-445, built bit-exactly on the host by nxs_mel_filters_f32) (:501-507) ->
// log10(clip(., 1e-10)) (:511) -> max(., reduce_max - 8) (:512) -> (. + 4) / 4 (:513).
// The filterbank is triangular, i.e. sparse: each mel bin touches a short run of FFT bins, so
// the "dot" is ~2 multiply-adds per FFT bin and the kernel is bound by reading z once
// (4 * fft_length bytes per frame: only the lower half-spectrum is touched; a one-sided
// spectrum from nxs_stft_onesided_f32_dev is accepted through z_ld).
//
// One warp per frame: the frame's power spectrum goes to shared memory, lane l then owns mel
// bins l, l+32, ...  The per-channel maximum (Nx.reduce_max over a vectorised tensor reduces
// per entry) is gathered with an ordered-integer atomicMax; a second, tiny kernel applies the
// dynamic-range clamp and the affine map.
#include <math.h>
#include <string.h>

#include "nxs_common.cuh"

namespace nxs {

struct MelArgs {
  const float2* z;  // [C][M][z_ld]
  int64_t M, z_ld, total_frames;
  int half;      // fft_length / 2 bins used
  int mel_bins;
  const float* wts;  // packed nonzero filter weights
  const int* start;  // [mel_bins] first FFT bin of each filter's run
  const int* count;  // [mel_bins] run length
  const int* offs;   // [mel_bins] offset of the run in wts
  float* out;        // [C][M][mel_bins]
  int* chmax;        // [C] ordered-int maximum of log_spec per channel
  // a warp walks tiles of kMelTile consecutive frames of one channel (one atomicMax per tile: with frames dealt
  // round-robin a warp changed channel on almost every frame when M is small, 360 k contended atomics on cfg5's shape)
  int64_t tiles_per_channel, total_tiles;
  FastDiv div_tpc;   // tile / tiles_per_channel when total_tiles < 2^31 (use_fastdiv)
  int use_fastdiv;
};
constexpr int kMelTile = 8;

__device__ __forceinline__ int float_key(float v) {
  const int b = __float_as_int(v);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) mel_logspec_kernel(const MelArgs a) {
  extern __shared__ float pw_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* pw = pw_all + (size_t)warp * a.half;
  for (int64_t tile = (int64_t)blockIdx.x * wpb + warp; tile < a.total_tiles; tile += (int64_t)gridDim.x * wpb) {
    const int64_t c = a.use_fastdiv ? (int64_t)a.div_tpc.div((int)tile) : tile / a.tiles_per_channel;
    const int64_t m0 = (tile - c * a.tiles_per_channel) * kMelTile;
    const int64_t m1 = m0 + kMelTile < a.M ? m0 + kMelTile : a.M;
    float cur_max = -INFINITY;
    for (int64_t f = c * a.M + m0; f < c * a.M + m1; ++f) {
      const float2* __restrict__ zf = a.z + f * a.z_ld;
      int k = lane;
      for (; k + 96 < a.half; k += 128) {  // four independent loads in flight per lane
        const float2 v0 = __ldcs(zf + k), v1 = __ldcs(zf + k + 32), v2 = __ldcs(zf + k + 64), v3 = __ldcs(zf + k + 96);
        pw[k] = v0.x * v0.x + v0.y * v0.y;  // Nx.abs(z) ** 2
        pw[k + 32] = v1.x * v1.x + v1.y * v1.y;
        pw[k + 64] = v2.x * v2.x + v2.y * v2.y;
        pw[k + 96] = v3.x * v3.x + v3.y * v3.y;
      }
      for (; k < a.half; k += 32) {
        const float2 v = __ldcs(zf + k);
        pw[k] = v.x * v.x + v.y * v.y;
      }
      __syncwarp();
      float m = -INFINITY;
      for (int j = lane; j < a.mel_bins; j += 32) {
        const int s = a.start[j], n = a.count[j];
        const float* __restrict__ w = a.wts + a.offs[j];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // independent partial sums (same order as the fused epilogue)
        const float* __restrict__ pp = pw + s;
        int i = 0;
        for (; i + 4 <= n; i += 4) {
          a0 = fmaf(pp[i], __ldg(w + i), a0);
          a1 = fmaf(pp[i + 1], __ldg(w + i + 1), a1);
          a2 = fmaf(pp[i + 2], __ldg(w + i + 2), a2);
          a3 = fmaf(pp[i + 3], __ldg(w + i + 3), a3);
        }
        for (; i < n; ++i) a0 = fmaf(pp[i], __ldg(w + i), a0);
        const float acc = (a0 + a1) + (a2 + a3);
        // Nx.log(Nx.clip(mel, 1e-10, inf)) / Nx.log(10)
        const float v = __log2f(fmaxf(acc, 1.0e-10f)) * 0.30102999566f;
        a.out[f * a.mel_bins + j] = v;
        m = fmaxf(m, v);
      }
  #pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      cur_max = fmaxf(cur_max, m);
      __syncwarp();
    }
    if (lane == 0) atomicMax(a.chmax + c, float_key(cur_max));
  }
}

// log_spec = max(log_spec, reduce_max(log_spec) - 8); (log_spec + 4) / 4   (lib/nx_signal.ex:512-513)
// blockIdx.y walks the channels (no per-element division), 128-bit accesses when the rows allow
__global__ void __launch_bounds__(256) mel_finalize_kernel(float* __restrict__ out, int64_t per_channel, int64_t channels,
                                                           const int* __restrict__ chmax, int vec4) {
  for (int64_t c = blockIdx.y; c < channels; c += gridDim.y) {
    const float floor_v = key_float(chmax[c]) - 8.0f;
    float* __restrict__ row = out + c * per_channel;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (vec4) {
      float4* __restrict__ r4 = reinterpret_cast<float4*>(row);
      for (int64_t i = tid; i < per_channel / 4; i += nthr) {
        float4 v = r4[i];
        v.x = (fmaxf(v.x, floor_v) + 4.0f) / 4.0f;
        v.y = (fmaxf(v.y, floor_v) + 4.0f) / 4.0f;
        v.z = (fmaxf(v.z, floor_v) + 4.0f) / 4.0f;
        v.w = (fmaxf(v.w, floor_v) + 4.0f) / 4.0f;
        r4[i] = v;
      }
    } else {
      for (int64_t i = tid; i < per_channel; i += nthr) row[i] = (fmaxf(row[i], floor_v) + 4.0f) / 4.0f;
    }
  }
}

static int launch_mel_finalize(nxs_ctx* ctx, float* out, int64_t per_channel, int64_t channels, const int* chmax,
                               cudaStream_t st) {
  const int vec4 = per_channel % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t work = vec4 ? per_channel / 4 : per_channel;
  int64_t gy = channels < 65535 ? channels : 65535;
  int64_t gx = (work + 255) / 256;
  const int64_t cap = (int64_t(ctx->sm_count) * 16 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  mel_finalize_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(out, per_channel, channels, chmax, vec4);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  return NXS_OK;
}

// sparse filterbank of one (fft_length, mel_bins, sampling_rate, max_mel, f_sp) on the device
static int get_mel_bank(nxs_ctx* ctx, int64_t nfft, int64_t mel_bins, double sr, double max_mel, double f_sp,
                        MelBank** out) {
  for (auto& b : ctx->mel_banks)
    if (b.nfft == nfft && b.mel_bins == mel_bins && b.sr == sr && b.max_mel == max_mel && b.f_sp == f_sp) {
      *out = &b;
      return NXS_OK;
    }
  std::vector<float> dense(size_t(mel_bins) * size_t(nfft));
  int rc = nxs_mel_filters_f32(nfft, mel_bins, sr, max_mel, f_sp, dense.data());
  if (rc) return rc;
  const int64_t half = nfft / 2;
  std::vector<int> start(mel_bins), count(mel_bins), offs(mel_bins);
  std::vector<float> wts;
  for (int64_t j = 0; j < mel_bins; ++j) {
    const float* row = dense.data() + j * nfft;
    int64_t lo = half, hi = -1;
    for (int64_t k = 0; k < half; ++k)
      if (row[k] != 0.0f) {  // NaN weights (degenerate filters) compare unequal and are kept
        if (k < lo) lo = k;
        hi = k;
      }
    offs[j] = (int)wts.size();
    if (hi < 0) {
      start[j] = 0;
      count[j] = 0;
    } else {
      start[j] = (int)lo;
      count[j] = (int)(hi - lo + 1);
      wts.insert(wts.end(), row + lo, row + hi + 1);
    }
  }
  if (wts.empty()) wts.push_back(0.f);
  MelBank b;
  b.nfft = nfft;
  b.mel_bins = mel_bins;
  b.sr = sr;
  b.max_mel = max_mel;
  b.f_sp = f_sp;
  b.nw = (int)wts.size();
  const size_t ib = size_t(mel_bins) * sizeof(int);
  NXS_CUDA(ctx, cudaMalloc(&b.d_wts, wts.size() * sizeof(float)));
  NXS_CUDA(ctx, cudaMalloc(&b.d_idx, 3 * ib));
  NXS_CUDA(ctx, cudaMemcpy(b.d_wts, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(b.d_idx, start.data(), ib, cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(b.d_idx + mel_bins, count.data(), ib, cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(b.d_idx + 2 * mel_bins, offs.data(), ib, cudaMemcpyHostToDevice));
  ctx->mel_banks.push_back(b);
  *out = &ctx->mel_banks.back();
  return NXS_OK;
}

int get_mel_layout(nxs_ctx* ctx, MelBank* bank, int B, const MelLayout** out) {
  for (auto& l : bank->layouts)
    if (l.B == B) {
      *out = &l;
      return NXS_OK;
    }
  MelLayout lay;
  lay.B = B;
  const int64_t nfft = bank->nfft, half = nfft / 2, mel_bins = bank->mel_bins;
  auto finish = [&]() {
    bank->layouts.push_back(lay);
    *out = &bank->layouts.back();
    return NXS_OK;
  };
  if (B < 2 || B > 16 || (B & 1) || half % B != 0 || mel_bins + 3 > (int64_t(1) << 20)) return finish();
  std::vector<float> dense(size_t(mel_bins) * size_t(nfft));
  int rc = nxs_mel_filters_f32(nfft, mel_bins, bank->sr, bank->max_mel, bank->f_sp, dense.data());
  if (rc) return rc;
  // per bin: the (at most two, consecutive) filters it feeds; jl must not decrease
  std::vector<int> jl(half);
  std::vector<float2> w2(half, make_float2(0.f, 0.f));
  int prev = -1;
  int64_t last_nz = -1;
  for (int64_t k = 0; k < half; ++k) {
    int nz[3], n = 0;
    for (int64_t j = 0; j < mel_bins && n < 3; ++j)
      if (dense[j * nfft + k] != 0.0f) nz[n++] = (int)j;  // NaN (degenerate filter) counts as a weight
    if (n == 0) {
      jl[k] = prev;
    } else if (n == 1) {
      const float w = dense[int64_t(nz[0]) * nfft + k];
      if (prev <= nz[0] - 1) {
        jl[k] = nz[0] - 1;
        w2[k].y = w;
      } else if (prev == nz[0]) {
        jl[k] = nz[0];
        w2[k].x = w;
      } else {
        return finish();
      }
    } else if (n == 2 && nz[1] == nz[0] + 1 && prev <= nz[0]) {
      jl[k] = nz[0];
      w2[k] = make_float2(dense[int64_t(nz[0]) * nfft + k], dense[int64_t(nz[1]) * nfft + k]);
    } else {
      return finish();
    }
    if (n > 0) last_nz = k;
    prev = jl[k];
  }
  // bins above the last weight form a segment of their own that no filter reads
  for (int64_t k = last_nz + 1; k < half; ++k) jl[k] = (int)mel_bins;
  const int64_t T = half / B;
  std::vector<int> desc(T), ps(mel_bins + 3);
  int pieces = 0;
  int seg_next = 0;  // segments below this one have their first piece recorded (segment = jl + 1)
  for (int64_t t = 0; t < T; ++t) {
    unsigned mask = 0;
    bool skip = true;
    const int base = pieces;
    for (int b = 0; b < B; ++b) {
      const int64_t k = t * B + b;
      const bool fresh = b == 0 || jl[k] != jl[k - 1];
      if (b > 0 && fresh) mask |= 1u << b;
      if (fresh) {
        if (k == 0 || jl[k] != jl[k - 1])
          for (; seg_next <= jl[k] + 1; ++seg_next) ps[seg_next] = pieces;
        ++pieces;
      }
      if (jl[k] != (int)mel_bins) skip = false;
    }
    desc[t] = base | (int)(mask << 12) | (skip ? (int)0x80000000u : 0);
  }
  for (; seg_next < (int)mel_bins + 3; ++seg_next) ps[seg_next] = pieces;
  if (pieces > half / 2 || pieces >= 4096) return finish();  // pieces live in half of the exchange buffer
  NXS_CUDA(ctx, cudaMalloc(&lay.d_w2, size_t(half) * sizeof(float2)));
  NXS_CUDA(ctx, cudaMalloc(&lay.d_desc, size_t(T) * sizeof(int)));
  NXS_CUDA(ctx, cudaMalloc(&lay.d_ps, ps.size() * sizeof(int)));
  // load-major order for the kernel's 128-bit reads: float4 i of thread t = bins B t + 2 i, B t + 2 i + 1
  // sits at float4 index i * T + t (consecutive lanes -> consecutive 16-byte words)
  std::vector<float2> w2s(half);
  for (int64_t t = 0; t < T; ++t)
    for (int i = 0; i < B / 2; ++i) {
      w2s[2 * (i * T + t)] = w2[t * B + 2 * i];
      w2s[2 * (i * T + t) + 1] = w2[t * B + 2 * i + 1];
    }
  NXS_CUDA(ctx, cudaMemcpy(lay.d_w2, w2s.data(), size_t(half) * sizeof(float2), cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(lay.d_desc, desc.data(), size_t(T) * sizeof(int), cudaMemcpyHostToDevice));
  NXS_CUDA(ctx, cudaMemcpy(lay.d_ps, ps.data(), ps.size() * sizeof(int), cudaMemcpyHostToDevice));
  lay.ok = true;
  return finish();
}

int launch_stft_to_mel(nxs_ctx* ctx, const float2* z, int64_t channels, int64_t num_frames, int64_t z_ld,
                       int64_t fft_length, int64_t mel_bins, double sampling_rate, double max_mel, double f_sp,
                       float* out, cudaStream_t st) {
  if (channels <= 0 || num_frames <= 0) return NXS_OK;
  const int64_t half = fft_length / 2;
  if (half > 49152 || mel_bins > (int64_t(1) << 20)) return NXS_EUNSUPPORTED;
  MelBank* bank = nullptr;
  int rc = get_mel_bank(ctx, fft_length, mel_bins, sampling_rate, max_mel, f_sp, &bank);
  if (rc) return rc;
  rc = ensure_scratch(ctx, size_t(channels) * sizeof(int));
  if (rc) return rc;
  int* chmax = (int*)ctx->d_scratch;
  // 0x80808080 is the ordered key of about -3.4e38: below every log10 value the kernel can produce
  NXS_CUDA(ctx, cudaMemsetAsync(chmax, 0x80, size_t(channels) * sizeof(int), st));
  MelArgs a;
  a.z = z;
  a.M = num_frames;
  a.z_ld = z_ld;
  a.total_frames = channels * num_frames;
  a.tiles_per_channel = (num_frames + kMelTile - 1) / kMelTile;
  a.total_tiles = a.tiles_per_channel * channels;
  a.use_fastdiv = a.total_tiles < (int64_t(1) << 31);
  a.div_tpc = FastDiv(a.use_fastdiv ? (int)a.tiles_per_channel : 1);
  a.half = (int)half;
  a.mel_bins = (int)mel_bins;
  a.wts = bank->d_wts;
  a.start = bank->d_idx;
  a.count = bank->d_idx + mel_bins;
  a.offs = bank->d_idx + 2 * mel_bins;
  a.out = out;
  a.chmax = chmax;
  int wpb = 8;
  while (wpb > 1 && size_t(wpb) * half * sizeof(float) > 96 * 1024) wpb >>= 1;
  const size_t smem = size_t(wpb) * (half > 0 ? half : 1) * sizeof(float);
  NXS_CUDA(ctx, cudaFuncSetAttribute(mel_logspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
  int64_t grid = (a.total_tiles + wpb - 1) / wpb;
  if (grid > int64_t(ctx->sm_count) * 8) grid = int64_t(ctx->sm_count) * 8;
  prof_begin(ctx, st);
  mel_logspec_kernel<<<(unsigned)grid, wpb * 32, smem, st>>>(a);
  ctx->launches++;
  NXS_CUDA(ctx, cudaGetLastError());
  rc = launch_mel_finalize(ctx, out, num_frames * mel_bins, channels, chmax, st);
  prof_end(ctx, st);  // both kernels count
  return rc;
}

// stft -> log-mel in one pass: the STFT kernel's epilogue reduces each frame's spectrum to mel
// powers on chip (the spectrum is never stored), then the same finalize pass as above.
int launch_stft_mel(nxs_ctx* ctx, const float* x, int64_t channels, int64_t length, int64_t x_ld,
                    const float* window, int64_t frame_length, int64_t hop, int64_t fft_length, const PadGeom& g,
                    int64_t num_frames, int scaling, double sampling_rate, int64_t mel_bins, double max_mel,
                    double f_sp, float* out, cudaStream_t st) {
  if (channels <= 0 || num_frames <= 0) return NXS_OK;
  if (mel_bins > (int64_t(1) << 16)) return NXS_EUNSUPPORTED;
  MelBank* bank = nullptr;
  int rc = get_mel_bank(ctx, fft_length, mel_bins, sampling_rate, max_mel, f_sp, &bank);
  if (rc) return rc;
  rc = ensure_scratch(ctx, size_t(channels) * sizeof(int));
  if (rc) return rc;
  int* chmax = (int*)ctx->d_scratch;
  NXS_CUDA(ctx, cudaMemsetAsync(chmax, 0x80, size_t(channels) * sizeof(int), st));
  MelEpilogue mel{bank, out, chmax};
  rc = launch_stft(ctx, x, channels, length, x_ld, window, frame_length, hop, fft_length, g, num_frames, scaling,
                   sampling_rate, nullptr, fft_length, 0, st, &mel);
  if (rc) return rc;
  return launch_mel_finalize(ctx, out, num_frames * mel_bins, channels, chmax, st);
}

}  // namespace nxs
