// pcie_probe.cu -- measures the pieces of the _host STFT pipeline in isolation on the GPU box:
// contiguous vs pitched D2H, zero-copy stores from a kernel, and the host mirror pass.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <immintrin.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void zc_store(const float2* __restrict__ src, float2* __restrict__ dst, long rows, int nout, int spitch, int dpitch) {
  for (long r = blockIdx.x; r < rows; r += gridDim.x)
    for (int k = threadIdx.x; k < nout; k += blockDim.x) dst[r * dpitch + k] = src[r * spitch + k];
}

static void mirror(float* z, long nfft, long r0, long r1) {
  const long kmax = nfft - (nfft / 2 + 1);
  const __m128 sign = _mm_castsi128_ps(_mm_set_epi32((int)0x80000000u, 0, (int)0x80000000u, 0));
  for (long r = r0; r < r1; ++r) {
    float* row = z + 2 * r * nfft;
    long k = 1;
    for (; k + 1 <= kmax; k += 2) {
      __m128 v = _mm_loadu_ps(row + 2 * k);
      v = _mm_shuffle_ps(v, v, _MM_SHUFFLE(1, 0, 3, 2));
      v = _mm_xor_ps(v, sign);
      _mm_stream_ps(row + 2 * (nfft - k - 1), v);
    }
    for (; k <= kmax; ++k) { row[2 * (nfft - k)] = row[2 * k]; row[2 * (nfft - k) + 1] = -row[2 * k + 1]; }
  }
}

__attribute__((target("avx512f,avx512dq"))) static void mirror512(float* z, long nfft, long r0, long r1) {
  // rows are 64-byte aligned; the mirror half is bins nout .. nfft-1; full 64-byte lines from bin 520 on (nfft = 1024)
  const long nout = nfft / 2 + 1;
  const __m512i rev = _mm512_set_epi64(0, 1, 2, 3, 4, 5, 6, 7);
  const __m512 sign = _mm512_castsi512_ps(_mm512_set_epi32((int)0x80000000u, 0, (int)0x80000000u, 0, (int)0x80000000u, 0, (int)0x80000000u, 0,
                                                         (int)0x80000000u, 0, (int)0x80000000u, 0, (int)0x80000000u, 0, (int)0x80000000u, 0));
  for (long r = r0; r < r1; ++r) {
    float* row = z + 2 * r * nfft;
    long j = nout;                       // destination bin
    const long jal = (nout + 7) / 8 * 8;  // first 64-byte aligned destination bin
    for (; j < jal; ++j) { row[2 * j] = row[2 * (nfft - j)]; row[2 * j + 1] = -row[2 * (nfft - j) + 1]; }
    for (; j + 8 <= nfft; j += 8) {
      // dest bins j..j+7  <- conj(src bins nfft-j .. nfft-j-7): load src bins [nfft-j-7, nfft-j], reverse
      __m512 v = _mm512_loadu_ps(row + 2 * (nfft - j - 7));
      v = _mm512_castpd_ps(_mm512_permutexvar_pd(rev, _mm512_castps_pd(v)));
      v = _mm512_xor_ps(v, sign);
      _mm512_stream_ps(row + 2 * j, v);
    }
  }
}
static void mirror_plain(float* z, long nfft, long r0, long r1) {
  const long kmax = nfft - (nfft / 2 + 1);
  for (long r = r0; r < r1; ++r) {
    float* row = z + 2 * r * nfft;
    for (long k = 1; k <= kmax; ++k) { row[2 * (nfft - k)] = row[2 * k]; row[2 * (nfft - k) + 1] = -row[2 * k + 1]; }
  }
}

int main(int argc, char** argv) {
  const long rows = argc > 1 ? atol(argv[1]) : 899976;
  const int nfft = 1024, nout = 513, spitch = 516;
  float2 *h, *d;
  CK(cudaMallocHost(&h, rows * nfft * sizeof(float2)));
  CK(cudaMalloc(&d, rows * spitch * sizeof(float2)));
  CK(cudaMemset(d, 1, rows * spitch * sizeof(float2)));
  cudaStream_t st, st2; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
  for (int rep = 0; rep < 2; ++rep) {
    double t = now();
    CK(cudaMemcpyAsync(h, d, rows * nout * sizeof(float2), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    double dt = now() - t; printf("contiguous D2H %.2f GB: %.1f ms  %.1f GB/s\n", rows * nout * 8 / 1e9, dt * 1e3, rows * nout * 8 / dt / 1e9);
    t = now();
    CK(cudaMemcpy2DAsync(h, nfft * 8, d, spitch * 8, nout * 8, rows, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    dt = now() - t; printf("pitched D2H (4104 B rows -> 8192 pitch): %.1f ms  %.1f GB/s\n", dt * 1e3, rows * nout * 8 / dt / 1e9);
    t = now();
    CK(cudaMemcpy2DAsync(h, nfft * 8, d, spitch * 8, 512 * 8, rows, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    dt = now() - t; printf("pitched D2H (4096 B rows -> 8192 pitch): %.1f ms  %.1f GB/s\n", dt * 1e3, rows * 512 * 8 / dt / 1e9);
    t = now();
    CK(cudaMemcpyAsync(d, h + rows * nfft / 2, 921600000, cudaMemcpyHostToDevice, st2)); CK(cudaDeviceSynchronize());
    dt = now() - t; printf("H2D 0.92 GB alone: %.1f ms  %.1f GB/s\n", dt * 1e3, 0.9216 / dt);
    for (long slab : {1L << 20, 4L << 20, 16L << 20}) {
      const long srows = slab / (nout * 8);
      t = now();
      for (long r0 = 0; r0 < rows; r0 += srows) {
        const long n = r0 + srows < rows ? srows : rows - r0;
        CK(cudaMemcpy2DAsync(h + r0 * nfft, nfft * 8, d + r0 * spitch, spitch * 8, nout * 8, n, cudaMemcpyDeviceToHost, st));
      }
      CK(cudaStreamSynchronize(st));
      dt = now() - t; printf("pitched D2H in %ld MiB slabs: %.1f ms  %.1f GB/s\n", slab >> 20, dt * 1e3, rows * nout * 8 / dt / 1e9);
    }
    t = now();
    zc_store<<<148 * 4, 256, 0, st>>>(d, h, rows, nout, spitch, nfft); CK(cudaStreamSynchronize(st));
    dt = now() - t; printf("zero-copy kernel stores (513 bins/row): %.1f ms  %.1f GB/s\n", dt * 1e3, rows * nout * 8 / dt / 1e9);
    t = now();
    CK(cudaMemcpyAsync(h, d, rows * nfft * sizeof(float2) / 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(d, h + rows * nfft / 2, 921600000, cudaMemcpyHostToDevice, st2)); CK(cudaDeviceSynchronize());
    dt = now() - t; printf("contiguous D2H 3.7 GB with concurrent H2D 0.92 GB: %.1f ms\n", dt * 1e3);
  }
  for (int T : {1, 2, 4, 8, 12, 16}) {
    double t = now();
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (long b = 0; b < (rows + 63) / 64; ++b) mirror((float*)h, nfft, b * 64, (b * 64 + 64 < rows) ? b * 64 + 64 : rows);
    double dt = now() - t; printf("mirror pass alone, %2d threads: %.1f ms  (%.1f GB/s written, same read)\n", T, dt * 1e3, rows * 511 * 8 / dt / 1e9);
  }
  for (int T : {8, 16}) {
    double t = now();
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (long b = 0; b < (rows + 63) / 64; ++b) mirror512((float*)h, nfft, b * 64, (b * 64 + 64 < rows) ? b * 64 + 64 : rows);
    double dt = now() - t; printf("mirror512 (64B NT stores) alone, %2d threads: %.1f ms\n", T, dt * 1e3);
    t = now();
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (long b = 0; b < (rows + 63) / 64; ++b) mirror_plain((float*)h, nfft, b * 64, (b * 64 + 64 < rows) ? b * 64 + 64 : rows);
    dt = now() - t; printf("mirror plain stores alone, %2d threads: %.1f ms\n", T, dt * 1e3);
    t = now();
    CK(cudaMemcpy2DAsync(h, nfft * 8, d, spitch * 8, nout * 8, rows, cudaMemcpyDeviceToHost, st));
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (long b = 0; b < (rows + 63) / 64; ++b) mirror512((float*)h, nfft, b * 64, (b * 64 + 64 < rows) ? b * 64 + 64 : rows);
    double dtm = now() - t;
    CK(cudaStreamSynchronize(st));
    dt = now() - t; printf("mirror512 (%d thr) concurrent with pitched D2H: mirror done %.1f ms, both done %.1f ms\n", T, dtm * 1e3, dt * 1e3);
  }
  {  // D2H concurrent with a pure streaming-read load and a pure streaming-write load
    double t = now();
    CK(cudaMemcpy2DAsync(h, nfft * 8, d, spitch * 8, nout * 8, rows, cudaMemcpyDeviceToHost, st));
    double acc = 0;
#pragma omp parallel for num_threads(16) reduction(+ : acc)
    for (long i = 0; i < rows * nfft / 2; ++i) acc += ((float*)h)[2 * i + rows * nfft];  // reads the upper half of the buffer
    double dtm = now() - t;
    CK(cudaStreamSynchronize(st));
    double dt = now() - t; printf("3.7 GB CPU read concurrent with D2H: cpu %.1f ms, both %.1f ms (%g)\n", dtm * 1e3, dt * 1e3, acc);
  }
  // mirror concurrent with a pitched D2H of the whole buffer
  for (int T : {8, 16}) {
    double t = now();
    CK(cudaMemcpy2DAsync(h, nfft * 8, d, spitch * 8, nout * 8, rows, cudaMemcpyDeviceToHost, st));
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (long b = 0; b < (rows + 63) / 64; ++b) mirror((float*)h, nfft, b * 64, (b * 64 + 64 < rows) ? b * 64 + 64 : rows);
    double dtm = now() - t;
    CK(cudaStreamSynchronize(st));
    double dt = now() - t; printf("mirror (%d thr) concurrent with pitched D2H: mirror done %.1f ms, both done %.1f ms\n", T, dtm * 1e3, dt * 1e3);
  }
  return 0;
}
