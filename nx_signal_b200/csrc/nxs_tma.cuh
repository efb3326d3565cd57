// nxs_tma.cuh -- 1-D bulk async copies (TMA, cp.async.bulk) signalled through mbarriers, and
// the per-frame-group named barrier, shared by the STFT / ISTFT / FIR kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nxs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int T>
struct GroupSync {
  int id;
  __device__ __forceinline__ void operator()() const {
    if constexpr (T > 32) asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(T) : "memory");
    else if constexpr (T == 32) __syncwarp();  // a group of one warp needs no barrier unit
    else {
      // several groups per warp: each synchronises over its OWN lanes only.  The groups of a warp walk segments of
      // different lengths, so their barrier counts differ; a full-warp __syncwarp would pair one group's barrier
      // with an unrelated one of its neighbour (harmless for the data, but not what the code means, and
      // compute-sanitizer's racecheck rightly cannot follow it)
      const unsigned lane = threadIdx.x & 31u;
      __syncwarp(((1u << T) - 1u) << (lane / T * T));
    }
  }
};

}  // namespace nxs
