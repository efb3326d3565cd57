// nxs_fft.cuh -- register/shared-memory Stockham FFT building blocks for sm_100a.
//
// The reference does every FFT as `Nx.fft` / `Nx.ifft` (lib/nx_signal.ex:102, :609;
// lib/nx_signal/transforms.ex:10,19), executed by Nx.BinaryBackend as a recursive
// radix-2 over Elixir lists.  Here one FFT of N complex points is computed by a
// group of T threads holding P = N/T points each in registers: radix-R butterflies
// in registers, one shared-memory exchange per pass (autosort, natural order in and
// out), bank-conflict-free padded layouts found by tools/smem_layout_search.py.
// Index algebra validated by tools/stockham_model.py.
//
// Pass p (radix R, NS = product of earlier radices), butterfly j = t + b*T:
//   in  : x[j + q*N/R]              q = 0..R-1   (stride-1 across threads)
//   tw  : W_{NS*R}^{q*(j mod NS)}
//   out : y[expand(j) + q*NS],  expand(j) = (j/NS)*NS*R + (j mod NS)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace nxs {

typedef float2 cpx;

// Complex arithmetic, scalar or (PK) on the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: one
// issue slot for the real and the imaginary lane).  The transforms here are bound by instruction issue, not by
// the FMA pipe, and a complex value already sits in an aligned register pair; the half swaps, broadcasts and
// sign flips below are operand modifiers of the packed instructions (R.F32x2.LO_HI.NP, R.F32), not
// instructions.  Every packed operation rounds exactly like its two scalar halves.  Whether it pays is a
// per-kernel matter -- where the register allocator has to move values into aligned pairs the moves eat the
// saving (profiles/r02x_packed_fp32x2.txt) -- so the choice rides on the plan (Plan::PK).
template <bool PK>
__device__ __forceinline__ cpx cadd_(cpx a, cpx b) {
  if constexpr (PK) return __fadd2_rn(a, b);
  else return make_float2(a.x + b.x, a.y + b.y);
}
template <bool PK>
__device__ __forceinline__ cpx csub_(cpx a, cpx b) {
  if constexpr (PK) return __fadd2_rn(a, make_float2(-b.x, -b.y));
  else return make_float2(a.x - b.x, a.y - b.y);
}
template <bool PK>
__device__ __forceinline__ cpx cmul_(cpx a, cpx b) {
  // the swizzled, half-negated operand goes FIRST, the broadcast second: that is the operand order FFMA2 has the
  // modifiers for (the other order costs a MOV and an FADD per product)
#if defined(NXS_CMUL_F1)  // A/B build: the operand order that costs a MOV and an FADD
  if constexpr (PK) return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), b));
#else
  if constexpr (PK) return __ffma2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y), __fmul2_rn(make_float2(a.x, a.x), b));
#endif
  else return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * (c - i s)
template <bool PK>
__device__ __forceinline__ cpx crot_(cpx a, float c, float s) {
  if constexpr (PK) return __ffma2_rn(make_float2(a.y, -a.x), make_float2(s, s), __fmul2_rn(a, make_float2(c, c)));
  else return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}
// a * s, s real
template <bool PK>
__device__ __forceinline__ cpx cscale_(cpx a, float s) {
  if constexpr (PK) return __fmul2_rn(a, make_float2(s, s));
  else return make_float2(a.x * s, a.y * s);
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cadd_<false>(a, b); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return csub_<false>(a, b); }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return cmul_<false>(a, b); }
__device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }

constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }
constexpr int bitrev(int k, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (k & 1); k >>= 1; }
  return r;
}

// a * W_16^I, I in [0,8), W_16 = exp(-2 pi i / 16); constants become FFMA immediates
template <int I, bool PK = false>
__device__ __forceinline__ cpx mul_w16(cpx a) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
  if constexpr (I == 0) return a;
  else if constexpr (I == 4) return make_float2(a.y, -a.x);
  else if constexpr (I == 2) return cscale_<PK>(cadd_<PK>(a, make_float2(a.y, -a.x)), H);
  else if constexpr (I == 6) return cscale_<PK>(csub_<PK>(make_float2(a.y, -a.x), a), H);
  else if constexpr (I == 1) return crot_<PK>(a, C1, S1);
  else if constexpr (I == 3) return crot_<PK>(a, S1, C1);
  else if constexpr (I == 5) return crot_<PK>(a, -S1, C1);
  else return crot_<PK>(a, -C1, S1);  // I == 7
}

// a * W_32^I, I in [0,16): the first stage of the radix-32 butterfly (odd I; even I are W_16 powers)
template <int I, bool PK = false>
__device__ __forceinline__ cpx mul_w32(cpx a) {
  if constexpr (I % 2 == 0) return mul_w16<I / 2, PK>(a);
  else {
    // cos(2 pi j / 32), j = 0 .. 8
    constexpr float c[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f};
    constexpr float co = I <= 8 ? c[I] : -c[16 - I];
    constexpr float si = I <= 8 ? c[8 - I] : c[I - 8];
    return crot_<PK>(a, co, si);
  }
}

// a * W_64^I, I in [0,32): the first stage of the radix-64 butterfly (odd I; even I are W_32 powers)
template <int I, bool PK = false>
__device__ __forceinline__ cpx mul_w64(cpx a) {
  if constexpr (I % 2 == 0) return mul_w32<I / 2, PK>(a);
  else {
    // cos(2 pi j / 64), j = 0 .. 16
    constexpr float c[17] = {1.0f, 0.99518472667219689f, 0.98078528040323043f, 0.95694033573220887f, 0.92387953251128674f,
                             0.88192126434835503f, 0.83146961230254524f, 0.77301045336273696f, 0.70710678118654752f,
                             0.63439328416364549f, 0.55557023301960218f, 0.47139673682599764f, 0.38268343236508977f,
                             0.29028467725446236f, 0.19509032201612825f, 0.09801714032956060f, 0.0f};
    constexpr float co = I <= 16 ? c[I] : -c[32 - I];
    constexpr float si = I <= 16 ? c[16 - I] : c[I - 16];
    return crot_<PK>(a, co, si);
  }
}

// In-place decimation-in-frequency DFT of R points (forward, e^{-i...}).
// Result X[k] is left at v[bitrev(k, log2 R)].
template <int R, int I, bool PK>
struct DifStage {
  static __device__ __forceinline__ void run(cpx* v) {
    if constexpr (I < R / 2) {
      cpx a = v[I], b = v[I + R / 2];
      v[I] = cadd_<PK>(a, b);
      if constexpr (R == 64) v[I + R / 2] = mul_w64<I, PK>(csub_<PK>(a, b));
      else if constexpr (R == 32) v[I + R / 2] = mul_w32<I, PK>(csub_<PK>(a, b));
      else v[I + R / 2] = mul_w16<I * (16 / R), PK>(csub_<PK>(a, b));
      DifStage<R, I + 1, PK>::run(v);
    }
  }
};

template <int R, bool PK = false>
struct Dft {
  static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32 || R == 64, "radix");
  static __device__ __forceinline__ void run(cpx* v) {
    if constexpr (R == 2) {
      cpx a = v[0], b = v[1];
      v[0] = cadd_<PK>(a, b);
      v[1] = csub_<PK>(a, b);
    } else {
      DifStage<R, 0, PK>::run(v);
      Dft<R / 2, PK>::run(v);
      Dft<R / 2, PK>::run(v + R / 2);
    }
  }
};

// ---------------------------------------------------------------------------------------
// Plan: N points, T threads, up to 4 passes (unused radices = 1).
// ---------------------------------------------------------------------------------------
template <int N_, int T_, int R0, int R1 = 1, int R2 = 1, int R3 = 1, bool PK_ = false>
struct Plan {
  static constexpr int N = N_, T = T_, P = N_ / T_;
  static constexpr bool PK = PK_;  // complex arithmetic on the packed fp32x2 instructions
  static constexpr int NP = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
  static_assert(R0 * R1 * R2 * R3 == N_, "radices must multiply to N");
  static_assert(N_ % T_ == 0, "T must divide N");
  static constexpr int R(int p) { return p == 0 ? R0 : p == 1 ? R1 : p == 2 ? R2 : R3; }
  static constexpr int NS(int p) { return p == 0 ? 1 : p == 1 ? R0 : p == 2 ? R0 * R1 : R0 * R1 * R2; }
  // exchange after pass p is laid out as pad(i) = i + (i >> A) * C   (A < 0: no padding)
  static constexpr int padA(int p) {
    return NS(p) >= 16 ? -1 : (ilog2(NS(p) * R(p)) > 4 ? ilog2(NS(p) * R(p)) : 4);
  }
  static constexpr int padC(int p) { return NS(p) >= 16 ? 0 : NS(p); }
  // complex elements one exchange buffer needs (largest padded layout, multiple of 16)
  static constexpr int bufNeed(int p) {
    return padA(p) < 0 ? N_ : N_ + ((N_ - 1) >> padA(p)) * padC(p) + 1;
  }
  static constexpr int bufMax() {
    int m = N_;
    for (int p = 0; p + 1 < NP; ++p) m = bufNeed(p) > m ? bufNeed(p) : m;
    return (m + 15) / 16 * 16;
  }
  static constexpr int BUF = bufMax();
  // per-pass twiddle table: for p >= 1, entries [(q-1)*NS + k], q = 1..R-1, k < NS
  static constexpr int twOffset(int p) {
    int o = 0;
    for (int i = 1; i < p; ++i) o += (R(i) - 1) * NS(i);
    return o;
  }
  static constexpr int TW_TOTAL = twOffset(NP);
  // compact table (TwDeriveC): only the power-of-two rows W^(2^r k), r < log2 R, per pass
  static constexpr int twcOffset(int p) {
    int o = 0;
    for (int i = 1; i < p; ++i) o += ilog2(R(i)) * NS(i);
    return o;
  }
  static constexpr int TWC_TOTAL = twcOffset(NP);
  // number of distinct twiddle sets a thread needs in pass p (1 when k = t mod NS for every b)
  static constexpr int twSets(int p) { return (T_ % NS(p) == 0) ? 1 : (P / R(p)); }
  static constexpr int twRegOffset(int p) {
    int o = 0;
    for (int i = 1; i < p; ++i) o += (R(i) - 1) * twSets(i);
    return o;
  }
  static constexpr int TW_REGS = twRegOffset(NP);
};

template <int A, int C>
__device__ __forceinline__ int pad_idx(int i) {
  if constexpr (A < 0) return i;
  else return i + (i >> A) * C;
}

// Twiddles held in registers for the lifetime of a persistent CTA.
template <class PL>
struct TwRegs {
  static constexpr bool PK = PL::PK;
  cpx w[PL::TW_REGS > 0 ? PL::TW_REGS : 1];
  template <int PASS>
  __device__ __forceinline__ void init_pass(const cpx* __restrict__ tab, int t) {
    if constexpr (PASS < PL::NP) {
      constexpr int R = PL::R(PASS), NS = PL::NS(PASS), SETS = PL::twSets(PASS);
#pragma unroll
      for (int s = 0; s < SETS; ++s) {
        int k = (t + s * PL::T) % NS;
#pragma unroll
        for (int q = 1; q < R; ++q)
          w[PL::twRegOffset(PASS) + s * (R - 1) + q - 1] = tab[PL::twOffset(PASS) + (q - 1) * NS + k];
      }
      init_pass<PASS + 1>(tab, t);
    }
  }
  __device__ __forceinline__ void init(const cpx* __restrict__ tab, int t) { init_pass<1>(tab, t); }
  template <int PASS>
  __device__ __forceinline__ cpx get(int b, int q, int /*k*/) const {
    constexpr int R = PL::R(PASS), SETS = PL::twSets(PASS);
    return w[PL::twRegOffset(PASS) + (SETS == 1 ? 0 : b) * (R - 1) + q - 1];
  }
  template <int PASS>
  __device__ __forceinline__ void fill(int b, int k, cpx* out) const {
#pragma unroll
    for (int q = 1; q < PL::R(PASS); ++q) out[q] = get<PASS>(b, q, k);
  }
  // v[q] *= W^(q k), q = 1 .. R-1
  template <int PASS>
  __device__ __forceinline__ void apply(int b, int k, cpx* v) const {
#pragma unroll
    for (int q = 1; q < PL::R(PASS); ++q) v[q] = cmul_<PK>(v[q], get<PASS>(b, q, k));
  }
};

// Twiddles read from a table (shared or global memory) on every use.
template <class PL>
struct TwTable {
  static constexpr bool PK = PL::PK;
  const cpx* tab;
  __device__ __forceinline__ void init(const cpx* t_, int) { tab = t_; }
  template <int PASS>
  __device__ __forceinline__ cpx get(int /*b*/, int q, int k) const {
    return tab[PL::twOffset(PASS) + (q - 1) * PL::NS(PASS) + k];
  }
  // w[q] = W^(q k), q = 1 .. R-1
  template <int PASS>
  __device__ __forceinline__ void fill(int b, int k, cpx* w) const {
#pragma unroll
    for (int q = 1; q < PL::R(PASS); ++q) w[q] = get<PASS>(b, q, k);
  }
  template <int PASS>
  __device__ __forceinline__ void apply(int b, int k, cpx* v) const {
#pragma unroll
    for (int q = 1; q < PL::R(PASS); ++q) v[q] = cmul_<PK>(v[q], get<PASS>(b, q, k));
  }
};

// Derived twiddles applied octet by octet: W^1 .. W^7 are formed once from W^1, W^2, W^4, then every
// further octet h = 8, 16, 24 takes its own factor W^h (a table row, or W^8 W^16) and the products
// W^h W^q as it goes.  The same multiplies as forming all R - 1 twiddles first, but only ~10 complex values
// are live at a time instead of R - 1: what lets the 32-point-per-lane plans fit their register budget.
// `row(r)` = the table row of W^(2^r).
template <int R, bool PK, class ROW>
__device__ __forceinline__ void apply_derived(const cpx* base, const ROW& row, cpx* v) {
  if constexpr (R == 2) {
    v[1] = cmul_<PK>(v[1], base[row(0)]);
  } else {
    cpx w[8];
    w[1] = base[row(0)];
    w[2] = base[row(1)];
    w[3] = cmul_<PK>(w[1], w[2]);
    v[1] = cmul_<PK>(v[1], w[1]);
    v[2] = cmul_<PK>(v[2], w[2]);
    v[3] = cmul_<PK>(v[3], w[3]);
    if constexpr (R >= 8) {
      w[4] = base[row(2)];
      v[4] = cmul_<PK>(v[4], w[4]);
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        w[4 + q] = cmul_<PK>(w[4], w[q]);
        v[4 + q] = cmul_<PK>(v[4 + q], w[4 + q]);
      }
    }
    if constexpr (R >= 16) {
      const cpx w8 = base[row(3)];
      v[8] = cmul_<PK>(v[8], w8);
#pragma unroll
      for (int q = 1; q < 8; ++q) v[8 + q] = cmul_<PK>(v[8 + q], cmul_<PK>(w8, w[q]));
      if constexpr (R >= 32) {
        const cpx w16 = base[row(4)];
        const cpx w24 = cmul_<PK>(w8, w16);
        v[16] = cmul_<PK>(v[16], w16);
        v[24] = cmul_<PK>(v[24], w24);
#pragma unroll
        for (int q = 1; q < 8; ++q) {
          v[16 + q] = cmul_<PK>(v[16 + q], cmul_<PK>(w16, w[q]));
          v[24 + q] = cmul_<PK>(v[24 + q], cmul_<PK>(w24, w[q]));
        }
        if constexpr (R >= 64) {
          const cpx w32 = base[row(5)];
#pragma unroll
          for (int h = 0; h < 4; ++h) {  // octets 32, 40, 48, 56: W^32 times 1, W^8, W^16, W^24
            const cpx wh = h == 0 ? w32 : cmul_<PK>(w32, h == 1 ? w8 : h == 2 ? w16 : w24);
            v[32 + 8 * h] = cmul_<PK>(v[32 + 8 * h], wh);
#pragma unroll
            for (int q = 1; q < 8; ++q) v[32 + 8 * h + q] = cmul_<PK>(v[32 + 8 * h + q], cmul_<PK>(wh, w[q]));
          }
        }
      }
    }
  }
}

// Twiddles derived from the table's power-of-two entries: W^k, W^2k, W^4k (and W^8k for radix
// 16) are loaded, the other powers are their products -- 3-4 shared-memory loads per butterfly
// instead of 7-15, paid for with complex multiplies on the (under-used) FMA pipe.  The shared-
// memory pipe is what bounds the large plans (ncu: LSU wavefronts ~80 % of peak).
template <class PL>
struct TwDerive {
  static constexpr bool PK = PL::PK;
  const cpx* tab;
  __device__ __forceinline__ void init(const cpx* t_, int) { tab = t_; }
  template <int PASS>
  __device__ __forceinline__ void fill(int /*b*/, int k, cpx* w) const {
    constexpr int R = PL::R(PASS), NS = PL::NS(PASS);
    const cpx* base = tab + PL::twOffset(PASS) + k;
    w[1] = base[0];
    if constexpr (R >= 4) {
      w[2] = base[1 * NS];
      w[3] = cmul_<PK>(w[1], w[2]);
    }
    if constexpr (R >= 8) {
      w[4] = base[3 * NS];
      w[5] = cmul_<PK>(w[4], w[1]);
      w[6] = cmul_<PK>(w[4], w[2]);
      w[7] = cmul_<PK>(w[4], w[3]);
    }
    if constexpr (R >= 16) {
      w[8] = base[7 * NS];
#pragma unroll
      for (int q = 1; q < 8; ++q) w[8 + q] = cmul_<PK>(w[8], w[q]);
    }
    if constexpr (R >= 32) {
      w[16] = base[15 * NS];
#pragma unroll
      for (int q = 1; q < 16; ++q) w[16 + q] = cmul_<PK>(w[16], w[q]);
    }
  }
  template <int PASS>
  __device__ __forceinline__ void apply(int /*b*/, int k, cpx* v) const {
    constexpr int NS = PL::NS(PASS);
    apply_derived<PL::R(PASS), PK>(tab + PL::twOffset(PASS) + k, [](int r) { return ((1 << r) - 1) * NS; }, v);
  }
};

// TwDerive over the compact table layout [pass][r = log2 q][k] (Plan::twcOffset): a quarter of
// the full table's shared memory for radix 16.
template <class PL>
struct TwDeriveC {
  static constexpr bool PK = PL::PK;
  const cpx* tab;
  __device__ __forceinline__ void init(const cpx* t_, int) { tab = t_; }
  template <int PASS>
  __device__ __forceinline__ void fill(int /*b*/, int k, cpx* w) const {
    constexpr int R = PL::R(PASS), NS = PL::NS(PASS);
    const cpx* base = tab + PL::twcOffset(PASS) + k;
    w[1] = base[0];
    if constexpr (R >= 4) {
      w[2] = base[1 * NS];
      w[3] = cmul_<PK>(w[1], w[2]);
    }
    if constexpr (R >= 8) {
      w[4] = base[2 * NS];
      w[5] = cmul_<PK>(w[4], w[1]);
      w[6] = cmul_<PK>(w[4], w[2]);
      w[7] = cmul_<PK>(w[4], w[3]);
    }
    if constexpr (R >= 16) {
      w[8] = base[3 * NS];
#pragma unroll
      for (int q = 1; q < 8; ++q) w[8 + q] = cmul_<PK>(w[8], w[q]);
    }
  }
  template <int PASS>
  __device__ __forceinline__ void apply(int /*b*/, int k, cpx* v) const {
    constexpr int NS = PL::NS(PASS);
    apply_derived<PL::R(PASS), PK>(tab + PL::twcOffset(PASS) + k, [](int r) { return r * NS; }, v);
  }
};

// host: fills the compact table, tw[twcOffset(p) + r*NS + k] = exp(-2 pi i 2^r k / (NS R))
template <class PL>
inline void build_compact_twiddles(cpx* tw) {
  for (int p = 1; p < PL::NP; ++p) {
    const int R = PL::R(p), NS = PL::NS(p);
    for (int r = 0; (1 << r) < R; ++r)
      for (int k = 0; k < NS; ++k) {
        const double ang = -2.0 * 3.14159265358979323846 * double(1 << r) * double(k) / double(NS * R);
        tw[PL::twcOffset(p) + r * NS + k] = make_float2((float)cos(ang), (float)sin(ang));
      }
  }
}

struct SyncBlock {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// named barrier over `count` threads (count % 32 == 0), id 1..15
struct SyncNamed {
  int id, count;
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
};
struct SyncWarp {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// ---------------------------------------------------------------------------------------
// One pass: (load from the previous exchange) -> twiddle -> butterflies -> (store to this
// pass's exchange).  The first pass expects v preloaded by the caller with
// v[b*R0 + q] = x[t + b*T + q*N/R0]; the last pass leaves X[t + b*T + q*N/R] in
// v[b*R + bitrev(q)].
// ---------------------------------------------------------------------------------------
// SINGLE: buf0 == buf1 (one exchange buffer): a barrier separates a pass's reads from its
// writes, at the price of one more sync per middle pass.
// PAIRED: the LAST pass assigns butterflies to threads in conjugate pairs (paired_bfly): slot
// b < B/2 takes butterfly j = t + b*T and slot b + B/2 the butterfly J - j that holds X[N - k] for
// every X[k] of the first, so a real-FFT split pass (or any k <-> N-k combination) needs no
// further exchange -- one shared-memory round trip less per transform.
template <class PL>
__device__ __forceinline__ int paired_bfly(int t, int b) {
  constexpr int RL = PL::R(PL::NP - 1), J = PL::N / RL, B = PL::P / RL;
  static_assert(B >= 2 && B % 2 == 0, "PAIRED needs an even number of last-pass butterflies per thread");
  if (b < B / 2) return t + b * PL::T;
  const int j = t + (b - B / 2) * PL::T;
  return j == 0 ? J / 2 : J - j;  // butterflies 0 and J/2 are their own partners: thread 0 takes both
}

// HOOK: called once, right after the LAST pass has read its inputs out of the exchange buffer (the buffer is
// then free as far as this thread is concerned: kernels whose exchange buffer doubles as their TMA stage
// re-arm the stage from there, behind a group barrier of their own).
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

template <class PL, int PASS, class TW, class SYNC, bool SINGLE = false, bool PAIRED = false, class HOOK = NoHook>
struct PassRunner {
  static __device__ __forceinline__ void run(cpx (&v)[PL::P], int t, cpx* buf0, cpx* buf1,
                                             const TW& tw, const SYNC& sync, const HOOK& hook = HOOK()) {
    constexpr int N = PL::N, T = PL::T, P = PL::P;
    constexpr int R = PL::R(PASS), NS = PL::NS(PASS), B = P / R;
    static_assert(P % R == 0, "radix must divide points per thread");
    if constexpr (PASS > 0) {
      constexpr int A = PL::padA(PASS - 1), C = PL::padC(PASS - 1);
      const cpx* in = ((PASS - 1) & 1) ? buf1 : buf0;
      constexpr int unit = (A < 0) ? 1 : (1 << (A < 0 ? 0 : A));
      // pad(t + X) splits into pad(t) + pad(X) when every X is a multiple of 2^A
      constexpr bool split = (A < 0) || (T % unit == 0 && (N / R) % unit == 0);
      constexpr bool LASTP = PAIRED && PASS + 1 == PL::NP;
      const int tb = pad_idx<A, C>(t);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        int jb = 0;
        if constexpr (LASTP) jb = paired_bfly<PL>(t, b);
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int X = b * T + q * (N / R);
          if constexpr (LASTP) v[b * R + q] = in[pad_idx<A, C>(jb + q * (N / R))];
          else if constexpr (split) v[b * R + q] = in[tb + pad_idx<A, C>(X)];
          else v[b * R + q] = in[pad_idx<A, C>(t + X)];
        }
      }
      if constexpr (SINGLE && PASS + 1 < PL::NP) sync();  // reads done before this pass overwrites the buffer
      if constexpr (PASS + 1 == PL::NP) hook();
#pragma unroll
      for (int b = 0; b < B; ++b) {
        int k = (t + b * T) % NS;
        if constexpr (LASTP) k = paired_bfly<PL>(t, b) % NS;
        tw.template apply<PASS>(b, k, &v[b * R]);
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) Dft<R, PL::PK>::run(&v[b * R]);
    if constexpr (PASS + 1 < PL::NP) {
      constexpr int A = PL::padA(PASS), C = PL::padC(PASS);
      constexpr int LR = ilog2(NS * R);
      cpx* out = (PASS & 1) ? buf1 : buf0;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int j = t + b * T;
        const int hi = j / NS, lo = j % NS;
        int base = hi * (NS * R) + lo;
        if constexpr (A >= 0) base += (A >= LR ? (hi >> (A - LR)) : (hi << (LR - A))) * C;
#pragma unroll
        for (int q = 0; q < R; ++q) out[base + q * NS] = v[b * R + bitrev(q, ilog2(R))];
      }
      sync();
      PassRunner<PL, PASS + 1, TW, SYNC, SINGLE, PAIRED, HOOK>::run(v, t, buf0, buf1, tw, sync, hook);
    }
  }
};

// Runs all passes.  buf0/buf1: two exchange buffers of PL::BUF complex each, private to
// the T-thread group.  The caller must order its own later use of buf0/buf1 against the
// reads of the last exchange (see the kernels).
template <class PL, class TW, class SYNC>
__device__ __forceinline__ void block_fft(cpx (&v)[PL::P], int t, cpx* buf0, cpx* buf1, const TW& tw,
                                          const SYNC& sync) {
  PassRunner<PL, 0, TW, SYNC>::run(v, t, buf0, buf1, tw, sync);
}
// one exchange buffer of PL::BUF complex (see PassRunner's SINGLE)
template <class PL, class TW, class SYNC, bool PAIRED = false>
__device__ __forceinline__ void block_fft_single(cpx (&v)[PL::P], int t, cpx* buf, const TW& tw, const SYNC& sync) {
  PassRunner<PL, 0, TW, SYNC, true, PAIRED>::run(v, t, buf, buf, tw, sync);
}
// one exchange buffer, with a hook behind the last pass's reads (see PassRunner's HOOK)
template <class PL, class TW, class SYNC, class HOOK>
__device__ __forceinline__ void block_fft_single_hook(cpx (&v)[PL::P], int t, cpx* buf, const TW& tw, const SYNC& sync,
                                                      const HOOK& hook) {
  PassRunner<PL, 0, TW, SYNC, true, false, HOOK>::run(v, t, buf, buf, tw, sync, hook);
}
// two exchange buffers, conjugate-paired last pass (see PassRunner's PAIRED)
template <class PL, class TW, class SYNC>
__device__ __forceinline__ void block_fft_paired(cpx (&v)[PL::P], int t, cpx* buf0, cpx* buf1, const TW& tw,
                                                 const SYNC& sync) {
  PassRunner<PL, 0, TW, SYNC, false, true>::run(v, t, buf0, buf1, tw, sync);
}

// logical index of the element the caller must preload into v[b*R0 + q]
template <class PL>
__device__ __forceinline__ int fft_in_index(int t, int b, int q) {
  return t + b * PL::T + q * (PL::N / PL::R(0));
}
// after block_fft: X[fft_out_index(t,b,q)] is at v[b*RL + bitrev(q)], RL = last radix
template <class PL>
__device__ __forceinline__ int fft_out_index(int t, int b, int q) {
  return t + b * PL::T + q * (PL::N / PL::R(PL::NP - 1));
}
template <class PL>
__device__ __forceinline__ constexpr int fft_out_reg(int b, int q) {
  return b * PL::R(PL::NP - 1) + bitrev(q, ilog2(PL::R(PL::NP - 1)));
}

}  // namespace nxs
