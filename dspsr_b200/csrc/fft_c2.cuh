// fft_c2.cuh -- compile-time-sized block FFT, TWO sequences per thread, affine shared-memory addressing.
//
// Second-generation core of the hot kernels (K1/K2/K3 fast paths).  Differences to fft_core.cuh's
// generic Stockham driver:
//   * every thread carries 16 points of TWO sequences (adjacent columns in K1, a row and its mirror
//     row in K2, the two polarisations of a channel in K3).  The pair is stored interleaved in
//     shared memory as one float4, so every exchange is a 128-bit LDS/STS (half the instructions of
//     64-bit exchanges) and stage twiddles are loaded once for both sequences;
//   * the bank-conflict remedy is PADDING (one 16-byte slot after every 16), not an XOR swizzle:
//     pad(i) = i + (i >> 4) is additive over the multiples of 16 that separate a thread's elements,
//     so every shared-memory address is  <one of three thread-constant bases> + <compile-time
//     constant>  and folds into the instruction's immediate -- no per-element integer arithmetic;
//   * all sizes are template parameters: radices 16 / 8 / 4 / 2, first radix always 16.
//
// Stockham autosort bookkeeping (same as fft_core.cuh): before every stage thread j of the T = L/16
// threads of a sequence pair holds v[e] = x[j + e*T]; a radix-R stage with sub-transform length Ns
// runs NB = 16/R butterflies b = j + q*T per thread, inputs v[q + r*NB], twiddles W_{Ns*R}^(r*k),
// k = b mod Ns, and output r' of butterfly b belongs at (b - k)*R + k + r'*Ns.  After the last
// stage register e holds natural-order element j + e*T.
//
// Everything is __host__ __device__: csrc/host_fft_emul.cu runs the same code thread by thread on
// the CPU (tests/test_host_logic.py).
#pragma once
#include "fft_core.cuh"

namespace b200 {
namespace c2 {

template <unsigned L> struct Plan;
#define B200_C2PLAN(LL, N, A, B, C, D)                    \
  template <> struct Plan<LL> {                           \
    static constexpr int nstage = N;                      \
    B200_HD static constexpr int radix(int s) { return s == 0 ? A : s == 1 ? B : s == 2 ? C : D; } \
  };
B200_C2PLAN(256, 2, 16, 16, 1, 1)
B200_C2PLAN(512, 3, 16, 8, 4, 1)
B200_C2PLAN(1024, 3, 16, 8, 8, 1)
B200_C2PLAN(2048, 3, 16, 16, 8, 1)
B200_C2PLAN(4096, 3, 16, 16, 16, 1)
B200_C2PLAN(8192, 4, 16, 16, 16, 2)
#undef B200_C2PLAN

template <unsigned L> B200_HD constexpr unsigned stage_ns(int s) {
  unsigned ns = 1;
  for (int i = 0; i < s; i++) ns *= (unsigned)Plan<L>::radix(i);
  return ns;
}
// offset (float2) of stage s's twiddle table: tw[off + (r-1)*Ns + k] = exp(-2 pi i r k / (Ns*R))
template <unsigned L> B200_HD constexpr unsigned stage_twoff(int s) {
  unsigned off = 0;
  for (int i = 1; i < s; i++) off += (unsigned)(Plan<L>::radix(i) - 1) * stage_ns<L>(i);
  return off;
}
template <unsigned L> B200_HD constexpr unsigned twiddle_count() { return stage_twoff<L>(Plan<L>::nstage) + 1; }
// float4 slots one sequence pair occupies in shared memory
template <unsigned L> B200_HD constexpr unsigned pair_slots() { return L + L / 16; }

B200_HD constexpr unsigned pad16(unsigned i) { return i + (i >> 4); }

// ---- butterflies of stage S on both sequences ------------------------------------------------
template <unsigned L, int S, bool INV>
B200_HD void stage_compute(float2* va, float2* vb, unsigned j, const float2* __restrict__ tw) {
  constexpr int R = Plan<L>::radix(S);
  constexpr unsigned Ns = stage_ns<L>(S);
  constexpr int NB = 16 / R;
  constexpr unsigned T = L / 16;
  constexpr unsigned OFF = stage_twoff<L>(S);
#pragma unroll
  for (int q = 0; q < NB; q++) {
    float2 ua[R], ub[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      ua[r] = va[q + r * NB];
      ub[r] = vb[q + r * NB];
    }
    if (Ns > 1) {
      const unsigned k = (Ns <= T) ? (j & (Ns - 1)) : (j + ((unsigned)q * T) % Ns);
      const float2* t = tw + OFF + k;
#pragma unroll
      for (int r = 1; r < R; r++) {
        const float2 w = tw_get<INV>(t, (unsigned)(r - 1) * Ns);
        ua[r] = cmul(ua[r], w);
        ub[r] = cmul(ub[r], w);
      }
    }
    dftR<R, INV>(ua);
    dftR<R, INV>(ub);
#pragma unroll
    for (int r = 0; r < R; r++) {
      va[q + r * NB] = ua[r];
      vb[q + r * NB] = ub[r];
    }
  }
}

// ---- exchange: scatter the outputs of stage S, then gather the inputs of stage S+1 --------------
template <unsigned L, int S>
B200_HD void scatter(float4* smem, const float2* va, const float2* vb, unsigned j) {
  constexpr int R = Plan<L>::radix(S);
  constexpr unsigned Ns = stage_ns<L>(S);
  constexpr int NB = 16 / R;
  constexpr unsigned T = L / 16;
  static_assert(T % 16 == 0, "threads per sequence pair must be a multiple of 16");
  static_assert(Ns > 1 || R == 16, "the first radix is 16");
  unsigned base;
  if (Ns == 1) base = 17u * j;
  else if (Ns <= T) {
    const unsigned k = j & (Ns - 1);
    base = ((j - k) * (unsigned)R / 16u) * 17u + k + (k >> 4);
  } else base = j + (j >> 4);
#pragma unroll
  for (int q = 0; q < NB; q++)
#pragma unroll
    for (int r = 0; r < R; r++) {
      unsigned off;
      if (Ns == 1) off = (unsigned)r;
      else if (Ns <= T) off = (unsigned)q * (T * (unsigned)R / 16u * 17u) + (unsigned)r * (Ns / 16u * 17u);
      else {
        const unsigned cq = ((unsigned)q * T) % Ns;
        const unsigned c = ((unsigned)q * T - cq) * (unsigned)R + cq + (unsigned)r * Ns;
        off = c + c / 16u;
      }
      const int e = q + r * NB;
      smem[base + off] = make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y);
    }
}

template <unsigned L>
B200_HD void gather(const float4* smem, float2* va, float2* vb, unsigned j) {
  constexpr unsigned T = L / 16;
  const unsigned base = j + (j >> 4);
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const float4 x = smem[base + (unsigned)e * (T / 16u * 17u)];
    va[e] = make_float2(x.x, x.y);
    vb[e] = make_float2(x.z, x.w);
  }
}

// natural-order store of the finished transform into the padded layout (slot pad16(j + e*T))
template <unsigned L>
B200_HD void store_natural(float4* smem, const float2* va, const float2* vb, unsigned j) {
  constexpr unsigned T = L / 16;
  const unsigned base = j + (j >> 4);
#pragma unroll
  for (int e = 0; e < 16; e++)
    smem[base + (unsigned)e * (T / 16u * 17u)] = make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y);
}

#ifdef __CUDACC__
// Whole transform of the pair held in (va, vb); `sync()` must synchronise the threads that share
// `smem` (one sequence pair or the whole CTA).  On return register e holds element j + e*T.
template <unsigned L, bool INV, int S = 0, typename SyncF>
__device__ __forceinline__ void fft_pair(float2* va, float2* vb, unsigned j, float4* smem,
                                         const float2* __restrict__ tw, SyncF sync) {
  stage_compute<L, S, INV>(va, vb, j, tw);
  if constexpr (S + 1 < Plan<L>::nstage) {
    if (S > 0) sync();               // every thread has gathered its inputs of this stage
    scatter<L, S>(smem, va, vb, j);
    sync();
    gather<L>(smem, va, vb, j);
    fft_pair<L, INV, S + 1>(va, vb, j, smem, tw, sync);
  }
}
#endif

// host-side table builder: fills tw[twiddle_count<L>()] (forward sign)
template <unsigned L> inline void fill_twiddles(float2* tw) {
  for (int s = 1; s < Plan<L>::nstage; s++) {
    const int R = Plan<L>::radix(s);
    const unsigned Ns = stage_ns<L>(s), off = stage_twoff<L>(s);
    for (int r = 1; r < R; r++)
      for (unsigned k = 0; k < Ns; k++) {
        const double a = -2.0 * 3.14159265358979323846 * double(r) * double(k) / (double(Ns) * R);
        tw[off + (unsigned)(r - 1) * Ns + k] = make_float2(float(cos(a)), float(sin(a)));
      }
  }
}

}  // namespace c2
}  // namespace b200
