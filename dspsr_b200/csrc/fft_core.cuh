// fft_core.cuh -- register-resident radix-2..32 butterflies and the Stockham stage
// bookkeeping shared by every FFT kernel of the library (sm_100a).
//
// Model: an N-point FFT is computed by T = N/EPT threads, each holding EPT complex points
// in registers.  A stage of radix R (R <= EPT) makes every thread perform EPT/R butterflies;
// between stages the points are exchanged through shared memory.  At the start of every
// stage thread j holds   v[e] = x[j + e*T],  e = 0..EPT-1      (stage-independent!)
// and after the butterflies element (q, r) (register q + r*EPT/R) belongs at
//   x'[ (b - k)*R + k + r*Ns ],  b = j + q*T,  k = b mod Ns     (Stockham autosort)
// where Ns is the product of the radices of the previous stages.
//
// All functions are __host__ __device__ so the index algebra is unit-tested on the CPU
// (tests/test_fft_core_host.py builds csrc/host_fft_emul.cu with nvcc and runs it here).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef B200_HD
#define B200_HD __host__ __device__ __forceinline__
#endif

namespace b200 {

// Complex arithmetic.  On the device it is written with Blackwell's packed FP32x2 instructions
// (PTX add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): a complex add is ONE issue slot and a
// complex multiply TWO (ptxas folds the lane swap, broadcast and sign into operand modifiers),
// half of what scalar FADD/FMUL/FFMA need.  The host versions are used by the CPU emulation test.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 up2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up2(add2(pk2(a.x, a.y), pk2(b.x, b.y))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up2(add2(pk2(a.x, a.y), pk2(-b.x, -b.y))); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  // (a.x b.x - a.y b.y, a.y b.x + a.x b.y)
  unsigned long long t = mul2(pk2(a.x, a.y), pk2(b.x, b.x));
  return up2(fma2(pk2(a.y, a.x), pk2(-b.y, b.y), t));
}
#else
B200_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B200_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B200_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
#endif
B200_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward transforms) or +i (inverse transforms)
template <bool INV> B200_HD float2 crot(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * (c - i s) forward,  a * (c + i s) inverse  (c, s compile-time constants)
template <bool INV> B200_HD float2 cmulc(float2 a, float c, float s) {
#ifdef __CUDA_ARCH__
  unsigned long long t = mul2(pk2(a.x, a.y), pk2(c, c));
  return INV ? up2(fma2(pk2(a.y, a.x), pk2(-s, s), t)) : up2(fma2(pk2(a.y, a.x), pk2(s, -s), t));
#else
  return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s)
             : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
#endif
}

#define B200_SQRT1_2 0.70710678118654752440f
#define B200_COS_PI_8 0.92387953251128675613f
#define B200_SIN_PI_8 0.38268343236508977173f
#define B200_COS_PI_16 0.98078528040323044913f
#define B200_SIN_PI_16 0.19509032201612826785f
#define B200_COS_3PI_16 0.83146961230254523708f
#define B200_SIN_3PI_16 0.55557023301960222474f

// ---- natural-order in, natural-order out DFTs on registers --------------------------------
template <bool INV> B200_HD void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV> B200_HD void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  float2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = crot<INV>(csub(v1, v3));
  v0 = cadd(a, c);
  v1 = cadd(b, d);
  v2 = csub(a, c);
  v3 = csub(b, d);
}

template <bool INV> B200_HD void dft8(float2* v /* stride 1, 8 entries */) {
  // even / odd 4-point transforms, then combine with W8^k
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4<INV>(e0, e1, e2, e3);
  dft4<INV>(o0, o1, o2, o3);
  o1 = cmulc<INV>(o1, B200_SQRT1_2, B200_SQRT1_2);
  o2 = crot<INV>(o2);
  o3 = cmulc<INV>(o3, -B200_SQRT1_2, B200_SQRT1_2);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <bool INV> B200_HD void dft16(float2* v) {
  // 4x4: n = 4*n1 + n2, k = k1 + 4*k2
  float2 y[4][4];   // y[n2][k1]
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    float2 a = v[n2], b = v[4 + n2], c = v[8 + n2], d = v[12 + n2];
    dft4<INV>(a, b, c, d);
    y[n2][0] = a; y[n2][1] = b; y[n2][2] = c; y[n2][3] = d;
  }
  // twiddle y[n2][k1] *= W16^(n2*k1)
  y[1][1] = cmulc<INV>(y[1][1], B200_COS_PI_8, B200_SIN_PI_8);
  y[1][2] = cmulc<INV>(y[1][2], B200_SQRT1_2, B200_SQRT1_2);
  y[1][3] = cmulc<INV>(y[1][3], B200_SIN_PI_8, B200_COS_PI_8);
  y[2][1] = cmulc<INV>(y[2][1], B200_SQRT1_2, B200_SQRT1_2);
  y[2][2] = crot<INV>(y[2][2]);
  y[2][3] = cmulc<INV>(y[2][3], -B200_SQRT1_2, B200_SQRT1_2);
  y[3][1] = cmulc<INV>(y[3][1], B200_SIN_PI_8, B200_COS_PI_8);
  y[3][2] = cmulc<INV>(y[3][2], -B200_SQRT1_2, B200_SQRT1_2);
  y[3][3] = cmulc<INV>(y[3][3], -B200_COS_PI_8, -B200_SIN_PI_8);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    float2 a = y[0][k1], b = y[1][k1], c = y[2][k1], d = y[3][k1];
    dft4<INV>(a, b, c, d);
    v[k1] = a; v[k1 + 4] = b; v[k1 + 8] = c; v[k1 + 12] = d;
  }
}

template <bool INV> B200_HD void dft32(float2* v) {
  float2 e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
  dft16<INV>(e);
  dft16<INV>(o);
  // W32^k, k = 0..15: angle k*pi/16
  const float c[16] = {1.f, B200_COS_PI_16, B200_COS_PI_8, B200_COS_3PI_16, B200_SQRT1_2, B200_SIN_3PI_16,
                       B200_SIN_PI_8, B200_SIN_PI_16, 0.f, -B200_SIN_PI_16, -B200_SIN_PI_8, -B200_SIN_3PI_16,
                       -B200_SQRT1_2, -B200_COS_3PI_16, -B200_COS_PI_8, -B200_COS_PI_16};
  const float s[16] = {0.f, B200_SIN_PI_16, B200_SIN_PI_8, B200_SIN_3PI_16, B200_SQRT1_2, B200_COS_3PI_16,
                       B200_COS_PI_8, B200_COS_PI_16, 1.f, B200_COS_PI_16, B200_COS_PI_8, B200_COS_3PI_16,
                       B200_SQRT1_2, B200_SIN_3PI_16, B200_SIN_PI_8, B200_SIN_PI_16};
#pragma unroll
  for (int k = 0; k < 16; k++) {
    float2 t = (k == 0) ? o[0] : (k == 8) ? crot<INV>(o[8]) : cmulc<INV>(o[k], c[k], s[k]);
    v[k] = cadd(e[k], t);
    v[k + 16] = csub(e[k], t);
  }
}

template <int R, bool INV> B200_HD void dftR(float2* u) {
  if (R == 2) dft2<INV>(u[0], u[1]);
  else if (R == 4) dft4<INV>(u[0], u[1], u[2], u[3]);
  else if (R == 8) dft8<INV>(u);
  else if (R == 16) dft16<INV>(u);
  else if (R == 32) dft32<INV>(u);
}

// ---- twiddle table access ------------------------------------------------------------------
// tw[m] = exp(-2 pi i m / NT), m < NT (forward sign); inverse transforms conjugate.
template <bool INV> B200_HD float2 tw_get(const float2* __restrict__ tw, unsigned m) {
#ifdef __CUDA_ARCH__
  float2 w = __ldg(tw + m);
#else
  float2 w = tw[m];
#endif
  return INV ? cconj(w) : w;
}

// Apply W^{r*k0}, r = 1..R-1, to u[r] where W^{m} = tw[m*?]; base index kk = k * stride.
// Loads the power-of-two powers and derives the rest with <= 3 chained multiplies.
template <int R, bool INV>
B200_HD void apply_stage_twiddles(float2* u, const float2* __restrict__ tw, unsigned kk) {
  if (R == 2) {
    u[1] = cmul(u[1], tw_get<INV>(tw, kk));
  } else if (R == 4) {
    float2 w1 = tw_get<INV>(tw, kk), w2 = tw_get<INV>(tw, 2 * kk);
    u[1] = cmul(u[1], w1);
    u[2] = cmul(u[2], w2);
    u[3] = cmul(u[3], cmul(w1, w2));
  } else if (R == 8) {
    float2 w1 = tw_get<INV>(tw, kk), w2 = tw_get<INV>(tw, 2 * kk), w4 = tw_get<INV>(tw, 4 * kk);
    float2 w3 = cmul(w1, w2);
    u[1] = cmul(u[1], w1);
    u[2] = cmul(u[2], w2);
    u[3] = cmul(u[3], w3);
    u[4] = cmul(u[4], w4);
    u[5] = cmul(u[5], cmul(w1, w4));
    u[6] = cmul(u[6], cmul(w2, w4));
    u[7] = cmul(u[7], cmul(w3, w4));
  } else {
    // R = 16 or 32
    float2 w[R];
    w[1] = tw_get<INV>(tw, kk);
    w[2] = tw_get<INV>(tw, 2 * kk);
    w[4] = tw_get<INV>(tw, 4 * kk);
    w[8] = tw_get<INV>(tw, 8 * kk);
    if (R == 32) w[16] = tw_get<INV>(tw, 16 * kk);
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[1], w[4]);
    w[6] = cmul(w[2], w[4]);
    w[7] = cmul(w[3], w[4]);
#pragma unroll
    for (int r = 9; r < 16; r++) w[r] = cmul(w[r - 8], w[8]);
    if (R == 32) {
#pragma unroll
      for (int r = 17; r < 32; r++) w[r] = cmul(w[r - 16], w[16]);
    }
#pragma unroll
    for (int r = 1; r < R; r++) u[r] = cmul(u[r], w[r]);
  }
}

// ---- one Stockham stage on the registers of thread j ---------------------------------------
// v[e] = x[j + e*T] on entry.  On exit v[q + r*(EPT/R)] is element r of butterfly b = j + q*T.
// tw is a table for size NT with NT a multiple of Ns*R (index stride NT/(Ns*R)).
template <int EPT, int R, bool INV>
B200_HD void stage_compute(float2* v, unsigned j, unsigned T, unsigned Ns, const float2* __restrict__ tw,
                           unsigned NT) {
  constexpr int NB = EPT / R;
  const unsigned stride = NT / (Ns * R);
#pragma unroll
  for (int q = 0; q < NB; q++) {
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; r++) u[r] = v[q + r * NB];
    if (Ns > 1) {
      unsigned k = (j + q * T) & (Ns - 1);
      apply_stage_twiddles<R, INV>(u, tw, k * stride);
    }
    dftR<R, INV>(u);
#pragma unroll
    for (int r = 0; r < R; r++) v[q + r * NB] = u[r];
  }
}

// destination index of element (q, r) after a radix-R stage with sub-transform length Ns
template <int EPT, int R>
B200_HD unsigned stage_dest(unsigned j, unsigned T, unsigned Ns, int q, int r) {
  unsigned b = j + q * T;
  unsigned k = b & (Ns - 1);
  return (b - k) * R + k + r * Ns;
}

// Radix plan for an N-point transform with EPT points per thread:
// EPT-radix stages while they fit, the remainder (2..EPT/2) last.
struct RadixPlan {
  int nstage;
  int radix[8];
};
B200_HD RadixPlan make_radix_plan(unsigned N, int EPT) {
  RadixPlan p;
  p.nstage = 0;
  unsigned rem = N;
  while (rem >= (unsigned)EPT) {
    p.radix[p.nstage++] = EPT;
    rem /= EPT;
  }
  if (rem > 1) p.radix[p.nstage++] = (int)rem;
  return p;
}

// ---- compile-time-sized variant ------------------------------------------------------------
// multiply by W32^m (forward) / W32^-m (inverse); m is a compile-time constant after unrolling,
// so the switch folds to one constant multiply (or a free rotation for m = 0, 8, 16, 24).
template <bool INV> B200_HD float2 mul_w32(float2 a, int m) {
  switch (m & 31) {
    case 0: return a;
    case 8: return crot<INV>(a);
    case 16: return make_float2(-a.x, -a.y);
    case 24: return crot<!INV>(a);
    case 1: return cmulc<INV>(a, B200_COS_PI_16, B200_SIN_PI_16);
    case 2: return cmulc<INV>(a, B200_COS_PI_8, B200_SIN_PI_8);
    case 3: return cmulc<INV>(a, B200_COS_3PI_16, B200_SIN_3PI_16);
    case 4: return cmulc<INV>(a, B200_SQRT1_2, B200_SQRT1_2);
    case 5: return cmulc<INV>(a, B200_SIN_3PI_16, B200_COS_3PI_16);
    case 6: return cmulc<INV>(a, B200_SIN_PI_8, B200_COS_PI_8);
    case 7: return cmulc<INV>(a, B200_SIN_PI_16, B200_COS_PI_16);
    case 9: return cmulc<INV>(a, -B200_SIN_PI_16, B200_COS_PI_16);
    case 10: return cmulc<INV>(a, -B200_SIN_PI_8, B200_COS_PI_8);
    case 11: return cmulc<INV>(a, -B200_SIN_3PI_16, B200_COS_3PI_16);
    case 12: return cmulc<INV>(a, -B200_SQRT1_2, B200_SQRT1_2);
    case 13: return cmulc<INV>(a, -B200_COS_3PI_16, B200_SIN_3PI_16);
    case 14: return cmulc<INV>(a, -B200_COS_PI_8, B200_SIN_PI_8);
    case 15: return cmulc<INV>(a, -B200_COS_PI_16, B200_SIN_PI_16);
    case 17: return cmulc<INV>(a, -B200_COS_PI_16, -B200_SIN_PI_16);
    case 18: return cmulc<INV>(a, -B200_COS_PI_8, -B200_SIN_PI_8);
    case 19: return cmulc<INV>(a, -B200_COS_3PI_16, -B200_SIN_3PI_16);
    case 20: return cmulc<INV>(a, -B200_SQRT1_2, -B200_SQRT1_2);
    case 21: return cmulc<INV>(a, -B200_SIN_3PI_16, -B200_COS_3PI_16);
    case 22: return cmulc<INV>(a, -B200_SIN_PI_8, -B200_COS_PI_8);
    case 23: return cmulc<INV>(a, -B200_SIN_PI_16, -B200_COS_PI_16);
    case 25: return cmulc<INV>(a, B200_SIN_PI_16, -B200_COS_PI_16);
    case 26: return cmulc<INV>(a, B200_SIN_PI_8, -B200_COS_PI_8);
    case 27: return cmulc<INV>(a, B200_SIN_3PI_16, -B200_COS_3PI_16);
    case 28: return cmulc<INV>(a, B200_SQRT1_2, -B200_SQRT1_2);
    case 29: return cmulc<INV>(a, B200_COS_3PI_16, -B200_SIN_3PI_16);
    case 30: return cmulc<INV>(a, B200_COS_PI_8, -B200_SIN_PI_8);
    default: return cmulc<INV>(a, B200_COS_PI_16, -B200_SIN_PI_16);   // 31
  }
}

// ---- compact per-stage twiddle tables (compile-time-sized path) -----------------------------
// For the stage with sub-transform length NS and radix R the butterflies of thread k need
// W_{NS*R}^(m*k) for a handful of multipliers m; they are stored structure-of-arrays,
//   stab[mi*NS + k] = exp(-2 pi i m_i k / (NS*R)),
// so that consecutive lanes (consecutive k) read consecutive 8-byte words: one or two L1
// wavefronts per load instead of the 8-32 of a strided gather from one big table.
//   R = 32: m = {8, 16, 1, 2, 4}   R = 16: m = {4, 8, 1, 2}   R = 8: {1, 2, 4}   R = 4: {1, 2}   R = 2: {1}
B200_HD constexpr int stage_radix(unsigned N, int EPT, unsigned NS) {
  return (N / NS >= (unsigned)EPT) ? EPT : (int)(N / NS);
}
B200_HD constexpr int stage_nmult(int R) { return R == 32 ? 5 : R == 16 ? 4 : R == 8 ? 3 : R == 4 ? 2 : 1; }
B200_HD constexpr int stage_mult(int R, int mi) {
  return R >= 16 ? (mi == 0 ? R / 4 : mi == 1 ? R / 2 : (1 << (mi - 2))) : (1 << mi);
}
// offset (in float2) of the table of the stage whose sub-transform length is NS (NS > 1)
B200_HD constexpr unsigned stage_offset(unsigned N, int EPT, unsigned NS) {
  unsigned off = 0, ns = 1;
  while (ns < NS) {
    int R = stage_radix(N, EPT, ns);
    if (ns > 1) off += (unsigned)stage_nmult(R) * ns;
    ns *= (unsigned)R;
  }
  return off;
}
B200_HD constexpr unsigned stage_table_size(unsigned N, int EPT) {
  unsigned off = 0, ns = 1;
  while (ns < N) {
    int R = stage_radix(N, EPT, ns);
    if (ns > 1) off += (unsigned)stage_nmult(R) * ns;
    ns *= (unsigned)R;
  }
  return off ? off : 1;
}

// Radix-R butterfly with the stage twiddles W^(r*k) folded into a 4 x (R/4) decomposition so
// that only a handful of runtime twiddles are live:  r = A*n1 + n2, A = R/4
//   R = 16: 4 x 4,  runtime twiddles w1,w2,w3 (second axis) and w4,w8,w12 (first axis)
//   R = 32: 4 x 8,  runtime twiddles w1..w7 and w8,w16,w24
// TW = false skips the runtime twiddles (first stage, NS = 1).  st points at this stage's compact
// table, k is the butterfly's position inside its sub-transform.
template <int R, bool INV, bool TW>
B200_HD void dft_tw(float2* u, const float2* __restrict__ st, unsigned k, unsigned NS) {
  if constexpr (R <= 8) {
    if (TW) {
      if (R == 2) {
        u[1] = cmul(u[1], tw_get<INV>(st, k));
      } else if (R == 4) {
        float2 w1 = tw_get<INV>(st, k), w2 = tw_get<INV>(st, NS + k);
        u[1] = cmul(u[1], w1);
        u[2] = cmul(u[2], w2);
        u[3] = cmul(u[3], cmul(w1, w2));
      } else {
        float2 w1 = tw_get<INV>(st, k), w2 = tw_get<INV>(st, NS + k), w4 = tw_get<INV>(st, 2 * NS + k);
        float2 w3 = cmul(w1, w2);
        u[1] = cmul(u[1], w1);
        u[2] = cmul(u[2], w2);
        u[3] = cmul(u[3], w3);
        u[4] = cmul(u[4], w4);
        u[5] = cmul(u[5], cmul(w1, w4));
        u[6] = cmul(u[6], cmul(w2, w4));
        u[7] = cmul(u[7], cmul(w3, w4));
      }
    }
    dftR<R, INV>(u);
  } else {
    constexpr int A = R / 4;        // second-axis length (4 or 8); first axis has 4 points
    float2 wa1, wa2, wa3;
    if (TW) {
      wa1 = tw_get<INV>(st, k);
      wa2 = tw_get<INV>(st, NS + k);
      wa3 = cmul(wa1, wa2);
    }
#pragma unroll
    for (int n2 = 0; n2 < A; n2++) {
      float2 a = u[n2], b = u[A + n2], c = u[2 * A + n2], d = u[3 * A + n2];
      if (TW) { b = cmul(b, wa1); c = cmul(c, wa2); d = cmul(d, wa3); }
      dft4<INV>(a, b, c, d);
      u[n2] = a; u[A + n2] = b; u[2 * A + n2] = c; u[3 * A + n2] = d;   // y[n2][k1] at u[k1*A + n2]
    }
    float2 w[A];
    if (TW) {
      w[1] = tw_get<INV>(st, 2 * NS + k);
      w[2] = tw_get<INV>(st, 3 * NS + k);
      w[3] = cmul(w[1], w[2]);
      if (A == 8) {
        w[4] = tw_get<INV>(st, 4 * NS + k);
        w[5] = cmul(w[1], w[4]);
        w[6] = cmul(w[2], w[4]);
        w[7] = cmul(w[3], w[4]);
      }
    }
#pragma unroll
    for (int n2 = 1; n2 < A; n2++) {
#pragma unroll
      for (int k1 = 0; k1 < 4; k1++) {
        float2 y = u[k1 * A + n2];
        if (TW) y = cmul(y, w[n2]);
        y = mul_w32<INV>(y, (32 / R) * n2 * k1);
        u[k1 * A + n2] = y;
      }
    }
    float2 o[R];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      float2 t[A];
#pragma unroll
      for (int n2 = 0; n2 < A; n2++) t[n2] = u[k1 * A + n2];
      dftR<A, INV>(t);
#pragma unroll
      for (int k2 = 0; k2 < A; k2++) o[k1 + 4 * k2] = t[k2];
    }
#pragma unroll
    for (int r = 0; r < R; r++) u[r] = o[r];
  }
}

// One Stockham stage with every size a compile-time constant (N points, EPT per thread,
// sub-transform length NS): same contract as stage_compute.  stw = compact stage tables of (N, EPT).
template <int EPT, int R, bool INV, unsigned N, unsigned NS>
B200_HD void stage_compute_ct(float2* v, unsigned j, const float2* __restrict__ stw) {
  constexpr int NB = EPT / R;
  constexpr unsigned T = N / EPT;
  constexpr unsigned OFF = stage_offset(N, EPT, NS);
#pragma unroll
  for (int q = 0; q < NB; q++) {
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; r++) u[r] = v[q + r * NB];
    if (NS > 1) {
      const unsigned k = (j + q * T) & (NS - 1);
      dft_tw<R, INV, true>(u, stw + OFF, k, NS);
    } else {
      dft_tw<R, INV, false>(u, stw, 0, 1);
    }
#pragma unroll
    for (int r = 0; r < R; r++) v[q + r * NB] = u[r];
  }
}

// ---- shared-memory index maps ----------------------------------------------------------------
// ROWS: one array per transform, XOR swizzle keeps the stride-R writes of the first stage
// conflict free for 64-bit accesses (16 lanes per phase).
struct MapRows {
  unsigned base;   // element offset of this transform's array
  unsigned sh;     // log2(first radix)
  unsigned skew;   // per-array XOR (0..15) so that lanes spanning arrays do not collide
  B200_HD unsigned operator()(unsigned idx) const { return base + ((idx ^ ((idx >> sh) & 15u)) ^ skew); }
};
// COLS: B transforms interleaved, element idx of transform b at idx*B + b; lanes span b.
struct MapCols {
  unsigned b;
  unsigned lb;     // log2(B), B <= 16
  unsigned sh;
  B200_HD unsigned operator()(unsigned idx) const {
    unsigned p = (idx << lb) + b;
    unsigned m = (16u >> lb) - 1u;          // 0 when B = 16
    return p ^ (((idx >> sh) & m) << lb);
  }
};

}  // namespace b200
