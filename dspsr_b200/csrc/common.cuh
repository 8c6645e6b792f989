// common.cuh -- shared device helpers and host-side error plumbing for libb200dsp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/b200dsp.h"
#include "fft_core.cuh"

namespace b200 {

// Tuning switches.  The product build (libb200dsp.so) has NONE: tune_flag / tune_int are constexpr and return the
// built-in choice, so every variant test below folds away at compile time and the library never reads the environment.
// `make dev` builds libb200dsp_dev.so with -DB200_TUNING, in which the same calls read B200_* variables once: that build
// serves the A/B measurements behind DESIGN.md section 6 (scratch/) and tests/test_gpu_variants.py.
#ifdef B200_TUNING
#include <cstdlib>
inline int tune_int(const char* name, int def) {
  const char* e = getenv(name);
  return e ? atoi(e) : def;
}
inline bool tune_flag(const char* name, bool def) { return tune_int(name, def ? 1 : 0) != 0; }
#else
constexpr int tune_int(const char*, int def) { return def; }
constexpr bool tune_flag(const char*, bool def) { return def; }
#endif

// ---- host-side error handling: no exceptions cross the C ABI --------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return b200::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define B200_REQUIRE(cond, ...)               \
  do {                                        \
    if (!(cond)) {                            \
      b200::set_error(__VA_ARGS__);           \
      return B200_ERR_INVALID;                \
    }                                         \
  } while (0)

// per-kernel-class device timing (bench.py's roofline): events are recorded around every launch
// of a class while timing is enabled and summed at read time.
enum KernelClass { KC_COLS_FWD = 0, KC_ROWS = 1, KC_INV = 2, KC_BINS = 3, KC_OTHER = 4, KC_COUNT = 5 };
struct TimedLaunch {
  int kc;
  cudaEvent_t start, stop;
};
struct Context {
  int device;
  cudaStream_t stream;
  int sm_count;
  int max_smem_optin;
  unsigned long long launches;   // kernels launched through this context (bench's gpu_launches)
  bool timing;
  std::vector<TimedLaunch>* timed;
  // 8 KiB of device memory owned by the context for the small tables that travel with stand-alone calls (the 256-entry
  // table of b200_unpack, the 2 x 513 levels of b200_unpack_twobit): execute-type calls never allocate
  float* d_tables;
};

// RAII helper: brackets one kernel launch with events when timing is on, and counts it
struct LaunchScope {
  Context* c;
  cudaEvent_t stop;
  LaunchScope(Context* ctx, int kc) : c(ctx), stop(nullptr) {
    c->launches++;
    if (c->timing) {
      TimedLaunch t;
      t.kc = kc;
      cudaEventCreate(&t.start);
      cudaEventCreate(&t.stop);
      cudaEventRecord(t.start, c->stream);
      stop = t.stop;
      c->timed->push_back(t);
    }
  }
  ~LaunchScope() {
    if (stop) cudaEventRecord(stop, c->stream);
  }
};

static inline unsigned ilog2(uint64_t x) {
  unsigned l = 0;
  while ((1ull << (l + 1)) <= x) l++;
  return l;
}
static inline bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }

// Device twiddle tables owned by a plan
struct TwiddleTable {
  float2* tw = nullptr;     // exp(-2 pi i m / n), m < n
  float2* stage = nullptr;  // compact per-stage tables for the EPT = 32 compile-time path (fft_core.cuh)
  float2* stage16 = nullptr;  // same for EPT = 16
  unsigned n = 0;
};
// two-level table for big transforms: W_n^m = hi[m >> 11] * lo[m & 2047]
struct BigTwiddle {
  float2* lo = nullptr;
  float2* hi = nullptr;
  unsigned nlo = 0, nhi = 0;
  uint64_t n = 0;
};
int make_twiddle(TwiddleTable& t, unsigned n, cudaStream_t s);
int make_big_twiddle(BigTwiddle& t, uint64_t n, cudaStream_t s);
void free_twiddle(TwiddleTable& t);
void free_big_twiddle(BigTwiddle& t);

#ifdef __CUDACC__

// W_n^m from the two-level table (forward sign); one complex multiply (<= 1.5 ulp)
template <bool INV>
__device__ __forceinline__ float2 big_twiddle(const float2* __restrict__ lo, const float2* __restrict__ hi,
                                              unsigned m) {
  float2 a = __ldg(lo + (m & 2047u));
  float2 b = __ldg(hi + (m >> 11));
  float2 w = cmul(a, b);
  return INV ? cconj(w) : w;
}

// Block-cooperative FFT.  Every thread of the CTA must call it (it contains __syncthreads()).
//   N      transform length (power of two, >= EPT unless EPT == N)
//   j, T   this thread's index within its transform and threads per transform (N / EPT)
//   map    shared-memory index map of this thread's transform
//   load   load(idx)  -> float2      called EPT times for idx = j + e*T
//   store  store(idx, value, e)      called EPT times; idx = j + e*T is the natural-order output
//                                     index of register e
template <int EPT, bool INV, typename Map, typename LoadF, typename StoreF>
__device__ __forceinline__ void block_fft(unsigned N, unsigned j, unsigned T, const Map& map, float2* smem,
                                          const float2* __restrict__ tw, unsigned NT, LoadF load,
                                          StoreF store) {
  float2 v[EPT];
#pragma unroll
  for (int e = 0; e < EPT; e++) v[e] = load(j + e * T);
  unsigned Ns = 1, rem = N;
#pragma unroll 1
  while (rem > 1) {
    const unsigned R = rem >= (unsigned)EPT ? (unsigned)EPT : rem;
    const bool last = (rem == R);
#define B200_STAGE(RR)                                                                   \
  {                                                                                      \
    stage_compute<EPT, RR, INV>(v, j, T, Ns, tw, NT);                                    \
    constexpr int NB = EPT / RR;                                                         \
    if (last) {                                                                          \
      _Pragma("unroll") for (int q = 0; q < NB; q++)                                     \
        _Pragma("unroll") for (int r = 0; r < RR; r++)                                   \
          store(stage_dest<EPT, RR>(j, T, Ns, q, r), v[q + r * NB], q + r * NB);         \
    } else {                                                                             \
      __syncthreads();                                                                   \
      _Pragma("unroll") for (int q = 0; q < NB; q++)                                     \
        _Pragma("unroll") for (int r = 0; r < RR; r++)                                   \
          smem[map(stage_dest<EPT, RR>(j, T, Ns, q, r))] = v[q + r * NB];                \
      __syncthreads();                                                                   \
      _Pragma("unroll") for (int e = 0; e < EPT; e++) v[e] = smem[map(j + e * T)];       \
    }                                                                                    \
  }
    if (R == (unsigned)EPT) B200_STAGE(EPT)
    else if (EPT > 16 && R == 16) B200_STAGE((EPT > 16 ? 16 : EPT))
    else if (EPT > 8 && R == 8) B200_STAGE((EPT > 8 ? 8 : EPT))
    else if (EPT > 4 && R == 4) B200_STAGE((EPT > 4 ? 4 : EPT))
    else if (EPT > 2 && R == 2) B200_STAGE((EPT > 2 ? 2 : EPT))
#undef B200_STAGE
    Ns *= R;
    rem /= R;
  }
}

// Compile-time-sized block FFT: N, EPT and therefore every radix / stride are constants, the
// stage loop is unrolled by template recursion.  Same contract as block_fft.
template <int EPT, bool INV, unsigned N, unsigned NS, typename Map, typename StoreF>
__device__ __forceinline__ void block_fft_ct_stages(float2* v, unsigned j, const Map& map, float2* smem,
                                                    const float2* __restrict__ tw /* stage tables */, StoreF store) {
  constexpr unsigned T = N / EPT;
  constexpr unsigned REM = N / NS;
  constexpr int R = REM >= (unsigned)EPT ? EPT : (int)REM;
  constexpr int NB = EPT / R;
  constexpr bool last = (REM == (unsigned)R);
  stage_compute_ct<EPT, R, INV, N, NS>(v, j, tw);
  if constexpr (last) {
#pragma unroll
    for (int q = 0; q < NB; q++)
#pragma unroll
      for (int r = 0; r < R; r++) {
        // last stage: NS = N/R, so element (q, r) = register e = q + r*NB lands at j + e*T
        store(j + (q + r * NB) * T, v[q + r * NB], q + r * NB);
      }
  } else {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NB; q++)
#pragma unroll
      for (int r = 0; r < R; r++) {
        const unsigned b = j + q * T, k = b & (NS - 1);
        smem[map((b - k) * R + k + r * NS)] = v[q + r * NB];
      }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; e++) v[e] = smem[map(j + e * T)];
    block_fft_ct_stages<EPT, INV, N, NS * R>(v, j, map, smem, tw, store);
  }
}

template <int EPT, bool INV, unsigned N, typename Map, typename LoadF, typename StoreF>
__device__ __forceinline__ void block_fft_ct(unsigned j, const Map& map, float2* smem,
                                             const float2* __restrict__ tw, LoadF load, StoreF store) {
  constexpr unsigned T = N / EPT;
  float2 v[EPT];
#pragma unroll
  for (int e = 0; e < EPT; e++) v[e] = load(j + e * T);
  block_fft_ct_stages<EPT, INV, N, 1>(v, j, map, smem, tw, store);
}

#endif  // __CUDACC__

}  // namespace b200
