// clusterconv.cuh -- what the one-kernel convolution path (clusterconv.cu) and the long-transform kernels (longconv.cu)
// share: the argument block that describes source, tables and sink, the per-format sample loader, small PTX wrappers.
#pragma once
#include "engine.cuh"
#include "fft_c2.cuh"

namespace b200 {

struct CcArgs {
  const void* src;
  uint64_t span, step, first;
  float scale;
  unsigned sample_swap;
  const float* lut;         // generic 8-bit: the 256-entry table
  const float2* H;          // [nchan_in][N] natural bin order, or null
  const float2* tw;         // c2 stage tables of Q
  const float2* blo;
  const float2* bhi;        // two-level table of W_N
  unsigned nchan_in, nb, ntiles, tiles_per_cluster;
  uint64_t part0;
  unsigned nfilt_pos, nkeep;
  FbSink sink;
  float4* xch;              // CC_GROUP: [groups][4][16][4096] exchange matrices (X0, X1, Y0, Y1)
  unsigned* bar;            // CC_GROUP: two arrival counters per group (32 words apart), zero at launch
};

__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

struct CcSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// complex sample n of (channel ic, polarisation pol) of the part (same formats and arithmetic as k_cols_fwd)
template <int SRC>
__device__ __forceinline__ float2 cc_load(const CcArgs& a, const float* s_lut, unsigned ic, unsigned pol, uint64_t part,
                                          unsigned n) {
  if (SRC == SRC_F32) {
    const float2* f = reinterpret_cast<const float2*>(static_cast<const float*>(a.src) + (uint64_t(ic) * 2 + pol) * a.span +
                                                      part * a.step);
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(f + n));
    return r;
  } else if (SRC == SRC_MEERKAT8) {
    // heaps of 256 samples, [heap][pol][chan][256 x (re, im) int8] (MeerKATUnpacker.C:196-229)
    uint64_t i = a.first + part * a.step + n;
    if (a.sample_swap == 2) i ^= 1ull;
    const uint64_t word = (((i >> 8) * 2 + pol) * a.nchan_in + ic) * 256ull + (i & 255ull);
    const unsigned short w = __ldg(static_cast<const unsigned short*>(a.src) + word);
    return make_float2(__fmul_rn(float(int(int8_t(w & 255u))) + 0.5f, a.scale),
                       __fmul_rn(float(int(int8_t(w >> 8))) + 0.5f, a.scale));
  } else if (SRC == SRC_GENERIC8) {
    // TFP bytes of complex samples: i*(nchan*npol*2) + 2*(npol*c + p) + d (BitUnpacker.C:56-75), through the table
    const uint64_t i = a.first + part * a.step + n;
    const uint64_t off = i * (uint64_t(a.nchan_in) * 4u) + 2u * (2u * ic + pol);
    const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(static_cast<const unsigned char*>(a.src) + off));
    return make_float2(s_lut[w & 255u], s_lut[w >> 8]);
  } else {
    // UWB: blocks of 2048 complex int16 samples per polarisation, offset binary (UWBUnpacker.C:177-218)
    const uint64_t i = a.first + part * a.step + n;
    const uint64_t word = ((i >> 11) * 2 + pol) * 2048ull + (i & 2047ull);
    const unsigned w = __ldg(static_cast<const unsigned*>(a.src) + word);
    return make_float2(float(short((w & 0xffffu) ^ 0x8000u)), float(short((w >> 16) ^ 0x8000u)));
  }
}

}  // namespace b200
