// phasebench.cu -- how much does running the FP32 (butterfly) phases of one half-CTA under the L1TEX phases
// (exchanges, global loads / stores) of the other half buy?  Synthetic K1-shaped tile loop on the real fft_c2 core.
//   mode 0  one 512-thread group, CTA-wide barriers (the round-1 structure)
//   mode 1  two independent 256-thread groups on named barriers (half tiles), no coupling
//   mode 2  mode 1 + a token handed back and forth: only one group computes butterflies at a time (ping-pong)
//   mode 3  mode 1 + a start skew of group 1
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../dspsr_b200/csrc phasebench.cu -o phasebench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fft_c2.cuh"

using namespace b200;

struct Args {
  const unsigned* src;
  float4* dst;
  const float2* tw;
  unsigned ntile_per_cta;
  unsigned long long src_mask, dst_mask;   // in elements
  int do_store, do_load, skew;
};

template <int ID, int N> __device__ __forceinline__ void bar_sync() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(N) : "memory"); }
__device__ __forceinline__ void bar_sync_r(unsigned id, unsigned n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_r(unsigned id, unsigned n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_phase(Args a) {
  extern __shared__ __align__(128) float4 smem4[];
  constexpr unsigned P = 2048, T = P / 16;
  constexpr unsigned RS = c2::pair_slots<P>() + 2;
  constexpr bool SPLIT = MODE != 0;
  constexpr unsigned GT = SPLIT ? 256 : 512;          // threads per group
  const unsigned grp = SPLIT ? threadIdx.x / 256 : 0;
  const unsigned tl = threadIdx.x % GT;
  const unsigned pair = tl / T + grp * (GT / T);      // pair region in shared memory
  const unsigned j = tl % T;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* sm = smem4 + pair * RS;
  const bool token = (MODE == 2);
  // named barriers: 1 + grp = group barrier; 3 + grp = "group grp may compute"
  auto gsync = [&]() { if (SPLIT) bar_sync_r(1 + grp, 256); else __syncthreads(); };
  bool have = (grp == 0);                              // group 0 starts with the token
  auto acq = [&]() {
    if (!token) return;
    if (!have) bar_sync_r(3 + grp, 512);
    have = false;
  };
  auto rel = [&]() { if (token) bar_arrive_r(3 + (grp ^ 1), 512); };
  if (MODE == 3 && grp == 1) __nanosleep(a.skew);

  float2 va[16], vb[16];
  unsigned w[16];
  unsigned long long tile = (unsigned long long)blockIdx.x * a.ntile_per_cta * 2 + grp;
  auto issue = [&](unsigned long long t) {
    // 8 rows x 16 bytes per warp request, rows 4 KiB apart (the raw CASPSR stream seen by K1)
    const unsigned long long base = (t * (2048ull * 1024ull)) & a.src_mask;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const unsigned long long row = (warp % 16) * 8 + lane / 4 + 128ull * e;
      w[e] = a.do_load ? __ldg(a.src + ((base + row * 1024ull + (lane % 4)) & a.src_mask)) : (unsigned)(row + lane);
    }
  };
  issue(tile);
  const float2 wa = make_float2(0.6f, 0.8f), wb = make_float2(0.8f, -0.6f);
  for (unsigned it = 0; it < a.ntile_per_cta; it++, tile += 2) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      va[e] = make_float2(float(w[e] & 255u) - 127.5f, float((w[e] >> 8) & 255u) - 127.5f);
      vb[e] = make_float2(float((w[e] >> 16) & 255u) - 127.5f, float(w[e] >> 24) - 127.5f);
    }
    gsync();                                   // the previous tile's gathers are done
    acq();
    c2::stage_compute<P, 0, false>(va, vb, j, a.tw);
    rel();
    c2::scatter<P, 0>(sm, va, vb, j);
    gsync();
    c2::gather<P>(sm, va, vb, j);
    acq();
    c2::stage_compute<P, 1, false>(va, vb, j, a.tw);
    rel();
    if (!token) gsync();
    c2::scatter<P, 1>(sm, va, vb, j);
    gsync();
    c2::gather<P>(sm, va, vb, j);
    acq();
    c2::stage_compute<P, 2, false>(va, vb, j, a.tw);
#pragma unroll
    for (int e = 0; e < 16; e++) { va[e] = cmul(cmul(va[e], wa), wb); vb[e] = cmul(cmul(vb[e], wb), wa); }
    rel();
    if (it + 1 < a.ntile_per_cta) issue(tile + 2);
    {
      const unsigned long long base = (tile * (2048ull * 512ull)) & a.dst_mask;   // float4 units: 2048 rows x 512 float4
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const unsigned long long row = (warp % 16) * 8 + lane / 4 + 128ull * e;
        const float4 v = make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y);
        if (a.do_store) __stcs(a.dst + ((base + row * 512ull + (lane % 4) + 4 * (tile & 127)) & a.dst_mask), v);
        else if (v.x == 12345.678f) a.dst[0] = v;
      }
    }
  }
  // group 0 started with the token: it consumes group 1's last hand-over so that no arrival is left pending
  if (token && grp == 0) bar_sync_r(3, 512);
}

template <int MODE> float run(Args a, int reps, size_t smem) {
  cudaFuncSetAttribute(k_phase<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_phase<MODE><<<148, 512, smem>>>(a);
  cudaEventRecord(e0);
  for (int r = 0; r < reps; r++) k_phase<MODE><<<148, 512, smem>>>(a);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main(int argc, char** argv) {
  const unsigned ntile = argc > 1 ? atoi(argv[1]) : 64;       // half tiles per group (mode 0: full tiles = ntile)
  std::vector<float2> h(c2::twiddle_count<2048>());
  c2::fill_twiddles<2048>(h.data());
  float2* tw; cudaMalloc(&tw, h.size() * sizeof(float2));
  cudaMemcpy(tw, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
  const size_t src_n = 1ull << 28, dst_n = 1ull << 27;        // 1 GiB of words, 2 GiB of float4
  unsigned* src; float4* dst;
  cudaMalloc(&src, src_n * 4); cudaMalloc(&dst, dst_n * 16);
  cudaMemset(src, 0x5a, src_n * 4);
  const size_t smem = 4 * (c2::pair_slots<2048>() + 2) * sizeof(float4);
  for (int cfg = 0; cfg < 3; cfg++) {
    Args a{src, dst, tw, ntile, src_n - 1, dst_n - 1, cfg != 1, cfg != 1, 2000};
    if (cfg == 2) a.do_load = 0;
    // mode 0 processes full 16 Ki-point tiles: the same number of points as two groups x ntile half tiles
    Args a0 = a;
    float t0 = run<0>(a0, 5, smem);
    float t1 = run<1>(a, 5, smem);
    float t2 = run<2>(a, 5, smem);
    float t3 = run<3>(a, 5, smem);
    const double tiles = 148.0 * ntile;     // 16 Ki-point tile equivalents
    printf("cfg %d (store %d load %d): per 16Ki tile-equivalent on one SM [us]: cta-sync %.3f | 2 groups %.3f | ping-pong %.3f | skewed %.3f\n",
           cfg, a.do_store, a.do_load, t0 * 1e3 / ntile, t1 * 1e3 / ntile, t2 * 1e3 / ntile, t3 * 1e3 / ntile);
    (void)tiles;
  }
  return 0;
}
