// fastpath.cu -- second-generation K1 / K2 / K3 for the transform sizes of the headline workloads.
//
// Same three-kernel decomposition as filterbank.cu (forward column pass, row pass + real split +
// response, per-channel inverse pass + epilogue) on the fft_c2.cuh core: 16 points of TWO sequences
// per thread, 128-bit padded shared-memory exchanges with immediate addressing, stage twiddles
// loaded once per pair.
//   K1  k1_c2   pair = two adjacent columns: one 32-bit load carries the 4 raw bytes of both, one
//               128-bit store writes both
//   K2  k2_c2   pair = row k1 and its mirror row P-k1 (real input) or two adjacent rows (complex)
//       k2_r32  Q = 1024 planned 32.32, one warp per row (the default for cfg1); writes Z in the tile-major order
//               zt_pos() that K3 reads with 256-bit loads, with the response pre-permuted into the same order
//   K3  k3_c2   pair = the two polarisations of an output channel: detection straight from registers; 8192 points
//               planned 32.16.16; fold epilogue by items of the bin plan (fold.cu k_bin_runs)
// All three are PERSISTENT: one 512-thread CTA per SM loops over tiles, and the global loads of the
// next tile are issued into the (by then dead) data registers before the current tile's epilogue,
// so load latency hides behind the store / split / fold phase instead of adding to it (one tile
// fills the SM's registers and most of its shared memory: there is no second CTA to overlap with).
// fb_run (filterbank.cu) dispatches here when the plan's sizes are instantiated below; every other
// shape keeps the generic kernels.  B200_FAST=0 disables the dispatch (A/B measurements).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "engine.cuh"
#include "fft_c2.cuh"

#ifndef B200_K1_NP
#define B200_K1_NP 4
#endif

namespace b200 {

// Ablation switches (B200_DBG1/2/3 bit flags: 1 skip the FFT, 2 skip the epilogue, 4 skip the loads) exist only in
// builds made with -DB200_ABLATION: in the product build the tests compile away.
#ifdef B200_ABLATION
#define B200_DBGF(a, bit) ((a).dbg & (bit))
#else
#define B200_DBGF(a, bit) 0
#endif

__device__ __forceinline__ float2 ldg_nc_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// 256-bit load of four consecutive float2 (sm_100: LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void ldg_nc_f2x4(const float2* p, float2& a, float2& b, float2& c, float2& d) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y), "=f"(c.x), "=f"(c.y), "=f"(d.x), "=f"(d.y) : "l"(p));
}
// Tile-image Z (see k2_g2): a K2 tile (8 mirror row pairs x Q = 1024 bins = 16384 float2 = 128 KiB) travels to K3 as
// the verbatim image of K2's shared-memory arrays, written by two 64 KiB bulk copies (one per half-CTA group).
// Offset, inside the tile, of element k2 of pair g (0..7), row `which` (0: row low, 1: its mirror row):
//   [g / 4][k2 / 4][which][ ((g % 4) * 4 + k2 % 4) ^ h(k2 / 4) ],   h(m) = ((m >> 3) ^ (m << 2)) & 15
// * the four k2 of one K3 channel and row stay one aligned 32-byte group (the XOR permutes them inside the group by
//   h & 3 and moves the group inside its 128-byte line by h >> 2): K3 fetches them with one 256-bit load;
// * the four rows g % 4 of a group fill a whole 128-byte line;
// * as a SHARED-MEMORY layout it is conflict free for every access pattern of K2 (64-bit accesses, 16 lanes per
//   wavefront): the radix-32 scatter (lane j writes element 32 j + r: bank pair = const ^ j), the gather (lane j
//   reads j + 32 e) and the split walk (4 consecutive k2 x 4 pairs per half warp).
__host__ __device__ __forceinline__ unsigned zi_h(unsigned m) { return ((m >> 3) ^ (m << 2)) & 15u; }
__host__ __device__ __forceinline__ unsigned zi_pos(unsigned g, unsigned which, unsigned k2) {
  const unsigned m = k2 >> 2;
  return (g >> 2) * 8192u + m * 32u + which * 16u + ((((g & 3u) << 2) | (k2 & 3u)) ^ zi_h(m));
}

// Cache policy of the once-written / once-read streams (spectrum scratch A and Z).  st.global.cs (evict
// first) for the Z stores of K2 measured -7.7 % on that kernel; B200_NO_STREAMING restores default policies.
#ifndef B200_ZST_FN
#define B200_ZST_FN __stcs
#endif
#ifndef B200_NO_STREAMING
#define B200_ZST(p, v) B200_ZST_FN(p, v)
#define B200_AST(p, v) B200_AST_IMPL(p, v)
#define B200_LDS1(p) B200_LDS1_IMPL(p)
#else
#define B200_ZST(p, v) (*(p) = (v))
#define B200_AST(p, v) (*(p) = (v))
#define B200_LDS1(p) ldg_nc_f2(p)
#endif
#ifndef B200_AST_FN
#define B200_AST_FN __stcs
#endif
#define B200_AST_IMPL(p, v) B200_AST_FN(p, v)
#ifdef B200_LD_STREAMING
#define B200_LDS1_IMPL(p) __ldcs(p)
#else
#define B200_LDS1_IMPL(p) ldg_nc_f2(p)
#endif

struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// ------------------------------------------------------------------------------------------
// K1: unpack + forward column pass.  tile = (column block of 2*NP columns, block = part x chan x pol)
// ------------------------------------------------------------------------------------------
struct K1Args {
  const void* src;
  uint64_t span, step;
  const float* lut;
  float2* dst;
  const float2* tw;      // c2 stage tables of P
  const float2* blo;
  const float2* bhi;
  unsigned Q, npol, nchan_in, Nc, nblk;
  uint64_t part0;
  unsigned dbg;
  unsigned skew_ns, nsm;
  int conv_ok;               // 8-bit table is RN(x*(conv_hi+conv_lo)): convert arithmetically, no gathers
  float conv_hi, conv_lo;
  int pad_;                  // keeps the fields below where ptxas allocates K1 without a spill (16 bytes of stack otherwise)
  int l2_prefetch;           // prefetch the next part's raw bytes into L2 as contiguous slices
  unsigned overlap;          // nsamp_overlap (samples)
};

// QC: row length Q fixed at compile time (0 = run-time a.Q): all row strides become immediates
template <int SRC, unsigned P, int NP, unsigned QC = 0>
__global__ void __launch_bounds__(NP*(P / 16), 512 / (NP * (P / 16)))
k1_c2(K1Args a) {
  extern __shared__ __align__(128) float4 smem4[];
  __shared__ float s_lut[256];
  __shared__ float4 s_h[16 * NP];   // [e][pair] = (W_N^(n2a*T*e), W_N^(n2b*T*e))
  constexpr unsigned T = P / 16;
  static_assert(c2::pair_slots<P>() % 8 == 0, "region skew assumes an 8-aligned pair size");
  constexpr unsigned RS = c2::pair_slots<P>() + 8 / NP;   // regions skewed so the NP pairs of a phase hit distinct banks
  const unsigned pair = threadIdx.x % NP, j = threadIdx.x / NP;
  const unsigned Q = QC ? QC : a.Q;
  const unsigned ncolblk = Q / (2 * NP);
  const unsigned ntiles = ncolblk * a.nblk;

  if (SRC == SRC_CASPSR8) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = a.lut[i];
    __syncthreads();
  }
  // co-resident CTAs would otherwise run their phases in lock step: stagger every other one
  if (a.skew_ns && (blockIdx.x / a.nsm) & 1) __nanosleep(a.skew_ns);

  // CASPSR: byte 8*(i/4) + 4*pol + i%4 (CASPSRUnpacker.C:141-187); columns n2, n2+1 = samples
  // 2*n2 .. 2*n2+3 = one 4-byte group of this polarisation
  auto raw_ptr = [&](unsigned t) -> const unsigned char* {
    // tile order: polarisation fastest, then column block -- the two polarisations of a column block
    // share every 32-byte sector of the raw stream, so they run side by side and the second reader hits L2
    const unsigned pol = t % a.npol, col0 = ((t / a.npol) % ncolblk) * (2 * NP);
    const unsigned rest = t / (a.npol * ncolblk);
    const uint64_t part = a.part0 + rest / a.nchan_in;
    return static_cast<const unsigned char*>(a.src) + 2ull * (part * a.step) + 4u * pol +
           4ull * (uint64_t(j) * Q + col0 + 2 * pair);
  };

  unsigned w[16];
  unsigned t = blockIdx.x;
  if (SRC == SRC_CASPSR8 && t < ntiles && !(B200_DBGF(a, 4))) {
    const unsigned char* raw = raw_ptr(t);
#pragma unroll
    for (int e = 0; e < 16; e++) w[e] = __ldg(reinterpret_cast<const unsigned*>(raw + 4ull * Q * T * e));
  }

  for (; t < ntiles; t += gridDim.x) {
    const unsigned pol = t % a.npol, col0 = ((t / a.npol) % ncolblk) * (2 * NP);
    const unsigned rest = t / (a.npol * ncolblk);            // (part, input channel)
    const unsigned n2 = col0 + 2 * pair;
    const unsigned ic = rest % a.nchan_in;
    const uint64_t part = a.part0 + rest / a.nchan_in;
    const unsigned blk = rest * a.npol + pol;

    // L2 prefetch of the NEXT part's raw bytes.  This kernel reads the raw stream as 32-byte granules 4 KiB
    // apart (one per FFT row), which DRAM serves at a quarter of its streaming rate; the same bytes requested
    // ahead of time as one contiguous slice per tile stream into L2 at full speed, and the scattered reads
    // of the next part then hit L2.  (tiles are ordered part-slowest: the 2*ncolblk tiles of a part each
    // prefetch 1/(2*ncolblk) of the following part.)
    if (SRC == SRC_CASPSR8 && a.l2_prefetch && threadIdx.x == 0 && a.nchan_in == 1) {
      const unsigned slices = a.npol * ncolblk;
      const unsigned slice = t % slices;
      const uint64_t next_part = a.part0 + rest + 1;
      if (rest + 1 < a.nblk / a.npol) {
        // new bytes of the next part: [ (next_part*step + overlap) , ((next_part+1)*step + overlap) ) samples x npol bytes
        const uint64_t lo = a.npol * (next_part * a.step + a.overlap), len = a.npol * a.step;
        const uint64_t per = ((len / slices) + 15) & ~15ull;
        const uint64_t off = slice * per;
        if (off < len) {
          const uint64_t n = min(per, len - off) & ~15ull;
          const unsigned char* ptr = static_cast<const unsigned char*>(a.src) + ((lo + off) & ~15ull);
          if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(ptr), "r"((unsigned)n) : "memory");
        }
      }
    }

    // twiddles of this tile (loads in flight while the samples are converted)
    float2 shv = make_float2(1.f, 0.f);
    if (threadIdx.x < 16 * NP * 2) {
      const unsigned e = threadIdx.x / (2 * NP), col = threadIdx.x % (2 * NP);
      shv = big_twiddle<false>(a.blo, a.bhi, ((col0 + col) * T * e) & (a.Nc - 1));
    }
    // W_N^(n2*k1), k1 = j + e*T, factorised into W_N^(n2*j) (per thread) and W_N^(n2*T*e) (s_h)
    const float2 wa = big_twiddle<false>(a.blo, a.bhi, (n2 * j) & (a.Nc - 1));
    const float2 wb = big_twiddle<false>(a.blo, a.bhi, ((n2 + 1) * j) & (a.Nc - 1));

    float2 va[16], vb[16];
    if (B200_DBGF(a, 4)) {
#pragma unroll
      for (int e = 0; e < 16; e++) { va[e] = make_float2(float(threadIdx.x + e), 1.f); vb[e] = make_float2(2.f, float(e)); }
    } else if (SRC == SRC_CASPSR8) {
      if (a.conv_ok) {
        // byte -> float without the table: (b ^ 0x80) dropped into bits 8..15 of the float 32768 reads
        // 32768 + 128 + int8(b); subtracting 32895.5 leaves x = int8(b) + 0.5 exactly, and fma(x, hi, x*lo)
        // is the table entry bit for bit (verified for all 256 entries on the host, lut_as_arithmetic)
        // two samples (one complex point) at a time with the packed FP32x2 instructions: same IEEE
        // operations per lane, half the issue slots
#ifdef __CUDA_ARCH__
        const unsigned long long off2 = pk2(-32895.5f, -32895.5f);
        const unsigned long long lo2 = pk2(a.conv_lo, a.conv_lo), hi2 = pk2(a.conv_hi, a.conv_hi);
        auto cv2 = [&](unsigned word, unsigned sel0, unsigned sel1) -> float2 {
          const unsigned long long x = add2(pk2(__uint_as_float(__byte_perm(word, 0x47000000u, sel0)),
                                                __uint_as_float(__byte_perm(word, 0x47000000u, sel1))), off2);
          return up2(fma2(x, hi2, mul2(x, lo2)));
        };
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const unsigned x = w[e] ^ 0x80808080u;
          va[e] = cv2(x, 0x7604, 0x7614);
          vb[e] = cv2(x, 0x7624, 0x7634);
        }
#endif
      } else {
#pragma unroll
        for (int e = 0; e < 16; e++) {
          va[e] = make_float2(s_lut[w[e] & 255u], s_lut[(w[e] >> 8) & 255u]);
          vb[e] = make_float2(s_lut[(w[e] >> 16) & 255u], s_lut[w[e] >> 24]);
        }
      }
    } else {
      const float2* f = reinterpret_cast<const float2*>(static_cast<const float*>(a.src) +
                                                        (uint64_t(ic) * a.npol + pol) * a.span + part * a.step) +
                        uint64_t(j) * Q + n2;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        va[e] = ldg_nc_f2(f + uint64_t(Q) * T * e);
        vb[e] = ldg_nc_f2(f + uint64_t(Q) * T * e + 1);
      }
    }
    __syncthreads();     // the previous tile's readers of s_h and of the exchange buffer are done
    if (threadIdx.x < 16 * NP * 2) reinterpret_cast<float2*>(s_h)[threadIdx.x] = shv;

    if (!(B200_DBGF(a, 1))) c2::fft_pair<P, false>(va, vb, j, smem4 + pair * RS, a.tw, CtaSync());
    else __syncthreads();

    // prefetch the raw words of the next tile; they land while this tile is twiddled and stored
    const unsigned tn = t + gridDim.x;
    if (SRC == SRC_CASPSR8 && tn < ntiles && !(B200_DBGF(a, 4))) {
      const unsigned char* raw = raw_ptr(tn);
#pragma unroll
      for (int e = 0; e < 16; e++) w[e] = __ldg(reinterpret_cast<const unsigned*>(raw + 4ull * Q * T * e));
    }

    {
      float2* dst = a.dst + uint64_t(blk) * a.Nc + uint64_t(j) * Q + n2;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const float4 h = s_h[e * NP + pair];
        const float2 xa = cmul(cmul(va[e], wa), make_float2(h.x, h.y));
        const float2 xb = cmul(cmul(vb[e], wb), make_float2(h.z, h.w));
        if (!(B200_DBGF(a, 2)) || xa.x == 12345.678f)
          B200_AST(reinterpret_cast<float4*>(dst + uint64_t(Q) * T * e), make_float4(xa.x, xa.y, xb.x, xb.y));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2: forward row pass + real-input split + response.  tile = (G row pairs, block)
// ------------------------------------------------------------------------------------------
struct K2Args {
  const float2* A;
  float2* Z;
  const float2* H;
  const float2* Ht;        // H in the tile-image order of Z (zi_pos), used when z_tiled
  const float2* tw;        // c2 stage tables of Q
  const float2* tw2Q;      // exp(-2 pi i m / (2Q))
  const float2* tw32;      // 32.32 plan: W_1024^(s j), s = 1..7, then W_1024^(8 m j), m = 1..3  ([10][32])
  const float2* b2lo;
  const float2* b2hi;
  unsigned Nc, npol, nchan_in, nblk;
  unsigned dbg;
  int z_tiled;             // Z leaves as the tile image (k2_g2 + the 32.16.16 K3 only)
};

template <unsigned P, unsigned Q, bool SPLIT>
__global__ void __launch_bounds__(512, 1) k2_c2(K2Args a) {
  extern __shared__ float4 smem4[];
  constexpr unsigned T = Q / 16;
  constexpr unsigned G = 512 / T;                          // sequence pairs per CTA
  constexpr unsigned RS = c2::pair_slots<Q>() | 1u;        // odd: phase 2 walks the pairs lane by lane
  constexpr unsigned TPB = SPLIT ? (P / 2) / G : P / (2 * G);   // tiles per block
  __shared__ float2 s_rowtw[G + 1];
  const unsigned g = threadIdx.x / T, j = threadIdx.x % T;
  const unsigned Nc = a.Nc;
  const unsigned ntiles = TPB * a.nblk;

  float2 va[16], vb[16];
  auto issue_loads = [&](unsigned tt) {
    const unsigned tile = tt % TPB, blk = tt / TPB;
    unsigned ra, rb;
    if (SPLIT) {
      const unsigned low = tile * G + g;
      ra = low;
      rb = low == 0 ? P / 2 : P - low;
    } else {
      ra = tile * 2 * G + 2 * g;
      rb = ra + 1;
    }
    const float2* pa = a.A + uint64_t(blk) * Nc + uint64_t(ra) * Q + j;
    const float2* pb = a.A + uint64_t(blk) * Nc + uint64_t(rb) * Q + j;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      if (B200_DBGF(a, 4)) { va[e] = make_float2(float(threadIdx.x + e), 1.f); vb[e] = make_float2(2.f, float(e)); continue; }
      va[e] = B200_LDS1(pa + e * T);
      vb[e] = B200_LDS1(pb + e * T);
    }
  };

  unsigned t = blockIdx.x;
  if (t < ntiles) issue_loads(t);
  for (; t < ntiles; t += gridDim.x) {
    const unsigned tile = t % TPB, blk = t / TPB;
    const unsigned ic = (blk / a.npol) % a.nchan_in;
    if (SPLIT && threadIdx.x <= G) {
      // W_2N^k, k = row + P*k2, factorises into W_2N^row (here) and W_2Q^k2 (tw2Q)
      const unsigned r = threadIdx.x < G ? tile * G + threadIdx.x : P / 2;
      s_rowtw[threadIdx.x] = big_twiddle<false>(a.b2lo, a.b2hi, r);
    }
    float4* sm = smem4 + g * RS;
    if (!(B200_DBGF(a, 1))) c2::fft_pair<Q, false>(va, vb, j, sm, a.tw, CtaSync());
    __syncthreads();
    c2::store_natural<Q>(sm, va, vb, j);
    __syncthreads();

    // the rows of the next tile stream in while this tile is split, multiplied and stored
    if (t + gridDim.x < ntiles) issue_loads(t + gridDim.x);

    if (!((B200_DBGF(a, 2)) && va[3].x != 12345.678f)) {
      // ---- phase 2: lanes walk the G pairs first (G consecutive bins = one 64-byte segment of Z) ----
      const float2* H = a.H ? a.H + uint64_t(ic) * Nc : nullptr;
      float2* Zblk = a.Z + uint64_t(blk) * Nc;
      const unsigned g2 = threadIdx.x % G, kk = threadIdx.x / G;   // kk < T
      const float4* sg = smem4 + g2 * RS;
      constexpr int SSTEP = T / 16 * 17;                           // pad16 advance of T elements
      constexpr int KSTEP = T * P;                                 // bin advance of T elements

      if (!SPLIT) {
        const unsigned k0 = tile * 2 * G + 2 * g2 + P * kk;
        const float4* sa = sg + c2::pad16(kk);
#pragma unroll 4
        for (int it = 0; it < 16; it++) {
          const float4 x = sa[SSTEP * it];
          const unsigned k = k0 + it * KSTEP;
          float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
          if (H) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(H + k));
            xa = cmul(xa, make_float2(h.x, h.y));
            xb = cmul(xb, make_float2(h.z, h.w));
          }
          *reinterpret_cast<float4*>(Zblk + k) = make_float4(xa.x, xa.y, xb.x, xb.y);
        }
      } else {
        // element k2 of row `low` is bin k = low + P*k2; its mirror N-k is element Q-1-k2 of row P-low:
        //   e = (z_k + conj z_m)/2, d = (z_k - conj z_m)/2, t = -i d W_2N^k;  X_k = e + t, X_(N-k) = conj(e - t)
        auto split = [&](float2 zk, float2 zmc, float2 w, float2& xk, float2& xm) {
          const float2 e = make_float2(0.5f * (zk.x + zmc.x), 0.5f * (zk.y + zmc.y));
          const float2 d = make_float2(0.5f * (zk.x - zmc.x), 0.5f * (zk.y - zmc.y));
          const float2 tt = cmul(make_float2(d.y, -d.x), w);
          xk = cadd(e, tt);
          xm = cconj(csub(e, tt));
        };
        const unsigned low = tile * G + g2;
        if (low != 0) {
          // Each thread takes bins k2 = kk + T*it of the lower half (k2 < Q/2) TOGETHER with their partners
          // Q-1-k2: the two float4 it reads hold (a[k2], b[k2]) and (a[Q-1-k2], b[Q-1-k2]), i.e. both
          // members of two mirror pairs -- no half of a shared-memory read is wasted.
          const float2 rw = s_rowtw[g2];
          const unsigned k0 = low + P * kk;                 // bin of a[k2]
          const unsigned k1b = low + P * (Q - 1 - kk);      // bin of a[Q-1-k2]
          const float4* sa = sg + c2::pad16(kk);
          const float4* sb = sg + c2::pad16(Q - 1 - kk);
          const float2* t2a = a.tw2Q + kk;
          const float2* t2b = a.tw2Q + (Q - 1 - kk);
          const float2* HA = H ? H + k0 : nullptr;          // item A: X[k0 + ..], mirror X[Nc - k0 - ..]
          const float2* HAm = H ? H + (Nc - k0) : nullptr;
          const float2* HB = H ? H + k1b : nullptr;         // item B: X[k1b - ..], mirror X[Nc - k1b + ..]
          const float2* HBm = H ? H + (Nc - k1b) : nullptr;
          float2* ZA = Zblk + k0;
          float2* ZAm = Zblk + (Nc - k0);
          float2* ZB = Zblk + k1b;
          float2* ZBm = Zblk + (Nc - k1b);
#pragma unroll 2
          for (int it = 0; it < 8; it++) {
            const float4 u = sa[SSTEP * it];                // (a[k2], b[k2])
            const float4 v = sb[-SSTEP * it];               // (a[Q-1-k2], b[Q-1-k2])
            const float2 wA = cmul(rw, __ldg(t2a + it * int(T)));
            const float2 wB = cmul(rw, __ldg(t2b - it * int(T)));
            float2 xk, xm, yk, ym;
            split(make_float2(u.x, u.y), make_float2(v.z, -v.w), wA, xk, xm);
            split(make_float2(v.x, v.y), make_float2(u.z, -u.w), wB, yk, ym);
            if (H) {
              xk = cmul(xk, __ldg(HA + it * KSTEP));
              xm = cmul(xm, __ldg(HAm - it * KSTEP));
              yk = cmul(yk, __ldg(HB - it * KSTEP));
              ym = cmul(ym, __ldg(HBm + it * KSTEP));
            }
            B200_ZST(ZA + it * KSTEP, xk);
            B200_ZST(ZAm - it * KSTEP, xm);
            B200_ZST(ZB - it * KSTEP, yk);
            B200_ZST(ZBm + it * KSTEP, ym);
          }
        } else {
          // rows 0 (sequence a) and P/2 (sequence b) mirror onto themselves
          const float2 rwh = s_rowtw[G];
          for (unsigned it = 0; it < 16; it++) {
            const unsigned k2 = kk + it * T;
            if (k2 <= Q / 2) {
              const unsigned km2 = (Q - k2) % Q;
              const float4 xa = sg[c2::pad16(k2)];
              const float4 xm4 = sg[c2::pad16(km2)];
              float2 xk, xm;
              split(make_float2(xa.x, xa.y), make_float2(xm4.x, -xm4.y), __ldg(a.tw2Q + k2), xk, xm);
              const unsigned k = P * k2, km = (Nc - k) & (Nc - 1);
              if (H) { xk = cmul(xk, __ldg(H + k)); xm = cmul(xm, __ldg(H + km)); }
              Zblk[k] = xk;
              if (km2 != k2) Zblk[km] = xm;
            }
            if (k2 < Q / 2) {
              const float4 xa = sg[c2::pad16(k2)];
              const float4 xm4 = sg[c2::pad16(Q - 1 - k2)];
              float2 xk, xm;
              split(make_float2(xa.z, xa.w), make_float2(xm4.z, -xm4.w), cmul(rwh, __ldg(a.tw2Q + k2)), xk, xm);
              const unsigned k = P / 2 + P * k2, km = Nc - k;
              if (H) { xk = cmul(xk, __ldg(H + k)); xm = cmul(xm, __ldg(H + km)); }
              Zblk[k] = xk;
              Zblk[km] = xm;
            }
          }
        }
      }
    }
    __syncthreads();       // phase-2 readers are done before the next tile's first scatter
  }
}

// ------------------------------------------------------------------------------------------
// K2, 32.32 plan (Q = 1024): every row is transformed by ONE WARP -- 32 threads x 32 points, two
// radix-32 stages -- so the single exchange of the transform is warp-local (__syncwarp, stride-33 padded
// 64-bit accesses) and the whole tile needs just two CTA barriers (before and after the split phase)
// instead of six.  Rows a / b of a mirror pair sit in separate float2 arrays (a at sequence 2g, b at 2g+1);
// the array stride 1057 makes the lane-by-pair walk of the split phase conflict free.  Writes Z in natural
// order (complex input, or real input when the tile-image path k2_g2 is not in use).
// ------------------------------------------------------------------------------------------
template <unsigned P, bool SPLIT>
__global__ void __launch_bounds__(512, 1) k2_r32(K2Args a) {
  extern __shared__ float4 smem4[];
  float2* sm2 = reinterpret_cast<float2*>(smem4);
  constexpr unsigned Q = 1024, G = 8, RSQ = 1057u;
  constexpr unsigned TPB = SPLIT ? (P / 2) / G : P / (2 * G);
  __shared__ float2 s_rowtw[G + 1];
  const unsigned seq = threadIdx.x >> 5, j = threadIdx.x & 31u;
  const unsigned g = seq >> 1, which = seq & 1u;
  const unsigned Nc = a.Nc;
  const unsigned ntiles = TPB * a.nblk;

  float2 v[32];
  auto issue_loads = [&](unsigned tt) {
    const unsigned tile = tt % TPB, blk = tt / TPB;
    unsigned row;
    if (SPLIT) {
      const unsigned low = tile * G + g;
      row = which ? (low == 0 ? P / 2 : P - low) : low;
    } else {
      row = tile * 2 * G + 2 * g + which;
    }
    const float2* src = a.A + uint64_t(blk) * Nc + uint64_t(row) * Q + j;
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = B200_LDS1(src + 32 * e);
  };

  unsigned t = blockIdx.x;
  if (t < ntiles) issue_loads(t);
  for (; t < ntiles; t += gridDim.x) {
    const unsigned tile = t % TPB, blk = t / TPB;
    const unsigned ic = (blk / a.npol) % a.nchan_in;
    if (SPLIT && threadIdx.x <= G) {
      const unsigned r = threadIdx.x < G ? tile * G + threadIdx.x : P / 2;
      s_rowtw[threadIdx.x] = big_twiddle<false>(a.b2lo, a.b2hi, r);
    }
    float2* sq = sm2 + seq * RSQ;
    // stage 0: thread j holds x[j + 32 e]; outputs r of butterfly j go to 32 j + r -> slot 33 j + r
    dft32<false>(v);
#pragma unroll
    for (int r = 0; r < 32; r++) sq[33u * j + r] = v[r];
    // stage-1 twiddles W_1024^(r j), r = 8 m + s, from W^(s j) (s = 1..7) and W^(8 m j) (m = 1..3)
    float2 ws[8], wm[4];
#pragma unroll
    for (int s1 = 1; s1 < 8; s1++) ws[s1] = __ldg(a.tw32 + (s1 - 1) * 32 + j);
#pragma unroll
    for (int m = 1; m < 4; m++) wm[m] = __ldg(a.tw32 + (6 + m) * 32 + j);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = sq[j + 33u * e];          // x1[j + 32 e] at slot j + 33 e
#pragma unroll
    for (int r = 1; r < 32; r++) {
      const int s1 = r & 7, m = r >> 3;
      const float2 w = m == 0 ? ws[s1] : (s1 == 0 ? wm[m] : cmul(wm[m], ws[s1]));
      v[r] = cmul(v[r], w);
    }
    dft32<false>(v);                                               // register e = X[j + 32 e]
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) sq[j + 33u * e] = v[e];           // natural order, slot k2 + (k2 >> 5)
    __syncthreads();

    // the rows of the next tile stream in while this tile is split, multiplied and stored
    if (t + gridDim.x < ntiles) issue_loads(t + gridDim.x);

    {
      const float2* H = a.H ? a.H + uint64_t(ic) * Nc : nullptr;
      float2* Zblk = a.Z + uint64_t(blk) * Nc;
      const unsigned g2 = threadIdx.x % G, kk = threadIdx.x / G;   // kk < 64; lanes walk the pairs first
      const float2* SA = sm2 + (2 * g2) * RSQ;                     // row a of pair g2
      const float2* SB = SA + RSQ;                                 // row b
      auto p33 = [](unsigned i) { return i + (i >> 5); };
      constexpr int SSTEP = 66;                                    // slot advance of 64 elements
      constexpr int KSTEP = 64 * P;                                // bin advance of 64 elements

      if (!SPLIT) {
        const unsigned k0 = tile * 2 * G + 2 * g2 + P * kk;
        const unsigned s0 = p33(kk);
#pragma unroll 4
        for (int it = 0; it < 16; it++) {
          float2 xa = SA[s0 + SSTEP * it], xb = SB[s0 + SSTEP * it];
          const unsigned k = k0 + it * KSTEP;
          if (H) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(H + k));
            xa = cmul(xa, make_float2(h.x, h.y));
            xb = cmul(xb, make_float2(h.z, h.w));
          }
          B200_ZST(reinterpret_cast<float4*>(Zblk + k), make_float4(xa.x, xa.y, xb.x, xb.y));
        }
      } else {
        auto split = [&](float2 zk, float2 zmc, float2 w, float2& xk, float2& xm) {
          const float2 e = make_float2(0.5f * (zk.x + zmc.x), 0.5f * (zk.y + zmc.y));
          const float2 d = make_float2(0.5f * (zk.x - zmc.x), 0.5f * (zk.y - zmc.y));
          const float2 tt = cmul(make_float2(d.y, -d.x), w);
          xk = cadd(e, tt);
          xm = cconj(csub(e, tt));
        };
        const unsigned low = tile * G + g2;
        if (low != 0) {
          const float2 rw = s_rowtw[g2];
          const unsigned k0 = low + P * kk;                 // bin of a[k2]
          const unsigned k1b = low + P * (Q - 1 - kk);      // bin of a[Q-1-k2]
          const unsigned sa = p33(kk), sb = p33(Q - 1 - kk);
          const float2* t2a = a.tw2Q + kk;
          const float2* t2b = a.tw2Q + (Q - 1 - kk);
          const float2* HA = H ? H + k0 : nullptr;
          const float2* HAm = H ? H + (Nc - k0) : nullptr;
          const float2* HB = H ? H + k1b : nullptr;
          const float2* HBm = H ? H + (Nc - k1b) : nullptr;
          float2* ZA = Zblk + k0;
          float2* ZAm = Zblk + (Nc - k0);
          float2* ZB = Zblk + k1b;
          float2* ZBm = Zblk + (Nc - k1b);
#pragma unroll 2
          for (int it = 0; it < 8; it++) {
            const float2 ua = SA[sa + SSTEP * it], ub = SB[sa + SSTEP * it];      // a[k2], b[k2]
            const float2 va2 = SA[sb - SSTEP * it], vb2 = SB[sb - SSTEP * it];    // a[Q-1-k2], b[Q-1-k2]
            const float2 wA = cmul(rw, __ldg(t2a + it * 64));
            const float2 wB = cmul(rw, __ldg(t2b - it * 64));
            float2 xk, xm, yk, ym;
            split(ua, make_float2(vb2.x, -vb2.y), wA, xk, xm);
            split(va2, make_float2(ub.x, -ub.y), wB, yk, ym);
            if (H) {
              xk = cmul(xk, __ldg(HA + it * KSTEP));
              xm = cmul(xm, __ldg(HAm - it * KSTEP));
              yk = cmul(yk, __ldg(HB - it * KSTEP));
              ym = cmul(ym, __ldg(HBm + it * KSTEP));
            }
            B200_ZST(ZA + it * KSTEP, xk);
            B200_ZST(ZAm - it * KSTEP, xm);
            B200_ZST(ZB - it * KSTEP, yk);
            B200_ZST(ZBm + it * KSTEP, ym);
          }
        } else {
          // rows 0 (array a) and P/2 (array b) mirror onto themselves
          const float2 rwh = s_rowtw[G];
          for (unsigned it = 0; it < 16; it++) {
            const unsigned k2 = kk + it * 64;
            if (k2 <= Q / 2) {
              const unsigned km2 = (Q - k2) % Q;
              const float2 xa = SA[p33(k2)], xm2 = SA[p33(km2)];
              float2 xk, xm;
              split(xa, make_float2(xm2.x, -xm2.y), __ldg(a.tw2Q + k2), xk, xm);
              const unsigned k = P * k2, km = (Nc - k) & (Nc - 1);
              if (H) { xk = cmul(xk, __ldg(H + k)); xm = cmul(xm, __ldg(H + km)); }
              Zblk[k] = xk;
              if (km2 != k2) Zblk[km] = xm;
            }
            if (k2 < Q / 2) {
              const float2 xb = SB[p33(k2)], xm2 = SB[p33(Q - 1 - k2)];
              float2 xk, xm;
              split(xb, make_float2(xm2.x, -xm2.y), cmul(rwh, __ldg(a.tw2Q + k2)), xk, xm);
              const unsigned k = P / 2 + P * k2, km = Nc - k;
              if (H) { xk = cmul(xk, __ldg(H + k)); xm = cmul(xm, __ldg(H + km)); }
              Zblk[k] = xk;
              Zblk[km] = xm;
            }
          }
        }
      }
    }
    __syncthreads();       // split-phase readers are done before the next tile's first scatter
  }
}

// ------------------------------------------------------------------------------------------
// K2, two half-CTA groups, in-place split, bulk stores (real input; the default for cfg1).  Same arithmetic as
// k2_r32, but
// * the 512 threads work as TWO independent 256-thread groups, each on four of the tile's eight mirror-row pairs,
//   synchronised by their own named barriers (bar.sync 1 + grp, 256): the groups drift apart and one's split phase
//   (L1TEX) runs under the other's row transforms (FP32 pipe);
// * the shared-memory arrays use the dense, XOR-swizzled tile-image layout zi_pos() -- conflict free for the
//   radix-32 scatter, the gather and the split walk -- and the split + response results are written back IN PLACE:
//   X[k] of row `low` replaces element k2 of that row, X[N-k] replaces element Q-1-k2 of the mirror row;
// * the finished 64 KiB image of a group then leaves the SM as cp.async.bulk (TMA) copies straight from shared
//   memory.  Global stores through the LSU pass the L1TEX data pipe at 32 bytes per cycle -- 4096 cycles per tile,
//   the longest single item of the old split phase -- the in-place 64-bit shared stores cost a quarter of that and
//   the copy itself runs under the next tile's first butterflies.  K3 reads the image as it is (Z travels as the
//   verbatim picture of K2's shared memory), the response is pre-permuted into the same order (k_tile_response).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_sync(unsigned id, unsigned n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }

template <unsigned P>
__global__ void __launch_bounds__(512, 1) k2_g2(K2Args a) {
  extern __shared__ __align__(128) float4 smem4[];
  float2* sm2 = reinterpret_cast<float2*>(smem4);
  constexpr unsigned Q = 1024, G = 8;
  constexpr unsigned TPB = (P / 2) / G;
  __shared__ float2 s_rowtw[G + 1];
  const unsigned grp = threadIdx.x >> 8, tl = threadIdx.x & 255u;
  const unsigned seq = threadIdx.x >> 5, j = threadIdx.x & 31u;
  const unsigned g = seq >> 1, which = seq & 1u;            // pair g (0..7; group grp owns 4 grp .. 4 grp + 3)
  const unsigned Nc = a.Nc;
  const unsigned ntiles = TPB * a.nblk;
  float2* sgrp = sm2 + grp * 8192u;                         // this group's half of the tile image (64 KiB)
  float2* s_tw2Q = sm2 + 16384u;                            // W_2Q^k2, k2 < Q, behind the image: read in the split walk
  for (unsigned i = threadIdx.x; i < Q; i += 512u) s_tw2Q[i] = a.tw2Q[i];
  __syncthreads();
  const float2 one = make_float2(1.f, 0.f);
  // zi_pos for this warp's row: element 32 j + r sits at s1[32 (r >> 2) + (jx ^ (r & 15))] (scatter of stage 0),
  // element j + 32 e at s2[256 e + (jx ^ (e & 15))] (gather, natural order)
  const unsigned jx = ((g & 3u) << 2) ^ (j & 15u);
  float2* s1 = sgrp + which * 16u + 256u * j;
  float2* s2 = sgrp + which * 16u + 32u * (j >> 2);
  unsigned long long zpolicy;                               // Z is written once and read once: evict first
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(zpolicy));

  float2 v[32];
  auto issue_loads = [&](unsigned tt) {
    const unsigned tile = tt % TPB, blk = tt / TPB;
    const unsigned low = tile * G + g;
    const unsigned row = which ? (low == 0 ? P / 2 : P - low) : low;
    {
      const float2* src = a.A + uint64_t(blk) * Nc + uint64_t(row) * Q + j;
#pragma unroll
      for (int e = 0; e < 32; e++) v[e] = B200_LDS1(src + 32 * e);
    }
  };

  unsigned t = blockIdx.x;
  if (t < ntiles) issue_loads(t);
  for (; t < ntiles; t += gridDim.x) {
    const unsigned tile = t % TPB, blk = t / TPB;
    const unsigned ic = (blk / a.npol) % a.nchan_in;
    // W_2N^row of this group's four pairs (+ row P/2, needed by pair 0 of tile 0 only: group 0)
    if (tl < 4) s_rowtw[4 * grp + tl] = big_twiddle<false>(a.b2lo, a.b2hi, tile * G + 4 * grp + tl);
    if (threadIdx.x == 4) s_rowtw[G] = big_twiddle<false>(a.b2lo, a.b2hi, P / 2);
    dft32<false>(v);
    // the bulk copy of the previous tile has finished reading this group's image
    if (tl == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    named_sync(1 + grp, 256);
#pragma unroll
    for (int r = 0; r < 32; r++) s1[32 * (r >> 2) + (jx ^ unsigned(r & 15))] = v[r];
    float2 ws[8], wm[4];
#pragma unroll
    for (int s1i = 1; s1i < 8; s1i++) ws[s1i] = __ldg(a.tw32 + (s1i - 1) * 32 + j);
#pragma unroll
    for (int m = 1; m < 4; m++) wm[m] = __ldg(a.tw32 + (6 + m) * 32 + j);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) v[e] = s2[256 * e + (jx ^ unsigned(e & 15))];
#pragma unroll
    for (int r = 1; r < 32; r++) {
      const int s1i = r & 7, m = r >> 3;
      const float2 w = m == 0 ? ws[s1i] : (s1i == 0 ? wm[m] : cmul(wm[m], ws[s1i]));
      v[r] = cmul(v[r], w);
    }
    dft32<false>(v);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) s2[256 * e + (jx ^ unsigned(e & 15))] = v[e];
    named_sync(1 + grp, 256);

    if (t + gridDim.x < ntiles) issue_loads(t + gridDim.x);

    {
      // 64 kk x 4 pairs per group; a half warp = 4 consecutive k2 of 4 pairs = 16 distinct bank pairs
      const unsigned gl = (tl >> 2) & 3u, q = tl & 3u, m0 = tl >> 4;     // k2 = q + 4 (m0 + 16 it)
      const unsigned g2 = 4u * grp + gl, kk = q + 4u * m0;
      const unsigned tile_off = tile * (2 * Q * G) + grp * 8192u;        // of this group's image inside the block
      const float2* Hg = a.Ht ? a.Ht + uint64_t(ic) * Nc + tile_off : nullptr;
      auto split = [&](float2 zk, float2 zmc, float2 w, float2& xk, float2& xm) {
        const float2 e = make_float2(0.5f * (zk.x + zmc.x), 0.5f * (zk.y + zmc.y));
        const float2 d = make_float2(0.5f * (zk.x - zmc.x), 0.5f * (zk.y - zmc.y));
        const float2 tt = cmul(make_float2(d.y, -d.x), w);
        xk = cadd(e, tt);
        xm = cconj(csub(e, tt));
      };
      const unsigned low = tile * G + g2;
      if (low != 0) {
        const float2 rw = s_rowtw[g2];
        const float2* t2a = s_tw2Q + kk;
        const float2* t2b = s_tw2Q + (Q - 1 - kk);
        const unsigned lq = (gl << 2) | q;
        // element k2 of rows a, b at sa, sa + 16; element Q-1-k2 (m -> 255 - m, q -> 3 - q: same swizzled low part)
        auto slots = [&](int it, unsigned& sa, unsigned& sb) {
          const unsigned m = m0 + 16u * it;
          const unsigned lo = lq ^ zi_h(m);
          sa = m * 32u + lo;
          sb = (255u - m) * 32u + lo;
        };
        // the response values of iteration it + 1 are requested before iteration it is computed (L2 latency)
        unsigned sa, sb;
        slots(0, sa, sb);
        float2 h0 = one, h1 = one, h2 = one, h3 = one;
        if (Hg) { h0 = __ldg(Hg + sa); h1 = __ldg(Hg + sb + 16); h2 = __ldg(Hg + sb); h3 = __ldg(Hg + sa + 16); }
#pragma unroll 2
        for (int it = 0; it < 8; it++) {
          unsigned san = 0, sbn = 0;
          float2 n0 = one, n1 = one, n2 = one, n3 = one;
          if (it < 7) {
            slots(it + 1, san, sbn);
            if (Hg) { n0 = __ldg(Hg + san); n1 = __ldg(Hg + sbn + 16); n2 = __ldg(Hg + sbn); n3 = __ldg(Hg + san + 16); }
          }
          const float2 ua = sgrp[sa], ub = sgrp[sa + 16];
          const float2 va2 = sgrp[sb], vb2 = sgrp[sb + 16];
          const float2 wA = cmul(rw, t2a[it * 64]);
          const float2 wB = cmul(rw, t2b[-it * 64]);
          float2 xk, xm, yk, ym;
          split(ua, make_float2(vb2.x, -vb2.y), wA, xk, xm);
          split(va2, make_float2(ub.x, -ub.y), wB, yk, ym);
          if (Hg) { xk = cmul(xk, h0); xm = cmul(xm, h1); yk = cmul(yk, h2); ym = cmul(ym, h3); }
          sgrp[sa] = xk;            // X[k],   k = low + P k2        (row low,     element k2)
          sgrp[sb + 16] = xm;       // X[N-k]                        (row P - low, element Q-1-k2)
          sgrp[sb] = yk;            // X[k'],  k' = low + P (Q-1-k2) (row low,     element Q-1-k2)
          sgrp[sa + 16] = ym;       // X[N-k']                       (row P - low, element k2)
          sa = san; sb = sbn; h0 = n0; h1 = n1; h2 = n2; h3 = n3;
        }
      } else {
        // rows 0 (array a) and P/2 (array b) mirror onto themselves (pair 0 of tile 0: group 0, gl = 0)
        const float2 rwh = s_rowtw[G];
        auto pa = [](unsigned k2) { return (k2 >> 2) * 32u + ((k2 & 3u) ^ zi_h(k2 >> 2)); };
        for (unsigned it = 0; it < 16; it++) {
          const unsigned k2 = kk + it * 64;
          if (k2 <= Q / 2) {
            const unsigned km2 = (Q - k2) % Q;
            const unsigned s0 = pa(k2), sm_ = pa(km2);
            const float2 xa = sgrp[s0], xm2 = sgrp[sm_];
            float2 xk, xm;
            split(xa, make_float2(xm2.x, -xm2.y), s_tw2Q[k2], xk, xm);
            if (Hg) { xk = cmul(xk, __ldg(Hg + s0)); xm = cmul(xm, __ldg(Hg + sm_)); }
            sgrp[s0] = xk;
            if (km2 != k2) sgrp[sm_] = xm;
          }
          if (k2 < Q / 2) {
            const unsigned s0 = pa(k2) + 16u, sm_ = pa(Q - 1 - k2) + 16u;
            const float2 xb = sgrp[s0], xm2 = sgrp[sm_];
            float2 xk, xm;
            split(xb, make_float2(xm2.x, -xm2.y), cmul(rwh, s_tw2Q[k2]), xk, xm);
            if (Hg) { xk = cmul(xk, __ldg(Hg + s0)); xm = cmul(xm, __ldg(Hg + sm_)); }
            sgrp[s0] = xk;
            sgrp[sm_] = xm;
          }
        }
      }
      // generic-proxy writes -> visible to the async proxy, then one thread hands the image to the TMA unit
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_sync(1 + grp, 256);
      if (tl == 0) {
        float2* dst = a.Z + uint64_t(blk) * Nc + tile_off;
        const unsigned src = (unsigned)__cvta_generic_to_shared(sgrp);
#pragma unroll
        for (unsigned c = 0; c < 4; c++)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                       :: "l"(dst + c * 2048u), "r"(src + c * 16384u), "r"(16384u), "l"(zpolicy) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tl == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// K3: per-channel inverse pass (both polarisations per thread) + discard + epilogue.
// tile = (CB output channels, part)
// ------------------------------------------------------------------------------------------
struct K3Args {
  const float2* Z;
  const float2* tw;        // c2 stage tables of F
  const float2* tw32;      // stage tables of the 32.16.16 plan (F = 8192): [15][32] then [15][512]
  unsigned C, Nc, nchan_in, nchan_out, npart;
  unsigned nfilt_pos, nkeep;
  uint64_t part0;
  FbSink sink;
  unsigned dbg;
  int z_tiled;             // Z is the tile image of k2_g2 (zi_pos)
};

// STATE >= 0: detection state fixed at compile time (no per-sample branches); -1: run-time state
template <int STATE>
__device__ __forceinline__ void detect4(int state, float2 p, float2 q, float* r) {
  if (STATE == B200_COHERENCE) {
    r[0] = __fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y));
    r[1] = __fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y));
    r[2] = __fadd_rn(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y));
    r[3] = __fsub_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x));
  } else {
    r[0] = r[1] = r[2] = r[3] = 0.f;
    detect_products(state, p, q, r);
  }
}

// R32 (F = 8192 only): the transform is planned 32.16.16 instead of 16.16.16.2 -- the first, twiddle-free
// stage runs as ONE 32-point butterfly per thread (threads 0-255 polarisation p, 256-511 polarisation q),
// which removes one of the three shared-memory exchanges and two of the six barriers.  The exchanges are
// 64-bit (one float2 array per polarisation, one pad slot after every 32: stride 33 is conflict free).
template <unsigned F, int EPI, int STATE, bool R32 = false>
__global__ void __launch_bounds__(512, 1) k3_c2(K3Args a) {
  static_assert(!R32 || F == 8192, "the 32.16.16 plan is built for 8192 points");
  extern __shared__ __align__(128) float4 smem4[];
  constexpr unsigned T = F / 16;
  constexpr unsigned CB = 512 / T;                     // channels per CTA
  constexpr unsigned RS = c2::pair_slots<F>();
  const unsigned cb = threadIdx.x / T, j = threadIdx.x % T;
  const unsigned tiles_per_part = a.nchan_out / CB;
  const unsigned ntiles = tiles_per_part * a.npart;
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;
  const int state = STATE >= 0 ? STATE : a.sink.state;
  const unsigned nprod = EPI == EPI_VOLT ? 1 : state_nprod(state, 2);
  const unsigned dndim = EPI == EPI_VOLT ? 1 : a.sink.dndim, dnpol = nprod / dndim;
  float4* sm = smem4 + cb * RS;

  float2 vp[16], vq[16];
  auto issue_loads = [&](unsigned tt) {
    const unsigned ch = (tt % tiles_per_part) * CB + cb, partl = tt / tiles_per_part;
    const unsigned ic = ch / a.C, csub = ch % a.C;
    const uint64_t blk = (uint64_t(partl) * a.nchan_in + ic) * 2;
    if (R32) {
      // stage-0 ownership: thread (pol, j0) holds x_pol[j0 + 256 e], e < 32, in vp[0..15], vq[0..15]
      const unsigned pol = threadIdx.x >> 8, j0 = threadIdx.x & 255u;
      if (a.z_tiled) {
        // bin f = j0 + 256 e of channel csub is element k2 = 4 csub + (e >> 3) of row r = j0 + 256 (e & 7): the four
        // k2 of a row are one aligned 32-byte group of the tile image (zi_pos) -> one 256-bit load; inside the group
        // they are permuted by XOR with hq = zi_h(csub) & 3, the same for every thread of the tile (one channel)
        const unsigned h = zi_h(csub), hq = h & 3u;
        const float2* zb = a.Z + (blk + pol) * a.Nc + 32u * csub;
        auto load_perm = [&](auto HQ) {
          constexpr unsigned hqc = decltype(HQ)::value;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const unsigned r = j0 + 256u * i;
            unsigned tile, which, g8;
            if (r < 1024u) { tile = r >> 3; which = 0; g8 = r & 7u; }
            else if (r == 1024u) { tile = 0; which = 1; g8 = 0; }
            else { const unsigned m = 2048u - r; tile = m >> 3; which = 1; g8 = m & 7u; }
            const float2* p = zb + tile * 16384u + (g8 >> 2) * 8192u + which * 16u + (((g8 & 3u) << 2) ^ (h & 12u));
            float2 x[4];
            ldg_nc_f2x4(p, x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int e = i + 8 * q;                       // element k2 = 4 csub + q sits at slot q ^ hq
              if (e < 16) vp[e] = x[q ^ hqc]; else vq[e - 16] = x[q ^ hqc];
            }
          }
        };
        switch (hq) {
          case 0: load_perm(std::integral_constant<unsigned, 0>()); break;
          case 1: load_perm(std::integral_constant<unsigned, 1>()); break;
          case 2: load_perm(std::integral_constant<unsigned, 2>()); break;
          default: load_perm(std::integral_constant<unsigned, 3>()); break;
        }
        return;
      }
      const float2* src = a.Z + (blk + pol) * a.Nc + uint64_t(csub) * F + j0;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        vp[e] = B200_LDS1(src + e * 256);
        vq[e] = B200_LDS1(src + (16 + e) * 256);
      }
      return;
    }
    const float2* srcp = a.Z + blk * a.Nc + uint64_t(csub) * F + j;
    const float2* srcq = srcp + a.Nc;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      if (B200_DBGF(a, 4)) { vp[e] = make_float2(float(threadIdx.x + e), 1.f); vq[e] = make_float2(2.f, float(e)); continue; }
      vp[e] = B200_LDS1(srcp + e * T);
      vq[e] = B200_LDS1(srcq + e * T);
    }
  };

  unsigned t = blockIdx.x;
  if (t < ntiles) issue_loads(t);
  for (; t < ntiles; t += gridDim.x) {
    const unsigned ch0 = (t % tiles_per_part) * CB, ch = ch0 + cb, partl = t / tiles_per_part;
    const uint64_t part = a.part0 + partl;
    const unsigned tn = t + gridDim.x;

    if (R32) {
      float2* arrP = reinterpret_cast<float2*>(smem4);
      float2* arrQ = arrP + (F + F / 32);
      {
        // stage 0: 32-point inverse butterfly of one polarisation, outputs r of butterfly j0 at 32 j0 + r
        const unsigned pol = threadIdx.x >> 8, j0 = threadIdx.x & 255u;
        float2 u[32];
#pragma unroll
        for (int e = 0; e < 16; e++) { u[e] = vp[e]; u[16 + e] = vq[e]; }
        dft32<true>(u);
        float2* dst = (pol ? arrQ : arrP) + 33u * j0;           // pad33(32 j0 + r) = 33 j0 + r
#pragma unroll
        for (int r = 0; r < 32; r++) dst[r] = u[r];
      }
      const unsigned k1 = j & 31u;
      float2 w[15];
#pragma unroll
      for (int r = 1; r < 16; r++) w[r - 1] = tw_get<true>(a.tw32, (unsigned)(r - 1) * 32u + k1);
      __syncthreads();
      const unsigned pj = j + (j >> 5);                           // pad33(j + 512 e) = pj + 528 e
#pragma unroll
      for (int e = 0; e < 16; e++) { vp[e] = arrP[pj + 528u * e]; vq[e] = arrQ[pj + 528u * e]; }
      // stage 1: radix 16, sub-transform length 32
#pragma unroll
      for (int r = 1; r < 16; r++) { vp[r] = cmul(vp[r], w[r - 1]); vq[r] = cmul(vq[r], w[r - 1]); }
      dft16<true>(vp);
      dft16<true>(vq);
      __syncthreads();
      {
        const unsigned base1 = (j - k1) / 2u * 33u + k1;          // pad33((j-k)*16 + k + 32 r) = base1 + 33 r
#pragma unroll
        for (int r = 0; r < 16; r++) { arrP[base1 + 33u * r] = vp[r]; arrQ[base1 + 33u * r] = vq[r]; }
      }
#pragma unroll
      for (int r = 1; r < 16; r++) w[r - 1] = tw_get<true>(a.tw32 + 15 * 32, (unsigned)(r - 1) * 512u + j);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 16; e++) { vp[e] = arrP[pj + 528u * e]; vq[e] = arrQ[pj + 528u * e]; }
      // stage 2: radix 16, sub-transform length 512: register e ends up holding element j + 512 e
#pragma unroll
      for (int r = 1; r < 16; r++) { vp[r] = cmul(vp[r], w[r - 1]); vq[r] = cmul(vq[r], w[r - 1]); }
      dft16<true>(vp);
      dft16<true>(vq);
    } else if (!(B200_DBGF(a, 1))) c2::fft_pair<F, true>(vp, vq, j, sm, a.tw, CtaSync());
    if ((B200_DBGF(a, 2)) && vp[3].x != 12345.678f) {
      __syncthreads();
      if (tn < ntiles) issue_loads(tn);
      continue;
    }

    if (EPI == EPI_VOLT) {
      float2* outp = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(ch) * 2) * a.sink.volt_span + part * a.sink.volt_step);
      float2* outq = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(ch) * 2 + 1) * a.sink.volt_span + part * a.sink.volt_step);
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const unsigned u = j + e * T - np0;             // unsigned: samples before nfilt_pos wrap to huge values
        if (u < nkeep) {
          outp[u] = vp[e];
          outq[u] = vq[e];
        }
      }
      __syncthreads();                                  // all gathers of the last stage are done
      if (tn < ntiles) issue_loads(tn);
      continue;
    }

    if (EPI == EPI_DETECT) {
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const unsigned u = j + e * T - np0;
        if (u < nkeep) {
          float r[4];
          detect4<STATE>(state, vp[e], vq[e], r);
          const uint64_t osamp = part * nkeep + u;
          for (unsigned pr = 0; pr < nprod; pr++) {
            float* out = a.sink.det + (uint64_t(ch) * dnpol + pr / dndim) * a.sink.det_span;
            out[osamp * dndim + pr % dndim] = r[pr];
          }
        }
      }
      __syncthreads();
      if (tn < ntiles) issue_loads(tn);
      continue;
    }

    // EPI_FOLD: detected products of the kept samples go back to shared memory in time order (slot pad16(t), t the
    // transform index); then every thread takes whole ITEMS of the bin plan -- at most 16 consecutive samples
    // of one phase bin inside a 16-aligned block of t, tabulated per part by k_bin_runs (fold.cu) -- sums each item
    // sequentially (the order of Fold.C:844-852) and adds it to the global PhaseSeries with one RED.ADD.F32 per
    // product.  No per-sample bin comparisons, no divergent flushes; valid for any pulse period (an item is one
    // sample long in the worst case).  (Shared-memory float atomics would compile to CAS loops.)
    const uint2* runs = a.sink.runs + partl * (uint64_t(nkeep) + 1);
    const unsigned nrun = __ldg(a.sink.nruns + partl);
    const unsigned total = CB * nrun;
    // (start, bin) and end of this thread's first two runs: requested now, needed after the detection pass
    auto run_of = [&](unsigned it, unsigned& t0, unsigned& t1, unsigned& bin) {
      const unsigned r = (CB == 1) ? it : it % nrun;
      const uint2 h = __ldg(runs + r);
      t0 = h.x; bin = h.y;
      t1 = __ldg(runs + r + 1).x;
    };
    unsigned rt0 = 0, rt1 = 0, rbin = 0, st0 = 0, st1 = 0, sbin = 0;
    if (threadIdx.x < total) run_of(threadIdx.x, rt0, rt1, rbin);
    if (threadIdx.x + 512u < total) run_of(threadIdx.x + 512u, st0, st1, sbin);
    __syncthreads();                                      // all gathers of the last stage are done
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const unsigned u = j + e * T - np0;
      if (u < nkeep) {
        float r[4];
        detect4<STATE>(state, vp[e], vq[e], r);
        sm[c2::pad16(j + e * T)] = make_float4(r[0], r[1], r[2], r[3]);   // staged by transform index: aligned lanes
      }
    }
    __syncthreads();
    // the item descriptors have long arrived.  They are looked at BEFORE the next tile's loads are queued -- a
    // sanity check of the table (an item is 1..16 samples) that also keeps the first item from waiting behind
    // those loads: the compiler tracks descriptor and spectrum loads on the same scoreboard
    if (rt1 - rt0 > 16u || st1 - st0 > 16u) asm volatile("trap;");
    // the spectra of the next tile stream in while this one is folded
    if (tn < ntiles) issue_loads(tn);
    {
      const unsigned nbin = a.sink.nbin;
      const uint64_t prof0 = uint64_t(ch0) * nbin * nprod;              // per channel [npol'][nbin][ndim']
      for (unsigned it = threadIdx.x; it < total; it += 512u) {
        const unsigned c = (CB == 1) ? 0 : it / nrun;
        if (it >= 1024u) run_of(it, rt0, rt1, rbin);
        else if (it >= 512u) { rt0 = st0; rt1 = st1; rbin = sbin; }
        // an item never crosses a 16-aligned block: its samples are consecutive slots of the padded layout
        const float4* det = smem4 + c * RS + c2::pad16(rt0 + np0);
        const unsigned n = rt1 - rt0;                       // 1..16
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = unsigned(i) < n ? det[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc[4] = {x[0].x, x[0].y, x[0].z, x[0].w};
#pragma unroll
        for (int i = 1; i < 8; i++) {
          if (unsigned(i) < n) { acc[0] += x[i].x; acc[1] += x[i].y; acc[2] += x[i].z; acc[3] += x[i].w; }
        }
        if (n > 8) {
#pragma unroll
          for (int i = 0; i < 8; i++) x[i] = unsigned(8 + i) < n ? det[8 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (unsigned(8 + i) < n) { acc[0] += x[i].x; acc[1] += x[i].y; acc[2] += x[i].z; acc[3] += x[i].w; }
          }
        }
        const uint64_t base = prof0 + uint64_t(c) * nbin * nprod;
        if (rbin < nbin)                                    // bin == nbin: samples of a flagged window (weights.cu)
          for (unsigned pr = 0; pr < nprod; pr++)
            profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, base + (uint64_t(pr / dndim) * nbin + rbin) * dndim + pr % dndim,
                        acc[pr]);
      }
    }
    __syncthreads();       // fold readers are done before the next tile's first scatter
  }
}

// Response in the tile-image order of Z: bin k = r + P k2 (row r, element k2) goes where k2_g2 leaves that bin.
__global__ void k_tile_response(const float2* __restrict__ H, float2* __restrict__ Ht, unsigned nchan_in) {
  constexpr unsigned P = 2048, Q = 1024, Nc = P * Q;
  const uint64_t n = uint64_t(nchan_in) * Nc;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    const unsigned k = unsigned(i % Nc), r = k % P, k2 = k / P;
    unsigned tile, which, g;
    if (r < P / 2) { tile = r >> 3; which = 0; g = r & 7u; }
    else if (r == P / 2) { tile = 0; which = 1; g = 0; }
    else { const unsigned m = P - r; tile = m >> 3; which = 1; g = m & 7u; }
    Ht[(i - k) + uint64_t(tile) * (Q * 16) + zi_pos(g, which, k2)] = H[i];
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <unsigned L> static int make_c2_table(float2** d_tw) {
  std::vector<float2> h(c2::twiddle_count<L>(), make_float2(1.f, 0.f));
  c2::fill_twiddles<L>(h.data());
  B200_CUDA(cudaMalloc(d_tw, sizeof(float2) * h.size()));
  B200_CUDA(cudaMemcpy(*d_tw, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
  return B200_OK;
}

template <typename K> static int opt_in_smem(K kernel, size_t dyn_bytes) {
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes));
  return B200_OK;
}

static unsigned dbg_flags(int k) {
  return (unsigned)tune_int(k == 1 ? "B200_DBG1" : k == 2 ? "B200_DBG2" : "B200_DBG3", 0);
}

static bool fast_enabled() {
  static const bool on = tune_flag("B200_FAST", true);
  return on;
}

static constexpr unsigned FP_P = 2048, FP_Q = 1024;
static constexpr int FP_NP = B200_K1_NP;
static size_t k1_smem() { return size_t(FP_NP) * (c2::pair_slots<FP_P>() + 8 / FP_NP) * sizeof(float4); }
static size_t k2_smem() { return size_t(512 / (FP_Q / 16)) * (c2::pair_slots<FP_Q>() | 1u) * sizeof(float4); }
static size_t k2r32_smem() { return size_t(16) * 1057 * sizeof(float2); }
static size_t k2g2_smem() { return size_t(17) * 1024 * sizeof(float2); }   // the tile image, dense, + W_2Q^k2
template <unsigned F> static size_t k3_smem() { return size_t(512 / (F / 16)) * c2::pair_slots<F>() * sizeof(float4); }

static bool k3_r32_enabled() {
  static const bool on = tune_flag("B200_K3_R32", true);
  return on;
}

template <unsigned F> static int k3_init(b200_fb_plan* pl) {
  constexpr unsigned CB = 512 / (F / 16);
  if (pl->nchan_out % CB != 0) return B200_OK;          // ragged channel count: generic kernels
  int rc;
  if ((rc = make_c2_table<F>(&pl->c2F)) != B200_OK) return rc;
  if (F == 8192) {
    // stage tables of the 32.16.16 plan: W_512^(r k), k < 32, then W_8192^(r k), k < 512 (forward sign)
    std::vector<float2> h(15 * 32 + 15 * 512);
    for (int r = 1; r < 16; r++) {
      for (unsigned k = 0; k < 32; k++) {
        const double ang = -2.0 * 3.14159265358979323846 * r * k / 512.0;
        h[(r - 1) * 32 + k] = make_float2(float(cos(ang)), float(sin(ang)));
      }
      for (unsigned k = 0; k < 512; k++) {
        const double ang = -2.0 * 3.14159265358979323846 * r * k / 8192.0;
        h[15 * 32 + (r - 1) * 512 + k] = make_float2(float(cos(ang)), float(sin(ang)));
      }
    }
    B200_CUDA(cudaMalloc(&pl->c2F32, sizeof(float2) * h.size()));
    B200_CUDA(cudaMemcpy(pl->c2F32, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    if ((rc = opt_in_smem(k3_c2<8192, EPI_VOLT, -1, true>, k3_smem<8192>())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k3_c2<8192, EPI_DETECT, -1, true>, k3_smem<8192>())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k3_c2<8192, EPI_FOLD, -1, true>, k3_smem<8192>())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k3_c2<8192, EPI_FOLD, B200_COHERENCE, true>, k3_smem<8192>())) != B200_OK) return rc;
  }
  if ((rc = opt_in_smem(k3_c2<F, EPI_VOLT, -1>, k3_smem<F>())) != B200_OK) return rc;
  if ((rc = opt_in_smem(k3_c2<F, EPI_DETECT, -1>, k3_smem<F>())) != B200_OK) return rc;
  if ((rc = opt_in_smem(k3_c2<F, EPI_FOLD, -1>, k3_smem<F>())) != B200_OK) return rc;
  if ((rc = opt_in_smem(k3_c2<F, EPI_FOLD, B200_COHERENCE>, k3_smem<F>())) != B200_OK) return rc;
  pl->fast_k3 = true;
  return B200_OK;
}

template <unsigned F> static void k3_launch(b200_fb_plan* pl, const K3Args& a, const FbSink& sk, unsigned nb) {
  Context* ctx = pl->ctx;
  constexpr unsigned CB = 512 / (F / 16);
  const unsigned ntiles = pl->nchan_out / CB * nb;
  dim3 grid(ntiles < (unsigned)ctx->sm_count ? ntiles : (unsigned)ctx->sm_count);
  if constexpr (F == 8192) {
    if (k3_r32_enabled() && pl->c2F32) {
      if (sk.kind == EPI_VOLT) k3_c2<8192, EPI_VOLT, -1, true><<<grid, 512, k3_smem<8192>(), ctx->stream>>>(a);
      else if (sk.kind == EPI_DETECT) k3_c2<8192, EPI_DETECT, -1, true><<<grid, 512, k3_smem<8192>(), ctx->stream>>>(a);
      else if (sk.state == B200_COHERENCE) k3_c2<8192, EPI_FOLD, B200_COHERENCE, true><<<grid, 512, k3_smem<8192>(), ctx->stream>>>(a);
      else k3_c2<8192, EPI_FOLD, -1, true><<<grid, 512, k3_smem<8192>(), ctx->stream>>>(a);
      return;
    }
  }
  if (sk.kind == EPI_VOLT) k3_c2<F, EPI_VOLT, -1><<<grid, 512, k3_smem<F>(), ctx->stream>>>(a);
  else if (sk.kind == EPI_DETECT) k3_c2<F, EPI_DETECT, -1><<<grid, 512, k3_smem<F>(), ctx->stream>>>(a);
  else if (sk.state == B200_COHERENCE) k3_c2<F, EPI_FOLD, B200_COHERENCE><<<grid, 512, k3_smem<F>(), ctx->stream>>>(a);
  else k3_c2<F, EPI_FOLD, -1><<<grid, 512, k3_smem<F>(), ctx->stream>>>(a);
}

// Z travels from K2 to K3 as the image of K2's shared memory (zi_pos) when both ends are the kernels that implement
// it: k2_g2 (real-input split) and the 32.16.16 K3, i.e. P = 2048, Q = 1024, freq_res = 8192 (cfg1).
static bool z_tiled(const b200_fb_plan* pl) {
  static const bool want = tune_flag("B200_Z_TILED", true);
  static const bool k2r32 = tune_flag("B200_K2_R32", true);
  static const bool g2 = tune_flag("B200_K2_G2", true);
  return want && k2r32 && g2 && k3_r32_enabled() && pl->fast_k2 && pl->fast_k3 && pl->c2Q32 && pl->c2F32 &&
         pl->desc.input_real && pl->P == 2048 && pl->Q == 1024 && pl->F == 8192;
}


int fast_plan_init(b200_fb_plan* pl) {
  pl->fast_k1 = pl->fast_k2 = pl->fast_k3 = false;
  pl->c2P = pl->c2Q = pl->c2F = nullptr;
  pl->c2F32 = nullptr;
  pl->c2Q32 = nullptr;
  pl->d_response_tiled = nullptr;
  if (!fast_enabled() || pl->conv_path) return B200_OK;
  int rc = B200_OK;
  if (pl->P == FP_P && pl->Q >= 2 * FP_NP && pl->Q % (2 * FP_NP) == 0) {
    if ((rc = make_c2_table<FP_P>(&pl->c2P)) != B200_OK) return rc;
    if ((rc = opt_in_smem(k1_c2<SRC_F32, FP_P, FP_NP>, k1_smem())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k1_c2<SRC_CASPSR8, FP_P, FP_NP>, k1_smem())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k1_c2<SRC_F32, FP_P, FP_NP, FP_Q>, k1_smem())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k1_c2<SRC_CASPSR8, FP_P, FP_NP, FP_Q>, k1_smem())) != B200_OK) return rc;
    pl->fast_k1 = true;
  }
  if (pl->Q == FP_Q && pl->P == FP_P) {
    if ((rc = make_c2_table<FP_Q>(&pl->c2Q)) != B200_OK) return rc;
    if ((rc = opt_in_smem(k2_c2<FP_P, FP_Q, true>, k2_smem())) != B200_OK) return rc;
    if ((rc = opt_in_smem(k2_c2<FP_P, FP_Q, false>, k2_smem())) != B200_OK) return rc;
    {
      std::vector<float2> h(10 * 32);
      for (unsigned jj = 0; jj < 32; jj++) {
        for (int s1 = 1; s1 < 8; s1++) {
          const double ang = -2.0 * 3.14159265358979323846 * s1 * jj / 1024.0;
          h[(s1 - 1) * 32 + jj] = make_float2(float(cos(ang)), float(sin(ang)));
        }
        for (int m = 1; m < 4; m++) {
          const double ang = -2.0 * 3.14159265358979323846 * 8 * m * jj / 1024.0;
          h[(6 + m) * 32 + jj] = make_float2(float(cos(ang)), float(sin(ang)));
        }
      }
      B200_CUDA(cudaMalloc(&pl->c2Q32, sizeof(float2) * h.size()));
      B200_CUDA(cudaMemcpy(pl->c2Q32, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
      if ((rc = opt_in_smem(k2_r32<FP_P, true>, k2r32_smem())) != B200_OK) return rc;
      if ((rc = opt_in_smem(k2_r32<FP_P, false>, k2r32_smem())) != B200_OK) return rc;
      if ((rc = opt_in_smem(k2_g2<FP_P>, k2g2_smem())) != B200_OK) return rc;
    }
    pl->fast_k2 = true;
  }
  // K3: any per-channel transform length the c2 core is planned for (CB = 8192/F channels per CTA)
  if (pl->desc.npol == 2) {
    switch (pl->F) {
      case 8192: rc = k3_init<8192>(pl); break;
      case 4096: rc = k3_init<4096>(pl); break;
      case 2048: rc = k3_init<2048>(pl); break;
      case 1024: rc = k3_init<1024>(pl); break;
      case 512: rc = k3_init<512>(pl); break;
      case 256: rc = k3_init<256>(pl); break;
      default: break;
    }
    if (rc != B200_OK) return rc;
  }
  if (z_tiled(pl) && pl->d_response) {
    const uint64_t n = uint64_t(pl->desc.input_nchan) * pl->Nc;
    B200_CUDA(cudaMalloc(&pl->d_response_tiled, n * sizeof(float2)));
    k_tile_response<<<1024, 256, 0, pl->ctx->stream>>>(pl->d_response, pl->d_response_tiled, pl->desc.input_nchan);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  }
  return B200_OK;
}

void fast_plan_free(b200_fb_plan* pl) {
  if (pl->d_response_tiled) cudaFree(pl->d_response_tiled);
  pl->d_response_tiled = nullptr;
  if (pl->c2P) cudaFree(pl->c2P);
  if (pl->c2Q) cudaFree(pl->c2Q);
  if (pl->c2F) cudaFree(pl->c2F);
  if (pl->c2F32) cudaFree(pl->c2F32);
  if (pl->c2Q32) cudaFree(pl->c2Q32);
  pl->c2Q32 = nullptr;
  pl->c2P = pl->c2Q = pl->c2F = nullptr;
  pl->c2F32 = nullptr;
}

static unsigned persistent_grid(Context* ctx, unsigned ntiles) {
  return ntiles < (unsigned)ctx->sm_count ? ntiles : (unsigned)ctx->sm_count;
}

int fast_k1(b200_fb_plan* pl, const FbSource& src, uint64_t part0, unsigned nb) {
  Context* ctx = pl->ctx;
  K1Args a;
  a.src = src.ptr; a.span = src.span; a.step = src.step; a.lut = src.d_lut;
  a.dst = pl->scratchA; a.tw = pl->c2P; a.blo = pl->bigN.lo; a.bhi = pl->bigN.hi;
  a.Q = pl->Q; a.npol = pl->desc.npol; a.nchan_in = pl->desc.input_nchan; a.Nc = pl->Nc; a.part0 = part0;
  a.nblk = nb * pl->desc.input_nchan * pl->desc.npol;
  a.dbg = dbg_flags(1);
  a.conv_ok = src.kind == SRC_CASPSR8 ? src.conv_ok : 0; a.conv_hi = src.conv_hi; a.conv_lo = src.conv_lo;
  const unsigned ntiles = pl->Q / (2 * FP_NP) * a.nblk;
  const unsigned cta_per_sm = 512 / (FP_NP * (FP_P / 16));
  a.nsm = (unsigned)ctx->sm_count;
  a.skew_ns = cta_per_sm > 1 ? (unsigned)tune_int("B200_K1_SKEW", 3000) : 0u;
  dim3 grid(std::min(ntiles, cta_per_sm * (unsigned)ctx->sm_count));
  dim3 block(FP_NP * (FP_P / 16));
  static const bool l2pf = tune_flag("B200_K1_L2PF", true);
  a.l2_prefetch = l2pf ? 1 : 0;
  a.overlap = pl->nsamp_overlap;
  LaunchScope ls(ctx, KC_COLS_FWD);
  {
    if (pl->Q == FP_Q) {
      if (src.kind == SRC_F32) k1_c2<SRC_F32, FP_P, FP_NP, FP_Q><<<grid, block, k1_smem(), ctx->stream>>>(a);
      else k1_c2<SRC_CASPSR8, FP_P, FP_NP, FP_Q><<<grid, block, k1_smem(), ctx->stream>>>(a);
    } else if (src.kind == SRC_F32) k1_c2<SRC_F32, FP_P, FP_NP><<<grid, block, k1_smem(), ctx->stream>>>(a);
    else k1_c2<SRC_CASPSR8, FP_P, FP_NP><<<grid, block, k1_smem(), ctx->stream>>>(a);
  }
  return B200_OK;
}

int fast_k2(b200_fb_plan* pl, unsigned nb) {
  Context* ctx = pl->ctx;
  K2Args a;
  a.A = pl->scratchA; a.Z = pl->scratchZ; a.H = pl->d_response; a.Ht = pl->d_response_tiled; a.tw = pl->c2Q; a.tw2Q = pl->tw2Q.tw;
  a.b2lo = pl->big2N.lo; a.b2hi = pl->big2N.hi;
  a.Nc = pl->Nc; a.npol = pl->desc.npol; a.nchan_in = pl->desc.input_nchan;
  a.nblk = nb * pl->desc.input_nchan * pl->desc.npol;
  a.dbg = dbg_flags(2);
  constexpr unsigned G = 512 / (FP_Q / 16);
  const bool split = pl->desc.input_real;
  const unsigned ntiles = (split ? (FP_P / 2) / G : FP_P / (2 * G)) * a.nblk;
  dim3 grid(persistent_grid(ctx, ntiles));
  LaunchScope ls(ctx, KC_ROWS);
  static const bool r32 = tune_flag("B200_K2_R32", true);
  a.tw32 = pl->c2Q32;
  a.z_tiled = z_tiled(pl) ? 1 : 0;
#ifdef B200_ABLATION
  if (tune_flag("B200_K2_NOH", false)) a.H = nullptr;   // timing experiment only (cost of the response stream): results are wrong
#endif
  if (r32 && pl->c2Q32 && split && a.z_tiled) {
    k2_g2<FP_P><<<grid, 512, k2g2_smem(), ctx->stream>>>(a);
    return B200_OK;
  }
  if (r32 && pl->c2Q32) {
    if (split) k2_r32<FP_P, true><<<grid, 512, k2r32_smem(), ctx->stream>>>(a);
    else k2_r32<FP_P, false><<<grid, 512, k2r32_smem(), ctx->stream>>>(a);
    return B200_OK;
  }
  if (split) k2_c2<FP_P, FP_Q, true><<<grid, 512, k2_smem(), ctx->stream>>>(a);
  else k2_c2<FP_P, FP_Q, false><<<grid, 512, k2_smem(), ctx->stream>>>(a);
  return B200_OK;
}

int fast_k3(b200_fb_plan* pl, const FbSink& sk, uint64_t part0, unsigned nb) {
  Context* ctx = pl->ctx;
  K3Args a;
  a.Z = pl->scratchZ; a.tw = pl->c2F; a.tw32 = pl->c2F32; a.C = pl->C; a.Nc = pl->Nc; a.nchan_in = pl->desc.input_nchan;
  a.nchan_out = pl->nchan_out; a.npart = nb;
  a.nfilt_pos = pl->desc.nfilt_pos; a.nkeep = pl->nkeep; a.part0 = part0; a.sink = sk;
  a.dbg = dbg_flags(3);
  a.z_tiled = z_tiled(pl) ? 1 : 0;
  B200_REQUIRE(sk.kind != EPI_FOLD || (sk.runs && sk.nruns), "fast_k3: the fold epilogue needs the run table (fold_build_runs)");
  LaunchScope ls(ctx, KC_INV);
  switch (pl->F) {
    case 8192: k3_launch<8192>(pl, a, sk, nb); break;
    case 4096: k3_launch<4096>(pl, a, sk, nb); break;
    case 2048: k3_launch<2048>(pl, a, sk, nb); break;
    case 1024: k3_launch<1024>(pl, a, sk, nb); break;
    case 512: k3_launch<512>(pl, a, sk, nb); break;
    default: k3_launch<256>(pl, a, sk, nb); break;
  }
  return B200_OK;
}

}  // namespace b200
