// filterbank.cu -- overlap-save coherent filterbank / convolution engine (sm_100a).
//
// Replaces the inner loops of dsp::Filterbank::filterbank (Signal/General/Filterbank.C:563-660)
// and dsp::Convolution::transformation (Convolution.C:389-458) -- forward FFT, multiply by the
// dsp::Response, inverse FFT per output channel, discard the wrap-around -- with three kernels
// around ONE spectrum round trip each:
//
//   K1 k_cols_fwd   raw bytes | float samples -> P-point column FFTs (n = Q*n1 + n2), times
//                   W_N^(n2*k1)                                   -> A[k1][n2]
//   K2 k_rows       Q-point row FFTs (-> k = k1 + P*k2), real-input split (mirror rows are
//                   co-resident), response multiply               -> Z[k] natural order
//                   (convolution path: + inverse Q-point row FFT, twiddle, in place)
//   K3 k_chan_inv   per output channel: inverse F-point FFT of both polarisations in shared
//                   memory, discard nfilt_pos/nfilt_neg, then voltages | detect | detect+fold
//      k_cols_inv   (convolution path, F = N > 8192) inverse P-point column FFTs + same epilogues
//
// No cuFFT.  FFTs are hand-written radix-16/8/4/2 Stockham stages on registers (fft_core.cuh)
// exchanged through shared memory.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.cuh"

namespace b200 {

__device__ __forceinline__ float2 ld_nc_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// Block FFT of compile-time size NCT (fully unrolled stages, EPT points per thread) or, when
// NCT == 0, of run-time size N (generic path).
template <int EPT, bool INV, unsigned NCT, typename Map, typename LoadF, typename StoreF>
__device__ __forceinline__ void fft_any(unsigned N, unsigned j, unsigned T, const Map& map, float2* smem,
                                        const float2* __restrict__ tw, const float2* __restrict__ stw,
                                        LoadF load, StoreF store) {
  if constexpr (NCT != 0) block_fft_ct<EPT, INV, NCT>(j, map, smem, stw, load, store);
  else block_fft<EPT, INV>(N, j, T, map, smem, tw, N, load, store);
}
// log2 of the first radix = swizzle shift of the shared-memory maps
template <int EPT> struct SwzShift { static constexpr unsigned value = EPT == 32 ? 5 : 4; };

// ------------------------------------------------------------------------------------------
// K1: forward column pass
// ------------------------------------------------------------------------------------------
struct ColsArgs {
  const void* src;
  uint64_t span, step;
  const float* lut;
  float2* dst;
  const float2* twP;
  const float2* twPs;
  const float2* blo;
  const float2* bhi;
  unsigned P, Q, lb, npol, nchan_in, Nc;
  uint64_t part0;
  uint64_t first;           // raw formats: first sample of part 0 within the stream at src
  float scale;
  unsigned sample_swap, ndim;
  const float2* H;          // Q == 1, complex input: the column pass IS the whole transform -- the response is
                            // applied here and the result goes straight to Z (no row pass)
  const float2* win;        // SRC_TWOBIT: (lo, hi) of every 512-sample window and digitizer
  unsigned lowsel, negsel;
};

template <int SRC, int EPT, unsigned PCT>
__global__ void __launch_bounds__((PCT && EPT == 32) ? 512 : 1024, 1) k_cols_fwd(ColsArgs a) {
  extern __shared__ float2 smem[];
  __shared__ float s_lut[256];
  __shared__ float2 s_h[32 * 16];   // W_N^(n2*T*e): output twiddle factors shared by the tile
  const unsigned P = PCT ? PCT : a.P;
  const unsigned B = 1u << a.lb;
  const unsigned b = threadIdx.x & (B - 1);
  const unsigned j = threadIdx.x >> a.lb;
  const unsigned T = P / EPT;
  const unsigned n2 = blockIdx.x * B + b;
  const unsigned blk = blockIdx.y;
  const unsigned pol = blk % a.npol;
  const unsigned ic = (blk / a.npol) % a.nchan_in;
  const uint64_t part = a.part0 + blk / (a.npol * a.nchan_in);
  // (a per-bank replicated table -- conflict-free gathers -- was measured slower: 0.71 vs 0.67 ms)
  if (SRC == SRC_CASPSR8 || SRC == SRC_GENERIC8)
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = a.lut[i];
  // Output twiddle W_N^(n2*k1) with k1 = j + e*T factorises into W_N^(n2*j) (one per thread) times
  // W_N^(n2*T*e) (EPT x B values per tile, staged here once): two table look-ups per thread
  // instead of two scattered ones per element.
  if (a.Q > 1)
    for (unsigned i = threadIdx.x; i < EPT * B; i += blockDim.x) {
      const unsigned e = i >> a.lb, bb = i & (B - 1);
      const unsigned m = ((blockIdx.x * B + bb) * T * e) & (a.Nc - 1);
      s_h[i] = big_twiddle<false>(a.blo, a.bhi, m);
    }
  __syncthreads();
  const float2* fsrc = nullptr;
  const unsigned char* raw = nullptr;
  uint64_t samp0 = 0;
  if (SRC == SRC_F32)
    fsrc = reinterpret_cast<const float2*>(static_cast<const float*>(a.src) + (uint64_t(ic) * a.npol + pol) * a.span +
                                           part * a.step) + n2;
  else {
    raw = static_cast<const unsigned char*>(a.src);
    // element n = Q*n1 + n2 of the part is the sample pair (2n, 2n+1) of a real stream, sample n of a complex one
    const bool real_in = (SRC == SRC_CASPSR8) || (SRC == SRC_TWOBIT) || (SRC == SRC_GENERIC8 && a.ndim == 1);
    samp0 = a.first + part * a.step + (real_in ? 2ull * n2 : uint64_t(n2));
  }
  MapCols map{b, a.lb, SwzShift<EPT>::value};
  auto load = [&](unsigned n1) -> float2 {
    if (SRC == SRC_F32) {
      return ld_nc_f2(fsrc + uint64_t(n1) * a.Q);
    } else if (SRC == SRC_MEERKAT8) {
      // heaps of 256 samples, [heap][pol][chan][256 x (re, im) int8] (MeerKATUnpacker.C:196-229); MKBFRo exchanges
      // odd and even samples: output sample i comes from input sample i ^ 1
      uint64_t i = samp0 + uint64_t(n1) * a.Q;
      if (a.sample_swap == 2) i ^= 1ull;
      const uint64_t word = (((i >> 8) * a.npol + pol) * a.nchan_in + ic) * 256ull + (i & 255ull);
      const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(raw) + word);
      return make_float2(__fmul_rn(float(int(int8_t(w & 255u))) + 0.5f, a.scale),
                         __fmul_rn(float(int(int8_t(w >> 8))) + 0.5f, a.scale));
    } else if (SRC == SRC_UWB16) {
      // blocks of 2048 complex int16 samples per polarisation, offset binary (UWBUnpacker.C:177-218)
      const uint64_t i = samp0 + uint64_t(n1) * a.Q;
      const uint64_t word = ((i >> 11) * a.npol + pol) * 2048ull + (i & 2047ull);
      const unsigned w = __ldg(reinterpret_cast<const unsigned*>(raw) + word);
      return make_float2(float(short((w & 0xffffu) ^ 0x8000u)), float(short((w >> 16) ^ 0x8000u)));
    } else if (SRC == SRC_TWOBIT) {
      // CPSR2 convention: four samples per byte, most significant first, digitizers interleaved byte by byte
      // (ExcisionUnpacker.C:258-266, BitTable.C:154-163); samples 2n, 2n+1 share a byte; the output levels of the
      // 512-sample window come from k_twobit_windows (0, 0 for an excised window) -- bit-identical to b200_unpack_twobit
      const uint64_t i0 = samp0 + 2ull * uint64_t(n1) * a.Q;
      const unsigned byte = __ldg(raw + (i0 >> 2) * a.npol + pol);
      const float2 lv = __ldg(a.win + (i0 >> 9) * a.npol + pol);
      const unsigned sh = 6u - 2u * unsigned(i0 & 3ull);
      const unsigned c0 = (byte >> sh) & 3u, c1 = (byte >> (sh - 2u)) & 3u;
      const float m0 = ((a.lowsel >> c0) & 1u) ? lv.x : lv.y, m1 = ((a.lowsel >> c1) & 1u) ? lv.x : lv.y;
      return make_float2(((a.negsel >> c0) & 1u) ? -m0 : m0, ((a.negsel >> c1) & 1u) ? -m1 : m1);
    } else if (SRC == SRC_GENERIC8) {
      // TFP bytes: i*(nchan*npol*ndim) + ndim*(npol*c + p) + d (BitUnpacker.C:56-75)
      if (a.ndim == 2) {
        const uint64_t i = samp0 + uint64_t(n1) * a.Q;
        const uint64_t off = i * (uint64_t(a.nchan_in) * a.npol * 2u) + 2u * (a.npol * ic + pol);
        const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(raw + off));
        return make_float2(s_lut[w & 255u], s_lut[w >> 8]);
      }
      const uint64_t i0 = samp0 + 2ull * uint64_t(n1) * a.Q;
      const uint64_t stride = uint64_t(a.nchan_in) * a.npol;
      const uint64_t off = i0 * stride + (a.npol * ic + pol);
      return make_float2(s_lut[__ldg(raw + off)], s_lut[__ldg(raw + off + stride)]);
    } else {
      // CASPSR: byte 8*(i/4) + 4*pol + i%4 (CASPSRUnpacker.C:141-187); samples 2n, 2n+1 are adjacent
      uint64_t i0 = samp0 + 2ull * uint64_t(n1) * a.Q;
      uint64_t off = 8ull * (i0 >> 2) + 4u * pol + (i0 & 3ull);
      unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(raw + off));
      return make_float2(s_lut[w & 255u], s_lut[w >> 8]);
    }
  };
  float2* dst = a.dst + uint64_t(blk) * a.Nc + n2;
  const float2 wbase = a.Q > 1 ? big_twiddle<false>(a.blo, a.bhi, n2 * j) : make_float2(1.f, 0.f);
  const float2* Hc = a.H ? a.H + uint64_t(ic) * a.Nc : nullptr;
  auto store = [&](unsigned k1, float2 v, int e) {
    if (a.Q > 1) v = cmul(cmul(v, wbase), s_h[(e << a.lb) + b]);
    else if (Hc) v = cmul(v, __ldg(Hc + k1));
    dst[uint64_t(k1) * a.Q] = v;
  };
  fft_any<EPT, false, PCT>(P, j, T, map, smem, a.twP, a.twPs, load, store);
}

// ------------------------------------------------------------------------------------------
// K2: row pass (+ real split, response multiply; convolution path: + inverse row pass)
// ------------------------------------------------------------------------------------------
struct RowsArgs {
  float2* A;
  float2* Z;
  const float2* H;
  const float2* twQ;
  const float2* twQs;
  const float2* tw2Q;      // exp(-2 pi i m / (2Q)), m < 2Q: column factor of the real-split twiddle
  const float2* blo;
  const float2* bhi;
  const float2* b2lo;
  const float2* b2hi;
  unsigned P, Q, G, Nc, npol, nchan_in;
};

template <bool SPLIT, bool CONV, int EPT, unsigned QCT>
__global__ void __launch_bounds__(QCT ? 512 : 1024, 1) k_rows(RowsArgs a) {
  extern __shared__ float2 smem[];
  __shared__ float2 s_rowtw[32];   // W_2N^(row) of every slot: row factor of the real-split twiddle
  __shared__ float2 s_wt[512];     // convolution path: W_N^-(row*T*e) of every slot, the part of the output
                                   // twiddle W_N^-(row*m2), m2 = j + T*e, that does not depend on the thread
  const unsigned Q = QCT ? QCT : a.Q;
  const unsigned T = EPT ? Q / (EPT ? EPT : 1) : 1;
  constexpr unsigned SH = SwzShift<EPT>::value;
  const unsigned slot = threadIdx.x / T;
  const unsigned j = threadIdx.x % T;
  const unsigned G = a.G;
  const unsigned tile = blockIdx.x;
  const unsigned blk = blockIdx.y;
  const unsigned ic = (blk / a.npol) % a.nchan_in;
  const unsigned P = a.P, Nc = a.Nc;

  auto slot_row = [&](unsigned s) -> unsigned {
    unsigned g = s % G;
    unsigned low = tile * G + g;
    if (!SPLIT || s < G) return low;
    return (low == 0) ? P / 2 : P - low;
  };
  auto smap = [&](unsigned s, unsigned idx) -> unsigned {
    if (!EPT) return s;
    return s * Q + ((idx ^ ((idx >> SH) & 15u)) ^ (((s % G) * 2u) & 15u));
  };

  const unsigned row = slot_row(slot);
  float2* Ablk = a.A + uint64_t(blk) * Nc;
  if (SPLIT && threadIdx.x < 2 * G) {
    // W_2N^k with k = row + P*k2 factorises into W_2N^row (here) times W_2Q^k2 (table tw2Q)
    unsigned r = slot_row(threadIdx.x);
    if (threadIdx.x == 0 && tile == 0) r = 0;
    s_rowtw[threadIdx.x] = big_twiddle<false>(a.b2lo, a.b2hi, r);
  }

  const unsigned nslots = blockDim.x / T;
  const bool wt_fact = CONV && EPT && nslots * (EPT ? EPT : 1) <= 512;
  float2 wj = make_float2(1.f, 0.f);
  if (wt_fact) {
    for (unsigned i = threadIdx.x; i < nslots * (EPT ? EPT : 1); i += blockDim.x) {
      const unsigned s = i / (EPT ? EPT : 1), e = i % (EPT ? EPT : 1);
      s_wt[i] = big_twiddle<true>(a.blo, a.bhi, slot_row(s) * T * e);
    }
    wj = big_twiddle<true>(a.blo, a.bhi, row * j);
  }

  // ---- phase 1: forward row FFT into shared memory (natural order) ----
  if (EPT) {
    MapRows map{slot * Q, SH, ((slot % G) * 2u) & 15u};
    const float2* src = Ablk + uint64_t(row) * Q;
    auto load = [&](unsigned idx) -> float2 { return src[idx]; };
    auto store = [&](unsigned idx, float2 v, int) { smem[map(idx)] = v; };
    fft_any<(EPT ? EPT : 2), false, QCT>(Q, j, T, map, smem, a.twQ, a.twQs, load, store);
  } else {
    smem[slot] = Ablk[row];
  }
  __syncthreads();

  // ---- phase 2: split / response ----
  const float2* H = a.H ? a.H + uint64_t(ic) * Nc : nullptr;
  float2* Zblk = CONV ? nullptr : a.Z + uint64_t(blk) * Nc;

  if (SPLIT) {
    // element (slot sA, column iA) is bin k = row(sA) + P*iA; its mirror N-k sits at (sB, iB)
    auto do_pair = [&](unsigned sA, unsigned iA, unsigned sB, unsigned iB, unsigned k) {
      const unsigned pa = smap(sA, iA), pb = smap(sB, iB);
      const bool same = (pa == pb);
      float2 zk = smem[pa], zm = cconj(smem[pb]);
      float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
      float2 d = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
      float2 o = make_float2(d.y, -d.x);   // -i * d
      float2 t = cmul(o, cmul(s_rowtw[sA], __ldg(a.tw2Q + iA)));
      float2 xk = cadd(e, t);
      float2 xm = cconj(csub(e, t));
      if (H) {
        xk = cmul(xk, __ldg(H + k));
        if (!same) xm = cmul(xm, __ldg(H + (Nc - k)));
      }
      if (CONV) {
        smem[pa] = xk;
        if (!same) smem[pb] = xm;
      } else {
        Zblk[k] = xk;
        if (!same) Zblk[Nc - k] = xm;
      }
    };
    // jobs are processed four at a time per thread so that the response / twiddle loads of all
    // four are in flight together (they come from L2 / HBM and nothing else hides their latency)
    const unsigned njobs = G * Q;
    constexpr int U = 4;
    for (unsigned it0 = threadIdx.x; it0 < njobs; it0 += blockDim.x * U) {
      unsigned sA[U], iA[U], sB[U], iB[U], kb[U];
      float2 hk[U], hm[U], tq[U];
      bool valid[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned it = it0 + u * blockDim.x;
        valid[u] = it < njobs;
        const unsigned itc = valid[u] ? it : threadIdx.x;
        const unsigned g = itc % G, k2 = itc / G;
        const unsigned low = tile * G + g;
        if (low != 0) {
          sA[u] = g; iA[u] = k2; sB[u] = G + g; iB[u] = Q - 1 - k2; kb[u] = low + P * k2;
        } else if (k2 <= Q / 2) {                                      // row 0 pairs with itself
          sA[u] = 0; iA[u] = k2; sB[u] = 0; iB[u] = (Q - k2) % Q; kb[u] = P * k2;
        } else {                                                       // row P/2 pairs with itself
          const unsigned kk = k2 - Q / 2 - 1;
          sA[u] = G; iA[u] = kk; sB[u] = G; iB[u] = Q - 1 - kk; kb[u] = P / 2 + P * kk;
        }
        tq[u] = __ldg(a.tw2Q + iA[u]);
        if (H) {
          hk[u] = __ldg(H + kb[u]);
          hm[u] = __ldg(H + ((Nc - kb[u]) & (Nc - 1)));
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!valid[u]) continue;
        const unsigned pa = smap(sA[u], iA[u]), pb = smap(sB[u], iB[u]);
        const bool same = (pa == pb);
        float2 zk = smem[pa], zm = cconj(smem[pb]);
        float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
        float2 d = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
        float2 o = make_float2(d.y, -d.x);   // -i * d
        float2 t = cmul(o, cmul(s_rowtw[sA[u]], tq[u]));
        float2 xk = cadd(e, t);
        float2 xm = cconj(csub(e, t));
        if (H) {
          xk = cmul(xk, hk[u]);
          xm = cmul(xm, hm[u]);
        }
        if (CONV) {
          smem[pa] = xk;
          if (!same) smem[pb] = xm;
        } else {
          Zblk[kb[u]] = xk;
          if (!same) Zblk[Nc - kb[u]] = xm;
        }
      }
    }
    if (tile == 0 && threadIdx.x == 0) {
      const unsigned kk = (Q / 2 ? Q / 2 : 1) - 1;                     // the one row-P/2 job not covered above
      do_pair(G, kk, G, Q - 1 - kk, P / 2 + P * kk);
    }
  } else {
    const unsigned njobs = G * Q;
    for (unsigned it = threadIdx.x; it < njobs; it += blockDim.x) {
      const unsigned g = it % G, k2 = it / G;
      const unsigned k = tile * G + g + P * k2;
      const unsigned pa = smap(g, k2);
      float2 x = smem[pa];
      if (H) x = cmul(x, __ldg(H + k));
      if (CONV) smem[pa] = x;
      else Zblk[k] = x;
    }
  }

  // ---- phase 3 (convolution path): inverse row FFT, twiddle, store in place ----
  if (CONV) {
    __syncthreads();
    float2* dst = Ablk + uint64_t(row) * Q;
    if (EPT) {
      MapRows map{slot * Q, SH, ((slot % G) * 2u) & 15u};
      auto load = [&](unsigned idx) -> float2 { return smem[map(idx)]; };
      const float2* wt = s_wt + slot * (EPT ? EPT : 1);
      auto store = [&](unsigned m2, float2 v, int e) {
        dst[m2] = wt_fact ? cmul(cmul(v, wj), wt[e]) : cmul(v, big_twiddle<true>(a.blo, a.bhi, row * m2));
      };
      fft_any<(EPT ? EPT : 2), true, QCT>(Q, j, T, map, smem, a.twQ, a.twQs, load, store);
    } else {
      dst[0] = smem[slot];
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3: per-channel inverse FFT + discard + epilogue
// ------------------------------------------------------------------------------------------
struct ChanArgs {
  const float2* Z;
  const float2* twF;
  const float2* twFs;
  unsigned F, C, Nc, npol, nchan_in, CB, npol_cta;
  unsigned nfilt_pos, nkeep;
  uint64_t part0;
  FbSink sink;
};

// K3 for short per-channel transforms (freq_res = 2 ... 16: cfg2's -F 4096:D has 8), voltage and detected-series sinks:
// ONE THREAD per output channel does the inverse transform of both polarisations in registers and walks PG consecutive
// parts.  What a warp reads per part is one contiguous stretch of the spectrum (32 channels x F bins); what the CTA
// writes is staged in shared memory as one row per (channel, output plane) -- PG parts = PG * nkeep consecutive
// samples -- and leaves as coalesced runs (scattered 4-byte stores cost eight times the L2 sector writes).
// Output planes: voltages (channel, pol) of 2 floats per sample; detected (channel, pr / dndim) of dndim floats.
template <unsigned F, int EPI>
__global__ void __launch_bounds__(128) k_chan_inv_small(ChanArgs a, unsigned npart, unsigned PG) {
  extern __shared__ float stage[];
  const unsigned cl = threadIdx.x, ch0 = blockIdx.x * 128u, ch = ch0 + cl;
  const unsigned nch = a.nchan_in * a.C;
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;
  const int state = a.sink.state;
  const unsigned nprod = EPI == EPI_VOLT ? 0 : state_nprod(state, a.npol);
  const unsigned edim = EPI == EPI_VOLT ? 2u : a.sink.dndim;               // floats per sample of a plane
  const unsigned nplane = EPI == EPI_VOLT ? a.npol : nprod / edim;         // planes per channel
  const unsigned RL = PG * nkeep * edim, RS = RL + 1u;                     // row length, odd-ish stride: no bank conflicts
  const unsigned p_begin = blockIdx.y * PG, p_end = min(npart, p_begin + PG);
  if (ch < nch) {
    const unsigned ic = ch / a.C, csub = ch % a.C;
    for (unsigned partl = p_begin; partl < p_end; partl++) {
      float2 v[2][F];
#pragma unroll
      for (unsigned pol = 0; pol < 2; pol++) {
        if (pol < a.npol) {
          const float2* src = a.Z + ((uint64_t(partl) * a.nchan_in + ic) * a.npol + pol) * a.Nc + uint64_t(csub) * F;
#pragma unroll
          for (unsigned i = 0; i < F; i += 2) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(src + i));   // L1: a lane's 8 F bytes span sectors
            v[pol][i] = make_float2(x.x, x.y);                                  // that its next load touches again
            v[pol][i + 1] = make_float2(x.z, x.w);
          }
          dftR<F, true>(v[pol]);
        } else {
#pragma unroll
          for (unsigned i = 0; i < F; i++) v[pol][i] = make_float2(0.f, 0.f);
        }
      }
      const unsigned s0 = (partl - p_begin) * nkeep;                           // first sample of this part in the row
#pragma unroll
      for (unsigned i = 0; i < F; i++) {
        if (i - np0 >= nkeep) continue;                   // unsigned: samples before nfilt_pos wrap to huge values
        const unsigned so = (s0 + (i - np0)) * edim;
        if (EPI == EPI_VOLT) {
#pragma unroll
          for (unsigned pol = 0; pol < 2; pol++)
            if (pol < a.npol) {
              float* row = stage + (cl * nplane + pol) * RS + so;
              row[0] = v[pol][i].x;
              row[1] = v[pol][i].y;
            }
        } else {
          float r[4] = {0.f, 0.f, 0.f, 0.f};
          detect_products(state, v[0][i], v[1][i], r);
          for (unsigned pr = 0; pr < nprod; pr++) stage[(cl * nplane + pr / edim) * RS + so + pr % edim] = r[pr];
        }
      }
    }
  }
  __syncthreads();
  // coalesced write-out: one warp per row, lanes along the row
  const unsigned nrow = min(128u, nch - ch0) * nplane, len = (p_end - p_begin) * nkeep * edim;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (unsigned row = warp; row < nrow; row += 4u) {
    const unsigned c = ch0 + row / nplane, pl = row % nplane;
    float* out;
    if (EPI == EPI_VOLT) out = a.sink.volt + (uint64_t(c) * a.npol + pl) * a.sink.volt_span + (a.part0 + p_begin) * a.sink.volt_step;
    else out = a.sink.det + (uint64_t(c) * nplane + pl) * a.sink.det_span + (a.part0 + p_begin) * uint64_t(nkeep) * edim;
    const float* srow = stage + row * RS;
    for (unsigned i = lane; i < len; i += 32u) out[i] = srow[i];
  }
}

template <int EPT, int EPI, unsigned FCT>
__global__ void __launch_bounds__((FCT && EPT == 32) ? 512 : 1024, 1) k_chan_inv(ChanArgs a) {
  extern __shared__ float2 smem[];
  const unsigned F = FCT ? FCT : a.F;
  const unsigned T = EPT ? F / (EPT ? EPT : 1) : 1;
  const unsigned f = threadIdx.x / T;              // transform within the CTA
  const unsigned j = threadIdx.x % T;
  const unsigned NF = a.CB * a.npol_cta;
  const unsigned cb = f / a.npol_cta;
  const unsigned pol = blockIdx.z * a.npol_cta + f % a.npol_cta;
  const unsigned ch0 = blockIdx.x * a.CB;          // first output channel of this CTA
  const unsigned ch = ch0 + cb;
  const unsigned ic = ch / a.C, csub = ch % a.C;
  const unsigned partl = blockIdx.y;
  const uint64_t part = a.part0 + partl;
  const unsigned blk = (partl * a.nchan_in + ic) * a.npol + pol;
  constexpr unsigned sh = SwzShift<EPT>::value;
  auto fmap = [&](unsigned ff, unsigned idx) -> unsigned {
    if (!EPT) return ff;
    if (F < 16) return ff * F + idx;
    return ff * F + (idx ^ ((idx >> sh) & 15u));
  };

  const unsigned nprod = (EPI == EPI_VOLT) ? 0 : state_nprod(a.sink.state, a.npol);

  const float2* src = a.Z + uint64_t(blk) * a.Nc + uint64_t(csub) * F;
  if (EPT) {
    if (F >= 16) {
      MapRows map{f * F, sh, 0};
      auto load = [&](unsigned idx) -> float2 { return ld_nc_f2(src + idx); };
      auto store = [&](unsigned idx, float2 v, int) { smem[map(idx)] = v; };
      fft_any<(EPT ? EPT : 2), true, FCT>(F, j, T, map, smem, a.twF, a.twFs, load, store);
    } else {
      // F = 2, 4, 8: one thread per transform, no shared-memory exchange needed
      MapRows map{f * F, 0, 0};
      auto load = [&](unsigned idx) -> float2 { return ld_nc_f2(src + idx); };
      auto store = [&](unsigned idx, float2 v, int) { smem[f * F + idx] = v; };
      block_fft<(EPT ? EPT : 2), true>(F, j, T, map, smem, a.twF, F, load, store);
    }
  } else {
    smem[f] = ld_nc_f2(src);   // freq_res == 1: no inverse transform (Filterbank.C:621-631)
  }
  __syncthreads();

  const unsigned nkeep = a.nkeep, np0 = a.nfilt_pos;
  if (EPI == EPI_VOLT) {
    const unsigned total = NF * nkeep;
    for (unsigned it = threadIdx.x; it < total; it += blockDim.x) {
      const unsigned ff = it / nkeep, m = it % nkeep;
      const unsigned c2 = ch0 + ff / a.npol_cta, p2 = blockIdx.z * a.npol_cta + ff % a.npol_cta;
      float2* out = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(c2) * a.npol + p2) * a.sink.volt_span +
                                              part * a.sink.volt_step);
      out[m] = smem[fmap(ff, np0 + m)];
    }
    return;
  }

  // detected epilogues need both polarisations in this CTA (npol_cta == npol)
  const unsigned dndim = a.sink.dndim;
  const unsigned dnpol = nprod / dndim;
  if (EPI == EPI_DETECT) {
    const unsigned total = a.CB * nkeep;
    for (unsigned it = threadIdx.x; it < total; it += blockDim.x) {
      const unsigned c = it / nkeep, m = it % nkeep;
      float2 p = smem[fmap(c * a.npol, np0 + m)];
      float2 q = a.npol > 1 ? smem[fmap(c * a.npol + 1, np0 + m)] : make_float2(0.f, 0.f);
      float r[4];
      detect_products(a.sink.state, p, q, r);
      const uint64_t osamp = part * nkeep + m;
      for (unsigned pr = 0; pr < nprod; pr++) {
        float* out = a.sink.det + (uint64_t(ch0 + c) * dnpol + pr / dndim) * a.sink.det_span;
        out[osamp * dndim + pr % dndim] = r[pr];
      }
    }
    return;
  }

  // EPI_FOLD.  Every thread walks L consecutive samples of one channel, summing sequentially while
  // the phase bin is unchanged (the order of the reference's per-bin +=, Fold.C:844-852).  A run
  // is added to the global profile with RED.ADD.F32 as soon as it ends (no shared-memory float
  // atomics: those compile to CAS loops).
  {
    const unsigned nbin = a.sink.nbin;
    const unsigned nthreads = blockDim.x;
    unsigned L = (a.CB * nkeep + nthreads - 1) / nthreads;
    L = (L + 3u) & ~3u;
    if (L < 4) L = 4;
    const unsigned nchunk = (nkeep + L - 1) / L;
    const unsigned total = a.CB * nchunk;
    const unsigned* plan = a.sink.bins + partl * uint64_t(nkeep);
    const uint64_t prof0 = uint64_t(ch0) * nbin * nprod;
    auto red_add = [&](unsigned key, const float* acc) {
      // key = c*(nbin+1) + bin; profile layout per channel [npol'][nbin][ndim'];  bin == nbin marks the samples of
      // a flagged window (weights.cu): dropped
      const unsigned c = key / (nbin + 1u), bin = key - c * (nbin + 1u);
      if (bin == nbin) return;
      const uint64_t base = prof0 + uint64_t(c) * nbin * nprod;
      for (unsigned pr = 0; pr < nprod; pr++)
        profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, base + (uint64_t(pr / dndim) * nbin + bin) * dndim + pr % dndim, acc[pr]);
    };
    const bool chunk_monotonic = a.sink.phase_per_sample > 0.0 && a.sink.phase_per_sample * double(L) < 0.25;
    const unsigned niter = (total + nthreads - 1) / nthreads;
    for (unsigned itr = 0; itr < niter; itr++) {
      const unsigned it = itr * nthreads + threadIdx.x;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      unsigned key = 0xffffffffu;
      if (it < total) {
        const unsigned c = (a.CB == 1) ? 0 : it / nchunk;
        const unsigned m0 = (it - c * nchunk) * L;
        const unsigned m1 = min(nkeep, m0 + L);
        const unsigned fp = c * a.npol;
        // When the phase advances by less than a quarter turn per chunk the bins of a chunk are
        // monotonic, so first == last means the whole chunk falls into one bin (the common case:
        // bins are usually many samples wide) and no per-sample bin look-up or compare is needed.
        const unsigned bfirst = __ldg(plan + m0), blast = __ldg(plan + m1 - 1);
        if (chunk_monotonic && bfirst == blast) {
          key = c * (nbin + 1u) + bfirst;
#pragma unroll 4
          for (unsigned m = m0; m < m1; m++) {
            float2 p = smem[fmap(fp, np0 + m)];
            float2 q = a.npol > 1 ? smem[fmap(fp + 1, np0 + m)] : make_float2(0.f, 0.f);
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            detect_products(a.sink.state, p, q, r);
#pragma unroll
            for (int pr = 0; pr < 4; pr++) acc[pr] += r[pr];
          }
        } else {
          for (unsigned m = m0; m < m1; m++) {
            const unsigned k = c * (nbin + 1u) + __ldg(plan + m);
            float2 p = smem[fmap(fp, np0 + m)];
            float2 q = a.npol > 1 ? smem[fmap(fp + 1, np0 + m)] : make_float2(0.f, 0.f);
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            detect_products(a.sink.state, p, q, r);
            if (k != key) {
              if (key != 0xffffffffu) red_add(key, acc);
              key = k;
#pragma unroll
              for (int pr = 0; pr < 4; pr++) acc[pr] = r[pr];
            } else {
#pragma unroll
              for (int pr = 0; pr < 4; pr++) acc[pr] += r[pr];
            }
          }
        }
      }
      // every walk adds its trailing run straight to the profile.  (An earlier version first combined equal
      // keys of neighbouring lanes with a shuffle scan; that is only valid while the keys of a warp are
      // contiguous, i.e. while the pulse period is longer than the samples a warp walks -- a randomised
      // sweep with 8-bin, few-sample periods caught it double counting.  Direct RED is also faster.)
      if (key != 0xffffffffu) red_add(key, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3' (convolution path): inverse column pass + discard + epilogue
// ------------------------------------------------------------------------------------------
struct ColsInvArgs {
  const float2* A;          // B[k1][m2] after K2<CONV>
  const float2* twP;
  unsigned P, Q, lb, npol, nchan_in, Nc;
  unsigned nfilt_pos, nkeep;
  uint64_t part0;
  FbSink sink;
};

template <int EPI>
__global__ void __launch_bounds__(1024, 1) k_cols_inv(ColsInvArgs a) {
  extern __shared__ float2 smem[];
  const unsigned B = 1u << a.lb;
  const unsigned T = a.P >> 4;
  const unsigned per_pol = T * B;
  const unsigned pol = threadIdx.x / per_pol;
  const unsigned t = threadIdx.x % per_pol;
  const unsigned b = t & (B - 1);
  const unsigned j = t >> a.lb;
  const unsigned m2 = blockIdx.x * B + b;
  const unsigned partl = blockIdx.y / a.nchan_in;
  const unsigned ic = blockIdx.y % a.nchan_in;
  const uint64_t part = a.part0 + partl;
  const unsigned blk = (partl * a.nchan_in + ic) * a.npol + pol;
  float2* spol = smem + uint64_t(pol) * a.P * B;
  MapCols map{b, a.lb, 4};
  const float2* src = a.A + uint64_t(blk) * a.Nc + m2;
  auto load = [&](unsigned k1) -> float2 { return ld_nc_f2(src + uint64_t(k1) * a.Q); };
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;

  if (EPI == EPI_VOLT) {
    float2* out = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(ic) * a.npol + pol) * a.sink.volt_span +
                                            part * a.sink.volt_step);
    auto store = [&](unsigned m1, float2 v, int) {
      const unsigned m = m1 * a.Q + m2;
      if (m >= np0 && m < np0 + nkeep) out[m - np0] = v;
    };
    block_fft<16, true>(a.P, j, T, map, spol, a.twP, a.P, load, store);
    return;
  }

  const unsigned nprod = state_nprod(a.sink.state, a.npol);
  const unsigned dndim = a.sink.dndim, dnpol = nprod / dndim;
  const unsigned nbin = a.sink.nbin;
  float* bins = reinterpret_cast<float*>(smem + uint64_t(a.npol) * a.P * B);
  if (EPI == EPI_FOLD)
    for (unsigned i = threadIdx.x; i < nbin * nprod; i += blockDim.x) bins[i] = 0.f;

  {
    auto store = [&](unsigned m1, float2 v, int) { spol[map(m1)] = v; };
    block_fft<16, true>(a.P, j, T, map, spol, a.twP, a.P, load, store);
  }
  __syncthreads();

  // items (m1, b): lanes span b, so a group of B lanes holds B consecutive time samples
  const unsigned total = a.P * B;
  const unsigned* plan = (EPI == EPI_FOLD) ? a.sink.bins + partl * uint64_t(nkeep) : nullptr;
  for (unsigned it0 = 0; it0 < total; it0 += blockDim.x) {
    const unsigned it = it0 + threadIdx.x;
    const bool in_range = it < total;
    const unsigned m1 = in_range ? it >> a.lb : 0, bb = it & (B - 1);
    const unsigned m = m1 * a.Q + blockIdx.x * B + bb;
    const bool valid = in_range && m >= np0 && m < np0 + nkeep;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      MapCols mp{bb, a.lb, 4};
      float2 p = smem[mp(m1)];
      float2 q = a.npol > 1 ? smem[uint64_t(a.P) * B + mp(m1)] : make_float2(0.f, 0.f);
      detect_products(a.sink.state, p, q, r);
    }
    if (EPI == EPI_DETECT) {
      if (valid) {
        const uint64_t osamp = part * nkeep + (m - np0);
        for (unsigned pr = 0; pr < nprod; pr++) {
          float* out = a.sink.det + (uint64_t(ic) * dnpol + pr / dndim) * a.sink.det_span;
          out[osamp * dndim + pr % dndim] = r[pr];
        }
      }
    } else {
      if (valid) {
        const unsigned bin = __ldg(plan + (m - np0));
        if (bin < nbin)                                      // nbin: flagged window
          for (unsigned pr = 0; pr < nprod; pr++)
            atomicAdd(bins + (uint64_t(pr / dndim) * nbin + bin) * dndim + pr % dndim, r[pr]);
      }
    }
  }
  if (EPI == EPI_FOLD) {
    __syncthreads();
    float* prof = a.sink.profile + uint64_t(ic) * nbin * nprod;
    for (unsigned i = threadIdx.x; i < nbin * nprod; i += blockDim.x) {
      float v = bins[i];
      if (v != 0.f) atomicAdd(prof + i, v);
    }
  }
}

// K3' with the fold epilogue.  A CTA's kept samples are P groups of B consecutive time samples (one group per
// m1, Q samples apart).  The detected products go back to shared memory in (m1, b) order (padded by one sample per
// group: conflict-free for the readers); then one thread per group walks its B consecutive samples, sums runs of
// equal phase bin sequentially (the order of Fold.C:844-852) and adds every finished run to the PhaseSeries with one
// RED.ADD.F32 per product.  (The first version accumulated per-sample into shared-memory bins with float atomics --
// CAS loops, 16 lanes on one address -- and then flushed ~one RED per sample anyway: 4.3 ms per 16 parts of cfg3
// against 0.5 ms for the transform itself.)
template <unsigned NPOL, unsigned NPROD>
__global__ void __launch_bounds__(1024, 1) k_cols_inv_fold(ColsInvArgs a) {
  extern __shared__ float2 smem[];
  const unsigned B = 1u << a.lb;
  const unsigned T = a.P >> 4;
  const unsigned per_pol = T * B;
  const unsigned pol = threadIdx.x / per_pol;
  const unsigned t = threadIdx.x % per_pol;
  const unsigned b = t & (B - 1);
  const unsigned j = t >> a.lb;
  const unsigned m2 = blockIdx.x * B + b;
  const unsigned partl = blockIdx.y / a.nchan_in;
  const unsigned ic = blockIdx.y % a.nchan_in;
  const unsigned blk = (partl * a.nchan_in + ic) * NPOL + pol;
  float2* spol = smem + uint64_t(pol) * a.P * B;
  MapCols map{b, a.lb, 4};
  const float2* src = a.A + uint64_t(blk) * a.Nc + m2;
  auto load = [&](unsigned k1) -> float2 { return ld_nc_f2(src + uint64_t(k1) * a.Q); };
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;
  {
    auto store = [&](unsigned m1, float2 v, int) { spol[map(m1)] = v; };
    block_fft<16, true>(a.P, j, T, map, spol, a.twP, a.P, load, store);
  }
  __syncthreads();
  // detect: NIT samples per thread, kept in registers while the transforms are overwritten
  constexpr unsigned NIT = 16 / NPOL;                       // P*B samples / (NPOL * P/16 * B) threads
  float r[NIT][NPROD];
#pragma unroll
  for (unsigned i = 0; i < NIT; i++) {
    const unsigned it = i * blockDim.x + threadIdx.x;
    const unsigned m1 = it >> a.lb, bb = it & (B - 1);
    MapCols mp{bb, a.lb, 4};
    const float2 p = smem[mp(m1)];
    const float2 q = NPOL > 1 ? smem[uint64_t(a.P) * B + mp(m1)] : make_float2(0.f, 0.f);
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    detect_products(a.sink.state, p, q, d);
#pragma unroll
    for (unsigned pr = 0; pr < NPROD; pr++) r[i][pr] = d[pr];
  }
  __syncthreads();
  float* stage = reinterpret_cast<float*>(smem);
#pragma unroll
  for (unsigned i = 0; i < NIT; i++) {
    const unsigned it = i * blockDim.x + threadIdx.x;
    float* s = stage + uint64_t(it + (it >> a.lb)) * NPROD;
    if (NPROD == 4) *reinterpret_cast<float4*>(s) = make_float4(r[i][0], r[i][1 % NPROD], r[i][2 % NPROD], r[i][3 % NPROD]);
    else if (NPROD == 2) *reinterpret_cast<float2*>(s) = make_float2(r[i][0], r[i][1 % NPROD]);
    else s[0] = r[i][0];
  }
  __syncthreads();
  const unsigned nbin = a.sink.nbin, dndim = a.sink.dndim;
  const unsigned* plan = a.sink.bins + partl * uint64_t(nkeep);
  const uint64_t prof = uint64_t(ic) * nbin * NPROD;               // per channel [npol'][nbin][ndim']
  // Groups of neighbouring lanes are Q samples apart in time: when a phase bin is wider than that (cfg4: 35 thousand
  // samples per bin) many groups end in the same bin, so the last run of every lane is first combined over
  // neighbouring lanes with equal bin (segmented scan over maximal runs of equal keys: every value is added exactly
  // once whatever the key pattern) and only the tail of each run of lanes issues the REDs.
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned g0 = threadIdx.x - lane; g0 < a.P; g0 += blockDim.x) {       // warp-uniform trip count
    const unsigned g = g0 + lane;
    float acc[NPROD];
#pragma unroll
    for (unsigned pr = 0; pr < NPROD; pr++) acc[pr] = 0.f;
    unsigned cur = 0xffffffffu;
    if (g < a.P) {
      const unsigned mbase = g * a.Q + blockIdx.x * B;
      const float* s = stage + uint64_t(g * B + g) * NPROD;
      for (unsigned bb = 0; bb < B; bb++) {
        const unsigned m = mbase + bb;
        const unsigned bin = (m >= np0 && m < np0 + nkeep) ? __ldg(plan + (m - np0)) : 0xffffffffu;
        float v[NPROD];
        if (NPROD == 4) {
          const float4 x = *reinterpret_cast<const float4*>(s + bb * NPROD);
          v[0] = x.x; v[1 % NPROD] = x.y; v[2 % NPROD] = x.z; v[3 % NPROD] = x.w;
        } else if (NPROD == 2) {
          const float2 x = *reinterpret_cast<const float2*>(s + bb * NPROD);
          v[0] = x.x; v[1 % NPROD] = x.y;
        } else v[0] = s[bb];
        if (bin != cur) {
          if (cur < nbin)                                      // nbin: flagged window; 0xffffffff: discarded sample
#pragma unroll
            for (unsigned pr = 0; pr < NPROD; pr++)
              profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, prof + (uint64_t(pr / dndim) * nbin + cur) * dndim + pr % dndim, acc[pr]);
          cur = bin;
#pragma unroll
          for (unsigned pr = 0; pr < NPROD; pr++) acc[pr] = v[pr];
        } else {
#pragma unroll
          for (unsigned pr = 0; pr < NPROD; pr++) acc[pr] += v[pr];
        }
      }
    }
    // the lane's last run (cur, acc): combine over neighbouring lanes
    const unsigned prev = __shfl_up_sync(0xffffffffu, cur, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != cur);
    const unsigned h = 31u - __clz(heads & (0xffffffffu >> (31u - lane)));    // first lane of this lane's run
#pragma unroll
    for (unsigned o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (unsigned pr = 0; pr < NPROD; pr++) {
        const float up = __shfl_up_sync(0xffffffffu, acc[pr], o);
        if (lane >= h + o) acc[pr] += up;
      }
    }
    const bool tail = lane == 31u || ((heads >> (lane + 1u)) & 1u);
    if (tail && cur < nbin)
#pragma unroll
      for (unsigned pr = 0; pr < NPROD; pr++)
        profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, prof + (uint64_t(pr / dndim) * nbin + cur) * dndim + pr % dndim, acc[pr]);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaFuncAttributes fa;
  B200_CUDA(cudaFuncGetAttributes(&fa, kernel));
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(bytes - fa.sharedSizeBytes)));
  return B200_OK;
}

static const size_t SMEM_TILE = 128 * 1024;   // FFT tile budget per CTA (elements * 8 B)

int fb_run(b200_fb_plan* pl, const FbSource& src, const FbSink& sink, uint64_t npart) {
  Context* ctx = pl->ctx;
  cudaStream_t st = ctx->stream;
  const unsigned npol = pl->desc.npol, nchan_in = pl->desc.input_nchan;
  const unsigned nblk1 = nchan_in * npol;
  const unsigned B = 1u << pl->lbB;
  const unsigned nprod = sink.kind == EPI_VOLT ? 0 : state_nprod(sink.state, npol);

  if (sink.kind != EPI_VOLT) {
    B200_REQUIRE(sink.state >= 0 && sink.state <= 3, "invalid detection state %d", sink.state);
    if (sink.state >= B200_COHERENCE) {
      B200_REQUIRE(npol == 2, "Coherence/Stokes detection requires npol == 2 (Detection.C:476-489)");
      B200_REQUIRE(sink.dndim == 1 || sink.dndim == 2 || sink.dndim == 4, "invalid detection ndim %u", sink.dndim);
    } else {
      B200_REQUIRE(sink.dndim == 1, "Intensity/PPQQ detection has ndim 1");
    }
  }

  const unsigned batch = (src.batch_override && src.batch_override < pl->batch) ? src.batch_override : pl->batch;
  for (uint64_t part0 = 0; part0 < npart; part0 += batch) {
    const unsigned nb = (unsigned)std::min<uint64_t>(batch, npart - part0);
    if (src.batch_ready) B200_CUDA(cudaStreamWaitEvent(st, src.batch_ready[part0 / batch], 0));
    // 16384- to 131072-point convolutions: one kernel for the whole transform pair and its epilogue (clusterconv.cu)
    if (cc_applies(pl, src, sink)) {
      FbSink sk = sink;
      if (sk.kind == EPI_FOLD) sk.bins = sink.bins + part0 * pl->nkeep;
      int rc = cc_run(pl, src, sk, part0, nb);
      if (rc == B200_OK) continue;
      if (rc != CC_NOT_RUN) return rc;
    }
    // a transform that fits one column pass (Q == 1) of complex input needs no row pass: K1 multiplies by the response
    // and writes Z (the row kernel would be P blocks of a handful of threads: 2.8 of 3.6 ms on the top UWL sub-bands)
    const bool skip_k2 = pl->Q == 1 && !pl->desc.input_real && !pl->conv_path;
    // ---- K1 ----
    const bool k1_fast = pl->fast_k1 && src.kind <= SRC_CASPSR8 && (src.kind == SRC_F32 || (src.step % 4 == 0 && (reinterpret_cast<uintptr_t>(src.ptr) & 3) == 0));
    // long convolutions (N > 131072): c2 kernels of longconv.cu; all three -> polarisations interleaved in the scratch
    const bool bc1 = bc_k1_applies(pl, src), bc2 = bc_k2_applies(pl), bc3 = bc_k3_applies(pl);
    const bool bc_il = bc1 && bc2 && bc3;
    B200_REQUIRE(bc_il || pl->Nc <= (1u << 22), "transforms of more than 2^22 points: source format %d is not built", src.kind);
    if (k1_fast) {
      int rc = fast_k1(pl, src, part0, nb);
      if (rc != B200_OK) return rc;
    } else if (bc1) {
      int rc = bc_k1(pl, src, part0, nb, bc_il);
      if (rc != B200_OK) return rc;
    } else {
      ColsArgs a;
      a.src = src.ptr; a.span = src.span; a.step = src.step; a.lut = src.d_lut;
      a.dst = skip_k2 ? pl->scratchZ : pl->scratchA; a.H = skip_k2 ? pl->d_response : nullptr;
      a.twP = pl->twP.tw; a.twPs = pl->twP.stage; a.blo = pl->bigN.lo; a.bhi = pl->bigN.hi;
      a.P = pl->P; a.Q = pl->Q; a.lb = pl->lbB; a.npol = npol; a.nchan_in = nchan_in; a.Nc = pl->Nc;
      a.part0 = part0;
      a.first = src.first; a.scale = src.scale; a.sample_swap = src.sample_swap; a.ndim = src.ndim;
      a.win = src.win; a.lowsel = src.lowsel; a.negsel = src.negsel;
      dim3 grid(pl->Q / B, nb * nblk1);
      const bool ct = (pl->P == 2048) && src.kind <= SRC_CASPSR8;   // compile-time-sized fast path (float / CASPSR sources)
      static const int k1_ept = tune_int("B200_K1_EPT", 32);
      const bool ct16 = ct && k1_ept == 16;
      if (ct16) a.twPs = pl->twP.stage16;
      dim3 block((pl->P / ((ct && !ct16) ? 32 : 16)) * B);
      size_t smem = size_t(pl->P) * B * sizeof(float2);
      if (tune_flag("B200_DEBUG", false) && part0 == 0) {
        int nb1 = 0, nb2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, k_cols_fwd<SRC_CASPSR8, 32, 2048>, block.x, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_cols_fwd<SRC_CASPSR8, 16, 0>, block.x, smem);
        fprintf(stderr, "[b200] K1 grid (%u,%u) block %u smem %zu occupancy ct=%d generic=%d\n", grid.x, grid.y, block.x, smem, nb1, nb2);
      }
      LaunchScope ls(ctx, KC_COLS_FWD);
      if (ct16) {
        if (src.kind == SRC_F32) k_cols_fwd<SRC_F32, 16, 2048><<<grid, block, smem, st>>>(a);
        else k_cols_fwd<SRC_CASPSR8, 16, 2048><<<grid, block, smem, st>>>(a);
      } else if (ct) {
        if (src.kind == SRC_F32) k_cols_fwd<SRC_F32, 32, 2048><<<grid, block, smem, st>>>(a);
        else k_cols_fwd<SRC_CASPSR8, 32, 2048><<<grid, block, smem, st>>>(a);
      } else {
        if (src.kind == SRC_F32) k_cols_fwd<SRC_F32, 16, 0><<<grid, block, smem, st>>>(a);
        else if (src.kind == SRC_MEERKAT8) k_cols_fwd<SRC_MEERKAT8, 16, 0><<<grid, block, smem, st>>>(a);
        else if (src.kind == SRC_UWB16) k_cols_fwd<SRC_UWB16, 16, 0><<<grid, block, smem, st>>>(a);
        else if (src.kind == SRC_GENERIC8) k_cols_fwd<SRC_GENERIC8, 16, 0><<<grid, block, smem, st>>>(a);
        else if (src.kind == SRC_TWOBIT) k_cols_fwd<SRC_TWOBIT, 16, 0><<<grid, block, smem, st>>>(a);
        else k_cols_fwd<SRC_CASPSR8, 16, 0><<<grid, block, smem, st>>>(a);
      }
    }
    // ---- K2 ----
    if (skip_k2) {
    } else if (pl->fast_k2) {
      int rc = fast_k2(pl, nb);
      if (rc != B200_OK) return rc;
    } else if (bc2) {
      int rc = bc_k2(pl, nb, bc_il);
      if (rc != B200_OK) return rc;
    } else {
      RowsArgs a;
      a.A = pl->scratchA; a.Z = pl->scratchZ; a.H = pl->d_response; a.twQ = pl->twQ.tw; a.twQs = pl->twQ.stage;
      a.tw2Q = pl->tw2Q.tw;
      a.blo = pl->bigN.lo; a.bhi = pl->bigN.hi; a.b2lo = pl->big2N.lo; a.b2hi = pl->big2N.hi;
      a.P = pl->P; a.Q = pl->Q; a.G = pl->G; a.Nc = pl->Nc; a.npol = npol; a.nchan_in = nchan_in;
      const bool split = pl->desc.input_real;
      const unsigned nslots = split ? 2 * pl->G : pl->G;
      const bool ct = (pl->Q == 1024 && !pl->conv_path);   // compile-time-sized fast path (EPT 32)
      const unsigned T = ct ? pl->Q / 32 : pl->Q >= 16 ? pl->Q / 16 : 1;
      dim3 grid(split ? (pl->P / 2) / pl->G : pl->P / pl->G, nb * nblk1);
      dim3 block(nslots * T);
      size_t smem = size_t(nslots) * pl->Q * sizeof(float2);
      LaunchScope ls(ctx, KC_ROWS);
      if (ct) {
        if (split) k_rows<true, false, 32, 1024><<<grid, block, smem, st>>>(a);
        else k_rows<false, false, 32, 1024><<<grid, block, smem, st>>>(a);
      } else if (pl->Q >= 16) {
        if (split) {
          if (pl->conv_path) k_rows<true, true, 16, 0><<<grid, block, smem, st>>>(a);
          else k_rows<true, false, 16, 0><<<grid, block, smem, st>>>(a);
        } else {
          if (pl->conv_path) k_rows<false, true, 16, 0><<<grid, block, smem, st>>>(a);
          else k_rows<false, false, 16, 0><<<grid, block, smem, st>>>(a);
        }
      } else {
        if (split) k_rows<true, false, 0, 0><<<grid, block, smem, st>>>(a);
        else k_rows<false, false, 0, 0><<<grid, block, smem, st>>>(a);
      }
    }
    // ---- K3 ----
    FbSink sk = sink;
    if (sk.kind == EPI_FOLD) {
      sk.bins = sink.bins + part0 * pl->nkeep;
      if (sink.runs) {
        sk.runs = sink.runs + part0 * (uint64_t(pl->nkeep) + 1);
        sk.nruns = sink.nruns + part0;
      }
    }
    if (pl->fast_k3) {
      int rc = fast_k3(pl, sk, part0, nb);
      if (rc != B200_OK) return rc;
    } else if (pl->conv_path && bc3) {
      int rc = bc_k3(pl, sk, part0, nb, bc_il);
      if (rc != B200_OK) return rc;
    } else if (pl->conv_path) {
      ColsInvArgs a;
      a.A = pl->scratchA; a.twP = pl->twP.tw;
      a.P = pl->P; a.Q = pl->Q; a.npol = npol; a.nchan_in = nchan_in; a.Nc = pl->Nc;
      a.nfilt_pos = pl->desc.nfilt_pos; a.nkeep = pl->nkeep; a.part0 = part0; a.sink = sk;
      // both polarisations share the CTA: halve the column tile if needed
      unsigned lb = pl->lbB;
      while (lb > 0 && size_t(npol) * pl->P * (1u << lb) * sizeof(float2) > SMEM_TILE) lb--;
      while (lb > 0 && npol * (pl->P / 16) * (1u << lb) > 1024) lb--;
      a.lb = lb;
      const unsigned Bi = 1u << lb;
      dim3 grid(pl->Q / Bi, nb * nchan_in);
      dim3 block(npol * (pl->P / 16) * Bi);
      size_t smem = size_t(npol) * pl->P * Bi * sizeof(float2);
      if (sk.kind == EPI_FOLD) { // the detected products are staged over the transforms, one padding sample per group
        smem = std::max(smem, (size_t(pl->P) * Bi + pl->P) * nprod * sizeof(float));
        B200_REQUIRE(block.x % 32 == 0, "convolution fold epilogue: %u threads per block are not whole warps", block.x);
      }
      LaunchScope ls(ctx, KC_INV);
      if (sk.kind == EPI_VOLT) k_cols_inv<EPI_VOLT><<<grid, block, smem, st>>>(a);
      else if (sk.kind == EPI_DETECT) k_cols_inv<EPI_DETECT><<<grid, block, smem, st>>>(a);
      else if (npol == 2 && nprod == 4) k_cols_inv_fold<2, 4><<<grid, block, smem, st>>>(a);
      else if (npol == 2 && nprod == 2) k_cols_inv_fold<2, 2><<<grid, block, smem, st>>>(a);
      else if (npol == 2) k_cols_inv_fold<2, 1><<<grid, block, smem, st>>>(a);
      else k_cols_inv_fold<1, 1><<<grid, block, smem, st>>>(a);
    } else {
      ChanArgs a;
      a.Z = pl->scratchZ; a.twF = pl->twF.tw; a.twFs = pl->twF.stage;
      a.F = pl->F; a.C = pl->C; a.Nc = pl->Nc; a.npol = npol; a.nchan_in = nchan_in;
      a.nfilt_pos = pl->desc.nfilt_pos; a.nkeep = pl->nkeep; a.part0 = part0; a.sink = sk;
      const unsigned F = pl->F;
      static const bool small_k3 = tune_flag("B200_K3_SMALL", true);
      const unsigned s_edim = sk.kind == EPI_VOLT ? 2u : sk.dndim;
      const unsigned s_nplane = sk.kind == EPI_VOLT ? npol : (sk.kind == EPI_DETECT ? state_nprod(sk.state, npol) / s_edim : 1u);
      if (small_k3 && F >= 2 && F <= 16 && sk.kind != EPI_FOLD && npol <= 2 && pl->nkeep > 0 &&
          size_t(128) * s_nplane * (pl->nkeep * s_edim + 1) * sizeof(float) <= 48 * 1024) {
        // one thread per channel, PG parts per thread: as many as the 48 KiB staging tile holds (rows of PG * nkeep
        // samples per output plane), fewer if the grid would not fill the machine a few times over
        const unsigned nch = pl->nchan_out;
        const unsigned gx = (nch + 127) / 128;
        const unsigned edim = sk.kind == EPI_VOLT ? 2u : sk.dndim;
        const unsigned nplane = sk.kind == EPI_VOLT ? npol : state_nprod(sk.state, npol) / edim;
        unsigned PG = 16;
        while (PG > 1 && size_t(128) * nplane * (PG * pl->nkeep * edim + 1) * sizeof(float) > 48 * 1024) PG /= 2;
        while (PG > 1 && uint64_t(gx) * ((nb + PG - 1) / PG) < 4ull * ctx->sm_count) PG /= 2;
        const size_t ssm = size_t(128) * nplane * (PG * pl->nkeep * edim + 1) * sizeof(float);
        dim3 grid(gx, (nb + PG - 1) / PG);
        LaunchScope ls(ctx, KC_INV);
#define B200_K3S(FF)                                                                               \
  if (sk.kind == EPI_VOLT) k_chan_inv_small<FF, EPI_VOLT><<<grid, 128, ssm, st>>>(a, nb, PG);         \
  else k_chan_inv_small<FF, EPI_DETECT><<<grid, 128, ssm, st>>>(a, nb, PG);
        switch (F) {
          case 2: B200_K3S(2) break;
          case 4: B200_K3S(4) break;
          case 8: B200_K3S(8) break;
          default: B200_K3S(16) break;
        }
#undef B200_K3S
        B200_CUDA(cudaGetLastError());
        continue;
      }
      const unsigned ept = F >= 16 ? 16 : F;       // 16, 8, 4, 2, 1
      const unsigned T = F >= 16 ? F / 16 : 1;
      // transforms per CTA: both polarisations of CB channels
      unsigned npol_cta = npol;
      if (size_t(npol) * F * sizeof(float2) > SMEM_TILE || npol * T > 1024) {
        B200_REQUIRE(sk.kind == EPI_VOLT,
                     "freq_res=%u: detection/fold fused epilogues need both polarisations on one SM "
                     "(freq_res <= 8192); use the voltage output + b200_detect/b200_fold", F);
        npol_cta = 1;
      }
      unsigned CB = 1;
      while (CB * 2 <= pl->C && size_t(CB) * 2 * npol_cta * F * sizeof(float2) <= SMEM_TILE / 2 &&
             CB * 2 * npol_cta * T <= 512)
        CB *= 2;
      a.CB = CB; a.npol_cta = npol_cta;
      size_t smem = size_t(CB) * npol_cta * F * sizeof(float2);
      dim3 grid(pl->nchan_out / CB, nb, npol / npol_cta);
      const bool ct = (F == 8192 && npol_cta == npol && CB == 1);   // compile-time-sized fast path
      static const int k3_ept = tune_int("B200_K3_EPT", 32);
      const bool ct16 = ct && k3_ept == 16;
      if (ct16) a.twFs = pl->twF.stage16;
      dim3 block(CB * npol_cta * ((ct && !ct16) ? F / 32 : T));
      LaunchScope ls(ctx, KC_INV);
#define B200_K3(E, FC)                                                                          \
  if (sk.kind == EPI_VOLT) k_chan_inv<E, EPI_VOLT, FC><<<grid, block, smem, st>>>(a);            \
  else if (sk.kind == EPI_DETECT) k_chan_inv<E, EPI_DETECT, FC><<<grid, block, smem, st>>>(a);   \
  else k_chan_inv<E, EPI_FOLD, FC><<<grid, block, smem, st>>>(a);
      if (ct16) {
        B200_K3(16, 8192)
      } else if (ct) {
        B200_K3(32, 8192)
      } else switch (ept) {
        case 16: B200_K3(16, 0) break;
        case 8: B200_K3(8, 0) break;
        case 4: B200_K3(4, 0) break;
        case 2: B200_K3(2, 0) break;
        default: B200_K3(0, 0) break;
      }
#undef B200_K3
    }
    B200_CUDA(cudaGetLastError());
  }
  return B200_OK;
}

static int plan_set_attributes(size_t maxs) {
  int rc;
#define SET(k) if ((rc = set_smem(k, maxs)) != B200_OK) return rc;
  SET((k_cols_fwd<SRC_F32, 16, 0>)) SET((k_cols_fwd<SRC_CASPSR8, 16, 0>))
  SET((k_cols_fwd<SRC_MEERKAT8, 16, 0>)) SET((k_cols_fwd<SRC_UWB16, 16, 0>)) SET((k_cols_fwd<SRC_GENERIC8, 16, 0>)) SET((k_cols_fwd<SRC_TWOBIT, 16, 0>))
  SET((k_cols_fwd<SRC_F32, 32, 2048>)) SET((k_cols_fwd<SRC_CASPSR8, 32, 2048>))
  SET((k_cols_fwd<SRC_F32, 16, 2048>)) SET((k_cols_fwd<SRC_CASPSR8, 16, 2048>))
  SET((k_chan_inv<16, EPI_VOLT, 8192>)) SET((k_chan_inv<16, EPI_DETECT, 8192>)) SET((k_chan_inv<16, EPI_FOLD, 8192>))
  SET((k_rows<true, true, 16, 0>)) SET((k_rows<true, false, 16, 0>)) SET((k_rows<false, true, 16, 0>))
  SET((k_rows<false, false, 16, 0>)) SET((k_rows<true, false, 0, 0>)) SET((k_rows<false, false, 0, 0>))
  SET((k_rows<true, false, 32, 1024>)) SET((k_rows<false, false, 32, 1024>))
  SET(k_cols_inv<EPI_VOLT>) SET(k_cols_inv<EPI_DETECT>)
  SET((k_cols_inv_fold<2, 4>)) SET((k_cols_inv_fold<2, 2>)) SET((k_cols_inv_fold<2, 1>)) SET((k_cols_inv_fold<1, 1>))
  SET((k_chan_inv<32, EPI_VOLT, 8192>)) SET((k_chan_inv<32, EPI_DETECT, 8192>)) SET((k_chan_inv<32, EPI_FOLD, 8192>))
  SET((k_chan_inv<16, EPI_VOLT, 0>)) SET((k_chan_inv<16, EPI_DETECT, 0>)) SET((k_chan_inv<16, EPI_FOLD, 0>))
  SET((k_chan_inv<8, EPI_VOLT, 0>)) SET((k_chan_inv<8, EPI_DETECT, 0>)) SET((k_chan_inv<8, EPI_FOLD, 0>))
  SET((k_chan_inv<4, EPI_VOLT, 0>)) SET((k_chan_inv<4, EPI_DETECT, 0>)) SET((k_chan_inv<4, EPI_FOLD, 0>))
  SET((k_chan_inv<2, EPI_VOLT, 0>)) SET((k_chan_inv<2, EPI_DETECT, 0>)) SET((k_chan_inv<2, EPI_FOLD, 0>))
  SET((k_chan_inv<0, EPI_VOLT, 0>)) SET((k_chan_inv<0, EPI_DETECT, 0>)) SET((k_chan_inv<0, EPI_FOLD, 0>))
#undef SET
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_fb_plan_create(b200_context* cctx, const b200_fb_desc* d, b200_fb_plan** out) {
  B200_REQUIRE(cctx && d && out, "b200_fb_plan_create: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  B200_REQUIRE(d->npol == 1 || d->npol == 2, "npol=%u unsupported (1 or 2)", d->npol);
  B200_REQUIRE(d->input_nchan >= 1 && d->nchan_subband >= 1 && d->freq_res >= 1, "invalid channelisation");
  B200_REQUIRE(is_pow2(d->nchan_subband) && is_pow2(d->freq_res),
               "nchan_subband=%u and freq_res=%u must be powers of two", d->nchan_subband, d->freq_res);
  const uint64_t Nc64 = uint64_t(d->nchan_subband) * d->freq_res;
  B200_REQUIRE(Nc64 >= 16 && Nc64 <= (1ull << 24), "forward transform of %llu complex points unsupported (16..2^24)",
               (unsigned long long)Nc64);
  B200_REQUIRE(d->nfilt_pos + d->nfilt_neg < d->freq_res || (d->freq_res == 1 && d->nfilt_pos + d->nfilt_neg == 0),
               "nfilt_pos+nfilt_neg=%u must be smaller than freq_res=%u", d->nfilt_pos + d->nfilt_neg, d->freq_res);
  B200_CUDA(cudaSetDevice(ctx->device));

  b200_fb_plan* pl = new b200_fb_plan();
  memset(pl, 0, sizeof(*pl));
  pl->ctx = ctx;
  pl->desc = *d;
  pl->desc.h_response = nullptr;
  pl->C = d->nchan_subband;
  pl->F = d->freq_res;
  pl->Nc = (unsigned)Nc64;
  const unsigned nfilt_tot = d->nfilt_pos + d->nfilt_neg;
  if (d->input_real) {                       // Filterbank.C:139-148
    pl->nsamp_fft = 2 * pl->Nc;
    pl->nsamp_overlap = 2 * nfilt_tot * pl->C;
  } else {
    pl->nsamp_fft = pl->Nc;
    pl->nsamp_overlap = nfilt_tot * pl->C;
  }
  pl->nsamp_step = pl->nsamp_fft - pl->nsamp_overlap;   // :155
  pl->nkeep = pl->F - nfilt_tot;                        // :409
  pl->nchan_out = d->input_nchan * pl->C;

  // ---- factorisation of the forward transform ----
  pl->conv_path = (pl->C == 1 && pl->F > 8192);
  if (pl->F > 16384 && !pl->conv_path) {
    set_error("freq_res=%u with nchan_subband=%u: per-channel inverse transforms above 16384 points are not built yet",
              pl->F, pl->C);
    delete pl;
    return B200_ERR_UNSUPPORTED;
  }
  const unsigned lgN = ilog2(pl->Nc);
  if (pl->Nc <= 8192 && !pl->conv_path) {
    pl->P = pl->Nc;
    pl->Q = 1;
  } else {
    unsigned lgP = (lgN + 1) / 2;
    const unsigned lgp_cap = (unsigned)tune_int("B200_LGP_CAP", 11);
    if (lgP > lgp_cap) lgP = lgp_cap;
    pl->P = 1u << lgP;
    pl->Q = pl->Nc / pl->P;
  }
  // column tile: B columns, P*B elements <= 16384, (P/16)*B threads <= 1024
  // (tuning overrides: B200_TILE_KB_COLS / B200_TILE_KB_ROWS = tile budget in KiB)
  size_t tile_cols = SMEM_TILE, tile_rows = SMEM_TILE;
  tile_cols = size_t(tune_int("B200_TILE_KB_COLS", int(SMEM_TILE / 1024))) * 1024;
  tile_rows = size_t(tune_int("B200_TILE_KB_ROWS", int(SMEM_TILE / 1024))) * 1024;
  {
    unsigned lb = 0;
    while ((2u << lb) <= pl->Q && size_t(pl->P) * (2u << lb) * sizeof(float2) <= tile_cols &&
           (pl->P / 16) * (2u << lb) <= 1024 && (2u << lb) <= 16)
      lb++;
    pl->lbB = lb;
  }
  // row tile: G (+G mirror) rows of Q points
  {
    const unsigned mult = d->input_real ? 2 : 1;
    const unsigned T = pl->Q >= 16 ? pl->Q / 16 : 1;
    const unsigned rows_avail = d->input_real ? pl->P / 2 : pl->P;
    unsigned G = 1;
    while (G * 2 <= rows_avail && size_t(G) * 2 * mult * pl->Q * sizeof(float2) <= tile_rows &&
           G * 2 * mult * T <= 1024 && G * 2 <= 8)
      G *= 2;
    pl->G = G;
  }
  // parts per internal batch: 16 (launch overhead < 2 %) unless one spectrum buffer would exceed 2 GiB -- or, when
  // it fits, the smallest count that makes the tiles of all three persistent kernels a multiple of the SM count
  // (cfg1: 256 tiles per part, 148 SMs -> 37 parts = 64 full waves; measured +4.8 % over batches of 16)
  {
    const uint64_t per_part = uint64_t(d->input_nchan) * d->npol * pl->Nc * sizeof(float2);
    const uint64_t cap = std::max<uint64_t>(1, (2ull << 30) / per_part);
    uint64_t b = std::min<uint64_t>(cap, 16);
    // short transforms (cfg2: 512 KiB per part; the upper UWL sub-bands): 16 parts would be a few hundred CTAs of a few
    // microseconds each.  Take as many parts as fill a spectrum buffer of 256 MiB (cfg2, 2048-part blocks: 30.7 GS/s
    // with 37 parts per batch, 36.2 with 64, 43.8 with 256, 47.3 with 1024: launch count and tails beat L2 residency)
    if (per_part * 16 < (256ull << 20)) b = std::min<uint64_t>(std::min<uint64_t>(cap, 4096), (256ull << 20) / per_part);
    if (!pl->conv_path && pl->Q >= 8 && pl->P >= 16) {
      auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
      const uint64_t sm = uint64_t(std::max(1, ctx->sm_count)), blk = uint64_t(d->input_nchan) * d->npol;
      const uint64_t tiles[3] = {pl->Q / 8 * blk, pl->P / 16 * blk,
                                 pl->F >= 8192 ? uint64_t(pl->nchan_out) : std::max<uint64_t>(1, uint64_t(pl->nchan_out) * pl->F / 8192)};
      uint64_t need = 1;
      for (uint64_t t : tiles) {
        const uint64_t g = sm / gcd(sm, t);
        need = need / gcd(need, g) * g;
        if (need > 64) break;
      }
      // ... as long as one spectrum buffer stays below 1.25 GiB (cfg1: 37 x 32 MiB = 1.16 GiB)
      // whole multiples of that count near the size chosen above
      if (need > 1 && need <= 64 && need <= cap && need * per_part <= (5ull << 28)) b = std::max<uint64_t>(1, b / need) * need;
    }
    pl->batch = d->max_npart ? d->max_npart : unsigned(b);
    // the generic kernels index (part, input channel, polarisation) blocks through grid.y (limit 65535): a
    // 4096-channel dual-pol input allows 7 parts per launch, not 16
    const uint64_t blocks_per_part = uint64_t(d->input_nchan) * d->npol;
    if (blocks_per_part > 65535) {
      b200_fb_plan_destroy(pl);
      set_error("input_nchan*npol = %llu exceeds the 65535 blocks one launch can index", (unsigned long long)blocks_per_part);
      return B200_ERR_UNSUPPORTED;
    }
    pl->batch = unsigned(std::min<uint64_t>(pl->batch, 65535 / blocks_per_part));
  }

  int rc = plan_set_attributes(size_t(ctx->max_smem_optin));
  if (rc == B200_OK) rc = make_twiddle(pl->twP, pl->P, ctx->stream);
  if (rc == B200_OK) rc = make_twiddle(pl->twQ, pl->Q, ctx->stream);
  if (rc == B200_OK) rc = make_twiddle(pl->twF, pl->F, ctx->stream);
  if (rc == B200_OK) rc = make_twiddle(pl->tw2Q, 2 * pl->Q, ctx->stream);
  if (rc == B200_OK) rc = make_big_twiddle(pl->bigN, pl->Nc, ctx->stream);
  if (rc == B200_OK) rc = make_big_twiddle(pl->big2N, 2ull * pl->Nc, ctx->stream);
  if (rc != B200_OK) { b200_fb_plan_destroy(pl); return rc; }

  const uint64_t nblk = uint64_t(pl->batch) * d->input_nchan * d->npol;
  const uint64_t sbytes = nblk * pl->Nc * sizeof(float2);
  cudaError_t e = cudaMalloc(&pl->scratchA, sbytes);
  if (e == cudaSuccess && !pl->conv_path) e = cudaMalloc(&pl->scratchZ, sbytes);
  if (e == cudaSuccess && d->h_response) {
    const uint64_t rbytes = uint64_t(d->input_nchan) * pl->Nc * sizeof(float2);
    e = cudaMalloc(&pl->d_response, rbytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pl->d_response, d->h_response, rbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess) {
    b200_fb_plan_destroy(pl);
    return cuda_fail(e, "plan allocation", __FILE__, __LINE__);
  }
  pl->scratch_bytes = sbytes * (pl->conv_path ? 1 : 2);
  rc = fast_plan_init(pl);
  if (rc == B200_OK) rc = cc_plan_init(pl);
  if (rc == B200_OK) rc = bc_plan_init(pl);
  if (rc == B200_OK && pl->Nc > (1u << 22) && !pl->bc_ok) {
    // 2^23 and 2^24 points exist only as the long-transform kernels of longconv.cu (rows of 4096 / 8192 points)
    set_error("transforms of more than 2^22 complex points are built for single-channel convolution of complex "
              "dual-polarisation input only (got %u points)", pl->Nc);
    rc = B200_ERR_UNSUPPORTED;
  }
  if (rc != B200_OK) { b200_fb_plan_destroy(pl); return rc; }
  *out = pl;
  return B200_OK;
}

int b200_fb_plan_info(const b200_fb_plan* pl, b200_fb_info* info) {
  B200_REQUIRE(pl && info, "b200_fb_plan_info: null argument");
  info->n_fft = pl->Nc;
  info->nsamp_fft = pl->nsamp_fft;
  info->nsamp_overlap = pl->nsamp_overlap;
  info->nsamp_step = pl->nsamp_step;
  info->nkeep = pl->nkeep;
  info->fft_rows = pl->P;
  info->fft_cols = pl->Q;
  info->batch_npart = pl->batch;
  info->scratch_bytes = pl->scratch_bytes;
  return B200_OK;
}

int b200_fb_plan_destroy(b200_fb_plan* pl) {
  if (!pl) return B200_OK;
  free_twiddle(pl->twP);
  free_twiddle(pl->twQ);
  free_twiddle(pl->twF);
  free_twiddle(pl->tw2Q);
  free_big_twiddle(pl->bigN);
  free_big_twiddle(pl->big2N);
  fast_plan_free(pl);
  cc_plan_free(pl);
  bc_plan_free(pl);
  if (pl->d_response) cudaFree(pl->d_response);
  if (pl->scratchA) cudaFree(pl->scratchA);
  if (pl->scratchZ) cudaFree(pl->scratchZ);
  delete pl;
  return B200_OK;
}

int b200_fb_perform(b200_fb_plan* pl, const float* d_in, uint64_t in_span, float* d_out, uint64_t out_span,
                    uint64_t npart, uint64_t in_step, uint64_t out_step) {
  B200_REQUIRE(pl && d_in && d_out, "b200_fb_perform: null argument");
  if (npart == 0) return B200_OK;
  const unsigned ndim = pl->desc.input_real ? 1 : 2;
  B200_REQUIRE(in_step == uint64_t(pl->nsamp_step) * ndim, "in_step=%llu != nsamp_step*ndim=%llu (Filterbank.C:517)",
               (unsigned long long)in_step, (unsigned long long)(uint64_t(pl->nsamp_step) * ndim));
  B200_REQUIRE(out_step == uint64_t(pl->nkeep) * 2, "out_step=%llu != nkeep*2=%llu (Filterbank.C:523)",
               (unsigned long long)out_step, (unsigned long long)(uint64_t(pl->nkeep) * 2));
  B200_REQUIRE(in_span % 2 == 0 && out_span % 2 == 0 && (reinterpret_cast<uintptr_t>(d_in) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(d_out) & 7) == 0,
               "time series planes must be 8-byte aligned with even spans");
  B200_REQUIRE(in_step % 2 == 0, "nsamp_step must be even for real input");
  FbSource src;
  memset(&src, 0, sizeof(src));
  src.kind = SRC_F32; src.ptr = d_in; src.span = in_span; src.step = in_step; src.d_lut = nullptr; src.batch_ready = nullptr; src.conv_ok = 0; src.batch_override = 0;
  FbSink sink;
  memset(&sink, 0, sizeof(sink));
  sink.kind = EPI_VOLT; sink.volt = d_out; sink.volt_span = out_span; sink.volt_step = out_step;
  return fb_run(pl, src, sink, npart);
}

}  // extern "C"
