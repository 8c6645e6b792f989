// fold.cu -- stand-alone unpack, detection and fold engines + the bin-plan expansion kernel.
//
//   b200_unpack   <- Unpacker device hook (CASPSR / generic 8-bit / MeerKAT / UWB)
//   b200_detect   <- dsp::Detection::Engine (Signal/General/dsp/Detection.h:98-106)
//   b200_fold_*   <- dsp::Fold::Engine      (Signal/Pulsar/dsp/Fold.h:249-312)
#include <cstring>
#include <vector>

#include "engine.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// unpack kernels: one thread per output (chan,pol) sample group; writes are coalesced along
// time, reads gather through the read-only path.
// ------------------------------------------------------------------------------------------
struct UnpackArgs {
  const unsigned char* raw;
  float* out;
  uint64_t span, ndat;
  unsigned nchan, npol, ndim;
  float scale;
  unsigned sample_swap;
};

// CASPSR: 4 samples of pol0 then 4 samples of pol1 (CASPSRUnpacker.C:141-187); LUT in smem.
__global__ void k_unpack_caspsr(UnpackArgs a, const float* __restrict__ lut) {
  __shared__ float s_lut[256];
  for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  // each thread converts one 8-byte group: 4 floats to each polarisation plane
  const uint64_t ngroup = a.ndat / 4;
  for (uint64_t g = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; g < ngroup; g += uint64_t(gridDim.x) * blockDim.x) {
    uint2 w = __ldg(reinterpret_cast<const uint2*>(a.raw) + g);
    float4 p0 = make_float4(s_lut[w.x & 255u], s_lut[(w.x >> 8) & 255u], s_lut[(w.x >> 16) & 255u], s_lut[w.x >> 24]);
    float4 p1 = make_float4(s_lut[w.y & 255u], s_lut[(w.y >> 8) & 255u], s_lut[(w.y >> 16) & 255u], s_lut[w.y >> 24]);
    reinterpret_cast<float4*>(a.out)[g] = p0;
    reinterpret_cast<float4*>(a.out + a.span)[g] = p1;
  }
}

// generic TFP 8-bit (BitUnpacker.C:56-75): byte idat*(nchan*npol*ndim) + ndim*(npol*ichan+ipol)+idim
__global__ void k_unpack_generic8(UnpackArgs a, const float* __restrict__ lut) {
  __shared__ float s_lut[256];
  for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const unsigned nplane = a.nchan * a.npol;
  const unsigned nskip = nplane * a.ndim;
  const uint64_t total = a.ndat * nskip;
  // consecutive threads read consecutive bytes (coalesced); the scattered 4-byte writes land in
  // nskip different planes and are merged by L2.
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t idat = i / nskip;
    const unsigned off = unsigned(i % nskip);
    const unsigned plane = off / a.ndim, idim = off % a.ndim;
    a.out[uint64_t(plane) * a.span + idat * a.ndim + idim] = s_lut[__ldg(a.raw + i)];
  }
}

// MeerKAT (MeerKATUnpacker.C:196-229): heaps of 256 samples, [heap][pol][chan][256 x (re,im) int8]
__global__ void k_unpack_meerkat(UnpackArgs a) {
  const uint64_t total = a.ndat * a.nchan * a.npol;   // complex samples
  const char2* from = reinterpret_cast<const char2*>(a.raw);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const unsigned isamp = unsigned(i & 255u);
    uint64_t r = i >> 8;
    const unsigned ichan = unsigned(r % a.nchan); r /= a.nchan;
    const unsigned ipol = unsigned(r % a.npol);
    const uint64_t iheap = r / a.npol;
    // sample_swap == 2 (MKBFRo) exchanges odd and even samples (MeerKATUnpacker.C:211-222)
    const unsigned osamp = (a.sample_swap == 2) ? (isamp ^ 1u) : isamp;
    char2 v = from[i];
    float2 o;
    o.x = __fmul_rn(float(v.x) + 0.5f, a.scale);
    o.y = __fmul_rn(float(v.y) + 0.5f, a.scale);
    float2* into = reinterpret_cast<float2*>(a.out + (uint64_t(ichan) * a.npol + ipol) * a.span) + iheap * 256 + osamp;
    *into = o;
  }
}

// UWB (UWBUnpacker.C:177-218): blocks of 2048 complex int16 samples per pol, offset binary
__global__ void k_unpack_uwb(UnpackArgs a) {
  const uint64_t total = a.ndat * a.npol;   // complex samples
  const short2* from = reinterpret_cast<const short2*>(a.raw);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const unsigned isamp = unsigned(i & 2047u);
    uint64_t r = i >> 11;
    const unsigned ipol = unsigned(r % a.npol);
    const uint64_t iblock = r / a.npol;
    short2 v = from[i];
    float2 o;
    o.x = float(short(v.x ^ short(0x8000)));
    o.y = float(short(v.y ^ short(0x8000)));
    reinterpret_cast<float2*>(a.out + uint64_t(ipol) * a.span)[iblock * 2048 + isamp] = o;
  }
}


// Two-bit excision unpacker (TwoBitCorrection / ExcisionUnpacker / TwoBitFour, see b200dsp.h):
// one warp per window of 512 samples.  Lane l owns bytes [4l, 4l+4) of every digitizer (bytes of the
// npol digitizers are interleaved), i.e. samples [16l, 16l+16) of the window.  Pass 1 counts the
// low-voltage states with a bit trick (a 2-bit code is "low" iff bit (lowsel >> code) & 1) and reduces
// over the warp; pass 2 writes sign * (low ? lo : hi)[clamped nlow] or zeros.
struct TwoBitArgs {
  const unsigned char* raw;
  float* out;
  uint64_t span, nwindow;
  unsigned npol, nlow_min, nlow_max;
  unsigned lowsel;      // bit c set: code c is a low-voltage state
  unsigned negsel;      // bit c set: code c is negative
  const float* levels;  // lo[513] then hi[513]
  unsigned* weights;
};

template <unsigned NPOL>
__global__ void k_unpack_twobit(TwoBitArgs a) {
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t warp0 = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
  const uint64_t nwarp = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t w = warp0; w < a.nwindow; w += nwarp) {
    // 4*NPOL consecutive bytes: byte b belongs to digitizer b % NPOL (ExcisionUnpacker.C:258-266)
    unsigned words[NPOL];
    const unsigned* src = reinterpret_cast<const unsigned*>(a.raw + (w * 128 + 4 * lane) * NPOL);
#pragma unroll
    for (unsigned i = 0; i < NPOL; i++) words[i] = __ldg(src + i);
    unsigned bytes[NPOL][4];
#pragma unroll
    for (unsigned b = 0; b < 4 * NPOL; b++) bytes[b % NPOL][b / NPOL] = (words[b / 4] >> (8 * (b % 4))) & 255u;
    bool zero_any = false;
#pragma unroll
    for (unsigned p = 0; p < NPOL; p++) {
      unsigned nlow = 0, any = 0;
#pragma unroll
      for (unsigned k = 0; k < 4; k++) {
        const unsigned byte = bytes[p][k];
        any |= byte;
#pragma unroll
        for (unsigned s4 = 0; s4 < 4; s4++) nlow += (a.lowsel >> ((byte >> (6 - 2 * s4)) & 3u)) & 1u;
      }
      nlow = __reduce_add_sync(0xffffffffu, nlow);
      any = __reduce_or_sync(0xffffffffu, any);
      // excision_unpack.h:79-97: all-zero bytes, or nlow outside the limits
      const bool bad = (any == 0) || nlow < a.nlow_min || nlow > a.nlow_max;
      zero_any |= bad;
      const unsigned row = min(max(nlow, a.nlow_min), a.nlow_max) - a.nlow_min;   // TwoBitFour.h:68-75
      const float lo = bad ? 0.f : __ldg(a.levels + row);
      const float hi = bad ? 0.f : __ldg(a.levels + 513 + row);
      float4* dst = reinterpret_cast<float4*>(a.out + uint64_t(p) * a.span + w * 512 + 16 * lane);
#pragma unroll
      for (unsigned k = 0; k < 4; k++) {
        const unsigned byte = bytes[p][k];
        float v[4];
#pragma unroll
        for (unsigned s4 = 0; s4 < 4; s4++) {
          const unsigned code = (byte >> (6 - 2 * s4)) & 3u;            // BitTable.C:154-163 MostToLeast
          const float mag = ((a.lowsel >> code) & 1u) ? lo : hi;
          v[s4] = ((a.negsel >> code) & 1u) ? -mag : mag;
        }
        dst[k] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    if (a.weights && lane == 0) a.weights[w] = zero_any ? 0u : 1u;     // WeightedTimeSeries::mask_weights
  }
}

// The per-window half of the two-bit unpacker alone: for every 512-sample window and digitizer the pair of output
// levels (lo, hi) the reference would use (0, 0 for an excised window) and the window's weight.  The filterbank's
// column kernel then converts the 2-bit codes itself (filterbank.cu, SRC_TWOBIT): no float time series is written.
template <unsigned NPOL>
__global__ void k_twobit_windows(TwoBitArgs a, float2* __restrict__ win) {
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t warp0 = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
  const uint64_t nwarp = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t w = warp0; w < a.nwindow; w += nwarp) {
    unsigned words[NPOL];
    const unsigned* src = reinterpret_cast<const unsigned*>(a.raw + (w * 128 + 4 * lane) * NPOL);
#pragma unroll
    for (unsigned i = 0; i < NPOL; i++) words[i] = __ldg(src + i);
    bool zero_any = false;
#pragma unroll
    for (unsigned p = 0; p < NPOL; p++) {
      unsigned nlow = 0, any = 0;
#pragma unroll
      for (unsigned k = 0; k < 4; k++) {
        const unsigned b = k * NPOL + p;                                  // byte b of the group belongs to digitizer b % NPOL
        const unsigned byte = (words[b / 4] >> (8 * (b % 4))) & 255u;
        any |= byte;
#pragma unroll
        for (unsigned s4 = 0; s4 < 4; s4++) nlow += (a.lowsel >> ((byte >> (6 - 2 * s4)) & 3u)) & 1u;
      }
      nlow = __reduce_add_sync(0xffffffffu, nlow);
      any = __reduce_or_sync(0xffffffffu, any);
      const bool bad = (any == 0) || nlow < a.nlow_min || nlow > a.nlow_max;   // excision_unpack.h:79-97
      zero_any |= bad;
      const unsigned row = min(max(nlow, a.nlow_min), a.nlow_max) - a.nlow_min;   // TwoBitFour.h:68-75
      if (lane == 0)
        win[w * NPOL + p] = bad ? make_float2(0.f, 0.f) : make_float2(__ldg(a.levels + row), __ldg(a.levels + 513 + row));
    }
    if (a.weights && lane == 0) a.weights[w] = zero_any ? 0u : 1u;     // WeightedTimeSeries::mask_weights
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone detection (Detection.C:218-320,322-421)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void detect4(int state, float2 p, float2 q, float* r) {
  float pp = __fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y));
  float qq = __fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y));
  float re = __fadd_rn(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y));
  float im = __fsub_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x));
  if (state == B200_INTENSITY) { r[0] = __fadd_rn(pp, qq); }
  else if (state == B200_PPQQ) { r[0] = pp; r[1] = qq; }
  else if (state == B200_COHERENCE) { r[0] = pp; r[1] = qq; r[2] = re; r[3] = im; }
  else { r[0] = __fadd_rn(pp, qq); r[1] = __fsub_rn(pp, qq); r[2] = __fmul_rn(2.f, re); r[3] = __fmul_rn(2.f, im); }
}

struct DetectArgs {
  const float* in;
  float* out;
  uint64_t in_span, out_span, ndat;
  unsigned nchan, npol, ndim_out, nprod;
  int state;
};

__global__ void k_detect(DetectArgs a) {
  const unsigned ichan = blockIdx.y;
  const float2* p = reinterpret_cast<const float2*>(a.in + uint64_t(ichan) * a.npol * a.in_span);
  const float2* q = a.npol > 1 ? reinterpret_cast<const float2*>(a.in + (uint64_t(ichan) * a.npol + 1) * a.in_span) : p;
  const unsigned dnpol = a.nprod / a.ndim_out;
  float* obase = a.out + uint64_t(ichan) * dnpol * a.out_span;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.ndat; i += uint64_t(gridDim.x) * blockDim.x) {
    float2 pv = p[i];
    float2 qv = a.npol > 1 ? q[i] : make_float2(0.f, 0.f);
    float r[4];
    detect4(a.state, pv, qv, r);
    // both reads happen before any write of this sample: safe in place for ndim_out == 2
    if (a.ndim_out == 4) {
      reinterpret_cast<float4*>(obase)[i] = make_float4(r[0], r[1], r[2], r[3]);
    } else if (a.ndim_out == 2) {
      reinterpret_cast<float2*>(obase)[i] = make_float2(r[0], r[1]);
      reinterpret_cast<float2*>(obase + a.out_span)[i] = make_float2(r[2], r[3]);
    } else {
      for (unsigned pr = 0; pr < a.nprod; pr++) obase[uint64_t(pr) * a.out_span + i] = r[pr];
    }
  }
}

// ------------------------------------------------------------------------------------------
// bin-plan expansion: segments -> bin of every sample (+ hits).  Exact integer arithmetic.
// ------------------------------------------------------------------------------------------
// Weighted input (Fold.C:687-716,746-763): sample idat belongs to window (idat + weight_idat) / ndatperweight; samples of
// a window whose flag is 0 get bin = nbin (every fold consumer skips that value) and give no hit.
struct BinWeights {
  const unsigned* w;         // null: unweighted input
  uint64_t nweights, weight_idat, idat_start;
  unsigned ndatperweight;
};

__global__ void k_expand_bins(const b200_phase_segment* __restrict__ seg, unsigned nseg, uint64_t ndat, unsigned nbin,
                              unsigned* __restrict__ bins, unsigned* __restrict__ hits_last,
                              unsigned* __restrict__ hits_total, BinWeights bw, unsigned iters) {
  const double double_nbin = double(nbin);
  // A warp takes chunks of 32 x iters consecutive samples, 32 at a time (coalesced stores; iters = 32 for long blocks,
  // fewer when the block would not fill the machine: cfg5's sub-bands).  Hits: the lowest lane holding
  // a bin in an iteration counts its peers, and keeps counting for as long as it leads the same bin -- one pair of
  // atomics per bin and chunk when bins are wide (cfg4: 35 thousand samples per bin; one pair per warp and iteration
  // made the kernel atomic-bound), at worst one pair per distinct bin of an iteration as before.
  const unsigned CHUNK = 32u * iters;
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5, nwarp = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t c0 = warp * CHUNK; c0 < ndat; c0 += nwarp * CHUNK) {           // warp-uniform trip count
    unsigned my_bin = 0xffffffffu, my_n = 0;
    // the lane's samples are 32 apart: inside a segment its phase numerator advances by 32 steps per iteration, and the
    // segment is looked up again only when the lane leaves it (142 warp instructions per 32 samples before, mostly the
    // search and ldexp)
    uint64_t seg_end = 0, a = 0, a_step = 0;
    double scale = 0.0;
    int sexp = 0;
    for (unsigned it = 0; it < iters; it++) {
      const uint64_t i = c0 + it * 32u + lane;
      unsigned ibin = 0xfffffffeu;                                              // beyond the end: counted by nobody
      if (i < ndat) {
        if (i >= seg_end) {
          // binary search for the segment containing sample i
          unsigned lo = 0, hi = nseg - 1;
          while (lo < hi) {
            unsigned mid = (lo + hi + 1) >> 1;
            if (seg[mid].start <= i) lo = mid;
            else hi = mid - 1;
          }
          const b200_phase_segment s = seg[lo];
          a = s.a0 + (i - s.start) * s.step;                         // < 2^53: exact in double
          a_step = 32u * s.step;
          seg_end = s.start + s.count;
          sexp = s.scale_exp;
          scale = sexp >= -1000 ? ldexp(1.0, sexp) : 0.0;            // 0: phases below 2^-900, scaled by ldexp itself
        } else a += a_step;
        // exact scaling by a power of two (the product of a 53-bit integer and 2^sexp is a normal double)
        const double phi = scale != 0.0 ? __dmul_rn(double(a), scale) : ldexp(double(a), sexp);
        const double double_ibin = __dmul_rn(phi, double_nbin);      // Fold.C:766
        ibin = unsigned(double_ibin);                                // Fold.C:767 (truncation)
        if (bw.w) {
          const uint64_t iw = (bw.idat_start + i + bw.weight_idat) / bw.ndatperweight;
          if (iw >= bw.nweights || bw.w[iw] == 0u) ibin = nbin;      // bad_data: binplan = folding_nbin (Fold.C:773-774)
        }
        bins[i] = ibin;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, ibin);
      if (ibin < nbin && lane == unsigned(__ffs(peers) - 1)) {
        const unsigned n = __popc(peers);
        if (ibin == my_bin) my_n += n;
        else {
          if (my_n) {
            atomicAdd(hits_last + my_bin, my_n);
            atomicAdd(hits_total + my_bin, my_n);
          }
          my_bin = ibin;
          my_n = n;
        }
      }
      if (c0 + (it + 1u) * 32u >= ndat) break;                                  // warp-uniform
    }
    if (my_n) {
      atomicAdd(hits_last + my_bin, my_n);
      atomicAdd(hits_total + my_bin, my_n);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fold work items of the bin plan, per part of nkeep samples: a new item starts wherever the phase bin changes and
// wherever (sample + align) is a multiple of 16, so an item is at most 16 consecutive samples of ONE bin inside one
// 16-aligned block of the consumer's staging buffer (align = nfilt_pos mod 16: K3 stages by transform index).  runs[part][r] = (first sample relative to the part, bin), runs[part][nruns] = (nkeep, 0).  The fused fold
// epilogue (fastpath.cu) gives every item to one thread: one sequential sum and one RED per product, no per-sample
// bin comparisons, balanced for wide bins (items of 16) and narrow ones alike.  One CTA per part.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bin_runs(const unsigned* __restrict__ bins, unsigned nkeep, unsigned align,
                                                   uint2* __restrict__ runs, unsigned* __restrict__ nruns) {
  __shared__ unsigned wsum[32];
  __shared__ unsigned s_total;
  const unsigned part = blockIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned* b = bins + uint64_t(part) * nkeep;
  uint2* r = runs + uint64_t(part) * (nkeep + 1);
  const unsigned per = (nkeep + 1023u) / 1024u;
  const unsigned t0 = min(nkeep, threadIdx.x * per), t1 = min(nkeep, t0 + per);
  unsigned cnt = 0;
  for (unsigned t = t0; t < t1; t++) cnt += (t == 0 || ((t + align) & 15u) == 0 || b[t] != b[t - 1]) ? 1u : 0u;
  unsigned incl = cnt;
  for (unsigned o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    const unsigned v = wsum[lane];
    unsigned iv = v;
    for (unsigned o = 1; o < 32; o <<= 1) {
      const unsigned x = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += x;
    }
    wsum[lane] = iv - v;
    if (lane == 31) s_total = iv;
  }
  __syncthreads();
  unsigned pos = wsum[w] + incl - cnt;
  for (unsigned t = t0; t < t1; t++)
    if (t == 0 || ((t + align) & 15u) == 0 || b[t] != b[t - 1]) r[pos++] = make_uint2(t, b[t]);
  if (threadIdx.x == 0) {
    nruns[part] = s_total;
    r[s_total] = make_uint2(nkeep, 0u);
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone fold (Fold.C:835-873): one CTA per (chan, pol, time slab); per-thread runs of
// consecutive samples summed sequentially, then added to shared bins, then to the profile.
// ------------------------------------------------------------------------------------------
struct FoldArgs {
  const float* in;
  uint64_t in_span, idat_start, ndat;
  const unsigned* bins;
  float* profile;
  uint64_t prof_span;            // floats between the planes of the profile (nbin*ndim when the handle owns it)
  unsigned nchan, npol, ndim, nbin, slab;
  unsigned smem_bins;
  long long* fix;                // deterministic mode: fixed-point accumulator (engine.cuh profile_add), else null
  float inv_lsb;
};

template <int NDIM>
__global__ void k_fold(FoldArgs a) {
  extern __shared__ float sbins[];
  const unsigned plane = blockIdx.y;               // ichan*npol + ipol
  const uint64_t s0 = uint64_t(blockIdx.x) * a.slab;
  const uint64_t s1 = min(a.ndat, s0 + a.slab);
  const float* tp = a.in + uint64_t(plane) * a.in_span + a.idat_start * NDIM;
  float* prof = a.profile + uint64_t(plane) * a.prof_span;
  long long* fixp = a.fix ? a.fix + uint64_t(plane) * a.prof_span : nullptr;
  if (a.smem_bins) {
    for (unsigned i = threadIdx.x; i < a.nbin * NDIM; i += blockDim.x) sbins[i] = 0.f;
    __syncthreads();
  }
  float* dst = a.smem_bins ? sbins : prof;
  auto add = [&](uint64_t idx, float v) {
    if (fixp) profile_add(prof, fixp, a.inv_lsb, idx, v);     // straight to the fixed-point accumulator
    else atomicAdd(dst + idx, v);
  };
  const unsigned L = 16;
  for (uint64_t c0 = s0 + uint64_t(threadIdx.x) * L; c0 < s1; c0 += uint64_t(blockDim.x) * L) {
    const uint64_t c1 = min(s1, c0 + L);
    float acc[NDIM];
    unsigned cur = 0xffffffffu;
    for (uint64_t i = c0; i < c1; i++) {
      const unsigned bin = __ldg(a.bins + i);
      float v[NDIM];
      if (NDIM == 4) {
        float4 t = reinterpret_cast<const float4*>(tp)[i];
        v[0] = t.x; v[1 % NDIM] = t.y; v[2 % NDIM] = t.z; v[3 % NDIM] = t.w;
      } else if (NDIM == 2) {
        float2 t = reinterpret_cast<const float2*>(tp)[i];
        v[0] = t.x; v[1 % NDIM] = t.y;
      } else {
        v[0] = tp[i];
      }
      if (bin != cur) {
        if (cur < a.nbin)                                    // cur == nbin: samples of a flagged window, dropped
          for (int d = 0; d < NDIM; d++) add(uint64_t(cur) * NDIM + d, acc[d]);
        cur = bin;
        for (int d = 0; d < NDIM; d++) acc[d] = v[d];
      } else {
        for (int d = 0; d < NDIM; d++) acc[d] += v[d];
      }
    }
    if (cur < a.nbin)
      for (int d = 0; d < NDIM; d++) add(uint64_t(cur) * NDIM + d, acc[d]);
  }
  if (a.smem_bins) {
    __syncthreads();
    for (unsigned i = threadIdx.x; i < a.nbin * NDIM; i += blockDim.x) {
      float v = sbins[i];
      if (v != 0.f) atomicAdd(prof + i, v);
    }
  }
}

}  // namespace b200

using namespace b200;

// fixed-point accumulator -> float profile (deterministic mode)
__global__ void k_fix_to_float(const long long* __restrict__ fix, float lsb, uint64_t n, float* __restrict__ out) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    out[i] = float(double(fix[i]) * double(lsb));
}

struct b200_fold {
  Context* ctx;
  unsigned nchan, npol, ndim, nbin;
  long long* d_fix;              // deterministic mode (b200_fold_set_deterministic): the accumulator; d_profile is
  float lsb;                     // refreshed from it whenever it is read
  float* d_profile;
  unsigned* d_hits_total;
  unsigned* d_hits_last;
  unsigned* d_bins;
  uint64_t bins_capacity;
  uint2* d_runs;                 // run table of the last set_bins (fold_build_runs), [npart][nkeep+1] (start, bin)
  unsigned* d_nruns;             // [npart]
  uint64_t runs_capacity, nruns_capacity;
  b200_phase_segment* d_seg;
  b200_phase_segment* h_seg;     // pinned
  uint64_t seg_capacity;
  uint64_t ndat, idat_start;     // of the last set_bins
  uint64_t ndat_total;
  cudaEvent_t seg_free;          // the pinned segment buffer may be rewritten after this event
  bool seg_pending;
  bool weighted;                 // some set_bins since the last zero skipped flagged samples: ndat_folded = sum of hits
};

extern "C" {

int b200_unpack(b200_context* cctx, const b200_unpack_desc* d, const void* d_raw, uint64_t ndat, float* d_out,
                uint64_t out_span) {
  B200_REQUIRE(cctx && d && d_raw && d_out, "b200_unpack: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  if (ndat == 0) return B200_OK;
  UnpackArgs a;
  a.raw = static_cast<const unsigned char*>(d_raw);
  a.out = d_out; a.span = out_span; a.ndat = ndat;
  a.nchan = d->nchan; a.npol = d->npol; a.ndim = d->ndim;
  a.scale = d->scale; a.sample_swap = d->sample_swap ? d->sample_swap : 1;
  if (d->format == B200_FMT_TWOBIT) {
    B200_REQUIRE(d->twobit, "b200_unpack: TWOBIT format needs desc->twobit");
    B200_REQUIRE(d->nchan == 1 && d->ndim == 1 && d->npol == d->twobit->npol,
                 "two-bit unpacker: nchan=1, ndim=1, npol matching the table");
    return b200_unpack_twobit(cctx, d->twobit, d_raw, ndat, d_out, out_span, nullptr);
  }
  const unsigned threads = 256;
  const unsigned maxgrid = ctx->sm_count * 16;
  float* d_lut = nullptr;
  if (d->format == B200_FMT_CASPSR8 || d->format == B200_FMT_GENERIC8) {
    // the table travels with the call: a stream-ordered 1 KiB copy into the context's table area (no allocation; the
    // copy of the next call is ordered behind this call's kernel)
    d_lut = ctx->d_tables;
    B200_CUDA(cudaMemcpyAsync(d_lut, d->lut, 256 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  LaunchScope ls(ctx, KC_OTHER);
  switch (d->format) {
    case B200_FMT_CASPSR8: {
      B200_REQUIRE(d->nchan == 1 && d->npol == 2 && d->ndim == 1, "CASPSR unpacker: nchan=1 npol=2 ndim=1 only");
      B200_REQUIRE(ndat % 4 == 0 && out_span % 4 == 0, "CASPSR unpacker: ndat and span must be multiples of 4");
      uint64_t ng = ndat / 4;
      unsigned grid = (unsigned)std::min<uint64_t>((ng + threads - 1) / threads, maxgrid);
      k_unpack_caspsr<<<grid, threads, 0, ctx->stream>>>(a, d_lut);
      break;
    }
    case B200_FMT_GENERIC8: {
      uint64_t total = ndat * d->nchan * d->npol * d->ndim;
      unsigned grid = (unsigned)std::min<uint64_t>((total + threads - 1) / threads, maxgrid);
      k_unpack_generic8<<<grid, threads, 0, ctx->stream>>>(a, d_lut);
      break;
    }
    case B200_FMT_MEERKAT8: {
      B200_REQUIRE(d->ndim == 2, "MeerKAT unpacker: ndim=2 only (MeerKATUnpacker.C:137-144)");
      B200_REQUIRE(ndat % 256 == 0 && out_span % 2 == 0, "MeerKAT unpacker: ndat must be a multiple of the 256-sample heap");
      uint64_t total = ndat * d->nchan * d->npol;
      unsigned grid = (unsigned)std::min<uint64_t>((total + threads - 1) / threads, maxgrid);
      k_unpack_meerkat<<<grid, threads, 0, ctx->stream>>>(a);
      break;
    }
    case B200_FMT_UWB16: {
      B200_REQUIRE(d->nchan == 1 && d->ndim == 2, "UWB unpacker: nchan=1 ndim=2 only (UWBUnpacker.C:140-147)");
      B200_REQUIRE(ndat % 2048 == 0 && out_span % 2 == 0, "UWB unpacker: ndat must be a multiple of the 2048-sample block");
      uint64_t total = ndat * d->npol;
      unsigned grid = (unsigned)std::min<uint64_t>((total + threads - 1) / threads, maxgrid);
      k_unpack_uwb<<<grid, threads, 0, ctx->stream>>>(a);
      break;
    }
    default:
      set_error("b200_unpack: unknown format %d", d->format);
      return B200_ERR_INVALID;
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_unpack_twobit(b200_context* cctx, const b200_twobit_desc* d, const void* d_raw, uint64_t ndat, float* d_out,
                       uint64_t out_span, unsigned* d_weights) {
  B200_REQUIRE(cctx && d && d_raw && d_out, "b200_unpack_twobit: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  B200_REQUIRE(d->npol == 1 || d->npol == 2, "two-bit unpacker: npol=%u (1 or 2 digitizers)", d->npol);
  B200_REQUIRE(d->table_type >= 0 && d->table_type <= 2, "two-bit unpacker: unknown table type %d", d->table_type);
  if (d->ndat_per_weight != 512) {
    set_error("two-bit unpacker: ndat_per_weight=%u (only the reference default 512 is built)", d->ndat_per_weight);
    return B200_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(ndat % 512 == 0, "two-bit unpacker: ndat=%llu is not a multiple of ndat_per_weight=512 "
               "(ExcisionUnpacker::get_resolution)", (unsigned long long)ndat);
  B200_REQUIRE(out_span % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(d_raw) & 3) == 0, "two-bit unpacker: unaligned buffers");
  B200_REQUIRE(d->nlow_min <= d->nlow_max && d->nlow_max <= 512, "two-bit unpacker: invalid nlow limits");
  if (ndat == 0) return B200_OK;
  float* d_levels = ctx->d_tables + 256;           // behind the 8-bit table; 2 x 513 floats (no allocation)
  B200_CUDA(cudaMemcpyAsync(d_levels, d->lo, 513 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(cudaMemcpyAsync(d_levels + 513, d->hi, 513 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  TwoBitArgs a;
  a.raw = static_cast<const unsigned char*>(d_raw);
  a.out = d_out; a.span = out_span; a.nwindow = ndat / 512; a.npol = d->npol;
  a.nlow_min = d->nlow_min; a.nlow_max = d->nlow_max; a.levels = d_levels; a.weights = d_weights;
  // TwoBitTable::generate_unique_values (TwoBitTable.C:42-75): which codes are low / negative
  static const unsigned lowsel[3] = {0x6u /* codes 1,2 */, 0x5u /* 0,2 */, 0x9u /* 0,3 */};
  static const unsigned negsel[3] = {0x3u /* codes 0,1 */, 0xcu /* 2,3 */, 0xcu /* 2,3 */};
  a.lowsel = lowsel[d->table_type];
  a.negsel = negsel[d->table_type];
  const unsigned threads = 256;
  unsigned grid = (unsigned)std::min<uint64_t>((a.nwindow * 32 + threads - 1) / threads, uint64_t(ctx->sm_count) * 16);
  {
    LaunchScope ls(ctx, KC_OTHER);
    if (d->npol == 1) k_unpack_twobit<1><<<grid, threads, 0, ctx->stream>>>(a);
    else k_unpack_twobit<2><<<grid, threads, 0, ctx->stream>>>(a);
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // extern "C"

namespace b200 {
// level pairs + weights of `nwindow` windows starting at d_raw (a 512-sample boundary); tables go through the
// context's table area like b200_unpack_twobit
int twobit_windows(Context* ctx, const b200_twobit_desc* d, const void* d_raw, uint64_t nwindow, float2* d_win,
                   unsigned* d_weights, unsigned* lowsel_out, unsigned* negsel_out) {
  static const unsigned lowsel[3] = {0x6u, 0x5u, 0x9u};
  static const unsigned negsel[3] = {0x3u, 0xcu, 0xcu};
  B200_REQUIRE(d->table_type >= 0 && d->table_type <= 2, "two-bit unpacker: unknown table type %d", d->table_type);
  *lowsel_out = lowsel[d->table_type];
  *negsel_out = negsel[d->table_type];
  if (nwindow == 0) return B200_OK;
  float* d_levels = ctx->d_tables + 256;
  B200_CUDA(cudaMemcpyAsync(d_levels, d->lo, 513 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(cudaMemcpyAsync(d_levels + 513, d->hi, 513 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  TwoBitArgs a;
  a.raw = static_cast<const unsigned char*>(d_raw);
  a.out = nullptr; a.span = 0; a.nwindow = nwindow; a.npol = d->npol;
  a.nlow_min = d->nlow_min; a.nlow_max = d->nlow_max; a.levels = d_levels; a.weights = d_weights;
  a.lowsel = *lowsel_out;
  a.negsel = *negsel_out;
  const unsigned threads = 256;
  unsigned grid = (unsigned)std::min<uint64_t>((nwindow * 32 + threads - 1) / threads, uint64_t(ctx->sm_count) * 16);
  LaunchScope ls(ctx, KC_OTHER);
  if (d->npol == 1) k_twobit_windows<1><<<grid, threads, 0, ctx->stream>>>(a, d_win);
  else k_twobit_windows<2><<<grid, threads, 0, ctx->stream>>>(a, d_win);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}
}  // namespace b200

extern "C" {

int b200_detect(b200_context* cctx, int state, unsigned ndim_out, const float* d_in, uint64_t in_span, unsigned nchan,
                unsigned npol, uint64_t ndat, float* d_out, uint64_t out_span) {
  B200_REQUIRE(cctx && d_in && d_out, "b200_detect: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  B200_REQUIRE(state >= 0 && state <= 3, "b200_detect: invalid state %d", state);
  if (state >= B200_COHERENCE) {
    // Detection::checks (Detection.C:476-489)
    B200_REQUIRE(npol == 2, "b200_detect: Coherence/Stokes need npol == 2, have %u", npol);
    B200_REQUIRE(ndim_out == 1 || ndim_out == 2 || ndim_out == 4, "b200_detect: invalid ndim=%u", ndim_out);
  } else {
    B200_REQUIRE(npol == 1 || npol == 2, "b200_detect: npol=%u", npol);
    ndim_out = 1;
  }
  B200_REQUIRE(!(d_in == d_out) || (state >= B200_COHERENCE && ndim_out == 2),
               "b200_detect: in-place detection only for ndim == 2 (Detection.C:358-361)");
  B200_REQUIRE(in_span % 2 == 0, "b200_detect: odd input span");
  if (ndat == 0) return B200_OK;
  DetectArgs a;
  a.in = d_in; a.out = d_out; a.in_span = in_span; a.out_span = out_span; a.ndat = ndat;
  a.nchan = nchan; a.npol = npol; a.ndim_out = ndim_out; a.state = state;
  a.nprod = state == B200_INTENSITY ? 1 : state == B200_PPQQ ? npol : 4;
  const unsigned threads = 256;
  unsigned gx = (unsigned)std::min<uint64_t>((ndat + threads - 1) / threads, 4096);
  LaunchScope ls(ctx, KC_OTHER);
  k_detect<<<dim3(gx, nchan), threads, 0, ctx->stream>>>(a);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_fold_create(b200_context* cctx, unsigned nchan, unsigned npol, unsigned ndim, unsigned nbin, b200_fold** out) {
  B200_REQUIRE(cctx && out, "b200_fold_create: null argument");
  B200_REQUIRE(nchan && npol && nbin && (ndim == 1 || ndim == 2 || ndim == 4), "b200_fold_create: invalid shape");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  B200_CUDA(cudaSetDevice(ctx->device));
  b200_fold* f = new b200_fold();
  memset(f, 0, sizeof(*f));
  f->ctx = ctx; f->nchan = nchan; f->npol = npol; f->ndim = ndim; f->nbin = nbin;
  const uint64_t nfloat = uint64_t(nchan) * npol * ndim * nbin;
  cudaError_t e = cudaMalloc(&f->d_profile, nfloat * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_hits_total, nbin * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_hits_last, nbin * sizeof(unsigned));
  f->seg_capacity = 1 << 16;
  if (e == cudaSuccess) e = cudaMalloc(&f->d_seg, f->seg_capacity * sizeof(b200_phase_segment));
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_seg, f->seg_capacity * sizeof(b200_phase_segment));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->seg_free, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_profile, 0, nfloat * sizeof(float), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_hits_total, 0, nbin * sizeof(unsigned), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_hits_last, 0, nbin * sizeof(unsigned), ctx->stream);
  if (e != cudaSuccess) {
    b200_fold_destroy(f);
    return cuda_fail(e, "b200_fold_create", __FILE__, __LINE__);
  }
  *out = f;
  return B200_OK;
}

int b200_fold_destroy(b200_fold* f) {
  if (!f) return B200_OK;
  if (f->d_profile) cudaFree(f->d_profile);
  if (f->d_fix) cudaFree(f->d_fix);
  if (f->d_hits_total) cudaFree(f->d_hits_total);
  if (f->d_hits_last) cudaFree(f->d_hits_last);
  if (f->d_bins) cudaFree(f->d_bins);
  if (f->d_runs) cudaFree(f->d_runs);
  if (f->d_nruns) cudaFree(f->d_nruns);
  if (f->d_seg) cudaFree(f->d_seg);
  if (f->h_seg) cudaFreeHost(f->h_seg);
  if (f->seg_free) cudaEventDestroy(f->seg_free);
  delete f;
  return B200_OK;
}

static int fold_set_bins(b200_fold* f, double phi, double pps, uint64_t ndat, uint64_t idat_start, uint64_t* ndat_folded,
                         const BinWeights& bw);

int b200_fold_set_bins(b200_fold* f, double phi, double pps, uint64_t ndat, uint64_t idat_start, uint64_t* ndat_folded) {
  BinWeights bw;
  memset(&bw, 0, sizeof bw);
  return fold_set_bins(f, phi, pps, ndat, idat_start, ndat_folded, bw);
}

int b200_fold_set_bins_weighted(b200_fold* f, double phi, double pps, uint64_t ndat, uint64_t idat_start,
                                const unsigned* d_weights, uint64_t nweights, unsigned ndatperweight, uint64_t weight_idat) {
  B200_REQUIRE(f, "b200_fold_set_bins_weighted: null fold");
  BinWeights bw;
  memset(&bw, 0, sizeof bw);
  if (d_weights && ndatperweight) {
    // Fold.C:693-706: the first and the last window of the block must exist
    const uint64_t last = ndat ? (idat_start + ndat - 1 + weight_idat) / ndatperweight : 0;
    B200_REQUIRE(last < nweights, "Fold: iweight=%llu >= nweights=%llu (Fold.C:699)", (unsigned long long)last,
                 (unsigned long long)nweights);
    bw.w = d_weights; bw.nweights = nweights; bw.weight_idat = weight_idat; bw.idat_start = idat_start;
    bw.ndatperweight = ndatperweight;
  }
  return fold_set_bins(f, phi, pps, ndat, idat_start, nullptr, bw);
}

static int fold_set_bins(b200_fold* f, double phi, double pps, uint64_t ndat, uint64_t idat_start, uint64_t* ndat_folded,
                         const BinWeights& bw) {
  B200_REQUIRE(f, "b200_fold_set_bins: null fold");
  Context* ctx = f->ctx;
  if (bw.w) f->weighted = true;
  f->ndat = ndat;
  f->idat_start = idat_start;
  if (ndat_folded) *ndat_folded = ndat;
  if (ndat == 0) return B200_OK;
  if (ndat > f->bins_capacity) {
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (f->d_bins) cudaFree(f->d_bins);
    f->d_bins = nullptr;
    f->bins_capacity = ndat + ndat / 4 + 1024;
    B200_CUDA(cudaMalloc(&f->d_bins, f->bins_capacity * sizeof(unsigned)));
  }
  if (f->seg_pending) {   // previous upload may still be reading the pinned buffer
    B200_CUDA(cudaEventSynchronize(f->seg_free));
    f->seg_pending = false;
  }
  int64_t nseg = b200_phase_segments(phi, pps, ndat, f->h_seg, f->seg_capacity, nullptr);
  if (nseg < 0) {
    // pathological phase_per_sample (a pulse period of a few samples): more binade segments than
    // the staging buffer holds.  Run the reference recurrence on the host and upload the bins.
    std::vector<unsigned> hb(ndat), hh(f->nbin, 0u), ht(f->nbin);
    b200_phase_bins_sequential(phi, pps, f->nbin, ndat, hb.data(), nullptr);
    if (bw.w) {
      std::vector<unsigned> hw(bw.nweights);
      B200_CUDA(cudaMemcpyAsync(hw.data(), bw.w, bw.nweights * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      B200_CUDA(cudaStreamSynchronize(ctx->stream));
      for (uint64_t i = 0; i < ndat; i++)
        if (hw[(idat_start + i + bw.weight_idat) / bw.ndatperweight] == 0u) hb[i] = f->nbin;
    }
    for (uint64_t i = 0; i < ndat; i++)
      if (hb[i] < f->nbin) hh[hb[i]]++;
    B200_CUDA(cudaMemcpyAsync(ht.data(), f->d_hits_total, f->nbin * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (unsigned b = 0; b < f->nbin; b++) ht[b] += hh[b];
    B200_CUDA(cudaMemcpyAsync(f->d_bins, hb.data(), ndat * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(f->d_hits_last, hh.data(), f->nbin * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(f->d_hits_total, ht.data(), f->nbin * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    f->ndat_total += ndat;
    return B200_OK;
  }
  B200_CUDA(cudaMemcpyAsync(f->d_seg, f->h_seg, nseg * sizeof(b200_phase_segment), cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(cudaEventRecord(f->seg_free, ctx->stream));
  f->seg_pending = true;
  B200_CUDA(cudaMemsetAsync(f->d_hits_last, 0, f->nbin * sizeof(unsigned), ctx->stream));
  const unsigned threads = 256;
  // a warp takes 32 x iters consecutive samples at a time: 1024 when that still leaves four warps' worth of chunks for
  // every warp slot of the machine, fewer for short blocks
  const uint64_t slots = uint64_t(ctx->sm_count) * 8 * (threads / 32) * 4;
  const unsigned iters = (unsigned)std::min<uint64_t>(32, std::max<uint64_t>(1, ndat / (32 * slots)));
  unsigned grid = (unsigned)std::min<uint64_t>((ndat + uint64_t(threads) * iters - 1) / (uint64_t(threads) * iters), ctx->sm_count * 8);
  {
    LaunchScope ls(ctx, KC_BINS);
    k_expand_bins<<<grid, threads, 0, ctx->stream>>>(f->d_seg, (unsigned)nseg, ndat, f->nbin, f->d_bins, f->d_hits_last,
                                                    f->d_hits_total, bw, iters);
  }
  f->ndat_total += ndat;
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_fold_get_bin_hits(b200_fold* f, unsigned* h_hits) {
  B200_REQUIRE(f && h_hits, "b200_fold_get_bin_hits: null argument");
  B200_CUDA(cudaMemcpyAsync(h_hits, f->d_hits_last, f->nbin * sizeof(unsigned), cudaMemcpyDeviceToHost, f->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return B200_OK;
}

static int fold_into(b200_fold* f, const float* d_in, uint64_t in_span, float* d_out, uint64_t out_span) {
  Context* ctx = f->ctx;
  if (f->ndat == 0) return B200_OK;
  B200_REQUIRE(f->d_bins, "b200_fold_fold: set_bins has not been called");
  FoldArgs a;
  a.in = d_in; a.in_span = in_span; a.idat_start = f->idat_start; a.ndat = f->ndat; a.bins = f->d_bins;
  a.profile = d_out; a.prof_span = out_span;
  a.nchan = f->nchan; a.npol = f->npol; a.ndim = f->ndim; a.nbin = f->nbin;
  // deterministic mode accumulates in the handle's fixed-point array (only when folding into the handle's own profile)
  a.fix = (f->d_fix && d_out == f->d_profile) ? f->d_fix : nullptr;
  a.inv_lsb = f->d_fix ? 1.0f / f->lsb : 0.f;
  const unsigned threads = 256;
  const unsigned nplane = f->nchan * f->npol;
  // slabs so that the grid fills the machine a few times over
  uint64_t want = std::max<uint64_t>(1, uint64_t(ctx->sm_count) * 8 / nplane);
  uint64_t slab = std::max<uint64_t>((f->ndat + want - 1) / want, uint64_t(threads) * 16);
  a.slab = (unsigned)std::min<uint64_t>(slab, 1u << 30);
  unsigned gx = (unsigned)((f->ndat + a.slab - 1) / a.slab);
  size_t smem = size_t(f->nbin) * f->ndim * sizeof(float);
  a.smem_bins = (smem <= 48 * 1024 && !a.fix) ? 1 : 0;
  if (!a.smem_bins) smem = 0;
  LaunchScope ls(ctx, KC_OTHER);
  // planes beyond the 65535 limit of grid.y go in further launches
  for (unsigned p0 = 0; p0 < nplane; p0 += 65535u) {
    const unsigned np = std::min(65535u, nplane - p0);
    FoldArgs b = a;
    b.in = a.in + uint64_t(p0) * in_span;
    b.profile = a.profile + uint64_t(p0) * out_span;
    if (a.fix) b.fix = a.fix + uint64_t(p0) * out_span;
    dim3 grid(gx, np);
    if (f->ndim == 4) k_fold<4><<<grid, threads, smem, ctx->stream>>>(b);
    else if (f->ndim == 2) k_fold<2><<<grid, threads, smem, ctx->stream>>>(b);
    else k_fold<1><<<grid, threads, smem, ctx->stream>>>(b);
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_fold_fold(b200_fold* f, const float* d_in, uint64_t in_span) {
  B200_REQUIRE(f && d_in, "b200_fold_fold: null argument");
  return fold_into(f, d_in, in_span, f->d_profile, uint64_t(f->nbin) * f->ndim);
}

int b200_fold_fold_into(b200_fold* f, const float* d_in, uint64_t in_span, float* d_out, uint64_t out_span) {
  B200_REQUIRE(f && d_in && d_out, "b200_fold_fold_into: null argument");
  B200_REQUIRE(out_span >= uint64_t(f->nbin) * f->ndim, "b200_fold_fold_into: out_span smaller than nbin*ndim");
  return fold_into(f, d_in, in_span, d_out, out_span);
}

// deterministic mode: bring the float view of the accumulator up to date
static int fold_refresh(b200_fold* f) {
  if (!f->d_fix) return B200_OK;
  const uint64_t nfloat = uint64_t(f->nchan) * f->npol * f->ndim * f->nbin;
  LaunchScope ls(f->ctx, KC_OTHER);
  const unsigned grid = unsigned(std::min<uint64_t>((nfloat + 255) / 256, uint64_t(f->ctx->sm_count) * 8));
  k_fix_to_float<<<grid, 256, 0, f->ctx->stream>>>(f->d_fix, f->lsb, nfloat, f->d_profile);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_fold_set_deterministic(b200_fold* f, float lsb) {
  B200_REQUIRE(f, "b200_fold_set_deterministic: null fold");
  const uint64_t nfloat = uint64_t(f->nchan) * f->npol * f->ndim * f->nbin;
  if (!(lsb > 0.f)) {                      // back to floating-point accumulation (from the current float view)
    if (f->d_fix) {
      int rc = fold_refresh(f);
      if (rc != B200_OK) return rc;
      B200_CUDA(cudaStreamSynchronize(f->ctx->stream));
      cudaFree(f->d_fix);
      f->d_fix = nullptr;
    }
    return B200_OK;
  }
  B200_REQUIRE(f->ndat_total == 0, "b200_fold_set_deterministic: switch modes on an empty (zeroed) PhaseSeries");
  if (!f->d_fix) B200_CUDA(cudaMalloc(&f->d_fix, nfloat * sizeof(long long)));
  B200_CUDA(cudaMemsetAsync(f->d_fix, 0, nfloat * sizeof(long long), f->ctx->stream));
  f->lsb = lsb;
  return B200_OK;
}

int b200_fold_synch(b200_fold* f, float* h_profile) {
  B200_REQUIRE(f && h_profile, "b200_fold_synch: null argument");
  const uint64_t nfloat = uint64_t(f->nchan) * f->npol * f->ndim * f->nbin;
  int rcr = fold_refresh(f);
  if (rcr != B200_OK) return rcr;
  B200_CUDA(cudaMemcpyAsync(h_profile, f->d_profile, nfloat * sizeof(float), cudaMemcpyDeviceToHost, f->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return B200_OK;
}

int b200_fold_get_hits(b200_fold* f, unsigned* h_hits, uint64_t* ndat_total) {
  B200_REQUIRE(f, "b200_fold_get_hits: null fold");
  if (h_hits) {
    B200_CUDA(cudaMemcpyAsync(h_hits, f->d_hits_total, f->nbin * sizeof(unsigned), cudaMemcpyDeviceToHost, f->ctx->stream));
    B200_CUDA(cudaStreamSynchronize(f->ctx->stream));
  }
  if (ndat_total) *ndat_total = f->ndat_total;
  return B200_OK;
}

int b200_fold_zero(b200_fold* f) {
  B200_REQUIRE(f, "b200_fold_zero: null fold");
  const uint64_t nfloat = uint64_t(f->nchan) * f->npol * f->ndim * f->nbin;
  B200_CUDA(cudaMemsetAsync(f->d_profile, 0, nfloat * sizeof(float), f->ctx->stream));
  if (f->d_fix) B200_CUDA(cudaMemsetAsync(f->d_fix, 0, nfloat * sizeof(long long), f->ctx->stream));
  B200_CUDA(cudaMemsetAsync(f->d_hits_total, 0, f->nbin * sizeof(unsigned), f->ctx->stream));
  B200_CUDA(cudaMemsetAsync(f->d_hits_last, 0, f->nbin * sizeof(unsigned), f->ctx->stream));
  f->ndat_total = 0;
  f->weighted = false;
  return B200_OK;
}

int b200_fold_weighted(const b200_fold* f) { return f && f->weighted ? 1 : 0; }

// deterministic mode: the float view is refreshed (stream ordered) before the pointer is handed out
float* b200_fold_device_profile(b200_fold* f) {
  if (!f) return nullptr;
  fold_refresh(f);
  return f->d_profile;
}
unsigned* b200_fold_device_hits(b200_fold* f) { return f ? f->d_hits_total : nullptr; }

}  // extern "C"

// internal accessors for pipeline.cu
namespace b200 {
const unsigned* fold_bins(b200_fold* f) { return f->d_bins; }
long long* fold_fix(b200_fold* f) { return f->d_fix; }
float fold_lsb(b200_fold* f) { return f->lsb; }
const uint2* fold_runs(b200_fold* f) { return f->d_runs; }
const unsigned* fold_nruns(b200_fold* f) { return f->d_nruns; }

// Sizes the bin plan (and, when nkeep != 0, the item table) for blocks of up to ndat output samples, so that no
// later set_bins / fold_build_runs has to allocate.
int fold_reserve(b200_fold* f, uint64_t ndat, unsigned nkeep) {
  Context* ctx = f->ctx;
  if (ndat > f->bins_capacity) {
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (f->d_bins) cudaFree(f->d_bins);
    f->d_bins = nullptr;
    f->bins_capacity = ndat + 1024;
    B200_CUDA(cudaMalloc(&f->d_bins, f->bins_capacity * sizeof(unsigned)));
  }
  if (nkeep) {
    const uint64_t npart = (ndat + nkeep - 1) / nkeep, need = npart * (uint64_t(nkeep) + 1);
    if (need > f->runs_capacity || npart > f->nruns_capacity) {
      B200_CUDA(cudaStreamSynchronize(ctx->stream));
      if (f->d_runs) cudaFree(f->d_runs);
      if (f->d_nruns) cudaFree(f->d_nruns);
      f->d_runs = nullptr;
      f->d_nruns = nullptr;
      f->runs_capacity = need;
      f->nruns_capacity = npart + 16;
      B200_CUDA(cudaMalloc(&f->d_runs, f->runs_capacity * sizeof(uint2)));
      B200_CUDA(cudaMalloc(&f->d_nruns, f->nruns_capacity * sizeof(unsigned)));
    }
  }
  return B200_OK;
}

// Builds the per-part run table of the bin plan set by the last b200_fold_set_bins (ndat = npart * nkeep).
int fold_build_runs(b200_fold* f, unsigned nkeep, unsigned align) {
  B200_REQUIRE(f && f->d_bins && nkeep && f->ndat % nkeep == 0, "fold_build_runs: the bin plan is not a whole number of parts");
  Context* ctx = f->ctx;
  const uint64_t npart = f->ndat / nkeep, need = npart * (uint64_t(nkeep) + 1);
  if (need > f->runs_capacity || npart > f->nruns_capacity) {
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    if (f->d_runs) cudaFree(f->d_runs);
    if (f->d_nruns) cudaFree(f->d_nruns);
    f->d_runs = nullptr;
    f->d_nruns = nullptr;
    f->runs_capacity = need + need / 4;
    f->nruns_capacity = npart + npart / 4 + 16;
    B200_CUDA(cudaMalloc(&f->d_runs, f->runs_capacity * sizeof(uint2)));
    B200_CUDA(cudaMalloc(&f->d_nruns, f->nruns_capacity * sizeof(unsigned)));
  }
  if (npart == 0) return B200_OK;
  {
    LaunchScope ls(ctx, KC_BINS);
    k_bin_runs<<<(unsigned)npart, 1024, 0, ctx->stream>>>(f->d_bins, nkeep, align & 15u, f->d_runs, f->d_nruns);
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}
}
