// longconv.cu -- convolutions of more than 131072 points (complex dual-polarisation input): N = P Q > 131072 (P = 512 ... 2048, Q = 512 ... 8192; cfg4: 2048 x 2048, 4 Mi points).
// No group of CTAs holds such a transform, so the spectrum makes its two round trips through HBM as in the generic
// three-kernel path (filterbank.cu k_cols_fwd / k_rows / k_cols_inv_fold) -- but the three kernels are built on the
// c2 core like the one-kernel path of shorter convolutions (clusterconv.cu): 16 points of BOTH polarisations per thread, compile-time transform sizes, 128-bit
// exchanges.  The generic kernels spend 120 instructions per point on run-time index arithmetic and two table
// look-ups per output twiddle (ncu, cfg4: issue-bound at 1.0 + 1.9 + 1.2 ms per 16 parts for 1 GiB of spectrum).
//   K1  k_bc_cols_fwd: NC adjacent columns n2 per CTA, P-point column transforms, times W_N^(n2 k1)
//   K2  k_bc_rows:     one row k1 per CTA: forward Q-point transform, response (from a TRANSPOSED copy, so the row's
//                      values are contiguous; staged with cp.async under the forward transform), inverse, W_N^(-k1 m2)
//   K3  k_bc_cols_inv: NC adjacent columns m2 per CTA, inverse P-point transforms, discard, detect, fold (the
//                      run-walk + segmented scan of k_cols_inv_fold) or the voltage / detected-series sinks
// When all three run, the scratch holds the two polarisations of a bin side by side (one float4: 16 NC-byte runs in
// the column kernels instead of 8 NC); any of them alone works on the generic layout (one plane per polarisation),
// which is how the parity tests isolate them (dev build: B200_BC_K1 / B200_BC_K2 / B200_BC_K3 = 0).
#include <algorithm>
#include <vector>

#include "clusterconv.cuh"

namespace b200 {

struct BcArgs {
  CcArgs c;             // source, two-level table of W_N, sink, part0, nb, nchan_in, nfilt_pos, nkeep (H, tw, xch, bar unused)
  float2* A;            // spectrum scratch: per (part, channel) 2 N float2
  const float2* Ht;     // response [channel][P][Q], or null
  const float2* twP;
  const float2* twQ;    // c2 stage tables
  unsigned P, Q;
  unsigned il;          // 1: float4 (pol 0, pol 1) per bin; 0: planes of N float2 per polarisation
  int conv_ok;          // generic 8-bit: the table is RN(x (conv_hi + conv_lo)), x = int8(b) + 0.5 (FbSource::conv_ok)
  float conv_hi, conv_lo;
};

constexpr unsigned BC_NC = 4;   // columns per CTA of the column kernels
template <unsigned P> struct Bc {
  static constexpr unsigned T = P / 16, NT = BC_NC * T;
  static constexpr unsigned RS = c2::pair_slots<P>() + 8 / BC_NC;    // pair regions skewed: the NC pairs of a warp hit distinct banks
  static constexpr size_t SMEM_FWD = size_t(BC_NC) * RS * sizeof(float4);
  // K3 stages the detected products of its P groups of NC samples over the transforms, one padding slot per group
  static constexpr size_t SMEM_INV = SMEM_FWD > size_t(P) * (BC_NC + 1) * sizeof(float4) ? SMEM_FWD : size_t(P) * (BC_NC + 1) * sizeof(float4);
  static_assert(c2::pair_slots<P>() % 8 == 0, "region skew assumes an 8-aligned pair size");
};

__device__ __forceinline__ float2 ld_cg_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

template <int SRC, unsigned P>
__global__ void __launch_bounds__(Bc<P>::NT, 512 / Bc<P>::NT) k_bc_cols_fwd(BcArgs a) {
  using B = Bc<P>;
  constexpr unsigned T = B::T, NC = BC_NC;
  extern __shared__ __align__(16) float4 buf[];
  __shared__ float s_lut[SRC == SRC_GENERIC8 ? 256 : 1];
  __shared__ float2 s_h[16 * NC];          // [e][col] = W_N^(n2 T e)
  const CcArgs& c = a.c;
  const unsigned tid = threadIdx.x, col = tid % NC, j = tid / NC;
  const unsigned Q = a.Q;
  const unsigned ncb = Q / NC;
  const unsigned cb = blockIdx.x % ncb, rest = blockIdx.x / ncb;      // rest = (part, channel)
  const unsigned ic = rest % c.nchan_in;
  const uint64_t part = c.part0 + rest / c.nchan_in;
  const unsigned n2 = cb * NC + col;
  if (SRC == SRC_GENERIC8) {
    for (unsigned i = tid; i < 256; i += B::NT) s_lut[i] = c.lut[i];
    __syncthreads();
  }
  // W_N^(n2 k1), k1 = j + T e, as W_N^(n2 j) (per thread) times W_N^(n2 T e) (table of the tile)
  if (tid < 16 * NC) s_h[tid] = big_twiddle<false>(c.blo, c.bhi, (cb * NC + tid % NC) * T * (tid / NC));
  const float2 wbase = big_twiddle<false>(c.blo, c.bhi, n2 * j);
  float2 va[16], vb[16];
  if (SRC == SRC_GENERIC8 && (reinterpret_cast<uintptr_t>(c.src) & 3u) == 0) {
    // TFP bytes: the two polarisations of a complex sample are one aligned 32-bit word (BitUnpacker.C:56-75 with
    // npol = ndim = 2): one load per point instead of two (the kernel's loads are 16-byte granules a row apart: the
    // number of requests, not of bytes, is what L1TEX pays for)
    const unsigned* words = static_cast<const unsigned*>(c.src) + (c.first + part * c.step) * c.nchan_in + ic;
    unsigned w[16];
#pragma unroll
    for (int e = 0; e < 16; e++) w[e] = __ldg(words + uint64_t(Q * (j + T * unsigned(e)) + n2) * c.nchan_in);
    if (a.conv_ok) {
      // two's-complement table that is RN(x (hi + lo)), x = int8(b) + 0.5, entry by entry (lut_as_arithmetic on the host):
      // converted arithmetically, bit for bit the table's values, instead of 64 shared-memory gathers per thread
      // (k1_c2's conversion: the byte, sign bit flipped, dropped into bits 8..15 of the float 32768 reads 32768 + 128 + int8(b))
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
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const unsigned n = Q * (j + T * unsigned(e)) + n2;
      va[e] = cc_load<SRC>(c, s_lut, ic, 0, part, n);
      vb[e] = cc_load<SRC>(c, s_lut, ic, 1, part, n);
    }
  }
  c2::fft_pair<P, false>(va, vb, j, buf + col * B::RS, a.twP, CcSync());
  const uint64_t N = uint64_t(P) * Q;
  float4* A4 = reinterpret_cast<float4*>(a.A) + uint64_t(rest) * N + n2;
  float2* Ap = a.A + uint64_t(rest) * 2 * N + n2;
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const float2 w = cmul(wbase, s_h[e * int(NC) + col]);
    const float2 u = cmul(va[e], w), v = cmul(vb[e], w);
    const uint64_t idx = uint64_t(j + T * unsigned(e)) * Q;
    if (a.il) A4[idx] = make_float4(u.x, u.y, v.x, v.y);
    else {
      Ap[idx] = u;
      Ap[N + idx] = v;
    }
  }
}

template <unsigned QQ>
__global__ void __launch_bounds__(QQ / 16, 512 / (QQ / 16)) k_bc_rows(BcArgs a) {
  constexpr unsigned Q = QQ, NT = Q / 16, PS = c2::pair_slots<Q>();
  extern __shared__ __align__(16) float4 buf[];
  float2* Hs = reinterpret_cast<float2*>(buf + PS);                  // the response values of this row: [Q]
  __shared__ float2 s_tw[16];                                        // W_N^(-k1 NT e)
  const CcArgs& c = a.c;
  const unsigned tid = threadIdx.x;
  // the parts of a (channel, row) run side by side: they share the response row (L2)
  const unsigned partl = blockIdx.x % c.nb, rowc = blockIdx.x / c.nb;
  const unsigned row = rowc % a.P, ic = rowc / a.P;
  const unsigned rest = partl * c.nchan_in + ic;
  if (a.Ht) {
    const float2* h = a.Ht + uint64_t(rowc) * Q;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(Hs);
    for (unsigned i = tid; i < Q / 2; i += NT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * i), "l"(h + 2u * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (tid < 16) s_tw[tid] = big_twiddle<true>(c.blo, c.bhi, row * NT * tid);
  const float2 wown = big_twiddle<true>(c.blo, c.bhi, row * tid);
  const uint64_t N = uint64_t(a.P) * Q;
  float4* A4 = reinterpret_cast<float4*>(a.A) + uint64_t(rest) * N + uint64_t(row) * Q + tid;
  float2* Ap = a.A + uint64_t(rest) * 2 * N + uint64_t(row) * Q + tid;
  float2 va[16], vb[16];
  if (a.il) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const float4 x = ld_cg_f4(A4 + NT * e);
      va[e] = make_float2(x.x, x.y);
      vb[e] = make_float2(x.z, x.w);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      va[e] = ld_cg_f2(Ap + NT * e);
      vb[e] = ld_cg_f2(Ap + N + NT * e);
    }
  }
  // forward and inverse through one copy of the transform code: inverse = conj(FFT(conj z)) (see k_conv64k phase B)
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    if (pass) {
      if (a.Ht) asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();                    // the response row has landed; every thread has gathered the forward transform
      if (a.Ht) {
        const float2* h = Hs + tid;
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const float2 hv = h[e * int(NT)];
          va[e] = cmul(va[e], hv);
          vb[e] = cmul(vb[e], hv);
        }
      }
#pragma unroll
      for (int e = 0; e < 16; e++) {
        va[e].y = -va[e].y;
        vb[e].y = -vb[e].y;
      }
    }
    c2::fft_pair<Q, false>(va, vb, tid, buf, a.twQ, CcSync());
  }
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const float2 w = cmul(wown, s_tw[e]);
    const float2 u = cmul(make_float2(va[e].x, -va[e].y), w), v = cmul(make_float2(vb[e].x, -vb[e].y), w);
    if (a.il) A4[NT * e] = make_float4(u.x, u.y, v.x, v.y);
    else {
      Ap[NT * e] = u;
      Ap[N + NT * e] = v;
    }
  }
}

template <unsigned P>
__global__ void __launch_bounds__(Bc<P>::NT, 512 / Bc<P>::NT) k_bc_cols_inv(BcArgs a) {
  using B = Bc<P>;
  constexpr unsigned T = B::T, NC = BC_NC;
  extern __shared__ __align__(16) float4 buf[];
  const CcArgs& c = a.c;
  const unsigned tid = threadIdx.x, col = tid % NC, j = tid / NC;
  const unsigned Q = a.Q;
  const unsigned ncb = Q / NC;
  const unsigned cb = blockIdx.x % ncb, rest = blockIdx.x / ncb;
  const unsigned ic = rest % c.nchan_in, partl = rest / c.nchan_in;
  const unsigned m2 = cb * NC + col;
  const uint64_t N = uint64_t(P) * Q;
  float2 va[16], vb[16];
  if (a.il) {
    const float4* A4 = reinterpret_cast<const float4*>(a.A) + uint64_t(rest) * N + m2;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const float4 x = ld_cg_f4(A4 + uint64_t(j + T * unsigned(e)) * Q);
      va[e] = make_float2(x.x, x.y);
      vb[e] = make_float2(x.z, x.w);
    }
  } else {
    const float2* Ap = a.A + uint64_t(rest) * 2 * N + m2;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      va[e] = ld_cg_f2(Ap + uint64_t(j + T * unsigned(e)) * Q);
      vb[e] = ld_cg_f2(Ap + N + uint64_t(j + T * unsigned(e)) * Q);
    }
  }
  c2::fft_pair<P, true>(va, vb, j, buf + col * B::RS, a.twP, CcSync());
  // register e = segment m1 = j + T e: sample Q m1 + m2 of the transform
  const unsigned np0 = c.nfilt_pos, nkeep = c.nkeep;
  const int state = c.sink.state;
  const unsigned dndim = c.sink.dndim, nbin = c.sink.nbin;
  const uint64_t part = c.part0 + partl;
  if (c.sink.kind == EPI_VOLT) {
    float2* outp = reinterpret_cast<float2*>(c.sink.volt + (uint64_t(ic) * 2) * c.sink.volt_span + part * c.sink.volt_step);
    float2* outq = reinterpret_cast<float2*>(c.sink.volt + (uint64_t(ic) * 2 + 1) * c.sink.volt_span + part * c.sink.volt_step);
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const unsigned u = Q * (j + T * unsigned(e)) + m2 - np0;
      if (u < nkeep) {
        outp[u] = va[e];
        outq[u] = vb[e];
      }
    }
    return;
  }
  const unsigned nprod = state_nprod(state, 2);
  if (c.sink.kind == EPI_DETECT) {
    const unsigned dnpol = nprod / dndim;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const unsigned u = Q * (j + T * unsigned(e)) + m2 - np0;
      if (u < nkeep) {
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        detect_products(state, va[e], vb[e], r);
        const uint64_t osamp = part * nkeep + u;
        for (unsigned pr = 0; pr < nprod; pr++)
          c.sink.det[(uint64_t(ic) * dnpol + pr / dndim) * c.sink.det_span + osamp * dndim + pr % dndim] = r[pr];
      }
    }
    return;
  }
  // fold: the CTA's samples are P groups (one per m1, Q samples apart) of NC consecutive samples; thread t walks the
  // groups t, t + NT, ... (16 / NC of them).  Their phase bins are requested now, all at once, and arrive while the
  // products are detected and staged
  constexpr unsigned NG = 16 / NC;
  static_assert(NG * Bc<P>::NT == P, "every thread walks 16 / NC groups");
  const unsigned* plan = c.sink.bins + uint64_t(partl) * nkeep;
  unsigned wb[NG][NC];
  // a group's first sample is a multiple of NC = 4 past -nfilt_pos: when that falls on a 16-byte boundary of the plan
  // (cfg4 does), one 128-bit load per group instead of four requests that touch 32 lines each
  static_assert(NC == 4, "vector bin loads assume four samples per group");
  const bool vec = ((reinterpret_cast<uintptr_t>(plan) >> 2) - np0) % 4u == 0;
#pragma unroll
  for (unsigned k = 0; k < NG; k++) {
    const unsigned u0 = (tid + k * B::NT) * Q + cb * NC - np0;           // unsigned: samples before nfilt_pos wrap to huge values
    if (vec && u0 < nkeep && u0 + 3u < nkeep) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(plan + u0));
      wb[k][0] = x.x; wb[k][1] = x.y; wb[k][2] = x.z; wb[k][3] = x.w;
    } else {
#pragma unroll
      for (unsigned bb = 0; bb < NC; bb++) {
        const unsigned u = u0 + bb;
        wb[k][bb] = u < nkeep ? __ldg(plan + u) : 0xffffffffu;
      }
    }
  }
  __syncthreads();                       // every thread has gathered the last stage: the pair buffers are free
#pragma unroll
  for (int e = 0; e < 16; e++) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    detect_products(state, va[e], vb[e], r);
    const unsigned g = j + T * unsigned(e);
    buf[g * (NC + 1) + col] = make_float4(r[0], r[1], r[2], r[3]);
  }
  __syncthreads();
  const uint64_t prof0 = uint64_t(ic) * nbin * nprod;
  uint64_t off[4];
#pragma unroll
  for (unsigned pr = 0; pr < 4; pr++) off[pr] = prof0 + uint64_t(pr / dndim) * nbin * dndim + pr % dndim;
  // one thread per group sums runs of equal phase bin in time order (Fold.C:844-852); neighbouring lanes are Q samples
  // apart, and a bin is often wider than that, so the last run of every lane is combined over neighbouring lanes of
  // equal bin (segmented scan, as in k_cols_inv_fold) and only the tail lane of each run of lanes issues the REDs
  const unsigned lane = tid & 31u;
  static_assert(P % Bc<P>::NT == 0 && Bc<P>::NT % 32 == 0, "whole warps walk whole groups");
#pragma unroll
  for (unsigned k = 0; k < NG; k++) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned cur = 0xffffffffu;
    const float4* sgrp = buf + (tid + k * B::NT) * (NC + 1);
#pragma unroll
    for (unsigned bb = 0; bb < NC; bb++) {
      const unsigned bin = wb[k][bb];
      const float4 x = sgrp[bb];
      if (bin != cur) {
        if (cur < nbin) {                                       // nbin: flagged window; 0xffffffff: discarded sample
#pragma unroll
          for (unsigned pr = 0; pr < 4; pr++)
            if (pr < nprod) profile_add(c.sink.profile, c.sink.fix, c.sink.inv_lsb, off[pr] + uint64_t(cur) * dndim, acc[pr]);
        }
        cur = bin;
        acc[0] = x.x; acc[1] = x.y; acc[2] = x.z; acc[3] = x.w;
      } else {
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
      }
    }
    // (combining the NG scans step by step -- 16 independent shuffle chains -- measured the same 0.77 ms at cfg4)
    const unsigned prev = __shfl_up_sync(0xffffffffu, cur, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != cur);
    const unsigned h = 31u - __clz(heads & (0xffffffffu >> (31u - lane)));    // first lane of this lane's run
#pragma unroll
    for (unsigned o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (unsigned pr = 0; pr < 4; pr++) {
        const float up = __shfl_up_sync(0xffffffffu, acc[pr], o);
        if (lane >= h + o) acc[pr] += up;
      }
    }
    const bool tail = lane == 31u || ((heads >> (lane + 1u)) & 1u);
    if (tail && cur < nbin) {
#pragma unroll
      for (unsigned pr = 0; pr < 4; pr++)
        if (pr < nprod) profile_add(c.sink.profile, c.sink.fix, c.sink.inv_lsb, off[pr] + uint64_t(cur) * dndim, acc[pr]);
    }
  }
}

__global__ void k_bc_transpose_response(const float2* __restrict__ H, float2* __restrict__ Ht, unsigned P, unsigned Q,
                                        unsigned nchan) {
  const uint64_t N = uint64_t(P) * Q, total = N * nchan;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const unsigned k2 = unsigned(i % Q);
    const uint64_t rowc = i / Q;
    const unsigned row = unsigned(rowc % P), ic = unsigned(rowc / P);
    Ht[i] = H[uint64_t(ic) * N + row + uint64_t(P) * k2];
  }
}

template <typename K> static int bc_optin(K kernel, size_t bytes) {
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200_OK;
}
template <unsigned L> static int bc_table(float2** out) {
  std::vector<float2> h(c2::twiddle_count<L>(), make_float2(1.f, 0.f));
  c2::fill_twiddles<L>(h.data());
  B200_CUDA(cudaMalloc(out, sizeof(float2) * h.size()));
  B200_CUDA(cudaMemcpy(*out, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
  return B200_OK;
}
template <unsigned Q> static constexpr size_t bc_rows_smem() { return size_t(c2::pair_slots<Q>()) * sizeof(float4) + size_t(Q) * sizeof(float2); }

template <unsigned P> static int bc_init_p(b200_fb_plan* pl) {
  int rc;
  if ((rc = bc_table<P>(&pl->bc_twP)) != B200_OK) return rc;
  if ((rc = bc_optin(k_bc_cols_fwd<SRC_F32, P>, Bc<P>::SMEM_FWD)) != B200_OK) return rc;
  if ((rc = bc_optin(k_bc_cols_fwd<SRC_MEERKAT8, P>, Bc<P>::SMEM_FWD)) != B200_OK) return rc;
  if ((rc = bc_optin(k_bc_cols_fwd<SRC_UWB16, P>, Bc<P>::SMEM_FWD)) != B200_OK) return rc;
  if ((rc = bc_optin(k_bc_cols_fwd<SRC_GENERIC8, P>, Bc<P>::SMEM_FWD)) != B200_OK) return rc;
  return bc_optin(k_bc_cols_inv<P>, Bc<P>::SMEM_INV);
}
template <unsigned Q> static int bc_init_q(b200_fb_plan* pl) {
  int rc;
  if ((rc = bc_table<Q>(&pl->bc_twQ)) != B200_OK) return rc;
  return bc_optin(k_bc_rows<Q>, bc_rows_smem<Q>());
}

int bc_plan_init(b200_fb_plan* pl) {
  pl->bc_ok = false;
  pl->bc_twP = pl->bc_twQ = pl->bc_Ht = nullptr;
  static const bool want = tune_flag("B200_BIG_CONV", true);
  if (!want || !pl->conv_path || pl->desc.input_real || pl->desc.npol != 2) return B200_OK;
  if (pl->Nc <= 16 * 8192 || size_t(pl->ctx->max_smem_optin) < bc_rows_smem<8192>() + 1024) return B200_OK;
  if (pl->Q % BC_NC) return B200_OK;
  int rc;
  switch (pl->P) {
    case 512: rc = bc_init_p<512>(pl); break;
    case 1024: rc = bc_init_p<1024>(pl); break;
    case 2048: rc = bc_init_p<2048>(pl); break;
    default: return B200_OK;
  }
  if (rc != B200_OK) return rc;
  switch (pl->Q) {
    case 512: rc = bc_init_q<512>(pl); break;
    case 1024: rc = bc_init_q<1024>(pl); break;
    case 2048: rc = bc_init_q<2048>(pl); break;
    case 4096: rc = bc_init_q<4096>(pl); break;
    case 8192: rc = bc_init_q<8192>(pl); break;
    default: return B200_OK;
  }
  if (rc != B200_OK) return rc;
  if (pl->d_response) {
    const uint64_t n = uint64_t(pl->desc.input_nchan) * pl->Nc;
    B200_CUDA(cudaMalloc(&pl->bc_Ht, n * sizeof(float2)));
    k_bc_transpose_response<<<1024, 256, 0, pl->ctx->stream>>>(pl->d_response, pl->bc_Ht, pl->P, pl->Q, pl->desc.input_nchan);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  }
  pl->bc_ok = true;
  return B200_OK;
}

void bc_plan_free(b200_fb_plan* pl) {
  if (pl->bc_twP) cudaFree(pl->bc_twP);
  if (pl->bc_twQ) cudaFree(pl->bc_twQ);
  if (pl->bc_Ht) cudaFree(pl->bc_Ht);
  pl->bc_twP = pl->bc_twQ = pl->bc_Ht = nullptr;
  pl->bc_ok = false;
}

bool bc_k1_applies(const b200_fb_plan* pl, const FbSource& src) {
  static const bool want = tune_flag("B200_BC_K1", true);
  return want && pl->bc_ok &&
         (src.kind == SRC_F32 || src.kind == SRC_MEERKAT8 || src.kind == SRC_UWB16 || (src.kind == SRC_GENERIC8 && src.ndim == 2));
}
bool bc_k2_applies(const b200_fb_plan* pl) {
  static const bool want = tune_flag("B200_BC_K2", true);
  return want && pl->bc_ok;
}
bool bc_k3_applies(const b200_fb_plan* pl) {
  static const bool want = tune_flag("B200_BC_K3", true);
  return want && pl->bc_ok;
}

static void bc_args(const b200_fb_plan* pl, BcArgs& a, uint64_t part0, unsigned nb, bool il) {
  a.c = CcArgs{};
  a.c.blo = pl->bigN.lo; a.c.bhi = pl->bigN.hi;
  a.c.nchan_in = pl->desc.input_nchan; a.c.nb = nb; a.c.part0 = part0;
  a.c.nfilt_pos = pl->desc.nfilt_pos; a.c.nkeep = pl->nkeep;
  a.A = pl->scratchA; a.Ht = pl->bc_Ht; a.twP = pl->bc_twP; a.twQ = pl->bc_twQ;
  a.P = pl->P; a.Q = pl->Q; a.il = il ? 1u : 0u;
  a.conv_ok = 0; a.conv_hi = a.conv_lo = 0.f;
}

template <unsigned P> static void bc_k1_launch(b200_fb_plan* pl, const FbSource& src, const BcArgs& a, unsigned nb) {
  const dim3 grid(pl->Q / BC_NC * nb * pl->desc.input_nchan), block(Bc<P>::NT);
  cudaStream_t st = pl->ctx->stream;
  if (src.kind == SRC_F32) k_bc_cols_fwd<SRC_F32, P><<<grid, block, Bc<P>::SMEM_FWD, st>>>(a);
  else if (src.kind == SRC_MEERKAT8) k_bc_cols_fwd<SRC_MEERKAT8, P><<<grid, block, Bc<P>::SMEM_FWD, st>>>(a);
  else if (src.kind == SRC_UWB16) k_bc_cols_fwd<SRC_UWB16, P><<<grid, block, Bc<P>::SMEM_FWD, st>>>(a);
  else k_bc_cols_fwd<SRC_GENERIC8, P><<<grid, block, Bc<P>::SMEM_FWD, st>>>(a);
}

int bc_k1(b200_fb_plan* pl, const FbSource& src, uint64_t part0, unsigned nb, bool il) {
  BcArgs a;
  bc_args(pl, a, part0, nb, il);
  a.c.src = src.ptr; a.c.span = src.span; a.c.step = src.step; a.c.first = src.first; a.c.scale = src.scale;
  a.c.sample_swap = src.sample_swap; a.c.lut = src.d_lut;
  a.conv_ok = src.kind == SRC_GENERIC8 ? src.conv_ok : 0; a.conv_hi = src.conv_hi; a.conv_lo = src.conv_lo;
  LaunchScope ls(pl->ctx, KC_COLS_FWD);
  switch (pl->P) {
    case 512: bc_k1_launch<512>(pl, src, a, nb); break;
    case 1024: bc_k1_launch<1024>(pl, src, a, nb); break;
    default: bc_k1_launch<2048>(pl, src, a, nb); break;
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int bc_k2(b200_fb_plan* pl, unsigned nb, bool il) {
  BcArgs a;
  bc_args(pl, a, 0, nb, il);
  const dim3 grid(pl->P * pl->desc.input_nchan * nb);
  cudaStream_t st = pl->ctx->stream;
  LaunchScope ls(pl->ctx, KC_ROWS);
  switch (pl->Q) {
    case 512: k_bc_rows<512><<<grid, 512 / 16, bc_rows_smem<512>(), st>>>(a); break;
    case 1024: k_bc_rows<1024><<<grid, 1024 / 16, bc_rows_smem<1024>(), st>>>(a); break;
    case 2048: k_bc_rows<2048><<<grid, 2048 / 16, bc_rows_smem<2048>(), st>>>(a); break;
    case 4096: k_bc_rows<4096><<<grid, 4096 / 16, bc_rows_smem<4096>(), st>>>(a); break;
    default: k_bc_rows<8192><<<grid, 8192 / 16, bc_rows_smem<8192>(), st>>>(a); break;
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int bc_k3(b200_fb_plan* pl, const FbSink& sk, uint64_t part0, unsigned nb, bool il) {
  BcArgs a;
  bc_args(pl, a, part0, nb, il);
  a.c.sink = sk;
  const dim3 grid(pl->Q / BC_NC * nb * pl->desc.input_nchan);
  cudaStream_t st = pl->ctx->stream;
  LaunchScope ls(pl->ctx, KC_INV);
  switch (pl->P) {
    case 512: k_bc_cols_inv<512><<<grid, Bc<512>::NT, Bc<512>::SMEM_INV, st>>>(a); break;
    case 1024: k_bc_cols_inv<1024><<<grid, Bc<1024>::NT, Bc<1024>::SMEM_INV, st>>>(a); break;
    default: k_bc_cols_inv<2048><<<grid, Bc<2048>::NT, Bc<2048>::SMEM_INV, st>>>(a); break;
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
