// clusterconv.cu -- the convolution path of 65536-point transforms (cfg3: MeerKAT, one coherent-dedispersion
// transform per input channel) as ONE kernel on thread-block clusters: unpack -> forward FFT -> response ->
// inverse FFT -> discard -> detect -> fold without a single spectrum round trip through HBM.
//
// dsp::Convolution::transformation (Convolution.C:389-458) per (channel, part): both polarisations, N = 65536
// complex points = 1 MiB of float2 -- more than one SM holds, exactly what a cluster of 16 CTAs holds in registers
// (16 x 256 threads x 16 points x 2 polarisations; two CTAs of different clusters share an SM, so one cluster's
// barrier waits and exchanges run under the other's butterflies).  Factorisation N = P Q, P = 16, Q = 4096
// (n = Q n1 + n2, k = k1 + P k2):
//   A  every thread owns ONE column n2 of both polarisations: 16 samples Q apart -> a 16-point DFT in registers,
//      times W_N^(n2 k1); element k1 goes to the CTA that owns row k1 (rank k1) through distributed shared
//      memory (st.shared::cluster, 128-bit: the two polarisations of a point travel together)
//   B  every CTA owns one row k1 (both polarisations = one fft_c2 sequence pair): forward 4096-point row
//      transform, times the response H[k1 + 16 k2] (the CTA's response row stays in shared memory for as long
//      as the cluster works on the same channel), inverse 4096-point row transform, times W_N^(-m2 k1)
//   C  element (k1, m2) goes back to the CTA that owns column m2 (rank m2 / 256)
//   D  every thread owns one column m2 again: inverse 16-point DFT over k1 -> y[Q m1 + m2], m1 < 16; detection of
//      both polarisations from registers; the detected products go to shared memory in time order (16 segments of
//      256 consecutive samples per CTA) and one thread per 16 samples walks them with the bin plan: runs of one
//      phase bin are summed sequentially (the order of Fold.C:844-852) and added with one RED.ADD.F32 per product.
// Four cluster barriers per tile order the two exchanges.  HBM traffic per (channel, part): the raw bytes, the bin
// plan (L2) and the REDs -- the three-kernel path moves the 1 MiB spectrum through HBM four times.
// Clusters are persistent; each takes a contiguous range of (channel, part) tiles so that the response rows are
// re-staged only when the channel changes.
#include <algorithm>
#include <vector>

#include "engine.cuh"
#include "fft_c2.cuh"

#ifndef CC_DBG
#define CC_DBG 0
#endif

namespace b200 {

namespace cc {
constexpr unsigned N = 65536, P = 16, Q = 4096, CL = 16, COLS = Q / CL, NT = 256;
constexpr unsigned T = Q / 16;                       // threads per row pair
constexpr unsigned PS = c2::pair_slots<Q>();         // float4 slots of one pair buffer (4352)
constexpr size_t SMEM = size_t(PS) * sizeof(float4) + size_t(Q) * sizeof(float2);
static_assert(PS == COLS * 16 + COLS, "the receive buffer of phase D (16 x 256 padded) is the pair buffer");
}  // namespace cc

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
};

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
#if CC_DBG & 8
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
#else
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
#endif
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}
__device__ __forceinline__ unsigned map_remote(unsigned saddr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote(unsigned addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
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

template <int SRC>
__global__ void __launch_bounds__(256, 2) k_conv64k(CcArgs a) {
  using namespace cc;
  extern __shared__ __align__(16) float4 buf[];                    // the pair buffer = the receive buffer of phase D
  float2* Hs = reinterpret_cast<float2*>(buf + PS);                // response row k1 = rank: [Q]
  __shared__ float2 s_tw[16];                                      // W_N^(-k1 256 e) of the CTA's row
  __shared__ float s_lut[SRC == SRC_GENERIC8 ? 256 : 1];
  if (SRC == SRC_GENERIC8) s_lut[threadIdx.x] = a.lut[threadIdx.x];
  const unsigned rank = cluster_ctarank();                         // = the row k1 this CTA owns
  const unsigned tid = threadIdx.x;
  const unsigned sbuf = (unsigned)__cvta_generic_to_shared(buf);
  if (tid < 16) s_tw[tid] = big_twiddle<true>(a.blo, a.bhi, (rank * 256u * tid) & (N - 1));
  const float2 wown = big_twiddle<true>(a.blo, a.bhi, rank * tid); // W_N^(-k1 j)
  const unsigned n2 = COLS * rank + tid;                           // phases A / D: this thread's column
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;
  const int state = a.sink.state;
  const unsigned nprod = state_nprod(state, 2), dndim = a.sink.dndim, nbin = a.sink.nbin;

  const unsigned t_begin = cluster_id_x() * a.tiles_per_cluster;
  const unsigned t_end = min(a.ntiles, t_begin + a.tiles_per_cluster);
  unsigned cur_ic = 0xffffffffu;
  cluster_sync();                           // every CTA of the cluster is running: its shared memory may be written
  for (unsigned t = t_begin; t < t_end; t++) {
    const unsigned ic = t / a.nb, partl = t % a.nb;
    const uint64_t part = a.part0 + partl;

    // ---- phase A: 16-point column transforms of both polarisations, scatter by row owner ----
    {
      float2 xp[16], xq[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; n1++) {
        xp[n1] = cc_load<SRC>(a, s_lut, ic, 0, part, n2 + Q * n1);
        xq[n1] = cc_load<SRC>(a, s_lut, ic, 1, part, n2 + Q * n1);
      }
      // W_N^(n2 k1), k1 < 16: four table values (k1 = 1, 2, 4, 8), the others as products over the bits of k1
      float2 w[4];
#pragma unroll
      for (int b = 0; b < 4; b++) w[b] = big_twiddle<false>(a.blo, a.bhi, (n2 << b) & (N - 1));
      if (a.H && ic != cur_ic) {
        // the response row of the new channel (the previous tile's readers passed three cluster barriers since)
        const float2* Hc = a.H + uint64_t(ic) * N + rank;
        for (unsigned i = tid; i < Q; i += NT) Hs[i] = __ldg(Hc + 16u * i);
      }
      cur_ic = ic;
      dft16<false>(xp);
      dft16<false>(xq);
      const unsigned slot = sbuf + c2::pad16(n2) * 16u;
      // barrier 4 of the previous tile (arrived after its last shared-memory read): every receive buffer of the
      // cluster has been read, this tile may overwrite them
      if (t != t_begin) cluster_wait();
#pragma unroll
      for (int k1 = 0; k1 < 16; k1++) {
        float2 u = xp[k1], v = xq[k1];
        if (k1) {
          float2 wk = make_float2(1.f, 0.f);
          bool first = true;
#pragma unroll
          for (int b = 3; b >= 0; b--)
            if (k1 & (1 << b)) {
              wk = first ? w[b] : cmul(wk, w[b]);
              first = false;
            }
          u = cmul(u, wk);
          v = cmul(v, wk);
        }
        st_remote(map_remote(slot, (CC_DBG & 4) ? rank : unsigned(k1)), make_float4(u.x, u.y, v.x, v.y));
      }
    }
    cluster_sync();

    // ---- phase B: forward row, response, inverse row ----
    float2 va[16], vb[16];
    {
      c2::gather<Q>(buf, va, vb, tid);
      __syncthreads();
      if (!(CC_DBG & 2)) c2::fft_pair<Q, false>(va, vb, tid, buf, a.tw, CcSync());
      if (a.H) {
        const float2* h = Hs + tid;
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const float2 hv = h[e * int(T)];
          va[e] = cmul(va[e], hv);
          vb[e] = cmul(vb[e], hv);
        }
      }
      __syncthreads();                      // every thread has gathered the last stage of the forward transform
      if (!(CC_DBG & 2)) c2::fft_pair<Q, true>(va, vb, tid, buf, a.tw, CcSync());
    }
    cluster_arrive();                       // barrier 2: this CTA no longer reads its pair buffer ...

    // ---- phase C: element (k1, m2 = j + 256 e) to the owner of column m2 (rank e) ----
    {
      const unsigned slot = sbuf + c2::pad16(rank * COLS + tid) * 16u;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const float2 w = cmul(wown, s_tw[e]);
        va[e] = cmul(va[e], w);
        vb[e] = cmul(vb[e], w);
      }
      cluster_wait();                       // ... nor does any other: the pair buffers become receive buffers
#pragma unroll
      for (int e = 0; e < 16; e++) st_remote(map_remote(slot, (CC_DBG & 4) ? rank : unsigned(e)), make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y));
    }
    cluster_arrive();                       // barrier 3: this CTA's columns are on their way

    // ---- phase D: inverse 16-point column transforms, detection, fold ----
    {
      // thread = 16 consecutive samples of segment m1 = tid / 16: transform index Q m1 + 256 rank + 16 (tid % 16) + i;
      // their phase bins are requested while the columns of the other CTAs arrive
      const unsigned seg = tid >> 4, b16 = tid & 15u;
      const unsigned t0 = Q * seg + COLS * rank + 16u * b16;
      const unsigned* plan = a.sink.bins + uint64_t(partl) * nkeep;
      unsigned bins16[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const unsigned u = t0 + unsigned(i) - np0;        // unsigned: samples before nfilt_pos wrap to huge values
        bins16[i] = u < nkeep ? __ldg(plan + u) : 0xfffffffeu;
      }
      cluster_wait();                       // barrier 3: all columns have arrived
      {
        float2 yp[16], yq[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
          const float4 x = buf[c2::pad16(unsigned(k1) * COLS + tid)];
          yp[k1] = make_float2(x.x, x.y);
          yq[k1] = make_float2(x.z, x.w);
        }
        dft16<true>(yp);
        dft16<true>(yq);
#pragma unroll
        for (int m1 = 0; m1 < 16; m1++) {
          float r[4] = {0.f, 0.f, 0.f, 0.f};
          detect_products(state, yp[m1], yq[m1], r);
          buf[c2::pad16(unsigned(m1) * COLS + tid)] = make_float4(r[0], r[1], r[2], r[3]);   // the slot this thread read
        }
      }
      __syncthreads();
      const float4* d = buf + c2::pad16(seg * COLS + 16u * b16);
      float4 x[16];
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = d[i];
      // barrier 4 (waited for in the next tile's phase A): the last shared-memory read of this tile is done
      if (t + 1 < t_end) cluster_arrive();
      const uint64_t prof0 = uint64_t(ic) * nbin * nprod;
      unsigned cur = 0xffffffffu;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      auto flush = [&]() {
        if (cur < nbin)                                   // nbin: samples of a flagged window; 0xffffffff: nothing yet
          for (unsigned pr = 0; pr < nprod; pr++)
            profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, prof0 + (uint64_t(pr / dndim) * nbin + cur) * dndim + pr % dndim,
                        acc[pr]);
      };
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const unsigned bin = bins16[i];
        if ((CC_DBG & 1) && x[i].x != 12345.678f) continue;
        if (bin == 0xfffffffeu) continue;                 // discarded by overlap-save
        if (bin != cur) {
          flush();
          cur = bin;
          acc[0] = x[i].x; acc[1] = x[i].y; acc[2] = x[i].z; acc[3] = x[i].w;
        } else {
          acc[0] += x[i].x; acc[1] += x[i].y; acc[2] += x[i].z; acc[3] += x[i].w;
        }
      }
      flush();
    }
  }
  cluster_sync();                           // no CTA leaves while its shared memory may still be written or read
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int SRC>
static int cc_prepare(int* max_clusters) {
  B200_CUDA(cudaFuncSetAttribute(k_conv64k<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc::SMEM));
  B200_CUDA(cudaFuncSetAttribute(k_conv64k<SRC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // clusters of 16
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cc::CL * 64);
  cfg.blockDim = dim3(cc::NT);
  cfg.dynamicSmemBytes = cc::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cc::CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  B200_CUDA(cudaOccupancyMaxActiveClusters(&n, k_conv64k<SRC>, &cfg));
  *max_clusters = n;
  return B200_OK;
}

int cc_plan_init(b200_fb_plan* pl) {
  pl->c2cc = nullptr;
  pl->cc_clusters = 0;
  static const bool want = tune_flag("B200_CLUSTER_CONV", true);
  if (!want || !pl->conv_path || pl->Nc != cc::N || pl->desc.input_real || pl->desc.npol != 2) return B200_OK;
  if (size_t(pl->ctx->max_smem_optin) < 2 * (cc::SMEM + 1024)) return B200_OK;
  std::vector<float2> h(c2::twiddle_count<cc::Q>(), make_float2(1.f, 0.f));
  c2::fill_twiddles<cc::Q>(h.data());
  B200_CUDA(cudaMalloc(&pl->c2cc, sizeof(float2) * h.size()));
  B200_CUDA(cudaMemcpy(pl->c2cc, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
  int n0 = 0, n1 = 0, n2 = 0, n3 = 0, rc;
  if ((rc = cc_prepare<SRC_F32>(&n0)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_MEERKAT8>(&n1)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_UWB16>(&n2)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_GENERIC8>(&n3)) != B200_OK) return rc;
  pl->cc_clusters = std::min(std::min(n0, n1), std::min(n2, n3));
  return B200_OK;
}

void cc_plan_free(b200_fb_plan* pl) {
  if (pl->c2cc) cudaFree(pl->c2cc);
  pl->c2cc = nullptr;
}

bool cc_applies(const b200_fb_plan* pl, const FbSource& src, const FbSink& sink) {
  return pl->cc_clusters > 0 && pl->c2cc && sink.kind == EPI_FOLD &&
         (src.kind == SRC_F32 || src.kind == SRC_MEERKAT8 || src.kind == SRC_UWB16 ||
          (src.kind == SRC_GENERIC8 && src.ndim == 2));
}

int cc_run(b200_fb_plan* pl, const FbSource& src, const FbSink& sk, uint64_t part0, unsigned nb) {
  Context* ctx = pl->ctx;
  CcArgs a;
  a.src = src.ptr; a.span = src.span; a.step = src.step; a.first = src.first; a.scale = src.scale; a.sample_swap = src.sample_swap;
  a.lut = src.d_lut;
  a.H = pl->d_response; a.tw = pl->c2cc; a.blo = pl->bigN.lo; a.bhi = pl->bigN.hi;
  a.nchan_in = pl->desc.input_nchan; a.nb = nb; a.ntiles = nb * pl->desc.input_nchan;
  a.part0 = part0; a.nfilt_pos = pl->desc.nfilt_pos; a.nkeep = pl->nkeep; a.sink = sk;
  const unsigned ncl = std::min<unsigned>(a.ntiles, (unsigned)pl->cc_clusters);
  a.tiles_per_cluster = (a.ntiles + ncl - 1) / ncl;
  const unsigned used = (a.ntiles + a.tiles_per_cluster - 1) / a.tiles_per_cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cc::CL * used);
  cfg.blockDim = dim3(cc::NT);
  cfg.dynamicSmemBytes = cc::SMEM;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cc::CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LaunchScope ls(ctx, KC_INV);
  if (src.kind == SRC_F32) B200_CUDA(cudaLaunchKernelEx(&cfg, k_conv64k<SRC_F32>, a));
  else if (src.kind == SRC_MEERKAT8) B200_CUDA(cudaLaunchKernelEx(&cfg, k_conv64k<SRC_MEERKAT8>, a));
  else if (src.kind == SRC_GENERIC8) B200_CUDA(cudaLaunchKernelEx(&cfg, k_conv64k<SRC_GENERIC8>, a));
  else B200_CUDA(cudaLaunchKernelEx(&cfg, k_conv64k<SRC_UWB16>, a));
  return B200_OK;
}

}  // namespace b200
