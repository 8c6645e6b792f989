// clusterconv.cu -- the convolution path of 16384- to 131072-point transforms (cfg3: 65536 points, MeerKAT, one
// coherent-dedispersion transform per input channel) as ONE kernel: unpack -> forward FFT -> response -> inverse FFT -> discard -> detect ->
// fold without a single spectrum round trip through HBM.
//
// dsp::Convolution::transformation (Convolution.C:389-458) per (channel, part): both polarisations, N = 65536
// complex points = 1 MiB of float2 -- more than one SM holds, exactly what a GROUP of 16 CTAs holds in registers
// (16 x 256 threads x 16 points x 2 polarisations; two CTAs of different groups share an SM, so one group's barrier
// waits and exchanges run under the other's butterflies).  Factorisation N = P Q, P = 16, Q = 4096
// (n = Q n1 + n2, k = k1 + P k2):
//   A  every thread owns ONE column n2 of both polarisations: 16 samples Q apart -> a 16-point DFT in registers,
//      times W_N^(n2 k1); element k1 goes to the CTA that owns row k1 (first exchange; the two polarisations of a point
//      travel together as one float4)
//   B  every CTA owns one row k1 (both polarisations = one fft_c2 sequence pair): forward 4096-point row
//      transform, times the response H[k1 + 16 k2] (the CTA's response row stays in shared memory for as long
//      as the group works on the same channel), inverse 4096-point row transform, times W_N^(-m2 k1)
//   C  element (k1, m2) goes back to the CTA that owns column m2 (second exchange)
//   D  every thread owns one column m2 again: inverse 16-point DFT over k1 -> y[Q m1 + m2], m1 < 16; detection of
//      both polarisations from registers; the detected products go to shared memory in time order (16 segments of
//      256 consecutive samples per CTA) and one thread per 16 samples walks them with the bin plan: runs of one
//      phase bin are summed sequentially (the order of Fold.C:844-852) and added with one RED.ADD.F32 per product.
//
// Two implementations of the group and its exchanges, both parity-tested, selected at compile time:
//   * product (CC_GROUP = 1): groups of 16 co-resident CTAs of a COOPERATIVE launch.  The exchanges go through two
//     pairs of 1 MiB matrices per group in global memory that never leave the 126 MB L2 (written with plain stores,
//     read with ld.global.cg); the group synchronises through an arrival counter in global memory (release fence +
//     atomic, acquire spin).  Double buffering makes TWO barriers per tile sufficient and lets phase A of the next
//     tile run between the arrive and the wait of the second one.  18 groups = 288 CTAs = 144 of the 148 SMs.
//   * -DCC_GROUP=0: hardware thread-block clusters of 16 CTAs, exchanges through DISTRIBUTED SHARED MEMORY
//     (st.shared::cluster.v4 + barrier.cluster arrive / wait, four per tile).  Measured on B200:
//     DSMEM moves 17 bytes per cycle and SM, a third of what the same SM gets from L2, and a 16-CTA cluster must sit
//     in one GPC, which leaves 14 clusters = 112 SMs resident: 4.67 ms per 16 parts x 128 channels against 3.4 ms
//     for the product variant (three-kernel path: 4.72 ms).
// HBM traffic per (channel, part) in both: the raw bytes, the bin plan (L2) and the REDs -- the three-kernel path moves
// the 1 MiB spectrum through HBM four times.  Groups are persistent; each takes a contiguous range of (channel, part)
// tiles so that the response rows are re-staged only when the channel changes.
#include <algorithm>
#include <vector>

#include "clusterconv.cuh"

#ifndef CC_DBG
#define CC_DBG 0
#endif
// CC_GROUP = 1 (product): cooperative groups of 16 CTAs, exchanges through L2-resident matrices, counter barriers;
// CC_GROUP = 0: hardware clusters, exchanges through distributed shared memory (see the header)
#ifndef CC_GROUP
#define CC_GROUP 1
#endif

namespace b200 {

// sizes of one transform length N = 16 Q (Q = 1024 ... 8192: N = 16384 ... 131072; cfg3: Q = 4096)
template <unsigned QQ>
struct Cc {
  static constexpr unsigned Q = QQ, N = 16 * QQ, P = 16, CL = 16;
  static constexpr unsigned NT = QQ / 16;              // threads per CTA = threads of the row pair = columns per CTA
  static constexpr unsigned COLS = NT, T = NT;
  static constexpr unsigned PS = c2::pair_slots<QQ>(); // float4 slots of the pair buffer
  static constexpr size_t SMEM = size_t(PS) * sizeof(float4) + size_t(QQ) * sizeof(float2);
  static_assert(PS == COLS * 16 + COLS, "the staging area of phase D (16 x NT padded) is the pair buffer");
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


// The 128-byte line at p is dead: drop it from L2 without writing it back (the exchange matrices are rewritten every
// other tile; without this their dirty lines are evicted to DRAM by the raw-data stream long before that)
__device__ __forceinline__ void l2_discard(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory");
}



template <int SRC, unsigned QQ>
__global__ void __launch_bounds__(QQ / 16, 512 / (QQ / 16)) k_conv64k(CcArgs a) {
  using C = Cc<QQ>;
  constexpr unsigned Q = C::Q, N = C::N, CL = C::CL, NT = C::NT, COLS = C::COLS, T = C::T, PS = C::PS;
  extern __shared__ __align__(16) float4 buf[];                    // the pair buffer of the row transforms; phase D staging
  float2* Hs = reinterpret_cast<float2*>(buf + PS);                // response row k1 = rank: [Q]
  __shared__ float2 s_tw[16];                                      // W_N^(-k1 NT e) of the CTA's row
  __shared__ float s_lut[SRC == SRC_GENERIC8 ? 256 : 1];
  if (SRC == SRC_GENERIC8)
    for (unsigned i = threadIdx.x; i < 256; i += NT) s_lut[i] = a.lut[i];
  const unsigned tid = threadIdx.x;
#if CC_GROUP
  const unsigned rank = blockIdx.x % CL, group = blockIdx.x / CL;  // rank = the row k1 this CTA owns
  // Two arrival counters per group, one per barrier kind (X complete / Y complete): a CTA arrives at the next barrier
  // of one kind only after it has waited for the previous one of that kind, so no CTA is ever more than one arrival
  // ahead on a counter and "count >= 16 k" means that all 16 have made their k-th arrival -- which a single counter
  // would not guarantee once two arrivals may precede a wait.
  unsigned* const ctr = a.bar + 64u * group;
  // arrive: every global write of this CTA so far is visible to whoever sees the count
  auto arrive = [&](unsigned kind) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(ctr + 32u * kind, 1u);
    }
  };
  // wait: all 16 CTAs of the group have made their k-th arrival of this kind
  auto wait = [&](unsigned kind, unsigned k) {
    if (tid == 0) {
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + 32u * kind) : "memory");
      } while (v < CL * k);
    }
    __syncthreads();
  };
  float4* const XY = a.xch + uint64_t(group) * (4u * 16u * Q);     // X0, X1, Y0, Y1
#else
  const unsigned rank = cluster_ctarank(), group = cluster_id_x(); // rank = the row k1 this CTA owns
  const unsigned sbuf = (unsigned)__cvta_generic_to_shared(buf);
#endif
  if (tid < 16) s_tw[tid] = big_twiddle<true>(a.blo, a.bhi, (rank * NT * tid) & (N - 1));
  const float2 wown = big_twiddle<true>(a.blo, a.bhi, rank * tid); // W_N^(-k1 j)
  const unsigned n2 = COLS * rank + tid;                           // phases A / D: this thread's column
  const unsigned np0 = a.nfilt_pos, nkeep = a.nkeep;
  const int state = a.sink.state;
  const unsigned nprod = a.sink.kind == EPI_VOLT ? 1 : state_nprod(state, 2), dndim = a.sink.kind == EPI_VOLT ? 1 : a.sink.dndim;
  const unsigned nbin = a.sink.nbin;
  const unsigned t_begin = group * a.tiles_per_cluster;
  const unsigned t_end = min(a.ntiles, t_begin + a.tiles_per_cluster);
  unsigned cur_ic = 0xffffffffu;

  // ---- phase A: 16-point column transforms of both polarisations; element k1 to the owner of row k1 ----
  // CC_GROUP: into X[(t - t_begin) & 1][k1][n2]; clusters: into the owner's pair buffer (natural order, padded);
  // `pre_store` runs between the transforms and the stores (clusters: wait for the receive buffers to be free)
  auto phase_a = [&](unsigned t, auto pre_store) {
    const unsigned ic = t / a.nb;
    const uint64_t part = a.part0 + t % a.nb;
    float2 xp[16], xq[16];
    if (SRC == SRC_MEERKAT8) {
      // heaps of 256 samples, [heap][pol][chan][256 x (re, im) int8] (MeerKATUnpacker.C:196-229): the 16 samples of a
      // column are Q = (Q / 256) heaps apart, so one address computation serves all of them
      uint64_t i = a.first + part * a.step + n2;
      if (a.sample_swap == 2) i ^= 1ull;
      const unsigned short* w0 = static_cast<const unsigned short*>(a.src) + ((i >> 8) * 2 * a.nchan_in + ic) * 256ull + (i & 255ull);
      const uint64_t dheap = uint64_t(Q / 256u) * 2u * a.nchan_in * 256u, dpol = uint64_t(a.nchan_in) * 256u;
#pragma unroll
      for (int n1 = 0; n1 < 16; n1++) {
        const unsigned short wp = __ldg(w0 + n1 * dheap), wq = __ldg(w0 + n1 * dheap + dpol);
        xp[n1] = make_float2(__fmul_rn(float(int(int8_t(wp & 255u))) + 0.5f, a.scale),
                             __fmul_rn(float(int(int8_t(wp >> 8))) + 0.5f, a.scale));
        xq[n1] = make_float2(__fmul_rn(float(int(int8_t(wq & 255u))) + 0.5f, a.scale),
                             __fmul_rn(float(int(int8_t(wq >> 8))) + 0.5f, a.scale));
      }
    } else {
#pragma unroll
      for (int n1 = 0; n1 < 16; n1++) {
        xp[n1] = cc_load<SRC>(a, s_lut, ic, 0, part, n2 + Q * n1);
        xq[n1] = cc_load<SRC>(a, s_lut, ic, 1, part, n2 + Q * n1);
      }
    }
    // W_N^(n2 k1), k1 < 16: four table values (k1 = 1, 2, 4, 8), the others as products over the bits of k1
    float2 w[4];
#pragma unroll
    for (int b = 0; b < 4; b++) w[b] = big_twiddle<false>(a.blo, a.bhi, (n2 << b) & (N - 1));
    if (a.H && ic != cur_ic) {
      // the response row of the new channel (every reader of the old one is behind a CTA barrier: the row transforms)
      const float2* Hc = a.H + uint64_t(ic) * N + rank;
      for (unsigned i = tid; i < Q; i += NT) Hs[i] = __ldg(Hc + 16u * i);
    }
    cur_ic = ic;
    dft16<false>(xp);
    dft16<false>(xq);
    pre_store();
#if CC_GROUP
    float4* X = XY + ((t - t_begin) & 1u) * (16u * Q) + n2;
#else
    const unsigned slot = sbuf + c2::pad16(n2) * 16u;
#endif
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
#if CC_GROUP
      X[unsigned(k1) * Q] = make_float4(u.x, u.y, v.x, v.y);
#else
      st_remote(map_remote(slot, (CC_DBG & 4) ? rank : unsigned(k1)), make_float4(u.x, u.y, v.x, v.y));
#endif
    }
  };

  // ---- phase B: forward row, response, inverse row, times W_N^(-m2 k1); register e = element m2 = tid + NT e ----
  auto phase_b = [&](unsigned t, float2* va, float2* vb) {
#if CC_GROUP
    const float4* X = XY + ((t - t_begin) & 1u) * (16u * Q) + rank * Q + tid;
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const float4 x = ld_cg_f4(X + NT * e);
      va[e] = make_float2(x.x, x.y);
      vb[e] = make_float2(x.z, x.w);
    }
#else
    c2::gather<Q>(buf, va, vb, tid);
    __syncthreads();
#endif
    // Both row transforms run through ONE copy of the forward code (the kernel is instruction-cache bound: two CTAs in
    // different phases share an SM): the inverse is conj(FFT(conj z)), which is the same IEEE operations as the
    // conjugate-twiddle inverse with the signs of the imaginary parts flipped -- bit-identical results.
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      if (pass) {
        if (a.H) {
          const float2* h = Hs + tid;
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const float2 hv = h[e * int(T)];
            va[e] = cmul(va[e], hv);
            vb[e] = cmul(vb[e], hv);
          }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) {
          va[e].y = -va[e].y;
          vb[e].y = -vb[e].y;
        }
        __syncthreads();                    // every thread has gathered the last stage of the forward transform
      }
      if (!(CC_DBG & 2)) c2::fft_pair<Q, false>(va, vb, tid, buf, a.tw, CcSync());
#if CC_GROUP
      // the row has been read (every lane's loads were consumed before the barriers of the transform): its lines are
      // dead -- a line is the 8 consecutive float4 of 8 consecutive threads, read by nobody else
      if (!pass && (tid & 7u) == 0) {
#pragma unroll
        for (int e = 0; e < 16; e++) l2_discard(X + NT * e);
      }
#endif
    }
#pragma unroll
    for (int e = 0; e < 16; e++) {
      va[e].y = -va[e].y;
      vb[e].y = -vb[e].y;
    }
  };
  auto twiddle_back = [&](float2* va, float2* vb) {
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const float2 w = cmul(wown, s_tw[e]);
      va[e] = cmul(va[e], w);
      vb[e] = cmul(vb[e], w);
    }
  };

  // ---- phase D: inverse 16-point column transforms, detection, fold ----
  // thread = 16 consecutive samples of segment m1 = tid / (NT / 16): transform index Q m1 + NT rank + 16 (tid % (NT / 16)) + i.
  // `columns_ready` runs after the phase bins have been requested and before the columns are read; `reads_done` after
  // the last shared-memory read
  auto phase_d = [&](unsigned t, auto columns_ready, auto reads_done) {
    const unsigned ic = t / a.nb, partl = t % a.nb;
    const unsigned seg = tid / (NT / 16u), b16 = tid % (NT / 16u);
    const unsigned t0 = Q * seg + COLS * rank + 16u * b16;
    const bool folding = a.sink.kind == EPI_FOLD;         // uniform over the grid
    const unsigned* plan = a.sink.bins + uint64_t(partl) * nkeep;
    columns_ready();
    {
      float2 yp[16], yq[16];
#if CC_GROUP
      const float4* Y = XY + (2u + ((t - t_begin) & 1u)) * (16u * Q) + n2;
#endif
#pragma unroll
      for (int k1 = 0; k1 < 16; k1++) {
#if CC_GROUP
        const float4 x = ld_cg_f4(Y + unsigned(k1) * Q);
#else
        const float4 x = buf[c2::pad16(unsigned(k1) * COLS + tid)];
#endif
        yp[k1] = make_float2(x.x, x.y);
        yq[k1] = make_float2(x.z, x.w);
      }
      dft16<true>(yp);
      dft16<true>(yq);
#if CC_GROUP
      __syncwarp();                         // every lane's column loads have been consumed
      if ((tid & 7u) == 0) {
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) l2_discard(Y + unsigned(k1) * Q);
      }
#endif
      if (!folding) {
        // voltages (Convolution::Engine::perform) or the detected series: this thread's 16 samples are Q apart, the
        // threads of a warp write 32 consecutive samples of a plane
        const uint64_t part = a.part0 + partl;
        if (a.sink.kind == EPI_VOLT) {
          float2* outp = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(ic) * 2) * a.sink.volt_span + part * a.sink.volt_step);
          float2* outq = reinterpret_cast<float2*>(a.sink.volt + (uint64_t(ic) * 2 + 1) * a.sink.volt_span + part * a.sink.volt_step);
#pragma unroll
          for (int m1 = 0; m1 < 16; m1++) {
            const unsigned u = Q * unsigned(m1) + n2 - np0;
            if (u < nkeep) {
              outp[u] = yp[m1];
              outq[u] = yq[m1];
            }
          }
        } else {
          const unsigned dnpol = nprod / dndim;
#pragma unroll
          for (int m1 = 0; m1 < 16; m1++) {
            const unsigned u = Q * unsigned(m1) + n2 - np0;
            if (u < nkeep) {
              float r[4] = {0.f, 0.f, 0.f, 0.f};
              detect_products(state, yp[m1], yq[m1], r);
              const uint64_t osamp = part * nkeep + u;
              for (unsigned pr = 0; pr < nprod; pr++)
                a.sink.det[(uint64_t(ic) * dnpol + pr / dndim) * a.sink.det_span + osamp * dndim + pr % dndim] = r[pr];
            }
          }
        }
        reads_done();
        return;
      }
#pragma unroll
      for (int m1 = 0; m1 < 16; m1++) {
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        detect_products(state, yp[m1], yq[m1], r);
        buf[c2::pad16(unsigned(m1) * COLS + tid)] = make_float4(r[0], r[1], r[2], r[3]);   // clusters: the slot this thread read
      }
    }
    __syncthreads();
    // one rolled loop (the kernel is instruction-cache bound): this thread's 16 samples are consecutive slots of the
    // padded staging area, their phase bins 64 consecutive bytes of the bin plan (L1 after the first look)
    const float4* d = buf + c2::pad16(seg * COLS + 16u * b16);
    // profile element of product pr and bin b: prof0 + (pr / dndim * nbin + b) * dndim + pr % dndim = off[pr] + b * dndim
    const uint64_t prof0 = uint64_t(ic) * nbin * nprod;
    uint64_t off[4];
#pragma unroll
    for (unsigned pr = 0; pr < 4; pr++) off[pr] = prof0 + uint64_t(pr / dndim) * nbin * dndim + pr % dndim;
    unsigned cur = 0xffffffffu;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (unsigned i = 0; i <= 16u; i++) {
      const unsigned u = t0 + i - np0;                    // unsigned: samples before nfilt_pos wrap to huge values
      // i = 16: flush the last run;  nbin: samples of a flagged window;  0xfffffffe: discarded by overlap-save
      const unsigned bin = i == 16u ? 0xffffffffu : (u < nkeep ? __ldg(plan + u) : 0xfffffffeu);
      if ((CC_DBG & 1) && bin != 12345678u) continue;
      if (bin == 0xfffffffeu) continue;
      if (bin != cur) {
        if (cur < nbin) {
#pragma unroll
          for (unsigned pr = 0; pr < 4; pr++)
            if (pr < nprod) profile_add(a.sink.profile, a.sink.fix, a.sink.inv_lsb, off[pr] + uint64_t(cur) * dndim, acc[pr]);
        }
        if (i == 16u) break;
        cur = bin;
        const float4 x = d[i];
        acc[0] = x.x; acc[1] = x.y; acc[2] = x.z; acc[3] = x.w;
      } else {
        const float4 x = d[i];
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
      }
    }
    reads_done();
  };
  auto nothing = [] {};

#if CC_GROUP
  // Buffers alternate with the tile, so a matrix is rewritten two tiles after it was read: by then every CTA of the
  // group has passed a barrier that followed its reads.  Barriers per tile: a1 (X complete) and a3 (Y complete); phase A
  // of the NEXT tile runs between the arrive and the wait of a3, under the group's skew.
  if (t_begin < t_end) {
    phase_a(t_begin, nothing);
    arrive(0);
  }
  for (unsigned t = t_begin; t < t_end; t++) {
    const unsigned k = t - t_begin + 1;       // this tile's barriers are the k-th of their kind
    float2 va[16], vb[16];
    wait(0, k);
    phase_b(t, va, vb);
    twiddle_back(va, vb);
    {
      float4* Y = XY + (2u + ((t - t_begin) & 1u)) * (16u * Q) + rank * Q + tid;
#pragma unroll
      for (int e = 0; e < 16; e++) Y[NT * e] = make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y);
    }
    arrive(1);
    if (t + 1 < t_end) {
      phase_a(t + 1, nothing);
      arrive(0);
    }
    phase_d(t, [&] { wait(1, k); }, nothing);
  }
#else
  cluster_sync();                           // every CTA of the cluster is running: its shared memory may be written
  for (unsigned t = t_begin; t < t_end; t++) {
    // barrier 4 of the previous tile (arrived after its last shared-memory read): every receive buffer of the cluster
    // has been read, this tile may overwrite them
    phase_a(t, [&] { if (t != t_begin) cluster_wait(); });
    cluster_sync();
    float2 va[16], vb[16];
    phase_b(t, va, vb);
    cluster_arrive();                       // barrier 2: this CTA no longer reads its pair buffer ...
    twiddle_back(va, vb);
    cluster_wait();                         // ... nor does any other: the pair buffers become receive buffers
    {
      // element (k1, m2 = tid + NT e) to the owner of column m2 (rank e)
      const unsigned slot = sbuf + c2::pad16(rank * COLS + tid) * 16u;
#pragma unroll
      for (int e = 0; e < 16; e++)
        st_remote(map_remote(slot, (CC_DBG & 4) ? rank : unsigned(e)), make_float4(va[e].x, va[e].y, vb[e].x, vb[e].y));
    }
    cluster_arrive();                       // barrier 3: this CTA's columns are on their way
    phase_d(t, [&] { cluster_wait(); }, [&] { if (t + 1 < t_end) cluster_arrive(); });
  }
  cluster_sync();                           // no CTA leaves while its shared memory may still be written or read
#endif
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int SRC, unsigned Q>
static int cc_prepare(int* max_groups) {
  using C = Cc<Q>;
  B200_CUDA(cudaFuncSetAttribute(k_conv64k<SRC, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  int n = 0;
#if CC_GROUP
  int dev = 0, sms = 0, per_sm = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv64k<SRC, Q>, C::NT, C::SMEM));
  n = per_sm * sms / int(C::CL);
#else
  B200_CUDA(cudaFuncSetAttribute(k_conv64k<SRC, Q>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // clusters of 16
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C::CL * 64);
  cfg.blockDim = dim3(C::NT);
  cfg.dynamicSmemBytes = C::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C::CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CUDA(cudaOccupancyMaxActiveClusters(&n, k_conv64k<SRC, Q>, &cfg));
#endif
  *max_groups = n;
  return B200_OK;
}

template <unsigned Q>
static int cc_init_q(b200_fb_plan* pl) {
  using C = Cc<Q>;
  if (size_t(pl->ctx->max_smem_optin) < C::SMEM + 1024) return B200_OK;
  std::vector<float2> h(c2::twiddle_count<Q>(), make_float2(1.f, 0.f));
  c2::fill_twiddles<Q>(h.data());
  B200_CUDA(cudaMalloc(&pl->c2cc, sizeof(float2) * h.size()));
  B200_CUDA(cudaMemcpy(pl->c2cc, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
  int n0 = 0, n1 = 0, n2 = 0, n3 = 0, rc;
  if ((rc = cc_prepare<SRC_F32, Q>(&n0)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_MEERKAT8, Q>(&n1)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_UWB16, Q>(&n2)) != B200_OK) return rc;
  if ((rc = cc_prepare<SRC_GENERIC8, Q>(&n3)) != B200_OK) return rc;
  pl->cc_clusters = std::min(std::min(n0, n1), std::min(n2, n3));
  if (CC_GROUP && pl->cc_clusters > 0) {
    // four exchange matrices of 16 Q float4 and two arrival counters (each on its own 128-byte line) per group
    B200_CUDA(cudaMalloc(&pl->cc_xch, size_t(pl->cc_clusters) * 4 * 16 * Q * sizeof(float4)));
    B200_CUDA(cudaMalloc(&pl->cc_bar, size_t(pl->cc_clusters) * 64 * sizeof(unsigned)));
  }
  return B200_OK;
}

int cc_plan_init(b200_fb_plan* pl) {
  pl->c2cc = nullptr;
  pl->cc_xch = nullptr;
  pl->cc_bar = nullptr;
  pl->cc_clusters = 0;
  static const bool want = tune_flag("B200_CLUSTER_CONV", true);
  if (!want || !pl->conv_path || pl->desc.input_real || pl->desc.npol != 2) return B200_OK;
  switch (pl->Nc) {
    case 16 * 1024: return cc_init_q<1024>(pl);
    case 16 * 2048: return cc_init_q<2048>(pl);
    case 16 * 4096: return cc_init_q<4096>(pl);
    case 16 * 8192: return cc_init_q<8192>(pl);
    default: return B200_OK;
  }
}

void cc_plan_free(b200_fb_plan* pl) {
  if (pl->c2cc) cudaFree(pl->c2cc);
  pl->c2cc = nullptr;
  if (pl->cc_xch) cudaFree(pl->cc_xch);
  pl->cc_xch = nullptr;
  if (pl->cc_bar) cudaFree(pl->cc_bar);
  pl->cc_bar = nullptr;
}

bool cc_applies(const b200_fb_plan* pl, const FbSource& src, const FbSink& sink) {
  return pl->cc_clusters > 0 && pl->c2cc &&
         (src.kind == SRC_F32 || src.kind == SRC_MEERKAT8 || src.kind == SRC_UWB16 ||
          (src.kind == SRC_GENERIC8 && src.ndim == 2));
}

template <unsigned Q>
static int cc_launch(b200_fb_plan* pl, const FbSource& src, CcArgs& a) {
  using C = Cc<Q>;
  Context* ctx = pl->ctx;
  const unsigned ncl = std::min<unsigned>(a.ntiles, (unsigned)pl->cc_clusters);
  a.tiles_per_cluster = (a.ntiles + ncl - 1) / ncl;
  const unsigned used = (a.ntiles + a.tiles_per_cluster - 1) / a.tiles_per_cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C::CL * used);
  cfg.blockDim = dim3(C::NT);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
#if CC_GROUP
  at[0].id = cudaLaunchAttributeCooperative;            // all CTAs resident: the counter barriers cannot deadlock
  at[0].val.cooperative = 1;
  B200_CUDA(cudaMemsetAsync(pl->cc_bar, 0, size_t(pl->cc_clusters) * 64 * sizeof(unsigned), ctx->stream));
#else
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C::CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
#endif
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
  {
    LaunchScope ls(ctx, KC_INV);
    if (src.kind == SRC_F32) e = cudaLaunchKernelEx(&cfg, k_conv64k<SRC_F32, Q>, a);
    else if (src.kind == SRC_MEERKAT8) e = cudaLaunchKernelEx(&cfg, k_conv64k<SRC_MEERKAT8, Q>, a);
    else if (src.kind == SRC_GENERIC8) e = cudaLaunchKernelEx(&cfg, k_conv64k<SRC_GENERIC8, Q>, a);
    else e = cudaLaunchKernelEx(&cfg, k_conv64k<SRC_UWB16, Q>, a);
  }
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
    // the device cannot keep the whole grid resident right now (shared with another context): nothing was launched;
    // this plan goes back to the three-kernel path for good
    cudaGetLastError();
    pl->cc_clusters = 0;
    return CC_NOT_RUN;
  }
  B200_CUDA(e);
  return B200_OK;
}

int cc_run(b200_fb_plan* pl, const FbSource& src, const FbSink& sk, uint64_t part0, unsigned nb) {
  CcArgs a;
  a.src = src.ptr; a.span = src.span; a.step = src.step; a.first = src.first; a.scale = src.scale; a.sample_swap = src.sample_swap;
  a.lut = src.d_lut;
  a.H = pl->d_response; a.tw = pl->c2cc; a.blo = pl->bigN.lo; a.bhi = pl->bigN.hi;
  a.nchan_in = pl->desc.input_nchan; a.nb = nb; a.ntiles = nb * pl->desc.input_nchan;
  a.xch = static_cast<float4*>(pl->cc_xch);
  a.bar = static_cast<unsigned*>(pl->cc_bar);
  a.part0 = part0; a.nfilt_pos = pl->desc.nfilt_pos; a.nkeep = pl->nkeep; a.sink = sk;
  switch (pl->Nc) {
    case 16 * 1024: return cc_launch<1024>(pl, src, a);
    case 16 * 2048: return cc_launch<2048>(pl, src, a);
    case 16 * 4096: return cc_launch<4096>(pl, src, a);
    default: return cc_launch<8192>(pl, src, a);
  }
}


}  // namespace b200
