// engine.cuh -- internal interface between the filterbank engine (filterbank.cu), the fold
// engine (fold.cu) and the fused pipeline (pipeline.cu).
#pragma once
#include "common.cuh"

namespace b200 {

enum SrcKind { SRC_F32 = 0, SRC_CASPSR8 = 1, SRC_MEERKAT8 = 2, SRC_UWB16 = 3, SRC_GENERIC8 = 4, SRC_TWOBIT = 5 };
enum Epilogue { EPI_VOLT = 0, EPI_DETECT = 1, EPI_FOLD = 2 };

// where the forward transform reads its samples
struct FbSource {
  int kind;                 // SrcKind
  const void* ptr;          // float planes (SRC_F32) or raw bytes (SRC_CASPSR8)
  uint64_t span;            // SRC_F32: floats between (chan,pol) planes
  uint64_t step;            // SRC_F32: floats between parts; raw: SAMPLES between parts
  const float* d_lut;       // raw 8-bit formats: device copy of the 256-entry table
  cudaEvent_t* batch_ready; // optional: event i must have fired before the i-th internal batch reads its input
  unsigned batch_override;  // parts per internal batch for this call (0 = the plan's), <= the plan's batch
  // two's-complement 8-bit tables that are exactly  lut[b] = RN(x * c), x = int8(b) + 0.5, c = conv_hi + conv_lo
  // (checked entry by entry on the host): the fast path converts arithmetically instead of gathering
  int conv_ok;
  float conv_hi, conv_lo;
  // raw formats unpacked inside the generic K1 (MeerKAT heaps, UWB blocks, generic TFP bytes): ptr is the start of a
  // stream that begins on a boundary of the format's resolution, `first` the first sample of part 0 in it
  uint64_t first;
  float scale;              // MeerKAT: (float(x) + 0.5) * scale
  unsigned sample_swap;     // MeerKAT: 2 = odd/even samples exchanged (MKBFRo)
  unsigned ndim;            // generic 8-bit: 1 real, 2 complex
  // two-bit (CPSR2 convention, real input) unpacked inside the generic K1: per-window level pairs (fold.cu
  // k_twobit_windows), [window][npol] from the 512-sample boundary at ptr; which codes are low / negative
  const float2* win;
  unsigned lowsel, negsel;
};

// Tries to express a 256-entry 8-bit table as the float evaluation fmaf(x, hi, x*lo); returns 1 and the
// constants iff every entry is reproduced bit for bit.
int lut_as_arithmetic(const float* lut256, float* hi, float* lo);

// what the last kernel does with the dedispersed samples
struct FbSink {
  int kind;                 // Epilogue
  // EPI_VOLT: complex voltages, Filterbank::Engine::perform's output TimeSeries
  float* volt;
  uint64_t volt_span;       // floats between (chan,pol) planes
  uint64_t volt_step;       // floats between parts (nkeep*2)
  // EPI_DETECT / EPI_FOLD
  int state;                // b200_state
  unsigned dndim;           // output ndim (1,2,4); npol' = nprod/dndim
  float* det;               // EPI_DETECT: detected planes
  uint64_t det_span;
  // EPI_FOLD
  const unsigned* bins;     // bin of every output sample of the block (npart*nkeep)
  const uint2* runs;        // run table of the bin plan, [npart][nkeep+1] of (start, bin) (fold.cu k_bin_runs), or null
  const unsigned* nruns;    // [npart]
  double phase_per_sample;  // of the block's fold call (0 = unknown): enables the one-bin-per-chunk path
  unsigned nbin;
  float* profile;           // [chan][npol'][nbin][dndim]
  // reproducible accumulation (b200_fold_set_deterministic): run sums are converted to fixed point (units of lsb)
  // and added to 64-bit integers -- integer addition is associative, so the result no longer depends on the order
  // in which the CTAs' reductions arrive
  long long* fix;           // same layout as profile, or null
  float inv_lsb;
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// detection products (cross_detect.ic:25-41, stokes_detect.ic:21-44, Detection.C:264-301);
// explicit _rn intrinsics: no FMA contraction, bit-identical to the CPU loops.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int detect_products(int state, float2 p, float2 q, float* r) {
  float pp = __fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y));
  float qq = __fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y));
  if (state == B200_INTENSITY) {
    r[0] = __fadd_rn(pp, qq);
    return 1;
  }
  if (state == B200_PPQQ) {
    r[0] = pp;
    r[1] = qq;
    return 2;
  }
  float re = __fadd_rn(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y));
  float im = __fsub_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x));
  if (state == B200_COHERENCE) {
    r[0] = pp; r[1] = qq; r[2] = re; r[3] = im;
  } else {
    r[0] = __fadd_rn(pp, qq); r[1] = __fsub_rn(pp, qq); r[2] = __fmul_rn(2.f, re); r[3] = __fmul_rn(2.f, im);
  }
  return 4;
}

// one run sum into the PhaseSeries: RED.ADD.F32, or RED.ADD.64 of the fixed-point value in deterministic mode
__device__ __forceinline__ void profile_add(float* profile, long long* fix, float inv_lsb, uint64_t idx, float v) {
  if (fix) atomicAdd(reinterpret_cast<unsigned long long*>(fix + idx), (unsigned long long)__float2ll_rn(v * inv_lsb));
  else atomicAdd(profile + idx, v);
}

__host__ __device__ inline unsigned state_nprod(int state, unsigned npol) {
  if (state == B200_INTENSITY) return 1;
  if (state == B200_PPQQ) return npol;
  return 4;
}

#endif

}  // namespace b200

struct b200_fb_plan {
  b200::Context* ctx;
  b200_fb_desc desc;
  // derived sizes (Filterbank.C:68-155,409)
  unsigned C, F, Nc, nsamp_fft, nsamp_overlap, nsamp_step, nkeep, nchan_out;
  // forward factorisation
  unsigned P, Q, lbB, G;
  bool conv_path;           // large single-channel transforms: inverse also two-pass
  unsigned batch;           // parts per internal batch
  b200::TwiddleTable twP, twQ, twF, tw2Q;
  b200::BigTwiddle bigN, big2N;
  float2* d_response;       // nchan_in*Nc complex or null
  float2* scratchA;         // batch*nblk*Nc
  float2* scratchZ;         // batch*nblk*Nc (unused on the conv path)
  uint64_t scratch_bytes;
  // second-generation kernels (fastpath.cu): which passes they cover for this plan + their stage tables
  bool fast_k1, fast_k2, fast_k3;
  float2 *c2P, *c2Q, *c2F, *c2F32, *c2Q32;
  float2* d_response_tiled;   // the response in the tile-image order of Z (fastpath.cu zi_pos), or null
  // one-kernel cluster path of 65536-point convolutions (clusterconv.cu)
  float2* c2cc;               // stage tables of its 4096-point row transforms, or null
  int cc_clusters;            // clusters of 16 CTAs the device keeps resident (0 = path not available)
  void* cc_xch;               // exchange matrices of the global-memory variant of its two transposes, or null
  void* cc_bar;               // arrival counters of the counter-barrier variant, or null
  // longer convolutions (N = P Q > 131072, longconv.cu): three kernels on the c2 core
  bool bc_ok;                 // the plan's P and Q are planned for and the kernels fit the device
  float2 *bc_twP, *bc_twQ;    // c2 stage tables of P and Q
  float2* bc_Ht;              // the response transposed to [channel][P][Q] (rows of the row pass contiguous), or null
};

namespace b200 {
// Runs the engine over npart parts: source -> (K1, K2, K3) -> sink.
int fb_run(b200_fb_plan* plan, const FbSource& src, const FbSink& sink, uint64_t npart);
// fold.cu
int twobit_windows(Context* ctx, const b200_twobit_desc* d, const void* d_raw, uint64_t nwindow, float2* d_win,
                   unsigned* d_weights, unsigned* lowsel, unsigned* negsel);
// fastpath.cu
int fast_plan_init(b200_fb_plan* plan);
void fast_plan_free(b200_fb_plan* plan);
int fast_k1(b200_fb_plan* plan, const FbSource& src, uint64_t part0, unsigned nb);
int fast_k2(b200_fb_plan* plan, unsigned nb);
int fast_k3(b200_fb_plan* plan, const FbSink& sink, uint64_t part0, unsigned nb);
// clusterconv.cu
int cc_plan_init(b200_fb_plan* plan);
void cc_plan_free(b200_fb_plan* plan);
bool cc_applies(const b200_fb_plan* plan, const FbSource& src, const FbSink& sink);
// B200_OK, an error, or CC_NOT_RUN: the launch was refused for lack of residency and the path is now disabled
enum { CC_NOT_RUN = 1000 };
int cc_run(b200_fb_plan* plan, const FbSource& src, const FbSink& sink, uint64_t part0, unsigned nb);
// longconv.cu: which of the three passes the c2 kernels take over (all three -> the spectrum
// scratch holds both polarisations of a bin side by side, otherwise the generic kernels' plane per polarisation)
int bc_plan_init(b200_fb_plan* plan);
void bc_plan_free(b200_fb_plan* plan);
bool bc_k1_applies(const b200_fb_plan* plan, const FbSource& src);
bool bc_k2_applies(const b200_fb_plan* plan);
bool bc_k3_applies(const b200_fb_plan* plan);
int bc_k1(b200_fb_plan* plan, const FbSource& src, uint64_t part0, unsigned nb, bool interleaved);
int bc_k2(b200_fb_plan* plan, unsigned nb, bool interleaved);
int bc_k3(b200_fb_plan* plan, const FbSink& sink, uint64_t part0, unsigned nb, bool interleaved);
}
