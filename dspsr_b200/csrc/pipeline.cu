// pipeline.cu -- fused path: raw bytes -> (K1, K2, K3 with detect+fold epilogue) -> PhaseSeries.
// What dspsr computes with [IOManager(unpack), Filterbank|Convolution, Detection, Fold]
// (Signal/Pulsar/LoadToFold1.C:117-599, SingleThread.C:405-431) when no operation sits between.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "engine.cuh"

extern "C" int b200_fold_weighted(const b200_fold* f);
namespace b200 {
const unsigned* fold_bins(b200_fold* f);
const uint2* fold_runs(b200_fold* f);
const unsigned* fold_nruns(b200_fold* f);
int fold_build_runs(b200_fold* f, unsigned nkeep, unsigned align);
long long* fold_fix(b200_fold* f);
float fold_lsb(b200_fold* f);
int fold_reserve(b200_fold* f, uint64_t ndat, unsigned nkeep);
}
using namespace b200;

struct b200_pipeline {
  Context* ctx;
  b200_pipeline_desc desc;
  b200_fb_plan* fb;
  b200_fold* fold;
  float* d_lut;
  // staging for execute_host: two device buffers used alternately, filled chunk by chunk on a private
  // copy stream so that the host->device transfer of batch i+1 (and of the next call) overlaps the
  // kernels of batch i
  void* d_stage[2];
  uint64_t stage_bytes[2];
  cudaStream_t copy_stream;
  cudaEvent_t stage_free[2];         // fired when the kernels that read d_stage[i] are done
  std::vector<cudaEvent_t>* chunk_ready;
  unsigned stage_turn;
  // unpacked float series for formats whose unpack is not fused into K1
  float* d_unpacked;
  uint64_t unpacked_floats;
  float2* d_win;             // fused two-bit path: (lo, hi) per 512-sample window and digitizer
  uint64_t win_capacity;
  unsigned nprod, dnpol, dndim;
  int conv_ok;
  float conv_hi, conv_lo;
  b200_twobit_desc twobit;
  bool bins_preset;        // execute_host already issued set_bins for the coming block
  // unfused tail (per-channel transforms too long for the fused epilogues): voltages, then detected series
  float* d_volt;
  uint64_t volt_floats;
  float* d_det;
  uint64_t det_floats;
  // WeightedTimeSeries flags of two-bit input: as unpacked / after convolve_weights + scrunch_weights, and the
  // scratch of the convolution (one word per transform + 1)
  unsigned* d_weights;
  unsigned* d_weights2;
  unsigned* d_wscratch;
  uint64_t weights_capacity, wscratch_capacity;
  // streaming input with block-edge carry (b200_pipeline_stream_begin / feed): two device buffers [carry | block]
  // used alternately; see the comment at b200_pipeline_feed_host
  unsigned char* d_stream[2];
  uint64_t stream_bytes;           // capacity of each
  cudaEvent_t stream_free[2];      // kernels that read d_stream[i] are done
  cudaEvent_t stream_loaded;       // the H2D of the current block has landed
  unsigned stream_turn;
  uint64_t stream_block;           // largest block (samples) a feed may bring
  uint64_t carry_samples;          // samples at the start of the current buffer that were kept from earlier feeds
  uint64_t carry_skip;             // of those, samples before the next part's first sample
  uint64_t stream_pos;             // observation sample index of the first sample of the current buffer
  bool streaming;
  // observation-driven folding (b200_pipeline_execute_obs): attributes of the raw input, of the series that
  // reaches Fold, the predictor, and the PhaseSeries attributes that Fold::transformation / Fold::fold maintain
  bool have_obs, have_poly;
  b200_observation raw_obs, fold_obs;
  b200_polyco poly;
  double folding_period, reference_phase;
  b200_mjd reference_epoch;
  b200_phase_series ps;
};

namespace b200 {
int lut_as_arithmetic(const float* lut, float* hi_out, float* lo_out) {
  // every entry constrains the double constant c to an interval: RN_float(x*c) == lut[b]
  double lo_c = -1e300, hi_c = 1e300;
  for (int b = 0; b < 256; b++) {
    const double x = double(int(int8_t(uint8_t(b)))) + 0.5;
    const float v = lut[b];
    const float up = std::nextafterf(v, INFINITY), dn = std::nextafterf(v, -INFINITY);
    double a = (double(v) + double(dn)) * 0.5 / x, c = (double(v) + double(up)) * 0.5 / x;
    if (a > c) std::swap(a, c);
    lo_c = std::max(lo_c, a);
    hi_c = std::min(hi_c, c);
  }
  if (!(lo_c < hi_c)) return 0;
  const double c = 0.5 * (lo_c + hi_c);
  const float hi = float(c), lo = float(c - double(hi));
  for (int b = 0; b < 256; b++) {
    const float x = float(int(int8_t(uint8_t(b)))) + 0.5f;
    volatile float t = x * lo;                 // rounded to float, as FMUL does
    const float r = std::fmaf(x, hi, t);
    if (std::memcmp(&r, &lut[b], sizeof(float)) != 0) return 0;
  }
  *hi_out = hi;
  *lo_out = lo;
  return 1;
}
}  // namespace b200

static unsigned fmt_resolution(int fmt) {
  switch (fmt) {
    case B200_FMT_CASPSR8: return 4;
    case B200_FMT_MEERKAT8: return 256;
    case B200_FMT_UWB16: return 2048;
    case B200_FMT_TWOBIT: return 512;      // ExcisionUnpacker::get_resolution = ndat_per_weight
    default: return 1;
  }
}
static unsigned fmt_nbit(int fmt) {
  return fmt == B200_FMT_UWB16 ? 16 : fmt == B200_FMT_FLOAT32 ? 32 : fmt == B200_FMT_TWOBIT ? 2 : 8;
}

extern "C" {

int b200_pipeline_create(b200_context* cctx, const b200_pipeline_desc* d, b200_pipeline** out) {
  B200_REQUIRE(cctx && d && out, "b200_pipeline_create: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  const int fmt = d->unpack.format;
  B200_REQUIRE(fmt >= B200_FMT_CASPSR8 && fmt <= B200_FMT_TWOBIT, "unknown input format %d", fmt);
  B200_REQUIRE(fmt != B200_FMT_TWOBIT || d->unpack.twobit, "TWOBIT input needs unpack.twobit");
  B200_REQUIRE(d->unpack.nchan == d->fb.input_nchan && d->unpack.npol == d->fb.npol,
               "unpacker nchan/npol (%u,%u) != filterbank input (%u,%u)", d->unpack.nchan, d->unpack.npol,
               d->fb.input_nchan, d->fb.npol);
  B200_REQUIRE(d->unpack.ndim == (d->fb.input_real ? 1u : 2u), "unpacker ndim=%u inconsistent with input state",
               d->unpack.ndim);
  B200_REQUIRE(d->detect_state >= 0 && d->detect_state <= 3, "invalid detection state");
  b200_pipeline* p = new b200_pipeline();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx;
  p->desc = *d;
  p->desc.fb.h_response = nullptr;
  if (fmt == B200_FMT_TWOBIT) {
    p->twobit = *d->unpack.twobit;           // the caller's descriptor need not outlive the call
    p->desc.unpack.twobit = &p->twobit;
  }
  if (d->detect_state >= B200_COHERENCE) {
    p->nprod = 4;
    p->dndim = d->detect_ndim;
    if (!(p->dndim == 1 || p->dndim == 2 || p->dndim == 4)) {
      delete p;
      set_error("invalid detection ndim %u", d->detect_ndim);
      return B200_ERR_INVALID;
    }
  } else {
    p->nprod = d->detect_state == B200_PPQQ ? d->fb.npol : 1;
    p->dndim = 1;
  }
  p->dnpol = p->nprod / p->dndim;
  int rc = b200_fb_plan_create(cctx, &d->fb, &p->fb);
  if (rc != B200_OK) { delete p; return rc; }
  if (d->nbin) {
    rc = b200_fold_create(cctx, p->fb->nchan_out, p->dnpol, p->dndim, d->nbin, &p->fold);
    if (rc != B200_OK) { b200_pipeline_destroy(p); return rc; }
  }
  if (fmt == B200_FMT_CASPSR8 || fmt == B200_FMT_GENERIC8) {
    static const bool arith = tune_flag("B200_LUT_ARITH", true);
    p->conv_ok = arith ? lut_as_arithmetic(d->unpack.lut, &p->conv_hi, &p->conv_lo) : 0;
    cudaError_t e = cudaMalloc(&p->d_lut, 256 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_lut, d->unpack.lut, 256 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { b200_pipeline_destroy(p); return cuda_fail(e, "lut upload", __FILE__, __LINE__); }
  }
  *out = p;
  return B200_OK;
}

int b200_pipeline_destroy(b200_pipeline* p) {
  if (!p) return B200_OK;
  if (p->fb) b200_fb_plan_destroy(p->fb);
  if (p->fold) b200_fold_destroy(p->fold);
  if (p->d_lut) cudaFree(p->d_lut);
  for (int i = 0; i < 2; i++) {
    if (p->d_stage[i]) cudaFree(p->d_stage[i]);
    if (p->stage_free[i]) cudaEventDestroy(p->stage_free[i]);
  }
  if (p->chunk_ready) {
    for (cudaEvent_t e : *p->chunk_ready) cudaEventDestroy(e);
    delete p->chunk_ready;
  }
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->d_unpacked) cudaFree(p->d_unpacked);
  if (p->d_win) cudaFree(p->d_win);
  if (p->d_volt) cudaFree(p->d_volt);
  if (p->d_det) cudaFree(p->d_det);
  for (int i = 0; i < 2; i++) {
    if (p->d_stream[i]) cudaFree(p->d_stream[i]);
    if (p->stream_free[i]) cudaEventDestroy(p->stream_free[i]);
  }
  if (p->stream_loaded) cudaEventDestroy(p->stream_loaded);
  if (p->d_weights) cudaFree(p->d_weights);
  if (p->d_weights2) cudaFree(p->d_weights2);
  if (p->d_wscratch) cudaFree(p->d_wscratch);
  delete p;
  return B200_OK;
}

int b200_pipeline_info(const b200_pipeline* p, b200_fb_info* info) {
  B200_REQUIRE(p, "null pipeline");
  return b200_fb_plan_info(p->fb, info);
}

b200_fold* b200_pipeline_fold(b200_pipeline* p) { return p ? p->fold : nullptr; }

static int pipeline_execute(b200_pipeline* p, const void* d_input, uint64_t input_span, uint64_t first_sample,
                            uint64_t npart, double phi, double pps, float* d_detected, uint64_t detected_span,
                            cudaEvent_t* batch_ready, unsigned batch_override = 0);

}  // extern "C"

// (re)allocation of a scratch array that must hold `need` elements; synchronises only when it has to grow
template <typename T>
static int grow(Context* ctx, T** ptr, uint64_t* capacity, uint64_t need) {
  if (need <= *capacity) return B200_OK;
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *capacity = 0;
  B200_CUDA(cudaMalloc(ptr, need * sizeof(T)));
  *capacity = need;
  return B200_OK;
}

// Two-bit input is unpacked inside the generic K1 (no float time series) whenever that kernel runs the column pass:
// real input (the CPSR2 convention), and no second-generation K1 for the plan's shape
static bool twobit_fused(const b200_pipeline* p) {
  static const bool want = b200::tune_flag("B200_TWOBIT_FUSED", true);
  return want && p->desc.unpack.format == B200_FMT_TWOBIT && p->desc.unpack.ndim == 1 && p->desc.unpack.nchan == 1 &&
         p->desc.fb.input_real && !p->fb->fast_k1 && p->twobit.ndat_per_weight == 512;
}

// scratch of a block of npart parts for everything execute needs besides the plan's own buffers
static int pipeline_scratch(b200_pipeline* p, uint64_t npart) {
  Context* ctx = p->ctx;
  b200_fb_plan* fb = p->fb;
  const int fmt = p->desc.unpack.format;
  const unsigned ndim = p->desc.unpack.ndim;
  const uint64_t ndat_out = npart * fb->nkeep;
  int rc = B200_OK;
  const bool fused_twobit = twobit_fused(p);
  const bool fused_unpack = fmt == B200_FMT_CASPSR8 ||
                            (!fb->fast_k1 && (fmt == B200_FMT_MEERKAT8 || fmt == B200_FMT_UWB16 || fmt == B200_FMT_GENERIC8));
  if (!fused_unpack && fmt != B200_FMT_FLOAT32) {
    const unsigned res = fmt_resolution(fmt);
    const uint64_t ndat_in = npart * fb->nsamp_step + fb->nsamp_overlap + 2 * res;
    if (fused_twobit)
      rc = grow(ctx, &p->d_win, &p->win_capacity, (ndat_in / p->twobit.ndat_per_weight + 2) * p->desc.unpack.npol);
    else
      rc = grow(ctx, &p->d_unpacked, &p->unpacked_floats, ndat_in * ndim * p->desc.unpack.nchan * p->desc.unpack.npol);
    if (rc == B200_OK && fmt == B200_FMT_TWOBIT) {
      const uint64_t nw = ndat_in / p->twobit.ndat_per_weight + 2;
      uint64_t cap2 = p->weights_capacity;
      rc = grow(ctx, &p->d_weights, &p->weights_capacity, nw);
      if (rc == B200_OK) rc = grow(ctx, &p->d_weights2, &cap2, nw);
      if (rc == B200_OK) rc = grow(ctx, &p->d_wscratch, &p->wscratch_capacity, npart + 2);
    }
  }
  if (rc == B200_OK && fb->F > 8192 && !fb->conv_path) {
    rc = grow(ctx, &p->d_volt, &p->volt_floats, uint64_t(fb->nchan_out) * fb->desc.npol * ndat_out * 2);
    if (rc == B200_OK && p->desc.nbin)
      rc = grow(ctx, &p->d_det, &p->det_floats, uint64_t(fb->nchan_out) * p->dnpol * ndat_out * p->dndim);
  }
  return rc;
}

extern "C" {

int b200_pipeline_reserve(b200_pipeline* p, uint64_t max_npart) {
  B200_REQUIRE(p && max_npart, "b200_pipeline_reserve: null pipeline or zero parts");
  int rc = pipeline_scratch(p, max_npart);
  if (rc == B200_OK && p->fold) {
    // the bin plan (and the item table of the fused fold) of a block of max_npart parts
    rc = b200_fold_set_bins(p->fold, 0.0, 0.0, 0, 0, nullptr);
    if (rc == B200_OK) rc = fold_reserve(p->fold, max_npart * p->fb->nkeep, p->fb->fast_k3 ? p->fb->nkeep : 0);
  }
  return rc;
}

int b200_pipeline_execute(b200_pipeline* p, const void* d_input, uint64_t input_span, uint64_t first_sample,
                          uint64_t npart, double phi, double pps, float* d_detected, uint64_t detected_span) {
  return pipeline_execute(p, d_input, input_span, first_sample, npart, phi, pps, d_detected, detected_span, nullptr);
}

static int pipeline_execute(b200_pipeline* p, const void* d_input, uint64_t input_span, uint64_t first_sample,
                            uint64_t npart, double phi, double pps, float* d_detected, uint64_t detected_span,
                            cudaEvent_t* batch_ready, unsigned batch_override) {
  B200_REQUIRE(p && d_input, "b200_pipeline_execute: null argument");
  if (npart == 0) return B200_OK;
  Context* ctx = p->ctx;
  b200_fb_plan* fb = p->fb;
  const int fmt = p->desc.unpack.format;
  const unsigned ndim = p->desc.unpack.ndim;
  const uint64_t ndat_out = npart * fb->nkeep;

  struct { const unsigned* d; uint64_t nweights, weight_idat; unsigned ndat_per_weight; } wt = {nullptr, 0, 0, 0};
  FbSource src;
  memset(&src, 0, sizeof(src));
  src.batch_ready = batch_ready;
  src.batch_override = batch_override;
  if (fmt == B200_FMT_CASPSR8) {
    B200_REQUIRE(first_sample % 2 == 0, "CASPSR input must start on an even sample");
    src.kind = SRC_CASPSR8;
    // fold the first-sample offset into whole 8-byte groups + a residual handled by the loader
    src.ptr = static_cast<const unsigned char*>(d_input) + 8 * (first_sample / 4);
    B200_REQUIRE(first_sample % 4 == 0, "CASPSR block must start on a 4-sample boundary");
    src.step = fb->nsamp_step;
    src.d_lut = p->d_lut;
    src.conv_ok = p->conv_ok; src.conv_hi = p->conv_hi; src.conv_lo = p->conv_lo;
  } else if (fmt == B200_FMT_FLOAT32) {
    src.kind = SRC_F32;
    src.ptr = static_cast<const float*>(d_input) + first_sample * ndim;
    src.span = input_span;
    src.step = uint64_t(fb->nsamp_step) * ndim;
    B200_REQUIRE(input_span % 2 == 0 && (first_sample * ndim) % 2 == 0, "float input planes must be 8-byte aligned");
  } else if (!fb->fast_k1 && (fmt == B200_FMT_MEERKAT8 || fmt == B200_FMT_UWB16 || fmt == B200_FMT_GENERIC8)) {
    // unpacked inside K1 (filterbank.cu k_cols_fwd): no float time series is ever written.  d_input starts on a
    // boundary of the format's resolution; the kernel seeks to first_sample itself
    src.kind = fmt == B200_FMT_MEERKAT8 ? SRC_MEERKAT8 : fmt == B200_FMT_UWB16 ? SRC_UWB16 : SRC_GENERIC8;
    src.ptr = d_input;
    src.first = first_sample;
    src.step = fb->nsamp_step;
    src.d_lut = p->d_lut;
    src.scale = p->desc.unpack.scale;
    src.sample_swap = p->desc.unpack.sample_swap ? p->desc.unpack.sample_swap : 1;
    src.ndim = ndim;
    if (fmt == B200_FMT_GENERIC8) { src.conv_ok = p->conv_ok; src.conv_hi = p->conv_hi; src.conv_lo = p->conv_lo; }
    if (fmt == B200_FMT_GENERIC8 && ndim == 1)
      B200_REQUIRE(first_sample % 2 == 0, "real 8-bit input must start on an even sample");
  } else {
    // unpack the enclosing resolution-aligned range, then seek (Unpacker.C:82-111)
    const unsigned res = fmt_resolution(fmt);
    const uint64_t ndat_in = npart * fb->nsamp_step + fb->nsamp_overlap;
    const uint64_t a0 = (first_sample / res) * res;
    const uint64_t a1 = ((first_sample + ndat_in + res - 1) / res) * res;
    const uint64_t span = (a1 - a0) * ndim;
    int rc = pipeline_scratch(p, npart);              // no-op when b200_pipeline_reserve covered this block size
    if (rc != B200_OK) return rc;
    const uint64_t bits_per_sample = uint64_t(p->desc.unpack.nchan) * p->desc.unpack.npol * ndim * fmt_nbit(fmt);
    const unsigned char* raw0 = static_cast<const unsigned char*>(d_input) + a0 * bits_per_sample / 8;
    const bool fused_twobit = twobit_fused(p);
    if (fmt == B200_FMT_TWOBIT) {
      // a WeightedTimeSeries: the flags of the windows travel with the data (weights.cu)
      if (fused_twobit)
        rc = twobit_windows(ctx, &p->twobit, raw0, (a1 - a0) / 512, p->d_win, p->d_weights, &src.lowsel, &src.negsel);
      else
        rc = b200_unpack_twobit(reinterpret_cast<b200_context*>(ctx), &p->twobit, raw0, a1 - a0, p->d_unpacked, span, p->d_weights);
      if (rc != B200_OK) return rc;
      wt.nweights = (a1 - a0) / p->twobit.ndat_per_weight;
      wt.ndat_per_weight = p->twobit.ndat_per_weight;
      wt.weight_idat = first_sample - a0;               // TimeSeries::seek moves weight_idat
      if (p->desc.nbin) {
        rc = b200_weights_convolve(reinterpret_cast<b200_context*>(ctx), p->d_weights, wt.nweights, wt.ndat_per_weight,
                                   wt.weight_idat, ndat_in, fb->nsamp_fft, fb->nsamp_step, p->d_weights2, p->d_wscratch);
        if (rc != B200_OK) return rc;
        // Filterbank.C:289,306: scrunch by nsamp_fft / freq_res; Convolution.C:317-318: by 2 for Nyquist input
        const unsigned tres = (fb->conv_path || fb->C == 1) ? (p->desc.fb.input_real ? 2u : 1u) : fb->nsamp_fft / fb->F;
        wt.d = p->d_weights2;
        if (tres > 1) {
          if (double(wt.ndat_per_weight) / double(tres) < 1.0) {
            // the reference scrunches `tres` FLAGS (not samples) into one and Fold then runs off the end of the array
            set_error("two-bit weights: time resolution ratio %u exceeds ndat_per_weight %u -- the reference's Fold throws "
                      "here (iweight >= nweights, Fold.C:699)", tres, wt.ndat_per_weight);
            return B200_ERR_INVALID;
          }
          rc = b200_weights_scrunch(reinterpret_cast<b200_context*>(ctx), wt.d, &wt.nweights, &wt.ndat_per_weight,
                                    &wt.weight_idat, tres, nullptr);
          if (rc != B200_OK) return rc;
        }
      }
    } else {
      rc = b200_unpack(reinterpret_cast<b200_context*>(ctx), &p->desc.unpack, raw0, a1 - a0, p->d_unpacked, span);
      if (rc != B200_OK) return rc;
    }
    if (fused_twobit) {
      // K1 converts the codes itself from the window levels: ptr is the 512-sample boundary a0, first the offset into it
      src.kind = SRC_TWOBIT;
      src.ptr = raw0;
      src.first = first_sample - a0;
      src.step = fb->nsamp_step;
      src.win = p->d_win;
      src.ndim = 1;
      B200_REQUIRE((first_sample - a0) % 2 == 0, "two-bit input must start on an even sample");
    } else {
      src.kind = SRC_F32;
      src.ptr = p->d_unpacked + (first_sample - a0) * ndim;
      src.span = span;
      src.step = uint64_t(fb->nsamp_step) * ndim;
      B200_REQUIRE(((first_sample - a0) * ndim) % 2 == 0, "unaligned block start");
    }
  }

  FbSink sink;
  memset(&sink, 0, sizeof(sink));
  sink.state = p->desc.detect_state;
  sink.dndim = p->dndim;

  // Per-channel inverse transforms above 8192 points do not fit one SM with both polarisations, so the
  // detection / fold epilogues cannot be fused: run the engine for voltages, then the stand-alone
  // detection and fold engines (what dspsr itself does: three separate operations).
  if (fb->F > 8192 && !fb->conv_path) {
    const unsigned npol = fb->desc.npol;
    int rcs = pipeline_scratch(p, npart);
    if (rcs != B200_OK) return rcs;
    sink.kind = EPI_VOLT;
    sink.volt = p->d_volt;
    sink.volt_span = ndat_out * 2;
    sink.volt_step = uint64_t(fb->nkeep) * 2;
    int rc = fb_run(fb, src, sink, npart);
    if (rc != B200_OK) return rc;
    float* det = d_detected;
    uint64_t det_span = detected_span;
    if (p->desc.nbin) {
      det = p->d_det;
      det_span = ndat_out * p->dndim;
    } else {
      B200_REQUIRE(d_detected, "b200_pipeline_execute: nbin == 0 needs an output buffer for the detected series");
    }
    rc = b200_detect(reinterpret_cast<b200_context*>(ctx), p->desc.detect_state, p->dndim, p->d_volt, ndat_out * 2,
                     fb->nchan_out, npol, ndat_out, det, det_span);
    if (rc != B200_OK || !p->desc.nbin) return rc;
    rc = b200_fold_set_bins_weighted(p->fold, phi, pps, ndat_out, 0, wt.d, wt.nweights, wt.ndat_per_weight, wt.weight_idat);
    if (rc != B200_OK) return rc;
    return b200_fold_fold(p->fold, det, det_span);
  }

  if (p->desc.nbin) {
    if (!p->bins_preset) {
      int rc = b200_fold_set_bins_weighted(p->fold, phi, pps, ndat_out, 0, wt.d, wt.nweights, wt.ndat_per_weight, wt.weight_idat);
      if (rc == B200_OK && fb->fast_k3) rc = fold_build_runs(p->fold, fb->nkeep, fb->desc.nfilt_pos);
      if (rc != B200_OK) return rc;
    }
    p->bins_preset = false;
    sink.kind = EPI_FOLD;
    sink.bins = fold_bins(p->fold);
    if (fb->fast_k3) {                     // item table of the fused fold epilogue (fastpath.cu); the generic K3 walks the bins
      sink.runs = fold_runs(p->fold);
      sink.nruns = fold_nruns(p->fold);
    }
    sink.nbin = p->desc.nbin;
    // one-bin-per-chunk shortcut of the fold epilogue: measured SLOWER than the per-sample walk on
    // B200 (0.83 vs 0.76 ms per 32 parts of cfg1), so it is opt-in for experiments only
    static const bool fold_fast = tune_flag("B200_FOLD_FAST", false);
    sink.phase_per_sample = (fold_fast && !wt.d) ? pps : 0.0;      // flagged samples break the one-bin-per-chunk shortcut
    sink.fix = fold_fix(p->fold);
    sink.inv_lsb = sink.fix ? 1.0f / fold_lsb(p->fold) : 0.f;
    sink.profile = sink.fix ? nullptr : b200_fold_device_profile(p->fold);
  } else {
    B200_REQUIRE(d_detected, "b200_pipeline_execute: nbin == 0 needs an output buffer for the detected series");
    sink.kind = EPI_DETECT;
    sink.det = d_detected;
    sink.det_span = detected_span;
  }
  return fb_run(fb, src, sink, npart);
}

int b200_pipeline_input_consumed(b200_pipeline* p) {
  B200_REQUIRE(p, "b200_pipeline_input_consumed: null pipeline");
  if (p->copy_stream) B200_CUDA(cudaStreamSynchronize(p->copy_stream));
  return B200_OK;
}

int b200_pipeline_execute_host(b200_pipeline* p, const void* h_input, uint64_t nbytes, uint64_t first_sample,
                               uint64_t npart, double phi, double pps, float* d_detected, uint64_t detected_span) {
  B200_REQUIRE(p && h_input, "b200_pipeline_execute_host: null argument");
  B200_REQUIRE(p->desc.unpack.format != B200_FMT_FLOAT32, "execute_host takes raw bytes");
  if (npart == 0) return B200_OK;
  Context* ctx = p->ctx;
  b200_fb_plan* fb = p->fb;
  // everything pipeline_execute would reject is rejected here, BEFORE the bin plan of the block is committed
  B200_REQUIRE(p->desc.nbin || d_detected, "b200_pipeline_execute_host: nbin == 0 needs an output buffer for the detected series");
  if (p->desc.unpack.format == B200_FMT_CASPSR8)
    B200_REQUIRE(first_sample % 4 == 0, "CASPSR block must start on a 4-sample boundary");
  {
    const uint64_t bits = uint64_t(p->desc.unpack.nchan) * p->desc.unpack.npol * p->desc.unpack.ndim * fmt_nbit(p->desc.unpack.format);
    const uint64_t last = first_sample + npart * fb->nsamp_step + fb->nsamp_overlap;
    B200_REQUIRE(nbytes >= (last * bits + 7) / 8, "b200_pipeline_execute_host: %llu bytes do not hold %llu samples",
                 (unsigned long long)nbytes, (unsigned long long)last);
  }
  if (!p->copy_stream) {
    B200_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) B200_CUDA(cudaEventCreateWithFlags(&p->stage_free[i], cudaEventDisableTiming));
    p->chunk_ready = new std::vector<cudaEvent_t>();
  }
  const unsigned turn = p->stage_turn++ & 1u;
  if (nbytes > p->stage_bytes[turn]) {
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    B200_CUDA(cudaStreamSynchronize(p->copy_stream));
    if (p->d_stage[turn]) cudaFree(p->d_stage[turn]);
    p->d_stage[turn] = nullptr;
    p->stage_bytes[turn] = nbytes;
    B200_CUDA(cudaMalloc(&p->d_stage[turn], nbytes));
  }
  // the copy stream may overwrite this buffer once the kernels of its previous use are done, and must
  // not run ahead of work already queued on the pipeline's stream that produced h_input (none: host memory)
  B200_CUDA(cudaStreamWaitEvent(p->copy_stream, p->stage_free[turn], 0));

  // one chunk per internal batch of the fused formats; everything at once when a separate unpack
  // pass reads the whole block first
  const int fmt = p->desc.unpack.format;
  // formats K1 unpacks itself read their parts straight from the staged bytes: transfer and compute pipeline per batch
  const bool chunked = fmt == B200_FMT_CASPSR8 ||
                       (!fb->fast_k1 && (fmt == B200_FMT_MEERKAT8 || fmt == B200_FMT_UWB16 || fmt == B200_FMT_GENERIC8));
  // The bin plan uploads a few KiB of phase segments on the pipeline's stream.  Host-to-device copies of all
  // streams share one copy engine and run in submission order, so that small copy must be SUBMITTED BEFORE
  // the bulk chunks -- queued behind them it would hold the first kernels back until the whole block had
  // arrived, serialising transfer and compute.
  if (p->desc.nbin && chunked && !(fb->F > 8192 && !fb->conv_path)) {
    int rc0 = b200_fold_set_bins(p->fold, phi, pps, npart * fb->nkeep, 0, nullptr);
    if (rc0 == B200_OK && fb->fast_k3) rc0 = fold_build_runs(p->fold, fb->nkeep, fb->desc.nfilt_pos);
    if (rc0 != B200_OK) return rc0;
    p->bins_preset = true;
  }
  // host-fed blocks are PCIe-bound: chunks (= kernel batches) of about 64 MB keep the pipeline fine-grained (cfg1: 8
  // parts, first kernels start after 1/4 of a 32-part block instead of 1/2) at a small cost in kernel efficiency
  const uint64_t bytes_per_part = std::max<uint64_t>(1, uint64_t(fb->nsamp_step) * p->desc.unpack.nchan * p->desc.unpack.npol *
                                                            p->desc.unpack.ndim * fmt_nbit(p->desc.unpack.format) / 8);
  const uint64_t host_batch = std::max<uint64_t>(1, (64ull << 20) / bytes_per_part);
  const uint64_t batch = chunked ? std::max<uint64_t>(1, std::min<uint64_t>(fb->batch, host_batch)) : fb->batch;
  const uint64_t nchunk = chunked ? (npart + batch - 1) / batch : 1;
  while (p->chunk_ready->size() < nchunk) {
    cudaEvent_t e;
    B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->chunk_ready->push_back(e);
  }
  const uint64_t bits_per_sample =
      uint64_t(p->desc.unpack.nchan) * p->desc.unpack.npol * p->desc.unpack.ndim * fmt_nbit(fmt);
  uint64_t done = 0;
  for (uint64_t c = 0; c < nchunk; c++) {
    uint64_t end = nbytes;
    if (c + 1 < nchunk) {
      // whole resolution units (a MeerKAT heap / UWB block interleaves its polarisations and channels)
      const unsigned res = fmt_resolution(fmt);
      const uint64_t last_sample = (first_sample + (c + 1) * batch * fb->nsamp_step + fb->nsamp_overlap + res - 1) / res * res;
      end = std::min<uint64_t>(nbytes, (last_sample * bits_per_sample / 8 + 4095) / 4096 * 4096);
    }
    if (end > done)
      B200_CUDA(cudaMemcpyAsync(static_cast<char*>(p->d_stage[turn]) + done, static_cast<const char*>(h_input) + done,
                                end - done, cudaMemcpyHostToDevice, p->copy_stream));
    done = std::max(done, end);
    B200_CUDA(cudaEventRecord((*p->chunk_ready)[c], p->copy_stream));
  }
  int rc;
  if (chunked) {
    rc = pipeline_execute(p, p->d_stage[turn], 0, first_sample, npart, phi, pps, d_detected, detected_span,
                          p->chunk_ready->data(), (unsigned)batch);
  } else {
    B200_CUDA(cudaStreamWaitEvent(ctx->stream, (*p->chunk_ready)[0], 0));
    rc = pipeline_execute(p, p->d_stage[turn], 0, first_sample, npart, phi, pps, d_detected, detected_span, nullptr);
  }
  if (rc != B200_OK) p->bins_preset = false;
  B200_CUDA(cudaEventRecord(p->stage_free[turn], ctx->stream));
  return rc;
}

// ---- observation-driven blocks ---------------------------------------------------------------------
int b200_pipeline_set_observation(b200_pipeline* p, const b200_observation* raw) {
  B200_REQUIRE(p && raw, "b200_pipeline_set_observation: null argument");
  B200_REQUIRE(raw->rate > 0, "b200_pipeline_set_observation: rate must be positive");
  B200_REQUIRE(raw->nchan == p->desc.fb.input_nchan && raw->npol == p->desc.fb.npol,
               "observation nchan/npol (%u,%u) != pipeline input (%u,%u)", raw->nchan, raw->npol,
               p->desc.fb.input_nchan, p->desc.fb.npol);
  b200_fb_plan* fb = p->fb;
  p->raw_obs = *raw;
  b200_observation o = *raw;
  o.nchan = fb->nchan_out;
  if (fb->conv_path || fb->C == 1) {
    // Convolution::prepare_output (Convolution.C:286-305): same rate (halved for Nyquist input, which the engine
    // turns into an analytic signal), scale *= nsamp_fft * n_fft
    o.rate = raw->rate * double(fb->F) / double(fb->nsamp_fft);
    o.scale = raw->scale * (double(fb->nsamp_fft) * double(fb->Nc));
  } else {
    // Filterbank::prepare_output (Filterbank.C:325-371)
    o.rate = raw->rate * (double(fb->F) / double(fb->nsamp_fft));
    o.scale = raw->scale * (double(fb->Nc) * double(fb->F));
    o.dc_centred = int(fb->F % 2);
    const bool dual = !p->desc.fb.input_real;
    if (dual) {
      if (raw->nchan > 1) o.nsub_swap = int(raw->nchan);
      else o.swap = 1;
    }
  }
  // Detection::transformation: state and shape of the detected series
  o.state = p->desc.detect_state;
  o.npol = p->dnpol;
  o.ndim = p->dndim;
  o.nbit = 32;
  p->fold_obs = o;
  p->have_obs = true;
  return B200_OK;
}

int b200_pipeline_set_predictor(b200_pipeline* p, const b200_polyco* pc, double reference_phase) {
  B200_REQUIRE(p && pc, "b200_pipeline_set_predictor: null argument");
  p->poly = *pc;
  p->have_poly = true;
  p->folding_period = 0.0;
  p->reference_phase = reference_phase;
  return B200_OK;
}

int b200_pipeline_set_folding_period(b200_pipeline* p, double period, double reference_phase, const b200_mjd* epoch) {
  B200_REQUIRE(p && period > 0, "b200_pipeline_set_folding_period: period must be positive");
  p->folding_period = period;
  p->reference_epoch = epoch ? *epoch : b200_mjd{0, 0, 0.0};
  p->have_poly = false;
  p->reference_phase = reference_phase;
  return B200_OK;
}

// attributes of the block that reaches Fold + its phase: Fold.C:650-657,718-720,943-958
static int obs_block(b200_pipeline* p, uint64_t npart, uint64_t obs_sample, b200_observation* blk, double* phi, double* pps) {
  B200_REQUIRE(p->have_obs, "b200_pipeline_execute_obs: call b200_pipeline_set_observation first");
  B200_REQUIRE(p->desc.nbin, "b200_pipeline_execute_obs: the pipeline has no fold stage");
  B200_REQUIRE(p->have_poly || p->folding_period > 0, "no polynomial and no period specified (Fold.C:638-640)");
  b200_fb_plan* fb = p->fb;
  *blk = p->fold_obs;
  // the block's first input sample, then change_start_time(nfilt_pos) in output samples (Filterbank.C:370)
  b200_mjd t = b200_mjd_add(&p->raw_obs.start_time, double(obs_sample) / p->raw_obs.rate);
  t = b200_mjd_add(&t, double(fb->desc.nfilt_pos) / blk->rate);
  blk->start_time = t;
  blk->ndat = npart * fb->nkeep;
  const b200_mjd mid = b200_mjd_add(&t, 0.5 / blk->rate);                    // midpoint of the first sample
  double pfold;
  if (p->folding_period > 0.0) {
    const double since = b200_mjd_diff(&mid, &p->reference_epoch);        // (start_time - reference_epoch).in_seconds()
    *phi = std::fmod(since, p->folding_period) / p->folding_period - p->reference_phase;
    pfold = p->folding_period;
  } else {
    *phi = b200_polyco_phase(&p->poly, mid.day, mid.sec, mid.frac, nullptr) - p->reference_phase;
    pfold = 1.0 / b200_polyco_frequency(&p->poly, mid.day, mid.sec, mid.frac);
  }
  *pps = (1.0 / blk->rate) / pfold;
  return B200_OK;
}

// Fold::transformation: get_output()->mixable(*input, nbin, idat_start, ndat_fold), tried on a copy of the attributes
// BEFORE any kernel runs, so that a refused block leaves the accumulator untouched
static int obs_mixable(b200_pipeline* p, const b200_observation* blk, b200_phase_series* trial) {
  *trial = p->ps;
  trial->data = nullptr;
  trial->hits = nullptr;
  if (trial->integration_length != 0.0 && trial->reference_phase != p->reference_phase) {
    set_error("reference phase changed within an integration (Fold.C:538-543)");
    return B200_ERR_INVALID;
  }
  if (!b200_phase_series_mixable(trial, blk, p->desc.nbin, 0, (int64_t)blk->ndat)) {
    set_error("PhaseSeries !mixable: the block differs from what has been folded so far");
    return B200_ERR_INVALID;
  }
  return B200_OK;
}

static int obs_commit(b200_pipeline* p, const b200_observation* blk, const b200_phase_series* trial) {
  p->ps = *trial;
  p->ps.folding_period = p->folding_period;
  p->ps.reference_phase = p->reference_phase;
  // Fold::fold: integration_length += ndat_folded / rate, ndat_total += ndat_fold (Fold.C:789-802)
  return b200_phase_series_folded(&p->ps, blk->ndat, blk->ndat);
}

int b200_pipeline_execute_obs(b200_pipeline* p, const void* d_input, uint64_t input_span, uint64_t first_sample,
                              uint64_t npart, uint64_t obs_sample) {
  B200_REQUIRE(p && d_input, "b200_pipeline_execute_obs: null argument");
  if (npart == 0) return B200_OK;
  b200_observation blk;
  double phi, pps;
  b200_phase_series trial;
  int rc = obs_block(p, npart, obs_sample, &blk, &phi, &pps);
  if (rc == B200_OK) rc = obs_mixable(p, &blk, &trial);
  if (rc == B200_OK) rc = pipeline_execute(p, d_input, input_span, first_sample, npart, phi, pps, nullptr, 0, nullptr);
  if (rc == B200_OK) rc = obs_commit(p, &blk, &trial);
  return rc;
}

int b200_pipeline_execute_host_obs(b200_pipeline* p, const void* h_input, uint64_t nbytes, uint64_t first_sample,
                                   uint64_t npart, uint64_t obs_sample) {
  B200_REQUIRE(p && h_input, "b200_pipeline_execute_host_obs: null argument");
  if (npart == 0) return B200_OK;
  b200_observation blk;
  double phi, pps;
  b200_phase_series trial;
  int rc = obs_block(p, npart, obs_sample, &blk, &phi, &pps);
  if (rc == B200_OK) rc = obs_mixable(p, &blk, &trial);
  if (rc == B200_OK) rc = b200_pipeline_execute_host(p, h_input, nbytes, first_sample, npart, phi, pps, nullptr, 0);
  if (rc == B200_OK) rc = obs_commit(p, &blk, &trial);
  return rc;
}

int b200_pipeline_get_phase_series(b200_pipeline* p, b200_phase_series* out) {
  B200_REQUIRE(p && out && p->fold, "b200_pipeline_get_phase_series: null argument or no fold stage");
  float* data = out->data;
  unsigned* hits = out->hits;
  *out = p->ps;
  out->data = data;
  out->hits = hits;
  if (!p->have_obs || p->ps.integration_length == 0.0) {
    // nothing folded through execute_obs: shape only
    out->obs.nchan = p->fb->nchan_out; out->obs.npol = p->dnpol; out->obs.ndim = p->dndim;
    out->nbin = p->desc.nbin; out->hits_nchan = 1;
  }
  int rc = B200_OK;
  if (data) rc = b200_fold_synch(p->fold, data);
  uint64_t ntot = 0;
  if (rc == B200_OK && hits) rc = b200_fold_get_hits(p->fold, hits, &ntot);
  if (rc == B200_OK && b200_fold_weighted(p->fold) && out->obs.rate > 0) {
    // flagged windows were skipped: time_folded = ndat_folded / rate with ndat_folded = the hits (Fold.C:783-789)
    std::vector<unsigned> tmp;
    const unsigned* h = hits;
    if (!h) {
      tmp.resize(p->desc.nbin);
      rc = b200_fold_get_hits(p->fold, tmp.data(), &ntot);
      h = tmp.data();
    }
    uint64_t folded = 0;
    for (unsigned b = 0; b < p->desc.nbin; b++) folded += h[b];
    if (rc == B200_OK) out->integration_length = double(folded) / out->obs.rate;
  }
  return rc;
}

int b200_pipeline_set_deterministic(b200_pipeline* p, float lsb) {
  B200_REQUIRE(p && p->fold, "b200_pipeline_set_deterministic: pipeline has no fold stage");
  // lsb < 0: a unit suited to input of unit variance (the 8-bit tables are scaled to it, BitTable.C:182-194): the
  // un-normalised transforms give detected samples of the order of n_fft * freq_res; 2^-20 of that per unit
  if (lsb < 0.f) lsb = float(double(p->fb->Nc) * double(p->fb->F) * std::ldexp(1.0, -20));
  return b200_fold_set_deterministic(p->fold, lsb);
}

int b200_pipeline_reset(b200_pipeline* p) {
  B200_REQUIRE(p && p->fold, "b200_pipeline_reset: pipeline has no fold stage");
  p->ps.integration_length = 0.0;
  p->ps.ndat_total = 0;
  return b200_fold_zero(p->fold);
}

// ---- streaming input with block-edge carry --------------------------------------------------
// The reference: IOManager loads blocks of the file; Filterbank/Convolution consume npart = (ndat - overlap) / step
// whole parts and tell their InputBuffering policy where the next block must start (set_next_start(step * npart),
// Filterbank.C:420-427); InputBuffering::set_next_start / pre_transformation (Kernel/Classes/InputBuffering.C:35-126)
// then copy the unconsumed tail (the overlap plus the samples of an incomplete part) in front of the next block.
// Here the tail stays on the device: a feed copies the new block straight behind the carried tail of the buffer the
// previous feed prepared, runs every whole part, and moves the new tail (from the last resolution boundary of the
// format at or before the next part) to the head of the other buffer with one device-to-device copy.  The
// host-to-device copy of feed k+1 runs on the copy stream while the kernels of feed k are still busy.
static uint64_t sample_bytes(const b200_pipeline* p, uint64_t nsamples) {
  const uint64_t bits = uint64_t(p->desc.unpack.nchan) * p->desc.unpack.npol * p->desc.unpack.ndim * fmt_nbit(p->desc.unpack.format);
  return nsamples * bits / 8;
}

int b200_pipeline_stream_begin(b200_pipeline* p, uint64_t max_block_samples, uint64_t obs_sample0) {
  B200_REQUIRE(p && max_block_samples, "b200_pipeline_stream_begin: null pipeline or empty block");
  B200_REQUIRE(p->desc.unpack.format != B200_FMT_FLOAT32, "streaming input takes raw bytes");
  Context* ctx = p->ctx;
  b200_fb_plan* fb = p->fb;
  const unsigned res = fmt_resolution(p->desc.unpack.format);
  B200_REQUIRE(max_block_samples % res == 0, "block length %llu is not a multiple of the format resolution %u",
               (unsigned long long)max_block_samples, res);
  // the carry never exceeds one step + the overlap + one resolution unit
  const uint64_t carry_max = uint64_t(fb->nsamp_step) + fb->nsamp_overlap + 2 * res;
  const uint64_t need = sample_bytes(p, carry_max + max_block_samples) + 256;
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!p->copy_stream) {
    B200_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) B200_CUDA(cudaEventCreateWithFlags(&p->stage_free[i], cudaEventDisableTiming));
    p->chunk_ready = new std::vector<cudaEvent_t>();
  }
  for (int i = 0; i < 2; i++) {
    if (need > p->stream_bytes) {
      if (p->d_stream[i]) cudaFree(p->d_stream[i]);
      p->d_stream[i] = nullptr;
      B200_CUDA(cudaMalloc(&p->d_stream[i], need));
    }
    if (!p->stream_free[i]) B200_CUDA(cudaEventCreateWithFlags(&p->stream_free[i], cudaEventDisableTiming));
  }
  if (!p->stream_loaded) B200_CUDA(cudaEventCreateWithFlags(&p->stream_loaded, cudaEventDisableTiming));
  p->stream_bytes = std::max(p->stream_bytes, need);
  p->stream_block = max_block_samples;
  p->carry_samples = p->carry_skip = 0;
  p->stream_pos = obs_sample0;
  p->stream_turn = 0;
  p->streaming = true;
  // scratch for the largest number of parts one feed can complete
  const uint64_t maxparts = (carry_max + max_block_samples) / fb->nsamp_step + 1;
  return b200_pipeline_reserve(p, maxparts);
}

static int pipeline_feed(b200_pipeline* p, const void* bytes, bool from_host, uint64_t nsamples, float* d_detected,
                         uint64_t detected_span, uint64_t* nparts_out) {
  B200_REQUIRE(p && (bytes || !nsamples), "b200_pipeline_feed: null argument");
  B200_REQUIRE(p->streaming, "b200_pipeline_feed: call b200_pipeline_stream_begin first");
  B200_REQUIRE(nsamples <= p->stream_block, "b200_pipeline_feed: block of %llu samples exceeds the %llu announced",
               (unsigned long long)nsamples, (unsigned long long)p->stream_block);
  Context* ctx = p->ctx;
  b200_fb_plan* fb = p->fb;
  const unsigned res = fmt_resolution(p->desc.unpack.format);
  B200_REQUIRE(nsamples % res == 0, "block length %llu is not a multiple of the format resolution %u",
               (unsigned long long)nsamples, res);
  if (nparts_out) *nparts_out = 0;
  const unsigned turn = p->stream_turn & 1u;
  unsigned char* buf = p->d_stream[turn];
  const uint64_t total = p->carry_samples + nsamples;
  const uint64_t avail = total - p->carry_skip;
  const uint64_t npart = avail > fb->nsamp_overlap ? (avail - fb->nsamp_overlap) / fb->nsamp_step : 0;
  B200_REQUIRE(p->desc.nbin || !npart || d_detected, "b200_pipeline_feed: nbin == 0 needs an output buffer");
  // 1. the new block lands behind the carried tail
  if (nsamples) {
    unsigned char* dst = buf + sample_bytes(p, p->carry_samples);
    if (from_host) {
      // the carried tail was written by the compute stream (step 3 of the previous feed); only the region behind it
      // is written here, and nothing reads that region before the wait below
      B200_CUDA(cudaMemcpyAsync(dst, bytes, sample_bytes(p, nsamples), cudaMemcpyHostToDevice, p->copy_stream));
      B200_CUDA(cudaEventRecord(p->stream_loaded, p->copy_stream));
      B200_CUDA(cudaStreamWaitEvent(ctx->stream, p->stream_loaded, 0));
    } else {
      B200_CUDA(cudaMemcpyAsync(dst, bytes, sample_bytes(p, nsamples), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  // 2. every whole part of [carry | block]
  int rc = B200_OK;
  if (npart) {
    const uint64_t obs_sample = p->stream_pos + p->carry_skip;
    if (p->desc.nbin && p->have_obs) {
      rc = b200_pipeline_execute_obs(p, buf, 0, p->carry_skip, npart, obs_sample);
    } else {
      B200_REQUIRE(!p->desc.nbin, "b200_pipeline_feed: a folding pipeline needs b200_pipeline_set_observation and a predictor");
      rc = pipeline_execute(p, buf, 0, p->carry_skip, npart, 0.0, 0.0, d_detected, detected_span, nullptr);
    }
    if (rc != B200_OK) return rc;
  }
  // 3. InputBuffering::set_next_start(step * npart): keep everything from the last resolution boundary at or before
  //    the next part's first sample
  const uint64_t next = p->carry_skip + npart * fb->nsamp_step;
  const uint64_t aligned = (next / res) * res;
  const uint64_t tail = total - aligned;
  const unsigned other = turn ^ 1u;
  // the other buffer may still be read by the kernels of the previous feed: they are earlier on this stream -- in order
  if (tail)
    B200_CUDA(cudaMemcpyAsync(p->d_stream[other], buf + sample_bytes(p, aligned), sample_bytes(p, tail),
                              cudaMemcpyDeviceToDevice, ctx->stream));
  B200_CUDA(cudaEventRecord(p->stream_free[turn], ctx->stream));
  // the NEXT feed's host-to-device copy writes into `other` behind the tail: it must wait for the tail copy and for
  // every kernel queued so far
  B200_CUDA(cudaStreamWaitEvent(p->copy_stream, p->stream_free[turn], 0));
  p->carry_samples = tail;
  p->carry_skip = next - aligned;
  p->stream_pos += aligned;
  p->stream_turn++;
  if (nparts_out) *nparts_out = npart;
  return B200_OK;
}

int b200_pipeline_feed_host(b200_pipeline* p, const void* h_bytes, uint64_t nsamples, float* d_detected,
                            uint64_t detected_span, uint64_t* nparts) {
  return pipeline_feed(p, h_bytes, true, nsamples, d_detected, detected_span, nparts);
}

int b200_pipeline_feed(b200_pipeline* p, const void* d_bytes, uint64_t nsamples, float* d_detected,
                       uint64_t detected_span, uint64_t* nparts) {
  return pipeline_feed(p, d_bytes, false, nsamples, d_detected, detected_span, nparts);
}

int b200_pipeline_synch(b200_pipeline* p, float* h_profile, unsigned* h_hits, uint64_t* ndat_total) {
  B200_REQUIRE(p && p->fold, "b200_pipeline_synch: pipeline has no fold stage");
  int rc = B200_OK;
  if (h_profile) rc = b200_fold_synch(p->fold, h_profile);
  if (rc == B200_OK) rc = b200_fold_get_hits(p->fold, h_hits, ndat_total);
  return rc;
}

int b200_pipeline_zero(b200_pipeline* p) {
  B200_REQUIRE(p && p->fold, "b200_pipeline_zero: pipeline has no fold stage");
  return b200_fold_zero(p->fold);
}

}  // extern "C"
