/* b200dsp.h -- C ABI of libb200dsp.so: dspsr's baseband hot path on NVIDIA B200 (sm_100a).
 *
 * unpack -> overlap-save coherent dedispersion / filterbank -> detection -> fold
 *
 * Every entry point is extern "C", takes plain pointers and sizes, returns a b200_status
 * (0 = success) and never throws; b200_last_error() holds the message of the last failure
 * on the calling thread.  Pointers named d_* are device pointers, h_* host pointers.
 * Nothing in execute-type calls allocates.  One b200_context == one CUDA stream == one dspsr
 * pipeline thread (reference: Signal/General/SingleThread.C:237-244 creates exactly one
 * stream per thread; engines obtain it from CUDA::DeviceMemory::get_stream()).
 *
 * Each group cites the reference interface it replaces (paths relative to demorest/dspsr).
 * The precedent for a C struct + C function boundary under the C++ engines is the
 * reference's own Signal/General/dsp/filterbank_engine.h:25-38 and
 * dsp/filterbank_cuda.h:64-66 (filterbank_cuda_perform).
 */
#ifndef B200DSP_H
#define B200DSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200_OK = 0,
  B200_ERR_INVALID = 1,     /* bad argument / state            (reference: Error(InvalidParam|InvalidState)) */
  B200_ERR_CUDA = 2,        /* CUDA runtime failure            (reference: Kernel/Classes/check_error.C)    */
  B200_ERR_UNSUPPORTED = 3, /* shape outside what is built                                                  */
  B200_ERR_NOMEM = 4
} b200_status;

typedef struct b200_context b200_context;
typedef struct b200_fb_plan b200_fb_plan;
typedef struct b200_fold b200_fold;
typedef struct b200_pipeline b200_pipeline;

/* ---------------------------------------------------------------------------------------
 * Library / context.  Replaces: CUDA::DeviceMemory(stream, device) and the cudaSetDevice /
 * cudaStreamCreate block of Signal/General/SingleThread.C:237-244; dsp::Memory::do_allocate /
 * do_free / do_zero / do_copy (Kernel/Classes/MemoryCUDA.C).
 * ------------------------------------------------------------------------------------- */
int b200_version(void);
const char* b200_last_error(void);

/* cuda_stream: a cudaStream_t to run on, or NULL to create a private non-blocking stream. */
int b200_context_create(int device, void* cuda_stream, b200_context** ctx);
int b200_context_destroy(b200_context* ctx);
int b200_context_synchronize(b200_context* ctx);
/* kernels launched through this context so far (bench.py's gpu_launches) */
unsigned long long b200_context_launch_count(const b200_context* ctx);
void* b200_context_stream(const b200_context* ctx);
/* Per-kernel-class device timing for roofline reports: while enabled, every launch is bracketed
 * by CUDA events on the context's stream.  read_timing synchronises and returns, per class
 * (0 forward column pass, 1 row pass, 2 inverse pass + epilogue, 3 bin plan, 4 other), the summed
 * milliseconds and launch counts since the last read (arrays of 5). */
int b200_context_set_timing(b200_context* ctx, int enable);
int b200_context_read_timing(b200_context* ctx, double* ms, unsigned long long* count);

int b200_malloc(b200_context* ctx, uint64_t nbytes, void** d_ptr);
int b200_free(b200_context* ctx, void* d_ptr);
int b200_malloc_host(b200_context* ctx, uint64_t nbytes, void** h_ptr); /* pinned */
int b200_free_host(b200_context* ctx, void* h_ptr);
int b200_memset(b200_context* ctx, void* d_ptr, int value, uint64_t nbytes);
/* dsp::Memory::do_copy on the device (CUDA::DeviceMemory::do_copy, Kernel/Classes/MemoryCUDA.C:90-106); stream ordered */
int b200_memcpy_d2d(b200_context* ctx, void* d_dst, const void* d_src, uint64_t nbytes);
int b200_memcpy_h2d(b200_context* ctx, void* d_dst, const void* h_src, uint64_t nbytes);
int b200_memcpy_d2h(b200_context* ctx, void* h_dst, const void* d_src, uint64_t nbytes);

/* ---------------------------------------------------------------------------------------
 * Unpacker device hook.  Replaces: Unpacker::unpack() on a device (Kernel/Classes/dsp/
 * Unpacker.h:57-61,100) for
 *   B200_FMT_CASPSR8   CASPSRUnpacker::unpack          Kernel/Formats/caspsr/CASPSRUnpacker.C:132-187
 *   B200_FMT_GENERIC8  BitUnpacker/EightBitUnpacker     Kernel/Classes/BitUnpacker.C:48-80
 *   B200_FMT_MEERKAT8  MeerKATUnpacker::unpack (FPT)    Kernel/Formats/kat/MeerKATUnpacker.C:196-229
 *   B200_FMT_UWB16     UWBUnpacker::unpack              Kernel/Formats/uwb/UWBUnpacker.C:177-218
 * Output is a dsp::TimeSeries in FPT order: plane(ichan,ipol) = d_out + (ichan*npol+ipol)*out_span
 * floats, element [idat*ndim + idim].  The 8-bit LUT formats index the HOST-BUILT 256-entry
 * table of dsp::BitTable (BitTable.C:165-218) so results are bit-identical to the CPU path.
 * ------------------------------------------------------------------------------------- */
typedef enum {
  B200_FMT_CASPSR8 = 0,
  B200_FMT_GENERIC8 = 1,
  B200_FMT_MEERKAT8 = 2,
  B200_FMT_UWB16 = 3,
  B200_FMT_FLOAT32 = 4, /* already a float TimeSeries (pipeline input only) */
  B200_FMT_TWOBIT = 5   /* 2-bit real-sampled, one digitizer per polarisation, bytes interleaved (CPSR2) */
} b200_format;

/* Two-bit excision unpacker.  Replaces: TwoBitCorrection::build/dig_unpack (Kernel/Classes/
 * TwoBitCorrection.C:89-151), ExcisionUnpacker::set_limits/unpack (ExcisionUnpacker.C:95-158,174-256),
 * excision_unpack (dsp/excision_unpack.h:21-106), TwoBitFour (dsp/TwoBitFour.h:42-89), TwoBitLookup::
 * lookup_build (TwoBitLookup.C:63-98).  Per digitizer and window of ndat_per_weight samples: nlow =
 * number of low-voltage states; the output levels (lo, hi) are those of Jenet & Anderson (1998) for
 * Phi = nlow/ndat_per_weight (clamped to [nlow_min, nlow_max]); windows that are all-zero bytes or whose
 * nlow falls outside the limits are zeroed and their weight set to 0.  The level table is built on the
 * HOST (b200_twobit_prepare) so that device output is bit-identical to the CPU path. */
typedef struct {
  int table_type;            /* 0 OffsetBinary (CPSR2TwoBitCorrection.C:21), 1 SignMagnitude, 2 TwosComplement */
  unsigned npol;             /* digitizers = polarisations (1 or 2) */
  unsigned ndat_per_weight;  /* TwoBitCorrection.C:31: 512 */
  unsigned nlow_min, nlow_max;
  float lo[513], hi[513];    /* levels of row nlow at index nlow - nlow_min */
} b200_twobit_desc;

/* threshold: sampling threshold in sigma (JenetAnderson98 optimal 4-level value 0.9674);
 * cutoff_sigma: ExcisionUnpacker default 10.0, 0 disables the limits */
int b200_twobit_prepare(double threshold, float cutoff_sigma, int table_type, unsigned npol,
                        unsigned ndat_per_weight, b200_twobit_desc* desc);
/* d_weights (nullable): ndat/ndat_per_weight entries, 1 = good, 0 = flagged in any polarisation
 * (WeightedTimeSeries::mask_weights).  ndat must be a multiple of ndat_per_weight. */
int b200_unpack_twobit(b200_context* ctx, const b200_twobit_desc* desc, const void* d_raw, uint64_t ndat,
                       float* d_out, uint64_t out_span, unsigned* d_weights);

typedef struct {
  int format;            /* b200_format */
  unsigned nchan, npol, ndim;
  float lut[256];        /* CASPSR8 / GENERIC8: BitTable::get_values() */
  float scale;           /* MEERKAT8: float(BitTable::get_scale()) */
  unsigned sample_swap;  /* MEERKAT8: 1 (MKBF) or 2 (MKBFRo) */
  const b200_twobit_desc* twobit; /* TWOBIT only (copied by b200_pipeline_create) */
} b200_unpack_desc;

int b200_unpack(b200_context* ctx, const b200_unpack_desc* desc, const void* d_raw, uint64_t ndat,
                float* d_out, uint64_t out_span);

/* ---------------------------------------------------------------------------------------
 * Filterbank / Convolution engine.  Replaces: dsp::Filterbank::Engine::{setup,set_scratch,
 * perform,finish} (Signal/General/dsp/FilterbankEngine.h:15-44; reference implementation
 * CUDA::FilterbankEngine, FilterbankCUDA.cu:60-304) and dsp::Convolution::Engine::{prepare,
 * set_scratch,perform} (Signal/General/dsp/Convolution.h:158-167; ConvolutionCUDA.cu).
 * Convolution is the nchan_subband = 1 case (freq_res = response ndat).
 * The response is the finished, matched dsp::Response buffer (host float pairs,
 * input_nchan*nchan_subband*freq_res complex) exactly as the reference hands it to its
 * engines (FilterbankCUDA.cu:134-165); NULL = no response (plain filterbank).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int input_real;          /* 1: Signal::Nyquist (ndim 1), 0: Signal::Analytic (ndim 2) */
  unsigned input_nchan;
  unsigned npol;
  unsigned nchan_subband;  /* output channels per input channel (Filterbank.C:68) */
  unsigned freq_res;       /* points per output channel = Response::get_ndat() (Filterbank.C:93) */
  unsigned nfilt_pos;      /* Response::get_impulse_pos() */
  unsigned nfilt_neg;      /* Response::get_impulse_neg() */
  const float* h_response; /* may be NULL */
  unsigned max_npart;      /* parts per internal batch (scratch sizing); 0 = library default */
} b200_fb_desc;

typedef struct {
  unsigned n_fft;          /* complex points of the forward transform (nchan_subband*freq_res) */
  unsigned nsamp_fft, nsamp_overlap, nsamp_step, nkeep;   /* Filterbank.C:131-155,409 */
  unsigned fft_rows, fft_cols;   /* two-pass factorisation P x Q of n_fft (Q = 1: single pass) */
  unsigned batch_npart;
  uint64_t scratch_bytes;
} b200_fb_info;

int b200_fb_plan_create(b200_context* ctx, const b200_fb_desc* desc, b200_fb_plan** plan);
int b200_fb_plan_info(const b200_fb_plan* plan, b200_fb_info* info);
int b200_fb_plan_destroy(b200_fb_plan* plan);

/* Filterbank::Engine::perform (FilterbankEngine.h:29-33): d_in/d_out are get_datptr(0,0) of
 * the device TimeSeries, *_span = floats between consecutive (chan,pol) planes
 * (DataSeries::get_nfloat_span), in_step = nsamp_step*ndim floats, out_step = nkeep*2 floats
 * (Filterbank.C:517-523).  For Convolution::Engine::perform pass the same quantities
 * (Convolution.C:373,387). */
int b200_fb_perform(b200_fb_plan* plan, const float* d_in, uint64_t in_span, float* d_out, uint64_t out_span,
                    uint64_t npart, uint64_t in_step, uint64_t out_step);

/* ---------------------------------------------------------------------------------------
 * Detection engine.  Replaces: dsp::Detection::Engine::{polarimetry,square_law}
 * (Signal/General/dsp/Detection.h:98-106; DetectionCUDA.cu:127-177,246-310).  Unlike the
 * reference CUDA engine (ndim 2 only) every Detection::get_result_pointers layout
 * (Detection.C:423-474, ndim 1|2|4) is produced.  In-place (d_in == d_out) is allowed for
 * ndim_out == 2 only, as in the reference (Detection.C:358-361).
 * ------------------------------------------------------------------------------------- */
typedef enum { B200_INTENSITY = 0, B200_PPQQ = 1, B200_COHERENCE = 2, B200_STOKES = 3 } b200_state;

int b200_detect(b200_context* ctx, int state, unsigned ndim_out, const float* d_in, uint64_t in_span,
                unsigned nchan, unsigned npol, uint64_t ndat, float* d_out, uint64_t out_span);

/* ---------------------------------------------------------------------------------------
 * Fold engine.  Replaces: dsp::Fold::Engine::{set_nbin,set_ndat,set_bins,get_bin_hits,
 * get_ndat_folded,fold,synch,zero} (Signal/Pulsar/dsp/Fold.h:249-312; CUDA::FoldEngine,
 * FoldCUDA.cu:84-152,586-697).  The engine owns the accumulating device PhaseSeries
 * [nchan][npol][nbin][ndim] floats (Fold.C:88-94).  set_bins reproduces the host's sequential
 * double-precision recurrence of Fold.C:765-768 bit for bit (see DESIGN.md "bin plan").
 * ------------------------------------------------------------------------------------- */
int b200_fold_create(b200_context* ctx, unsigned nchan, unsigned npol, unsigned ndim, unsigned nbin,
                     b200_fold** fold);
int b200_fold_destroy(b200_fold* fold);
/* Fold::Engine::set_bins(phi, phase_per_sample, ndat, idat_start) (Fold.h:261). */
int b200_fold_set_bins(b200_fold* fold, double phi, double phase_per_sample, uint64_t ndat,
                       uint64_t idat_start, uint64_t* ndat_folded);
/* The same for a WeightedTimeSeries input whose per-window flags live on the device (Fold.C:687-716,746-763): sample
 * idat belongs to window (idat + weight_idat) / ndatperweight; samples of windows flagged 0 are not folded and give no
 * hit.  d_weights NULL or ndatperweight 0: identical to b200_fold_set_bins.  The number of samples folded is the sum
 * of the hits (b200_fold_get_hits). */
int b200_fold_set_bins_weighted(b200_fold* fold, double phi, double phase_per_sample, uint64_t ndat, uint64_t idat_start,
                                const unsigned* d_weights, uint64_t nweights, unsigned ndatperweight,
                                uint64_t weight_idat);
/* 1 when some set_bins since the last zero skipped flagged samples */
int b200_fold_weighted(const b200_fold* fold);
/* Fold::Engine::get_bin_hits for every bin of the last set_bins (Fold.C:731-735). */
int b200_fold_get_bin_hits(b200_fold* fold, unsigned* h_hits);
/* Fold::Engine::fold(): d_in = input->get_datptr(0,0), in_span = get_nfloat_span (Fold.C:989-990). */
int b200_fold_fold(b200_fold* fold, const float* d_in, uint64_t in_span);
/* Same, accumulating into a caller-owned device PhaseSeries instead of the handle's own: d_out = the engine-owned
 * PhaseSeries' get_datptr(0,0) and out_span = its get_nfloat_span(), exactly the `output` / `output_span` that
 * Fold::Engine::setup hands every engine (Fold.C:992-996; CUDA::FoldEngine folds into d_profiles the same way,
 * FoldCUDA.cu:122,586-697).  Plane (ichan, ipol) starts at d_out + (ichan*npol + ipol)*out_span. */
int b200_fold_fold_into(b200_fold* fold, const float* d_in, uint64_t in_span, float* d_out, uint64_t out_span);
/* Fold::Engine::synch(PhaseSeries*): copies the device profiles to the host (idempotent). */
int b200_fold_synch(b200_fold* fold, float* h_profile);
/* accumulated hits of every set_bins since the last zero (PhaseSeries::get_hits) */
int b200_fold_get_hits(b200_fold* fold, unsigned* h_hits, uint64_t* ndat_total);
int b200_fold_zero(b200_fold* fold);
/* Reproducible accumulation.  By default run sums reach the PhaseSeries as float reductions (RED.ADD.F32) from many
 * CTAs: the result depends on their arrival order in the last bits (the reference's own multi-threaded result
 * depends on the thread count in the same way).  With lsb > 0 every run sum is rounded to a multiple of lsb and
 * added to a 64-bit integer accumulator instead: integer addition is associative, so the profile is bit-identical
 * from run to run and independent of the launch order.  lsb must be far below the size of a sample (2^-20 of a
 * typical detected sample keeps the rounding below that of the float path) and large enough that the total stays
 * below 2^63 * lsb.  lsb <= 0 returns to float accumulation.  Only on a zeroed PhaseSeries. */
int b200_fold_set_deterministic(b200_fold* fold, float lsb);
/* device pointer of the accumulating profile (for NCCL reductions at sub-integration ends) */
float* b200_fold_device_profile(b200_fold* fold);
unsigned* b200_fold_device_hits(b200_fold* fold);

/* ---------------------------------------------------------------------------------------
 * Fused path: raw bytes -> folded profile without materialising the unpacked, filtered or
 * detected time series (one spectrum round trip).  It is what the four engines above
 * compute when dspsr wires [IOManager, Filterbank|Convolution, Detection, Fold]
 * (Signal/Pulsar/LoadToFold1.C:117-599) with nothing in between.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  b200_unpack_desc unpack; /* format FLOAT32: d_input is a float TimeSeries (span = input_span) */
  b200_fb_desc fb;
  int detect_state;        /* b200_state */
  unsigned detect_ndim;    /* 1, 2 or 4 for COHERENCE / STOKES */
  unsigned nbin;           /* 0: no fold -- detected series is written to d_detected */
} b200_pipeline_desc;

int b200_pipeline_create(b200_context* ctx, const b200_pipeline_desc* desc, b200_pipeline** pipe);
int b200_pipeline_destroy(b200_pipeline* pipe);
int b200_pipeline_info(const b200_pipeline* pipe, b200_fb_info* info);

/* One block = one Fold::transformation call: npart overlap-save parts whose first sample is
 * sample `first_sample` of d_input (d_input itself must start on a boundary of the format's
 * resolution: 4 samples CASPSR, 256-sample heap MeerKAT, 2048-sample block UWB); phi /
 * phase_per_sample as Fold::fold computes them (Fold.C:650-657,718-720) for the block's first
 * OUTPUT sample.  input_span: floats between (chan,pol) planes when the format is FLOAT32,
 * ignored for raw formats.  d_detected (nbin == 0 only): planes of npart*nkeep*ndim' floats
 * with span detected_span. */
int b200_pipeline_execute(b200_pipeline* pipe, const void* d_input, uint64_t input_span, uint64_t first_sample,
                          uint64_t npart, double phi, double phase_per_sample, float* d_detected,
                          uint64_t detected_span);
/* Sizes every scratch array for blocks of up to max_npart parts (what Convolution::reserve / Filterbank::reserve do
 * for the reference's scratch space): afterwards execute-type calls with npart <= max_npart do not allocate.
 * Without it the arrays grow on first use (one synchronisation each time a larger block arrives). */
int b200_pipeline_reserve(b200_pipeline* pipe, uint64_t max_npart);
/* Same, from HOST memory (pinned or pageable; File::load_bytes_device, Kernel/Classes/File.C:213-272): the bytes
 * are copied chunk by chunk on a private copy stream into one of two staging buffers while the kernels of the
 * previous chunk run.  The call RETURNS BEFORE THE COPIES HAVE FINISHED: h_input must stay untouched until
 * b200_pipeline_input_consumed (or any later synchronisation of the pipeline's stream) has returned. */
int b200_pipeline_execute_host(b200_pipeline* pipe, const void* h_input, uint64_t nbytes, uint64_t first_sample,
                               uint64_t npart, double phi, double phase_per_sample, float* d_detected,
                               uint64_t detected_span);
/* Blocks the calling thread until every host buffer handed to b200_pipeline_execute_host so far has been read
 * (the copy stream is idle); kernels may still be running. */
int b200_pipeline_input_consumed(b200_pipeline* pipe);
int b200_pipeline_synch(b200_pipeline* pipe, float* h_profile, unsigned* h_hits, uint64_t* ndat_total);
int b200_pipeline_zero(b200_pipeline* pipe);
b200_fold* b200_pipeline_fold(b200_pipeline* pipe);

/* ---------------------------------------------------------------------------------------
 * Host-side helpers with no device work (exported so that hosts in any language share the
 * exact arithmetic): the bin-plan recurrence and its closed form.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t start;      /* first sample of the segment */
  uint64_t count;      /* samples in the segment */
  uint64_t a0;         /* phase of the first sample in units of 2^scale_exp */
  uint64_t step;       /* phase increment per sample in the same units */
  int scale_exp;       /* phase = (a0 + i*step) * 2^scale_exp, exactly */
  int pad;
} b200_phase_segment;

/* Splits the recurrence  phi -= floor(phi); ibin = unsigned(phi*nbin); phi += pps  (Fold.C:765-768)
 * over ndat samples into segments inside which phi is an exact arithmetic progression.
 * Returns the number of segments written (<= max_segments) or -1 if more are needed.
 * phi_end receives phi after the last sample. */
int64_t b200_phase_segments(double phi, double phase_per_sample, uint64_t ndat, b200_phase_segment* segments,
                            uint64_t max_segments, double* phi_end);
/* The plain sequential recurrence (reference behaviour), for hosts and tests. */
void b200_phase_bins_sequential(double phi, double phase_per_sample, unsigned nbin, uint64_t ndat,
                                unsigned* bins, double* phi_end);

/* ---------------------------------------------------------------------------------------
 * Host-side restatements of what the reference computes on the CPU around the engines
 * (no device work).  Inside a real dspsr tree these come from dspsr / PSRCHIVE; they are
 * exported for stand-alone hosts (bench.py, the Python handles, host/b200_demo.cpp).
 * ------------------------------------------------------------------------------------- */

/* dsp::BitTable(8, TwosComplement|OffsetBinary): 256-entry value table and get_scale()
 * (Kernel/Classes/BitTable.C:121-218). */
int b200_bittable8(int twos_complement, float* lut256, double* scale);

/* dsp::Dedispersion (Signal/General/Dedispersion.C) + Response::match (Response.C:132-181). */
typedef struct {
  double centre_frequency;       /* MHz */
  double bandwidth;              /* MHz, signed */
  double dispersion_measure;     /* pc cm^-3 */
  unsigned input_nchan;          /* channels of the input Observation */
  unsigned nchan;                /* channels of the response = output channels */
  int input_dual_sideband;       /* Observation::get_dual_sideband (Observation.C:80-87) */
  int input_dc_centred;
  int input_swap;
  unsigned frequency_resolution; /* 0: Response::set_optimal_ndat; else the -x nfft override */
  /* filled by b200_dedispersion_prepare: */
  unsigned impulse_pos, impulse_neg, ndat;
} b200_dedispersion;

/* Dedispersion::prepare (Dedispersion.C:216-248) and the ndat choice of Dedispersion::build
 * (:296-308) via optimal_fft_length (optimize_fft.c:63-127). */
int b200_dedispersion_prepare(b200_dedispersion* d);
/* Dedispersion::build (:310-331,478-556) then Response::match and the DC zap (:278):
 * h_response receives nchan*ndat complex floats in the order the FFT produces them. */
int b200_dedispersion_build(const b200_dedispersion* d, float* h_response);
/* The rows [first_chan, first_chan + nchan_local) of the same matched response: the slice one rank of a
 * channel-sharded run needs (channels are independent, Convolution.C:389-391).  Requires
 * nchan == input_nchan and no input swap (B200_ERR_UNSUPPORTED otherwise). */
int b200_dedispersion_build_channels(const b200_dedispersion* d, unsigned first_chan, unsigned nchan_local,
                                     float* h_response);
int64_t b200_optimal_fft_length(uint64_t nbadperfft, uint64_t nfft_max);

/* TEMPO polyco block (what Pulsar::Predictor::phase/frequency evaluate for Fold.C:943-958). */
typedef struct {
  int tmid_day;
  double tmid_sec;
  double rphase_int, rphase_frac;
  double f0, span_min, obsfreq, dm;
  int ncoef;
  double coef[32];
} b200_polyco;

int b200_polyco_parse(const char* text, b200_polyco* pc);
/* fractional turns in [0,1) at MJD (day, sec, frac); *turns (nullable) gets the integer part */
double b200_polyco_phase(const b200_polyco* pc, int day, int sec, double frac, double* turns);
double b200_polyco_frequency(const b200_polyco* pc, int day, int sec, double frac);

/* digifil tail.  Replaces dsp::Rescale::transformation / compute_various
 * (Signal/General/Rescale.C:165-412; FPT order, default mode: statistics over `interval_samples`, 0 = the
 * length of the first block; offset = -mean, scale = 1/sqrt(variance) on the first call and at every interval
 * end) and dsp::SigProcDigitizer::pack for 8 bits (Kernel/Formats/sigproc/SigProcDigitizer.C:80-160,244-300:
 * TPF bytes clip(int(x*digi_scale + 127.5 + 0.5), 0, 255), digi_scale = (127.5/6)/(input_scale*scale_fac),
 * channels re-ordered by ChannelSort :38-70). */
typedef struct b200_rescale b200_rescale;
int b200_rescale_create(b200_context* ctx, unsigned nchan, unsigned npol, uint64_t interval_samples, int constant,
                        b200_rescale** out);
int b200_rescale_destroy(b200_rescale* r);
/* d_in / d_out: planes (ichan*npol+ipol)*span of ndat floats; in place allowed */
int b200_rescale_transform(b200_rescale* r, const float* d_in, uint64_t in_span, uint64_t ndat, float* d_out,
                           uint64_t out_span);
/* current offset / scale, [nchan][npol] each (synchronises) */
int b200_rescale_get(b200_rescale* r, float* h_offset, float* h_scale);
/* d_out: ndat*npol*nchan bytes in TPF order; flip_band: input bandwidth > 0; swap_band: Observation::get_swap */
int b200_sigproc_digitize8(b200_context* ctx, const float* d_in, uint64_t in_span, unsigned nchan, unsigned npol,
                           uint64_t ndat, float digi_scale, float digi_mean, float xpol_offset, int flip_band,
                           int swap_band, unsigned char* d_out);

/* ---------------------------------------------------------------------------------------
 * WeightedTimeSeries flags on the device.  Replaces: WeightedTimeSeries::convolve_weights and
 * scrunch_weights (Kernel/Classes/WeightedTimeSeries.C:582-690,692-780) as called by Filterbank::prepare_output
 * (Filterbank.C:302-307) and Convolution::prepare_output (Convolution.C:312-319).  Flags: one uint32 per window of
 * ndat_per_weight samples, 0 = bad.  The fused pipeline applies them itself to two-bit input.
 * ------------------------------------------------------------------------------------- */
/* d_out (distinct from d_weights) receives the flags after the convolution of a block of ndat samples processed as
 * overlap-save transforms of nfft samples every nkeep samples; d_scratch: (ndat / nkeep + 2) words.  Requires
 * nkeep >= ndat_per_weight. */
int b200_weights_convolve(b200_context* ctx, const unsigned* d_weights, uint64_t nweights, unsigned ndat_per_weight,
                          uint64_t weight_idat, uint64_t ndat, unsigned nfft, unsigned nkeep, unsigned* d_out,
                          unsigned* d_scratch);
/* nweights / ndat_per_weight / weight_idat are updated like the members of the reference.  When ndat_per_weight >=
 * nscrunch only they change (d_out may be NULL); otherwise d_out (distinct) receives the scrunched flags. */
int b200_weights_scrunch(b200_context* ctx, const unsigned* d_weights, uint64_t* nweights, unsigned* ndat_per_weight,
                         uint64_t* weight_idat, unsigned nscrunch, unsigned* d_out);

/* ---------------------------------------------------------------------------------------
 * PhaseSeries on the host: the accumulator's attributes and its merge / unload rules.
 * Replaces: dsp::PhaseSeries::mixable / combine (Signal/Pulsar/PhaseSeries.C:336-418,442-480) with the
 * attribute comparison of dsp::Observation::combinable (Kernel/Classes/Observation.C:139-310), the
 * bookkeeping of Fold::fold (Fold.C:796-812: integration_length += ndat_folded / rate, ndat_total += ndat_fold)
 * and the normalisation applied when a PhaseSeries is archived (dsp::Archiver::set, Signal/Pulsar/Archiver.C:
 * 773-895: amplitude / (scale * hits), bins without hits set to the mean of the others, non-finite profiles
 * zeroed with weight 0).  Times are split MJDs (integer day, integer second, fraction) like PSRCHIVE's MJD.
 * ------------------------------------------------------------------------------------- */
typedef struct { int day; int sec; double frac; } b200_mjd;

/* the dsp::Observation attributes that travel with a block of data and decide whether two can be combined */
typedef struct {
  char telescope[32], receiver[32], source[32], mode[32], machine[32], format[32];
  double centre_frequency, bandwidth;      /* MHz */
  double rate;                             /* samples per second of this series */
  double scale;                            /* product of the un-normalised FFT lengths so far (Observation::scale) */
  double dispersion_measure, rotation_measure;
  unsigned nchan, npol, ndim, nbit;
  int state;                               /* Signal::State as b200_state, or 16 + ndim_in for undetected data */
  int type, basis, swap, nsub_swap, dc_centred;
  b200_mjd start_time;                     /* time of sample 0 */
  uint64_t ndat;
} b200_observation;

typedef struct {
  b200_observation obs;                    /* attributes of the folded series (copied from the first block mixed in) */
  unsigned nbin, hits_nchan;               /* hits_nchan = 1: hits are common to all channels (no zeroed data) */
  double integration_length;               /* seconds */
  uint64_t ndat_total, ndat_expected;
  b200_mjd end_time;                       /* obs.start_time .. end_time bound the folded data */
  double folding_period, reference_phase;
  float* data;                             /* [nchan][npol][nbin][ndim], caller-owned host memory */
  unsigned* hits;                          /* [hits_nchan][nbin] */
} b200_phase_series;

/* Observation::combinable: 1 / 0; `reason` (nullable, reason_len bytes) receives the reference's explanation */
int b200_observation_combinable(const b200_observation* a, const b200_observation* b, char* reason, unsigned reason_len);
/* seconds from a to b */
double b200_mjd_diff(const b200_mjd* b, const b200_mjd* a);
b200_mjd b200_mjd_add(const b200_mjd* t, double seconds);
/* PhaseSeries::mixable(obs, nbin, istart, fold_ndat): prepares an empty PhaseSeries (attributes copied, data and
 * hits zeroed, ndat_total kept) or checks combinable + nbin and widens [start_time, end_time].  Returns 1 / 0. */
int b200_phase_series_mixable(b200_phase_series* ps, const b200_observation* obs, unsigned nbin, int64_t istart,
                              int64_t fold_ndat);
/* The bookkeeping of one Fold::fold call after mixable (Fold.C:796-812). */
int b200_phase_series_folded(b200_phase_series* ps, uint64_t ndat_folded, uint64_t ndat_fold);
/* PhaseSeries::combine: B200_OK, or B200_ERR_INVALID ("PhaseSeries !mixable"). */
int b200_phase_series_combine(b200_phase_series* ps, const b200_phase_series* other);
/* Archiver::set for every (ichan, ipol, idim): h_profiles [nchan][npol][ndim][nbin] floats, h_weights
 * [nchan][npol][ndim] (1, or 0 for a corrupted profile); *corrupted (nullable) counts the latter. */
int b200_phase_series_normalise(const b200_phase_series* ps, float* h_profiles, float* h_weights, unsigned* corrupted);
/* Self-describing dump of a sub-integration (the reference writes PSRFITS through PSRCHIVE, which
 * cannot be built here): a 4096-byte ASCII header of `KEY value` lines (DADA style), then the normalised profiles
 * (float32 [nchan][npol][ndim][nbin]), the weights (float32 [nchan][npol][ndim]), the hits (uint32 [hits_nchan][nbin])
 * and the raw accumulated sums (float32 [nchan][npol][nbin][ndim]). */
int b200_phase_series_unload(const b200_phase_series* ps, const char* path);
/* Reads the header of such a file into *ps (data / hits pointers untouched) and, when the buffers are given, the
 * arrays (any may be NULL). */
int b200_phase_series_load(const char* path, b200_phase_series* ps, float* h_profiles, float* h_weights,
                           unsigned* h_hits, float* h_raw);

/* ---------------------------------------------------------------------------------------
 * The fused path driven like dsp::Fold::transformation drives it: the library itself evaluates the predictor for
 * every block (Fold.C:650-657 get_phi / get_pfold at the midpoint of the block's first output sample, :718-720
 * phase_per_sample), derives the attributes of the series that reaches Fold from those of the raw input
 * (Filterbank.C:325-371 / Convolution.C:286-305: rate *= freq_res / nsamp_fft, scale *= n_fft * freq_res or
 * nsamp_fft * n_fft, start_time += nfilt_pos samples; Detection: state, npol, ndim) and keeps the PhaseSeries
 * attributes (mixable, integration_length, ndat_total, start / end) next to the device accumulator.
 * ------------------------------------------------------------------------------------- */
/* attributes of the RAW input (rate = samples per second of one input channel, start_time = time of sample 0) */
int b200_pipeline_set_observation(b200_pipeline* pipe, const b200_observation* raw_obs);
int b200_pipeline_set_predictor(b200_pipeline* pipe, const b200_polyco* polyco, double reference_phase);
/* constant-period folding instead (Fold::set_folding_period, Fold::set_reference_epoch; reference_epoch NULL = MJD 0,
 * the reference's default): phi = fmod((t - epoch) seconds, period) / period - reference_phase (Fold.C:943-950) */
int b200_pipeline_set_folding_period(b200_pipeline* pipe, double period_seconds, double reference_phase,
                                     const b200_mjd* reference_epoch);
/* One block whose first sample (sample `first_sample` of the buffer) is sample `obs_sample` of the observation.
 * B200_ERR_INVALID ("PhaseSeries !mixable") when the block cannot be added to what has been folded so far. */
int b200_pipeline_execute_obs(b200_pipeline* pipe, const void* d_input, uint64_t input_span, uint64_t first_sample,
                              uint64_t npart, uint64_t obs_sample);
int b200_pipeline_execute_host_obs(b200_pipeline* pipe, const void* h_input, uint64_t nbytes, uint64_t first_sample,
                                   uint64_t npart, uint64_t obs_sample);
/* Streaming input with block-edge carry.  Replaces the block loop of IOManager (Kernel/Classes/
 * IOManager.C:322-470) with InputBuffering::set_next_start / pre_transformation (InputBuffering.C:35-126) as driven
 * by Filterbank::transformation (Filterbank.C:420-427): blocks of ANY length (a multiple of the format's resolution)
 * are appended to the samples left over from earlier blocks; every whole overlap-save part is processed, the rest
 * (overlap + incomplete part) is carried on the device.  Folding pipelines must have an observation and a predictor
 * (each feed is one Fold call whose phase the library evaluates); for nbin == 0 the detected samples of the *nparts
 * completed parts go to d_detected.  stream_begin sizes every buffer for blocks of up to max_block_samples. */
int b200_pipeline_stream_begin(b200_pipeline* pipe, uint64_t max_block_samples, uint64_t obs_sample0);
int b200_pipeline_feed_host(b200_pipeline* pipe, const void* h_bytes, uint64_t nsamples, float* d_detected,
                            uint64_t detected_span, uint64_t* nparts);
int b200_pipeline_feed(b200_pipeline* pipe, const void* d_bytes, uint64_t nsamples, float* d_detected,
                       uint64_t detected_span, uint64_t* nparts);
/* Fold::Engine::synch + the attributes: fills *ps; when ps->data / ps->hits are non-NULL they receive the
 * accumulated sums [nchan][npol][nbin][ndim] and hits [nbin] (synchronises the stream). */
int b200_pipeline_get_phase_series(b200_pipeline* pipe, b200_phase_series* ps);
/* b200_fold_set_deterministic on the pipeline's fold stage; lsb < 0 picks a unit from the transform sizes for input
 * of unit variance (n_fft * freq_res * 2^-20). */
int b200_pipeline_set_deterministic(b200_pipeline* pipe, float lsb);
/* Fold::reset -> Engine::zero + PhaseSeries::zero: clears the sums, the hits and integration_length / ndat_total */
int b200_pipeline_reset(b200_pipeline* pipe);

/* SIGPROC filterbank (.fil) header / file: the end of digifil.  Replaces SigProcOutputFile::write_header
 * (Kernel/Formats/sigproc/SigProcOutputFile.C:35-60) -> SigProcObservation::unload_global (SigProcObservation.C:
 * 228-275) -> filterbank_header (filterbank_header.c:44-100, send_stuff.c). */
typedef struct {
  char rawdatafile[80], source_name[80];
  int machine_id, telescope_id, nchans, nbits, nifs, nbeams, ibeam;
  double src_raj, src_dej, az_start, za_start, fch1, foff, tstart, tsamp;
} b200_sigproc_header;
/* header fields of the BitSeries SigProcDigitizer::pack makes of the detected series `detected`
 * (SigProcDigitizer.C:84-86: bandwidth = -|bandwidth|, channels in descending frequency order) */
int b200_sigproc_header_from_observation(const b200_observation* detected, unsigned nbit, b200_sigproc_header* h);
/* returns the number of bytes written to buf, or -1 (buffer too small) */
int64_t b200_sigproc_header_write(const b200_sigproc_header* h, unsigned char* buf, uint64_t buflen);
/* header + TPF bytes (append != 0: bytes only, added to an existing file) */
int b200_sigproc_file_write(const char* path, const b200_sigproc_header* h, const unsigned char* h_bytes,
                            uint64_t nbytes, int append);

/* Sub-integration boundaries.  Replaces the arithmetic of dsp::TimeDivide::set_bounds / set_boundaries
 * (Signal/Pulsar/TimeDivide.C:132-330,349-425) for divisions given in seconds (dspsr -L), as driven by
 * dsp::Subint<Fold>::transformation (Signal/Pulsar/dsp/Subint.h:235-305): each input block is cut at the
 * division boundaries k*L measured from the observation start; the caller folds [idat_start, idat_start+ndat)
 * (Fold::Engine::set_ndat) and, when end_reached, unloads and zeroes the PhaseSeries.  Times are seconds
 * since the observation start. */
typedef struct {
  double division_seconds;
  double lower, upper;       /* boundaries of the current division */
  double current_end;        /* end of the data folded so far */
  int is_valid;
  uint64_t division;         /* index of the current division */
} b200_time_divide;

typedef struct {
  int is_valid;              /* 0: the block ends before the current division starts */
  int new_division;          /* a new division was started by this call */
  int end_reached;           /* the block reaches the end of the division: unload + zero */
  int in_next;               /* the block extends beyond the division: call set_bounds again */
  uint64_t idat_start, ndat; /* slice of the block that belongs to the division */
  uint64_t division;
} b200_time_bounds;

int b200_time_divide_init(b200_time_divide* td, double division_seconds);
int b200_time_divide_set_bounds(b200_time_divide* td, double input_start, double rate, uint64_t input_ndat,
                                b200_time_bounds* out);

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
