// oracle/dspsr_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of dspsr's baseband hot path (unpack -> overlap-save coherent
// dedispersion / filterbank -> detection -> fold), written from the behaviour of the
// reference (demorest/dspsr) with every function citing the reference file:line it
// follows.  It exists to CHECK the CUDA product (tests/, __graft_entry__.smoke(),
// bench.py's cpu_baseline / --impl reference legs).  Nothing in dspsr_b200/ links,
// imports or executes it.
//
// PARITY STATUS: PINNED to reference code, stage by stage, by tests/test_ref_pin.py -- oracle/ref.mk compiles the
// reference's sources where they lie under /root/reference into oracle/_ref/*.so (whole files: BitTable.C, the format
// and 8-bit / 2-bit unpackers, Dedispersion.C, Response.C, Shape.C, optimize_fft.c, cross_detect.c, stokes_detect.c,
// filterbank_header.c, ascii_header.c; statement blocks cut from the text: the overlap-save loop nests of
// Filterbank.C:563-660 and Convolution.C:389-458, the loops of Fold.C:687-716,744-787,835-873, the bodies of
// WeightedTimeSeries::convolve_weights / scrunch_weights) and every function below that restates one of them must
// agree with it bit for bit.  The reference holds no golden vectors for this path (its test_*.C are file-driven
// smoke/timing drivers) and cannot be built as a whole here (needs PSRCHIVE, autotools).
// PARITY UNPINNED for the third-party arithmetic that lives in PSRCHIVE (version un-pinned by configure.ac:73, source
// not in the tree); it is restated from its published definition:
//   * FTransform frc1d/fcc1d/bcc1d  -> orc_fft.cpp (FFTW conventions, unnormalised)
//   * JenetAnderson98::get_optimal_spacing -> Table of optimal n-bit thresholds (JA98)
//   * NormalDistribution::cumulative_distribution -> 0.5*(1+erf(x/sqrt 2))
//   * Pulsar::Predictor (TEMPO polyco) phase/frequency -> TEMPO polyco definition (checked on Benchmark/vela.polyco)
//   * MJD -> (day, second, fraction) triple
// and for the size bookkeeping of Filterbank.C:68-155 / Convolution.C:105-221 and TimeDivide.C (restated; values of
// SURVEY App. B asserted).
// Compiled with -ffp-contract=off so double-precision results are reproducible.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" {
void orc_fft_fcc1d(unsigned n, float* out, const float* in);
void orc_fft_bcc1d(unsigned n, float* out, const float* in);
void orc_fft_frc1d(unsigned n, float* out, const float* in);
}

namespace {
template <typename T> inline T sqr(T x) { return x * x; }
}

extern "C" {

// ---------------------------------------------------------------------------------------
// a1  BitTable  (Kernel/Classes/BitTable.C:121-218)
// ---------------------------------------------------------------------------------------

// JenetAnderson98::get_optimal_spacing (PSRCHIVE; called BitTable.C:171): optimal input
// threshold spacing, in units of sigma, for an n-bit uniform quantiser
// (Jenet & Anderson 1998, PASP 110, 1467, Table 3).  UNPINNED: PSRCHIVE source absent.
double orc_ja98_optimal_spacing(unsigned nbit) {
  switch (nbit) {
    case 2: return 0.9674;
    case 3: return 0.5605;
    case 4: return 0.3188;
    case 5: return 0.1789;
    case 6: return 0.09925;
    case 7: return 0.05445;
    case 8: return 0.02957;
    default: return 0.0;
  }
}

// NormalDistribution::cumulative_distribution (PSRCHIVE; called BitTable.C:182)
double orc_normal_cdf(double x) { return 0.5 * (1.0 + std::erf(x / std::sqrt(2.0))); }

// BitTable::generate_unique_values (BitTable.C:165-218).  Returns BitTable::get_scale().
double orc_bittable_unique_values(unsigned nbit, int twos_complement, float* values) {
  const unsigned unique_values = 1u << nbit;
  double output_spacing = 1.0 / double(unique_values);
  double output_middle = double(unique_values - 1) / 2.0;
  unsigned input_middle = unique_values / 2;
  double input_spacing = orc_ja98_optimal_spacing(nbit);
  unsigned input_offset = twos_complement ? unique_values / 2 : 0;
  double cumulative_probability = 0.0;
  double variance = 0.0;
  for (unsigned i = 0; i < unique_values; i++) {
    double output = (double(i) - output_middle) * output_spacing;
    values[(i + input_offset) % unique_values] = output;
    if (i < input_middle) {
      double threshold = double(int(i + 1) - int(input_middle)) * input_spacing;
      double cumulative = orc_normal_cdf(threshold);
      double interval = cumulative - cumulative_probability;
      cumulative_probability = cumulative;
      variance += output * output * interval;
    }
  }
  variance *= 2.0;
  double scale = 1.0 / std::sqrt(variance);
  for (unsigned i = 0; i < unique_values; i++) values[i] *= scale;   // float *= double
  scale *= output_spacing;
  return scale;
}

// BitTable::generate for nbit=8 (one value per byte; BitTable.C:121-145)
double orc_bittable8(int twos_complement, float* lut256) {
  return orc_bittable_unique_values(8, twos_complement, lut256);
}

// ---------------------------------------------------------------------------------------
// a2-a5  unpackers.  Output layout is dsp::TimeSeries FPT order (DataSeries.C:246-259):
//   plane(ichan, ipol) = out + (ichan*npol + ipol) * span,  element [idat*ndim + idim]
// ---------------------------------------------------------------------------------------

// a2 CASPSRUnpacker::unpack (Kernel/Formats/caspsr/CASPSRUnpacker.C:132-187)
void orc_unpack_caspsr(const uint8_t* raw, uint64_t ndat, const float* lut, float* out, uint64_t span) {
  for (unsigned ipol = 0; ipol < 2; ipol++) {
    const uint8_t* from = raw + 4 * ipol;
    float* into = out + ipol * span;
    for (uint64_t idat = 0; idat < ndat; idat += 4) {
      into[0] = lut[from[0]];
      into[1] = lut[from[1]];
      into[2] = lut[from[2]];
      into[3] = lut[from[3]];
      from += 8;
      into += 4;
    }
  }
}

// a3 BitUnpacker::unpack + EightBitUnpacker::unpack (BitUnpacker.C:48-80, EightBitUnpacker.C:25-49)
void orc_unpack_generic8(const uint8_t* raw, uint64_t ndat, unsigned nchan, unsigned npol, unsigned ndim,
                         const float* lut, float* out, uint64_t span, uint64_t* hist /* nullable: [ndig][256] */) {
  const unsigned nskip = npol * nchan * ndim;
  const unsigned fskip = ndim;
  unsigned offset = 0;
  for (unsigned ichan = 0; ichan < nchan; ichan++)
    for (unsigned ipol = 0; ipol < npol; ipol++)
      for (unsigned idim = 0; idim < ndim; idim++) {
        const uint8_t* from = raw + offset;
        float* into = out + (uint64_t(ichan) * npol + ipol) * span + idim;
        uint64_t* h = hist ? hist + uint64_t(offset) * 256 : nullptr;
        for (uint64_t idat = 0; idat < ndat; idat++) {
          if (h) h[*from]++;
          *into = lut[*from];
          from += nskip;
          into += fskip;
        }
        offset++;
      }
}

// a4 MeerKATUnpacker::unpack, OrderFPT branch (Kernel/Formats/kat/MeerKATUnpacker.C:196-229)
void orc_unpack_meerkat(const int8_t* raw, uint64_t ndat, unsigned nchan, unsigned npol, float scale,
                        unsigned sample_swap, float* out, uint64_t span) {
  const int16_t* from = reinterpret_cast<const int16_t*>(raw);
  const unsigned ndim = 2;
  const unsigned nsamp_per_heap = 256;
  const uint64_t nheap = ndat / nsamp_per_heap;
  for (uint64_t iheap = 0; iheap < nheap; iheap++)
    for (unsigned ipol = 0; ipol < npol; ipol++)
      for (unsigned ichan = 0; ichan < nchan; ichan++) {
        float* into = out + (uint64_t(ichan) * npol + ipol) * span + iheap * nsamp_per_heap * ndim;
        for (unsigned isamp = 0; isamp < nsamp_per_heap; isamp += sample_swap)
          for (unsigned iswap = 0; iswap < sample_swap; iswap++) {
            int16_t from16 = from[isamp + (sample_swap - 1 - iswap)];
            int8_t from8[2];
            std::memcpy(from8, &from16, 2);
            into[2 * (isamp + iswap) + 0] = (float(from8[0]) + 0.5) * scale;
            into[2 * (isamp + iswap) + 1] = (float(from8[1]) + 0.5) * scale;
          }
        from += nsamp_per_heap;
      }
}

// a5 UWBUnpacker::unpack (Kernel/Formats/uwb/UWBUnpacker.C:177-218)
void orc_unpack_uwb(const int16_t* raw, uint64_t ndat, unsigned npol, float* out, uint64_t span) {
  const unsigned ndim = 2;
  const unsigned nsamp_block = 2048;
  const uint64_t nblock = ndat / nsamp_block;
  const unsigned into_stride = nsamp_block * ndim;
  const unsigned from_pol_stride = into_stride;
  const unsigned from_stride = from_pol_stride * npol;
  for (unsigned ipol = 0; ipol < npol; ipol++) {
    const int16_t* from = raw + ipol * from_pol_stride;
    float* into = out + ipol * span;
    for (uint64_t iblock = 0; iblock < nblock; iblock++) {
      for (unsigned isamp = 0; isamp < nsamp_block * ndim; isamp += 2) {
        int16_t re = from[isamp + 0] ^ 0x8000;
        into[isamp + 0] = float(re);
        int16_t im = from[isamp + 1] ^ 0x8000;
        into[isamp + 1] = float(im);
      }
      into += into_stride;
      from += from_stride;
    }
  }
}

// ---------------------------------------------------------------------------------------
// a6  two-bit excision unpacker: TwoBitCorrection::build / dig_unpack (Kernel/Classes/
//     TwoBitCorrection.C:89-151), ExcisionUnpacker::set_limits / unpack (ExcisionUnpacker.C:95-158,
//     174-256), excision_unpack (dsp/excision_unpack.h:21-106), TwoBitFour::prepare / unpack /
//     nlow_build (dsp/TwoBitFour.h:42-89, TwoBitFour.C:25-40), TwoBitLookup::lookup_build
//     (TwoBitLookup.C:63-98), TwoBitTable::generate_unique_values (TwoBitTable.C:42-75), BitTable::
//     generate MostToLeast (BitTable.C:121-163).  CPSR2 convention: OffsetBinary, polarisations
//     interleaved byte by byte (cpsr2/CPSR2TwoBitCorrection.C:11-23, ExcisionUnpacker.C:258-277).
//
//     JenetAnderson98 lives in PSRCHIVE (not in the reference tree); restated from Jenet & Anderson
//     (1998, PASP 110, 1467) section 6 with sigma = 1:
//       alpha = ierf(Phi)                                          (Eq. 45)
//       lo^2  = 1 - (2 alpha / sqrt(pi)) exp(-alpha^2) / Phi        (Eq. 41: <x^2 | |x| < threshold>)
//       hi^2  = 1 + (2 alpha / sqrt(pi)) exp(-alpha^2) / (1 - Phi)  (Eq. 40: <x^2 | |x| > threshold>)
//     i.e. the power-preserving levels of the dynamic level-setting scheme (NOT the Lloyd-Max
//     conditional means 0.4528 / 1.510); PARITY UNPINNED against PSRCHIVE's own JenetAnderson98.C.
//       mean_Phi = erf(threshold / sqrt 2), var_Phi = mean_Phi (1 - mean_Phi); optimal threshold 0.9674
// ---------------------------------------------------------------------------------------
static double orc_ierf(double y) {
  // inverse error function: Newton iterations on erf from a rational starting point
  if (y <= -1.0) return -INFINITY;
  if (y >= 1.0) return INFINITY;
  const double a = 0.147;
  const double ln1 = std::log(1.0 - y * y);
  const double t = 2.0 / (M_PI * a) + 0.5 * ln1;
  double x = std::copysign(std::sqrt(std::sqrt(t * t - ln1 / a) - t), y);
  for (int i = 0; i < 4; i++) x -= (std::erf(x) - y) / (2.0 / std::sqrt(M_PI) * std::exp(-x * x));
  return x;
}

void orc_ja98_levels(double Phi, double* lo, double* hi) {
  const double root_pi = std::sqrt(M_PI);
  const double alpha = orc_ierf(Phi);
  const double expon = std::exp(-alpha * alpha);
  *lo = std::sqrt(1.0 - (2.0 * alpha / root_pi) * (expon / Phi));
  *hi = std::sqrt(1.0 + (2.0 * alpha / root_pi) * (expon / (1.0 - Phi)));
}

// ExcisionUnpacker::set_limits (ExcisionUnpacker.C:95-158)
void orc_twobit_limits(double threshold, float cutoff_sigma, unsigned ndat_per_weight, unsigned* nlow_min,
                       unsigned* nlow_max) {
  if (cutoff_sigma == 0.0) {
    *nlow_min = 0;
    *nlow_max = ndat_per_weight;
    return;
  }
  const double mean_Phi = std::erf(threshold / std::sqrt(2.0));
  const double var_Phi = mean_Phi * (1.0 - mean_Phi);
  float fsample = ndat_per_weight;
  float nlo_mean = fsample * mean_Phi;
  float nlo_variance = fsample * var_Phi;
  float nlo_sigma = sqrt(nlo_variance);
  *nlow_max = unsigned(nlo_mean + (cutoff_sigma * nlo_sigma));
  if (*nlow_max >= ndat_per_weight) *nlow_max = ndat_per_weight - 1;
  if (cutoff_sigma * nlo_sigma >= nlo_mean + 1.0) *nlow_min = 1;
  else *nlow_min = unsigned(nlo_mean - (cutoff_sigma * nlo_sigma));
}

// TwoBitTable::generate_unique_values (TwoBitTable.C:42-75); type 0 OffsetBinary, 1 SignMagnitude,
// 2 TwosComplement
static void twobit_unique_values(int type, float lo_val, float hi_val, float* vals) {
  switch (type) {
    case 0: vals[0] = -hi_val; vals[1] = -lo_val; vals[2] = lo_val; vals[3] = hi_val; break;
    case 1: vals[0] = lo_val; vals[1] = hi_val; vals[2] = -lo_val; vals[3] = -hi_val; break;
    default: vals[0] = lo_val; vals[1] = hi_val; vals[2] = -hi_val; vals[3] = -lo_val; break;
  }
}

// BitTable::generate for nbit = 2, MostToLeast (BitTable.C:121-163): 4 floats per unique byte
static void twobit_generate(int type, float lo_val, float hi_val, float* table /* 256*4 */) {
  float vals[4];
  twobit_unique_values(type, lo_val, hi_val, vals);
  for (unsigned byte = 0; byte < 256; byte++)
    for (unsigned samp = 0; samp < 4; samp++) table[byte * 4 + samp] = vals[(byte >> (6 - 2 * samp)) & 3];
}

struct orc_twobit {
  int table_type;
  unsigned ndat_per_weight, nlow_min, nlow_max;
  std::vector<float>* lookup;     // (nlow_max-nlow_min+1) blocks of 256*4 floats (TwoBitLookup::lookup_build)
  char nlow_lookup[256];          // TwoBitFour::nlow_build
};

orc_twobit* orc_twobit_create(int table_type, double threshold, float cutoff_sigma, unsigned ndat_per_weight) {
  orc_twobit* t = new orc_twobit();
  t->table_type = table_type;
  t->ndat_per_weight = ndat_per_weight;
  orc_twobit_limits(threshold, cutoff_sigma, ndat_per_weight, &t->nlow_min, &t->nlow_max);
  t->lookup = new std::vector<float>(size_t(t->nlow_max - t->nlow_min + 1) * 1024);
  float* lookup = t->lookup->data();
  for (unsigned nlo = t->nlow_min; nlo <= t->nlow_max; nlo++) {   // TwoBitLookup.C:75-97
    unsigned use_nlow = nlo;
    if (nlo == 0) use_nlow = 1;
    // TwoBitLookup.C:83-84 writes `if (nlow == ndat)` (the member, 0 at build time) where it means `nlo`: with
    // cutoff_sigma = 0 the reference's own row nlo = ndat is NaN (checked against the compiled reference in
    // tests/test_ref_pin.py).  Oracle and product clamp as the reference intends; documented deviation.
    if (nlo == ndat_per_weight) use_nlow = ndat_per_weight - 1;
    float p_in = (float)use_nlow / (float)ndat_per_weight;
    double lo, hi;
    orc_ja98_levels(p_in, &lo, &hi);
    twobit_generate(table_type, float(lo), float(hi), lookup);
    lookup += 1024;
  }
  // TwoBitFour::nlow_build (TwoBitFour.C:25-40): count the low-voltage states of every byte
  float fv[1024];
  twobit_generate(table_type, 1.0f, 0.75f, fv);
  for (unsigned byte = 0; byte < 256; byte++) {
    t->nlow_lookup[byte] = 0;
    for (unsigned i = 0; i < 4; i++)
      if (fv[byte * 4 + i] * fv[byte * 4 + i] == 1.0f) t->nlow_lookup[byte]++;
  }
  return t;
}
void orc_twobit_destroy(orc_twobit* t) {
  if (t) {
    delete t->lookup;
    delete t;
  }
}
void orc_twobit_info(const orc_twobit* t, unsigned* nlow_min, unsigned* nlow_max) {
  *nlow_min = t->nlow_min;
  *nlow_max = t->nlow_max;
}
// lo / hi of table row nlow (for the product's compact table)
void orc_twobit_levels(const orc_twobit* t, unsigned nlow, float* lo, float* hi) {
  const float* blk = t->lookup->data() + size_t(nlow - t->nlow_min) * 1024;
  float a = std::fabs(blk[0]), b = a;   // byte 0 and the other magnitude
  for (unsigned i = 0; i < 1024; i++) {
    a = std::min(a, std::fabs(blk[i]));
    b = std::max(b, std::fabs(blk[i]));
  }
  *lo = a;
  *hi = b;
}

// ExcisionUnpacker::unpack for real-sampled (ndim 1) data, one digitizer per polarisation:
// out plane p = out + p*span, weights[p*nweights + w] start at 1 and are masked (AND over
// polarisations, WeightedTimeSeries::mask_weights) at the end.
void orc_unpack_twobit(const orc_twobit* t, const uint8_t* raw, uint64_t ndat, unsigned npol, float* out,
                       uint64_t span, unsigned* weights, uint64_t nweights) {
  const unsigned ndat_per_weight = t->ndat_per_weight;
  const uint64_t n_weights = ndat / ndat_per_weight;
  for (uint64_t i = 0; i < nweights * npol; i++) weights[i] = 1;
  for (unsigned idig = 0; idig < npol; idig++) {
    const uint8_t* input = raw + idig;            // get_input_offset = idig, get_input_incr = npol
    float* output = out + uint64_t(idig) * span;
    unsigned* w = weights + uint64_t(idig) * nweights;
    for (uint64_t wt = 0; wt < n_weights; wt++) {
      // TwoBitFour::prepare
      const unsigned nbyte = ndat_per_weight / 4;
      unsigned total = 0, nlow = 0;
      for (unsigned bt = 0; bt < nbyte; bt++) {
        nlow += t->nlow_lookup[input[bt * npol]];
        total += input[bt * npol];
      }
      const bool bad = (total == 0);
      // TwoBitFour::unpack
      const unsigned n_low = nlow;
      if (nlow < t->nlow_min) nlow = t->nlow_min;
      else if (nlow > t->nlow_max) nlow = t->nlow_max;
      const float* lookup = t->lookup->data() + size_t(nlow - t->nlow_min) * 1024;
      for (unsigned bt = 0; bt < nbyte; bt++) {
        const float* fourval = lookup + input[bt * npol] * 4;
        for (unsigned pt = 0; pt < 4; pt++) output[bt * 4 + pt] = fourval[pt];
      }
      input += uint64_t(nbyte) * npol;
      // excision_unpack.h:79-97
      if (bad || n_low < t->nlow_min || n_low > t->nlow_max || w[wt] == 0) {
        w[wt] = 0;
        for (unsigned i = 0; i < ndat_per_weight; i++) output[i] = 0.0;
      }
      output += ndat_per_weight;
    }
  }
  // WeightedTimeSeries::mask_weights: a window flagged in any polarisation is flagged in all
  for (uint64_t wt = 0; wt < n_weights; wt++) {
    unsigned all = 1;
    for (unsigned p = 0; p < npol; p++) all &= weights[p * nweights + wt] != 0;
    for (unsigned p = 0; p < npol; p++) weights[p * nweights + wt] = all;
  }
}

// ---------------------------------------------------------------------------------------
// a9  optimal_fft_length (Signal/General/optimize_fft.c:63-127)
// ---------------------------------------------------------------------------------------
int64_t orc_optimal_fft_length(uint64_t nbadperfft, uint64_t nfft_max) {
  if (!nbadperfft) return -1;
  uint64_t nfft_min = (uint64_t)std::pow(2.0, std::ceil(std::log((double)nbadperfft) / std::log(2.0)));
  if (nfft_max && nfft_max < nfft_min) return -1;
  uint64_t nfft = nfft_min;
  double order_fft = (double)nfft * std::log((double)nfft);
  double timescale = order_fft / (double)(nfft - nbadperfft);
  double prev_timescale;
  while (nfft_max == 0 || nfft * 2 < nfft_max) {
    prev_timescale = timescale;
    nfft *= 2;
    order_fft = (double)nfft * std::log((double)nfft);
    timescale = order_fft / (double)(nfft - nbadperfft);
    if (timescale > prev_timescale) {
      nfft /= 2;
      break;
    }
  }
  return (int64_t)nfft;
}

// ---------------------------------------------------------------------------------------
// a7/a8  Dedispersion + Response::match
// ---------------------------------------------------------------------------------------
struct orc_dedisp {
  // inputs
  double centre_frequency;       // MHz
  double bandwidth;              // MHz (sign = band sense)
  double dispersion_measure;
  double doppler_shift;          // 1.0
  unsigned input_nchan;          // channels of the input Observation
  unsigned nchan;                // channels of the response (= output channels)
  int input_dual_sideband;       // Observation::get_dual_sideband (Observation.C:80-87)
  int input_dc_centred;
  int input_swap;
  unsigned frequency_resolution; // 0 = choose optimal (-x nfft override otherwise)
  // outputs of prepare
  unsigned impulse_pos, impulse_neg, ndat;
  unsigned unsupported_channels;
};

static const double dm_dispersion = 2.41e-4;       // Dedispersion.C:28
static const double smearing_buffer = 0.1;         // Dedispersion.C:30
static const unsigned smearing_samples_threshold = 16 * 1024 * 1024;   // Dedispersion.C:214

// Dedispersion::delay_time (Dedispersion.C:348-356)
static double delay_time(double dm, double freq1, double freq2) {
  double dispersion = dm / dm_dispersion;
  return dispersion * (1.0 / sqr(freq1) - 1.0 / sqr(freq2));
}

// Dedispersion::smearing_time(int half) (Dedispersion.C:383-430)
static double smearing_time(const orc_dedisp* d, int half, unsigned nunsupported) {
  double abs_bw = std::fabs(d->bandwidth);
  double ch_abs_bw = abs_bw / double(d->nchan);
  double lower_ch_cfreq = d->centre_frequency - (abs_bw - ch_abs_bw) / 2.0;
  for (unsigned ichan = 0; ichan < nunsupported; ichan++) lower_ch_cfreq += ch_abs_bw;
  if (half) {
    ch_abs_bw /= 2.0;
    lower_ch_cfreq += double(half) * ch_abs_bw;
  }
  // smearing_time (cfreq, bw) (Dedispersion.C:343-346)
  return delay_time(d->dispersion_measure, lower_ch_cfreq - std::fabs(0.5 * ch_abs_bw),
                    lower_ch_cfreq + std::fabs(0.5 * ch_abs_bw));
}

// Dedispersion::smearing_samples (Dedispersion.C:432-475)
static unsigned smearing_samples(const orc_dedisp* d, int half, unsigned nunsupported) {
  double tsmear = smearing_time(d, half, nunsupported);
  double ch_abs_bw = std::fabs(d->bandwidth) / double(d->nchan);
  double sampling_rate = ch_abs_bw * 1e6;
  tsmear *= (1.0 + smearing_buffer);
  return unsigned(std::ceil(tsmear * sampling_rate));
}

// Response::get_minimum_ndat (Response.C:259-275)
static unsigned minimum_ndat(unsigned impulse_pos, unsigned impulse_neg) {
  double impulse_tot = impulse_pos + impulse_neg;
  if (impulse_tot == 0) return 0;
  unsigned min = unsigned(std::pow(2.0, std::ceil(std::log(impulse_tot) / std::log(2.0))));
  while (min <= impulse_tot) min *= 2;
  return min;
}

// Dedispersion::prepare (Dedispersion.C:216-248) + the ndat choice of Dedispersion::build
// (Dedispersion.C:296-308) / Response::set_optimal_ndat (Response.C:282-311).
// Returns 0 on success, <0 mirroring the reference's Error throws.
int orc_dedisp_prepare(orc_dedisp* d) {
  unsigned threshold = smearing_samples_threshold / d->nchan;
  unsigned ichan = 0;
  while ((d->impulse_neg = smearing_samples(d, -1, ichan)) > threshold) {
    ichan++;
    if (ichan == d->nchan) return -1;   // "smearing samples=%u exceeds threshold=%u"
  }
  d->unsupported_channels = ichan;
  d->impulse_pos = smearing_samples(d, 1, ichan);
  if (d->frequency_resolution) {
    d->ndat = d->frequency_resolution;
    if (d->ndat < minimum_ndat(d->impulse_pos, d->impulse_neg)) return -2;   // Response::check_ndat
  } else {
    int64_t n = orc_optimal_fft_length(uint64_t(d->impulse_pos) + d->impulse_neg, 0);
    if (n < 0) return -3;
    d->ndat = unsigned(n);
  }
  return 0;
}

// Response::doswap (Response.C:649-700), on an interleaved complex buffer of nchan*ndat points
static void doswap(float* buffer, unsigned nchan, unsigned ndat, unsigned divisions) {
  const unsigned ndim = 2;
  unsigned half_npts = (ndat * ndim * nchan) / (2 * divisions);
  float* ptr1 = buffer;
  float* ptr2 = ptr1 + half_npts;
  for (unsigned idiv = 0; idiv < divisions; idiv++) {
    for (unsigned ipt = 0; ipt < half_npts; ipt++) {
      float temp = *ptr1;
      *ptr1 = *ptr2; ptr1++;
      *ptr2 = temp; ptr2++;
    }
    ptr1 += half_npts;
    ptr2 += half_npts;
  }
}

// Dedispersion::build (Dedispersion.C:291-331,478-556), then Response::match
// (Response.C:132-181) and the DC zap of Dedispersion::match (Dedispersion.C:278).
// H: nchan*ndat complex floats (interleaved).
int orc_dedisp_build(const orc_dedisp* d, float* H) {
  const unsigned _ndat = d->ndat, _nchan = d->nchan;
  std::vector<float> phases(uint64_t(_ndat) * _nchan);
  const bool dc_centred = d->input_dc_centred;   // Dedispersion::prepare copies it (Dedispersion.C:179)

  double centrefreq = d->centre_frequency / d->doppler_shift;
  double bw = d->bandwidth / d->doppler_shift;
  double sign = bw / std::fabs(bw);
  double chanwidth = bw / double(_nchan);
  double binwidth = chanwidth / double(_ndat);
  double lower_cfreq = centrefreq - 0.5 * bw;
  if (!dc_centred) lower_cfreq += 0.5 * chanwidth;
  double dispersion_per_MHz = 1e6 * d->dispersion_measure / dm_dispersion;

  for (unsigned ichan = 0; ichan < _nchan; ichan++) {
    double chan_cfreq = lower_cfreq + double(ichan) * chanwidth;
    double coeff = -sign * 2 * M_PI * dispersion_per_MHz / sqr(chan_cfreq);
    unsigned spt = ichan * _ndat;
    for (unsigned ipt = 0; ipt < _ndat; ipt++) {
      double freq = double(ipt) * binwidth - 0.5 * chanwidth;
      double delay_phase = -2.0 * M_PI * freq * 0.0;   // fractional_delay off (delay = 0)
      phases[spt + ipt] = coeff * sqr(freq) / (chan_cfreq + freq) + delay_phase;
    }
  }

  std::complex<float>* phasors = reinterpret_cast<std::complex<float>*>(H);
  uint64_t npt = uint64_t(_ndat) * _nchan;
  for (uint64_t ipt = 0; ipt < npt; ipt++) phasors[ipt] = std::polar(float(1.0), phases[ipt]);
  phasors[0] = 0;   // always zap DC channel (Dedispersion.C:323)

  // Response::match (Response.C:132-181).  The dc_centred rotation branch (:152-161) is
  // unreachable here because Dedispersion::prepare already copied input.dc_centred.
  if (d->input_nchan == 1) {
    if (d->input_dual_sideband) doswap(H, _nchan, _ndat, 1);
  } else {
    if (d->input_dual_sideband) doswap(H, _nchan, _ndat, d->input_nchan);
    if (d->input_swap) doswap(H, _nchan, _ndat, 1);
  }
  H[0] = H[1] = 0.0;   // Dedispersion::match (Dedispersion.C:278)
  return 0;
}

// ---------------------------------------------------------------------------------------
// a8  Response::operate (Response.C:385-444): spectrum *= H, float arithmetic
// ---------------------------------------------------------------------------------------
static void response_operate(float* spectrum, const float* f_p, uint64_t npts) {
  for (uint64_t ipt = 0; ipt < npts; ipt++) {
    float d_r = spectrum[0], d_i = spectrum[1];
    float f_r = f_p[0], f_i = f_p[1];
    spectrum[0] = f_r * d_r - f_i * d_i;
    spectrum[1] = f_i * d_r + f_r * d_i;
    spectrum += 2;
    f_p += 2;
  }
}

// the same loop behind a C door, for tests/test_ref_pin.py (checked against the reference's Response::operate)
void orc_response_operate(float* spectrum, const float* H, uint64_t npts) { response_operate(spectrum, H, npts); }

// ---------------------------------------------------------------------------------------
// a10  Filterbank (Signal/General/Filterbank.C:55-263 sizes, :389-430 npart, :563-660 loop)
// ---------------------------------------------------------------------------------------
struct orc_fb {
  int input_real;                // Signal::Nyquist (1) or Signal::Analytic (0)
  unsigned input_nchan, npol;
  unsigned nchan;                // output channels
  unsigned freq_res;
  unsigned nfilt_pos, nfilt_neg;
  // derived (filled by orc_fb_sizes)
  unsigned nchan_subband, n_fft, nsamp_fft, nsamp_overlap, nsamp_step, nkeep;
};

void orc_fb_sizes(orc_fb* f) {
  f->nchan_subband = f->nchan / f->input_nchan;                 // Filterbank.C:68
  f->n_fft = f->nchan_subband * f->freq_res;                    // :107
  unsigned nfilt_tot = f->nfilt_pos + f->nfilt_neg;             // :131
  if (f->input_real) {                                          // :139-148
    f->nsamp_fft = 2 * f->n_fft;
    f->nsamp_overlap = 2 * nfilt_tot * f->nchan_subband;
  } else {
    f->nsamp_fft = f->n_fft;
    f->nsamp_overlap = nfilt_tot * f->nchan_subband;
  }
  f->nsamp_step = f->nsamp_fft - f->nsamp_overlap;              // :155
  f->nkeep = f->freq_res - nfilt_tot;                           // :409
}

// Filterbank::resize_output npart (Filterbank.C:401-402)
uint64_t orc_fb_npart(const orc_fb* f, uint64_t ndat) {
  if (ndat > f->nsamp_overlap) return (ndat - f->nsamp_overlap) / f->nsamp_step;
  return 0;
}

// Filterbank::filterbank CPU branch (Filterbank.C:563-660).
// in:  planes (ichan*npol+ipol)*in_span, ndim = input_real ? 1 : 2
// out: planes (ochan*npol+ipol)*out_span, complex, npart*nkeep samples
// H:   nchan*freq_res complex (nullable)
// ipart0/npart: range of parts to compute (for multi-threaded callers); output offsets are
// absolute (ipart*out_step), as in the reference.
void orc_filterbank_parts(const orc_fb* f, const float* in, uint64_t in_span, const float* H,
                          float* out, uint64_t out_span, uint64_t ipart0, uint64_t npart) {
  const unsigned ndim = f->input_real ? 1 : 2;
  const uint64_t in_step = uint64_t(f->nsamp_step) * ndim;      // :517
  const uint64_t out_step = uint64_t(f->nkeep) * 2;             // :523
  unsigned bigfftsize = f->nchan_subband * f->freq_res * 2;     // :480
  if (f->input_real) bigfftsize += 256;
  std::vector<float> c_spectrum(bigfftsize);
  std::vector<float> c_time(2 * f->freq_res);
  for (unsigned input_ichan = 0; input_ichan < f->input_nchan; input_ichan++)
    for (uint64_t ipart = ipart0; ipart < ipart0 + npart; ipart++) {
      uint64_t in_offset = ipart * in_step;
      uint64_t out_offset = ipart * out_step;
      for (unsigned ipol = 0; ipol < f->npol; ipol++) {
        const float* time_dom_ptr = in + (uint64_t(input_ichan) * f->npol + ipol) * in_span + in_offset;
        if (f->input_real) orc_fft_frc1d(f->nsamp_fft, c_spectrum.data(), time_dom_ptr);   // :591
        else orc_fft_fcc1d(f->nsamp_fft, c_spectrum.data(), time_dom_ptr);                 // :593
        if (H)                                                                              // :611-613
          response_operate(c_spectrum.data(),
                           H + uint64_t(input_ichan) * f->nchan_subband * f->freq_res * 2,
                           uint64_t(f->nchan_subband) * f->freq_res);
        unsigned jchan = input_ichan * f->nchan_subband;
        if (f->freq_res == 1) {                                                             // :621-631
          for (unsigned ichan = 0; ichan < f->nchan_subband; ichan++) {
            float* data_into = out + (uint64_t(jchan + ichan) * f->npol + ipol) * out_span + out_offset;
            data_into[0] = c_spectrum[2 * ichan];
            data_into[1] = c_spectrum[2 * ichan + 1];
          }
          continue;
        }
        const float* freq_dom_ptr = c_spectrum.data();
        for (unsigned ichan = 0; ichan < f->nchan_subband; ichan++) {                       // :640-652
          orc_fft_bcc1d(f->freq_res, c_time.data(), freq_dom_ptr);
          freq_dom_ptr += f->freq_res * 2;
          float* data_into = out + (uint64_t(jchan + ichan) * f->npol + ipol) * out_span + out_offset;
          const float* data_from = c_time.data() + f->nfilt_pos * 2;
          std::memcpy(data_into, data_from, sizeof(float) * 2 * f->nkeep);
        }
      }
    }
}

// ---------------------------------------------------------------------------------------
// a11  Convolution (Signal/General/Convolution.C:105-221 sizes, :389-458 loop)
// ---------------------------------------------------------------------------------------
struct orc_conv {
  int input_real;
  unsigned nchan, npol;
  unsigned n_fft;                // response ndat
  unsigned nfilt_pos, nfilt_neg;
  unsigned nsamp_fft, nsamp_overlap, nsamp_step;   // derived
};

void orc_conv_sizes(orc_conv* c) {
  unsigned nfilt_tot = c->nfilt_pos + c->nfilt_neg;
  if (c->input_real) {                                          // Convolution.C:170-175
    c->nsamp_fft = c->n_fft * 2;
    c->nsamp_overlap = nfilt_tot * 2;
  } else {                                                      // :176-180
    c->nsamp_fft = c->n_fft;
    c->nsamp_overlap = nfilt_tot;
  }
  c->nsamp_step = c->nsamp_fft - c->nsamp_overlap;              // :229
}

// Convolution::prepare_output npart (Convolution.C:236-238)
uint64_t orc_conv_npart(const orc_conv* c, uint64_t ndat) {
  if (ndat >= c->nsamp_fft) return (ndat - c->nsamp_overlap) / c->nsamp_step;
  return 0;
}

// Convolution::transformation CPU branch (Convolution.C:389-458).
// out planes are complex with npart*(n_fft - nfilt_tot) samples.
void orc_convolution_parts(const orc_conv* c, const float* in, uint64_t in_span, const float* H,
                           float* out, uint64_t out_span, uint64_t ipart0, uint64_t npart) {
  const unsigned ndim = c->input_real ? 1 : 2;
  const uint64_t step = uint64_t(c->nsamp_step) * ndim;         // :387
  const uint64_t nbytes_step = step * sizeof(float);            // :373
  std::vector<float> spectrum(uint64_t(c->n_fft) * 2 + 4);
  std::vector<float> complex_time(uint64_t(c->n_fft) * 2);
  for (unsigned ichan = 0; ichan < c->nchan; ichan++)
    for (unsigned ipol = 0; ipol < c->npol; ipol++)
      for (uint64_t ipart = ipart0; ipart < ipart0 + npart; ipart++) {
        uint64_t offset = ipart * step;
        const float* ptr = in + (uint64_t(ichan) * c->npol + ipol) * in_span + offset;
        if (c->input_real) orc_fft_frc1d(c->nsamp_fft, spectrum.data(), ptr);   // :412
        else orc_fft_fcc1d(c->nsamp_fft, spectrum.data(), ptr);                 // :415
        response_operate(spectrum.data(), H + uint64_t(ichan) * c->n_fft * 2, c->n_fft);   // :430
        orc_fft_bcc1d(c->n_fft, complex_time.data(), spectrum.data());          // :446
        float* optr = out + (uint64_t(ichan) * c->npol + ipol) * out_span + offset;   // :449
        std::memcpy(optr, complex_time.data() + c->nfilt_pos * 2, nbytes_step);       // :455
      }
}

// ---------------------------------------------------------------------------------------
// a12  Detection (Signal/General/Detection.C:218-320,322-421,423-474;
//      cross_detect.ic:25-41, stokes_detect.ic:21-44)
// state: 0 Intensity, 1 PPQQ, 2 Coherence, 3 Stokes   (npol in = 2, Analytic)
// ---------------------------------------------------------------------------------------
void orc_detect(int state, unsigned ndim_out, const float* in, uint64_t in_span, unsigned nchan,
                unsigned npol, uint64_t ndat, float* out, uint64_t out_span) {
  if (state == 0 || state == 1) {
    // Detection::square_law, OrderFPT, Analytic input (Detection.C:240-281)
    const unsigned out_npol = (state == 0) ? 1 : npol;
    for (unsigned ichan = 0; ichan < nchan; ichan++) {
      std::vector<float> tmp(uint64_t(npol) * ndat);
      for (unsigned ipol = 0; ipol < npol; ipol++) {
        const float* in_ptr = in + (uint64_t(ichan) * npol + ipol) * in_span;
        float* out_ptr = tmp.data() + uint64_t(ipol) * ndat;
        for (uint64_t i = 0; i < ndat; i++) {
          out_ptr[i] = in_ptr[0] * in_ptr[0];
          out_ptr[i] += in_ptr[1] * in_ptr[1];
          in_ptr += 2;
        }
      }
      if (state == 0 && npol == 2)                       // pscrunch (Detection.C:283-301)
        for (uint64_t i = 0; i < ndat; i++) tmp[i] += tmp[ndat + i];
      for (unsigned ipol = 0; ipol < out_npol; ipol++)
        std::memcpy(out + (uint64_t(ichan) * out_npol + ipol) * out_span, tmp.data() + uint64_t(ipol) * ndat,
                    sizeof(float) * ndat);
    }
    return;
  }
  // Detection::polarimetry, out of place (Detection.C:385-412) with get_result_pointers (:423-474)
  const unsigned out_npol = 4 / ndim_out;
  for (unsigned ichan = 0; ichan < nchan; ichan++) {
    const float* p = in + (uint64_t(ichan) * 2 + 0) * in_span;
    const float* q = in + (uint64_t(ichan) * 2 + 1) * in_span;
    float* r[4];
    float* base = out + uint64_t(ichan) * out_npol * out_span;
    switch (ndim_out) {
      case 1: r[0] = base; r[1] = base + out_span; r[2] = base + 2 * out_span; r[3] = base + 3 * out_span; break;
      case 2: r[0] = base; r[1] = r[0] + 1; r[2] = base + out_span; r[3] = r[2] + 1; break;
      default: r[0] = base; r[1] = r[0] + 1; r[2] = r[1] + 1; r[3] = r[2] + 1; break;
    }
    const unsigned span = ndim_out;
    for (uint64_t j = 0; j < ndat; j++) {
      float p_r = p[0], p_i = p[1], q_r = q[0], q_i = q[1];
      p += 2; q += 2;
      if (state == 3) {   // stokes_detect.ic
        float pp = p_r * p_r + p_i * p_i;
        float qq = q_r * q_r + q_i * q_i;
        *r[0] = pp + qq;
        *r[1] = pp - qq;
        *r[2] = 2.0 * (p_r * q_r + p_i * q_i);
        *r[3] = 2.0 * (p_r * q_i - p_i * q_r);
      } else {            // cross_detect.ic
        *r[0] = p_r * p_r + p_i * p_i;
        *r[1] = q_r * q_r + q_i * q_i;
        *r[2] = p_r * q_r + p_i * q_i;
        *r[3] = p_r * q_i - p_i * q_r;
      }
      r[0] += span; r[1] += span; r[2] += span; r[3] += span;
    }
  }
}

// ---------------------------------------------------------------------------------------
// a13  Fold (Signal/Pulsar/Fold.C:744-788 plan, :835-873 accumulate)
// ---------------------------------------------------------------------------------------

// The sequential double-precision phase recurrence (Fold.C:765-768); hits[ibin]++ (:783).
// Returns ndat_folded.  phi_out (nullable) receives the phase after the last sample.
uint64_t orc_fold_plan(double phi, double phase_per_sample, unsigned nbin, uint64_t ndat,
                       unsigned* binplan, unsigned* hits, double* phi_out) {
  const double double_nbin = double(nbin);
  uint64_t ndat_folded = 0;
  for (uint64_t idat = 0; idat < ndat; idat++) {
    phi -= std::floor(phi);
    double double_ibin = phi * double_nbin;
    unsigned ibin = unsigned(double_ibin);
    phi += phase_per_sample;
    binplan[idat] = ibin;
    if (hits) hits[ibin]++;
    ndat_folded++;
  }
  if (phi_out) *phi_out = phi;
  return ndat_folded;
}

// The same loop with a WeightedTimeSeries input (Fold.C:687-716 set-up, :746-763 per sample): samples of a
// window whose weight is zero get binplan = nbin (not folded, no hit).  Returns ndat_folded, or
// UINT64_MAX where the reference throws ("iweight >= nweights").
uint64_t orc_fold_plan_weighted(double phi, double phase_per_sample, unsigned nbin, uint64_t idat_start, uint64_t ndat,
                                const unsigned* weights, uint64_t nweights, unsigned ndatperweight,
                                uint64_t weight_idat, unsigned* binplan, unsigned* hits, double* phi_out) {
  if (!ndatperweight || !weights) return orc_fold_plan(phi, phase_per_sample, nbin, ndat, binplan, hits, phi_out);
  const double double_nbin = double(nbin);
  uint64_t ndat_folded = 0;
  uint64_t iweight = (idat_start + weight_idat) / ndatperweight;
  uint64_t idat_nextweight = (iweight + 1) * ndatperweight - weight_idat;
  if (iweight >= nweights) return UINT64_MAX;
  bool bad_data = weights[iweight] == 0;
  for (uint64_t idat = idat_start; idat < idat_start + ndat; idat++) {
    if (idat >= idat_nextweight) {
      iweight++;
      if (iweight >= nweights) return UINT64_MAX;       // assert (iweight < nweights), Fold.C:751
      bad_data = weights[iweight] == 0;
      idat_nextweight += ndatperweight;
    }
    phi -= std::floor(phi);
    double double_ibin = phi * double_nbin;
    unsigned ibin = unsigned(double_ibin);
    phi += phase_per_sample;
    binplan[idat - idat_start] = ibin;
    if (bad_data)
      binplan[idat - idat_start] = nbin;
    else {
      if (hits) hits[ibin]++;
      ndat_folded++;
    }
  }
  if (phi_out) *phi_out = phi;
  return ndat_folded;
}

// WeightedTimeSeries::convolve_weights (Kernel/Classes/WeightedTimeSeries.C:582-690): a transform that
// contains a bad window is flagged as a whole; the flagging of transform i is applied while looking at
// transform i+1 ("flag only the previous data set"), so that it does not influence the next test.
// Returns 0, or -1 where the reference throws (end_weight > nweights).
int orc_convolve_weights(unsigned* weights, uint64_t nweights_tot, unsigned ndat_per_weight, uint64_t weight_idat,
                         uint64_t ndat, unsigned nfft, unsigned nkeep) {
  if (ndat_per_weight >= nfft) return 0;
  if (ndat + nkeep < nfft) return 0;
  if (ndat_per_weight == 0) return 0;
  const double weights_per_dat = 1.0 / ndat_per_weight;
  const uint64_t blocks = (ndat + nkeep - nfft) / nkeep;
  const uint64_t end_idat = blocks * nkeep;
  uint64_t zero_start = 0, zero_end = 0;
  for (uint64_t start_idat = 0; start_idat < end_idat; start_idat += nkeep) {
    const uint64_t wt_idat = start_idat + weight_idat;
    const uint64_t start_weight = uint64_t(wt_idat * weights_per_dat);
    const uint64_t end_weight = uint64_t(std::ceil((wt_idat + nfft) * weights_per_dat));
    if (end_weight > nweights_tot) return -1;
    uint64_t zero_weights = 0;
    for (uint64_t iweight = start_weight; iweight < end_weight; iweight++)
      if (weights[iweight] == 0) zero_weights++;
    for (uint64_t iweight = zero_start; iweight < zero_end; iweight++) weights[iweight] = 0;
    if (zero_weights == 0)
      zero_start = zero_end = 0;
    else {
      zero_start = start_weight;
      zero_end = uint64_t(std::ceil((start_idat + nkeep) * weights_per_dat));
    }
  }
  for (uint64_t iweight = zero_start; iweight < zero_end; iweight++) weights[iweight] = 0;
  return 0;
}

// WeightedTimeSeries::scrunch_weights (WeightedTimeSeries.C:692-780), in place; the three attributes are
// updated like the members of the reference.
void orc_scrunch_weights(unsigned* weights, uint64_t* nweights, unsigned* ndat_per_weight, uint64_t* weight_idat,
                         unsigned nscrunch) {
  const uint64_t nweights_tot = *nweights;
  if (!*ndat_per_weight) return;
  const double points_per_weight = double(*ndat_per_weight) / double(nscrunch);
  if (points_per_weight >= 1.0) {
    *ndat_per_weight = unsigned(points_per_weight);
    const bool leftover = (*weight_idat % *ndat_per_weight != 0);
    *weight_idat /= *ndat_per_weight;
    if (leftover) (*weight_idat)++;
    return;
  }
  uint64_t new_nweights = nweights_tot / nscrunch;
  const uint64_t extra = nweights_tot % nscrunch;
  if (extra) new_nweights++;
  for (uint64_t iwt = 0; iwt < new_nweights; iwt++) {
    unsigned* indi_weight = weights + iwt * nscrunch;
    if ((iwt + 1) * nscrunch > nweights_tot) nscrunch = unsigned(extra);
    for (unsigned ivt = 0; ivt < nscrunch; ivt++) {
      if (*indi_weight == 0) {
        weights[iwt] = 0;
        break;
      } else if (ivt == 0)
        weights[iwt] = *indi_weight;
      else
        weights[iwt] += *indi_weight;
      indi_weight++;
    }
    weights[iwt] /= nscrunch;
  }
  *nweights = new_nweights;
  *ndat_per_weight = 1;
}

// Fold::fold accumulate, OrderFPT (Fold.C:835-873).  profile planes:
// (ichan*npol+ipol)*nbin*ndim, element [ibin*ndim + idim]; float +=, sequential.
void orc_fold(const float* in, uint64_t in_span, unsigned nchan, unsigned npol, unsigned ndim,
              uint64_t idat_start, uint64_t ndat_fold, const unsigned* binplan, unsigned nbin,
              float* profile) {
  for (unsigned ichan = 0; ichan < nchan; ichan++)
    for (unsigned ipol = 0; ipol < npol; ipol++) {
      const float* timep = in + (uint64_t(ichan) * npol + ipol) * in_span + idat_start * ndim;
      float* phasep = profile + (uint64_t(ichan) * npol + ipol) * nbin * ndim;
      for (uint64_t idat = 0; idat < ndat_fold; idat++) {
        if (binplan[idat] != nbin) {
          float* phdimp = phasep + uint64_t(binplan[idat]) * ndim;
          for (unsigned idim = 0; idim < ndim; idim++) phdimp[idim] += timep[idim];
        }
        timep += ndim;
      }
    }
}

// ---------------------------------------------------------------------------------------
// a16  TEMPO polyco predictor (Pulsar::Predictor::phase/frequency; called Fold.C:949,957).
//      Format: SURVEY A.7 / Benchmark/vela.polyco.  UNPINNED (PSRCHIVE source absent).
// ---------------------------------------------------------------------------------------
struct orc_polyco {
  int tmid_day;          // integer MJD of TMID
  double tmid_sec;       // seconds of day of TMID (incl. fraction)
  double rphase_int;     // integer turns of RPHASE
  double rphase_frac;    // fractional turns of RPHASE
  double f0;             // Hz
  double span_min;
  double obsfreq, dm;
  int ncoef;
  double coef[32];
};

static double parse_fortran_double(const std::string& tok) {
  std::string t = tok;
  for (auto& ch : t) if (ch == 'D' || ch == 'd') ch = 'e';
  return std::strtod(t.c_str(), nullptr);
}

// Parse the first polyco block of a TEMPO polyco text.  Returns 0 on success.
int orc_polyco_parse(const char* text, orc_polyco* pc) {
  std::vector<std::string> tok;
  {
    std::string cur;
    for (const char* p = text; ; p++) {
      if (*p == 0 || *p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') {
        if (!cur.empty()) tok.push_back(cur);
        cur.clear();
        if (*p == 0) break;
      } else cur.push_back(*p);
    }
  }
  if (tok.size() < 13) return -1;
  // line 1: name date utc tmid dm doppler log10rms
  const std::string& tmid = tok[3];
  size_t dot = tmid.find('.');
  pc->tmid_day = std::atoi(tmid.substr(0, dot).c_str());
  std::string fr = "0" + (dot == std::string::npos ? std::string(".0") : tmid.substr(dot));
  pc->tmid_sec = std::strtod(fr.c_str(), nullptr) * 86400.0;
  pc->dm = std::strtod(tok[4].c_str(), nullptr);
  // line 2: rphase f0 site span ncoef obsfreq
  const std::string& rp = tok[7];
  dot = rp.find('.');
  pc->rphase_int = std::strtod(rp.substr(0, dot).c_str(), nullptr);
  std::string rf = "0" + (dot == std::string::npos ? std::string(".0") : rp.substr(dot));
  pc->rphase_frac = std::strtod(rf.c_str(), nullptr);
  if (!rp.empty() && rp[0] == '-') pc->rphase_frac = -pc->rphase_frac;
  pc->f0 = std::strtod(tok[8].c_str(), nullptr);
  pc->span_min = std::strtod(tok[10].c_str(), nullptr);
  pc->ncoef = std::atoi(tok[11].c_str());
  pc->obsfreq = std::strtod(tok[12].c_str(), nullptr);
  if (pc->ncoef < 1 || pc->ncoef > 32 || tok.size() < size_t(13 + pc->ncoef)) return -2;
  for (int i = 0; i < pc->ncoef; i++) pc->coef[i] = parse_fortran_double(tok[13 + i]);
  return 0;
}

// DT in minutes between MJD (day, sec, frac) and TMID
static double polyco_dt_min(const orc_polyco* pc, int day, int sec, double frac) {
  double dsec = double(day - pc->tmid_day) * 86400.0 + (double(sec) - pc->tmid_sec) + frac;
  return dsec / 60.0;
}

// phase = RPHASE + DT*60*F0 + sum c_i DT^i ; returns the fractional turns in [0,1)
// (Phase::fracturns of a positive phase) and the integer turns through *turns.
double orc_polyco_phase(const orc_polyco* pc, int day, int sec, double frac, double* turns) {
  double dt = polyco_dt_min(pc, day, sec, frac);
  double poly = 0.0, poweroft = 1.0;
  for (int i = 0; i < pc->ncoef; i++) {
    poly += pc->coef[i] * poweroft;
    poweroft *= dt;
  }
  double spin = dt * 60.0 * pc->f0;
  double spin_int = std::floor(spin);
  double poly_int = std::floor(poly);
  double f = (spin - spin_int) + (poly - poly_int) + pc->rphase_frac;
  double fi = std::floor(f);
  if (turns) *turns = pc->rphase_int + spin_int + poly_int + fi;
  return f - fi;
}

// frequency = F0 + (1/60) sum_{i>=1} i c_i DT^(i-1)   [Hz]
double orc_polyco_frequency(const orc_polyco* pc, int day, int sec, double frac) {
  double dt = polyco_dt_min(pc, day, sec, frac);
  double dpoly = 0.0, poweroft = 1.0;
  for (int i = 1; i < pc->ncoef; i++) {
    dpoly += double(i) * pc->coef[i] * poweroft;
    poweroft *= dt;
  }
  return pc->f0 + dpoly / 60.0;
}

// ---------------------------------------------------------------------------------------
// a14  PhaseSeries::combine (Signal/Pulsar/PhaseSeries.C:442-480)
// ---------------------------------------------------------------------------------------
void orc_phaseseries_combine(float* data, unsigned* hits, double* integration_length, uint64_t* ndat_total,
                             const float* odata, const unsigned* ohits, double ointegration_length,
                             uint64_t ondat_total, uint64_t nfloat, unsigned nhits) {
  for (uint64_t i = 0; i < nfloat; i++) data[i] += odata[i];
  for (unsigned i = 0; i < nhits; i++) hits[i] += ohits[i];
  *integration_length += ointegration_length;
  *ndat_total += ondat_total;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Whole-path driver used by the end-to-end parity tests and by bench.py's cpu_baseline /
// --impl reference legs.  Mirrors `dspsr -t P` (Signal/General/MultiThread.C:90-148,274-379):
// P independent pipelines each take a contiguous run of overlap-save parts of the same raw
// byte stream (re-reading nsamp_overlap samples at the left edge), run
//   unpack -> Filterbank|Convolution -> Detection -> Fold        (SingleThread.C:405-431)
// into a private PhaseSeries, and the partial profiles are summed at the end
// (MultiThread.C:329-342 -> PhaseSeries::combine).
// ---------------------------------------------------------------------------------------
extern "C" {

struct orc_pipe {
  int unpack_fmt;            // 0 CASPSR 8-bit, 1 generic 8-bit TFP, 2 MeerKAT, 3 UWB 16-bit, 5 two-bit (CPSR2 convention)
  unsigned input_nchan, npol, ndim;
  const float* lut;          // 256-entry table (fmt 0,1)
  float scale;               // fmt 2
  int use_filterbank;        // 1: Filterbank (orc_fb), 0: Convolution (orc_conv)
  orc_fb fb;
  orc_conv conv;
  const float* H;
  int detect_state;          // 0 Intensity 1 PPQQ 2 Coherence 3 Stokes
  unsigned detect_ndim;      // 1,2,4 (Coherence/Stokes)
  unsigned nbin;             // 0 = no fold
  const orc_twobit* twobit;  // fmt 5: two-bit excision unpacker (WeightedTimeSeries: weights travel to Fold)
};

static uint64_t raw_bits_per_sample(const orc_pipe* p) {
  unsigned nbit = (p->unpack_fmt == 3) ? 16 : (p->unpack_fmt == 5) ? 2 : 8;
  return uint64_t(p->input_nchan) * p->npol * p->ndim * nbit;
}

static void pipe_unpack(const orc_pipe* p, const uint8_t* raw, uint64_t ndat, float* out, uint64_t span,
                        unsigned* weights, uint64_t nweights) {
  switch (p->unpack_fmt) {
    case 0: orc_unpack_caspsr(raw, ndat, p->lut, out, span); break;
    case 1: orc_unpack_generic8(raw, ndat, p->input_nchan, p->npol, p->ndim, p->lut, out, span, nullptr); break;
    case 2: orc_unpack_meerkat(reinterpret_cast<const int8_t*>(raw), ndat, p->input_nchan, p->npol, p->scale, 1, out, span); break;
    case 3: orc_unpack_uwb(reinterpret_cast<const int16_t*>(raw), ndat, p->npol, out, span); break;
    case 5: orc_unpack_twobit(p->twobit, raw, ndat, p->npol, out, span, weights, nweights); break;
  }
}

// Process parts [ipart0, ipart0+npart) of the stream as ONE block (one Fold call).
// raw points at the first byte of the whole stream.  phi/pps: phase of the midpoint of the
// block's first output sample and phase advance per output sample (Fold.C:650-657,718-720).
// profile: [out_nchan][out_npol][nbin][out_ndim] (+=), hits: [nbin] (+=).
// detected (nullable): receives the detected block, planes of nkeep*npart*out_ndim floats.
// Two-bit input is a WeightedTimeSeries: the per-window weights go through
// convolve_weights/scrunch_weights (Filterbank.C:279-307, Convolution.C:312-319) and Fold
// skips the samples of flagged windows (Fold.C:687-716,746-763).  Returns the number of samples
// folded (ndat_folded), or UINT64_MAX where the reference would throw.
uint64_t orc_pipe_block(const orc_pipe* p, const uint8_t* raw, uint64_t ipart0, uint64_t npart,
                        double phi, double pps, float* profile, unsigned* hits, float* detected) {
  const unsigned step = p->use_filterbank ? p->fb.nsamp_step : p->conv.nsamp_step;
  const unsigned overlap = p->use_filterbank ? p->fb.nsamp_overlap : p->conv.nsamp_overlap;
  const uint64_t ndat_in = npart * step + overlap;
  // The raw layouts are periodic in `res` samples (Unpacker resolution: CASPSR 4, MeerKAT
  // 256-sample heaps, UWB 2048-sample blocks, two-bit ndat_per_weight).  As IOManager/Unpacker::transformation
  // do (Unpacker.C:82-111), unpack the enclosing aligned range and seek to the requested sample.
  const unsigned res = p->unpack_fmt == 0 ? 4 : p->unpack_fmt == 2 ? 256 : p->unpack_fmt == 3 ? 2048
                       : p->unpack_fmt == 5 ? p->twobit->ndat_per_weight : 1;
  const uint64_t s0 = ipart0 * step;
  const uint64_t a0 = (s0 / res) * res;
  const uint64_t a1 = ((s0 + ndat_in + res - 1) / res) * res;
  const uint64_t seek = s0 - a0;
  const uint64_t in_span = (a1 - a0) * p->ndim;
  std::vector<float> unpacked_v(uint64_t(p->input_nchan) * p->npol * in_span);
  std::vector<unsigned> weights;
  uint64_t nweights = 0, weight_idat = 0;
  unsigned ndat_per_weight = 0;
  if (p->unpack_fmt == 5) {
    ndat_per_weight = p->twobit->ndat_per_weight;
    nweights = (a1 - a0) / ndat_per_weight;
    weights.resize(nweights * p->npol);
    weight_idat = seek;                       // TimeSeries::seek -> WeightedTimeSeries::seek moves weight_idat
  }
  pipe_unpack(p, raw + a0 * raw_bits_per_sample(p) / 8, a1 - a0, unpacked_v.data(), in_span, weights.data(), nweights);
  struct { float* p; float* data() { return p; } } unpacked{unpacked_v.data() + seek * p->ndim};

  unsigned out_nchan, nkeep, nsamp_fft, tres_ratio;
  if (p->use_filterbank) {
    out_nchan = p->fb.nchan; nkeep = p->fb.nkeep; nsamp_fft = p->fb.nsamp_fft;
    tres_ratio = p->fb.nsamp_fft / p->fb.freq_res;         // Filterbank.C:289
  } else {
    out_nchan = p->conv.nchan; nkeep = p->conv.n_fft - p->conv.nfilt_pos - p->conv.nfilt_neg; nsamp_fft = p->conv.nsamp_fft;
    tres_ratio = p->conv.input_real ? 2 : 1;                 // Convolution.C:317-318 (Nyquist input -> scrunch 2)
  }
  if (ndat_per_weight) {
    // polarisation 0's weights (identical in all polarisations after mask_weights; npol_weight = 1)
    if (orc_convolve_weights(weights.data(), nweights, ndat_per_weight, weight_idat, ndat_in, nsamp_fft, step) != 0)
      return UINT64_MAX;
    if (tres_ratio > 1 || p->use_filterbank) orc_scrunch_weights(weights.data(), &nweights, &ndat_per_weight, &weight_idat, tres_ratio);
  }
  const uint64_t ndat_out = npart * nkeep;
  const uint64_t v_span = ndat_out * 2;
  std::vector<float> volt(uint64_t(out_nchan) * p->npol * v_span);
  if (p->use_filterbank)
    orc_filterbank_parts(&p->fb, unpacked.data(), in_span, p->H, volt.data(), v_span, 0, npart);
  else
    orc_convolution_parts(&p->conv, unpacked.data(), in_span, p->H, volt.data(), v_span, 0, npart);
  std::vector<float>().swap(unpacked_v);

  unsigned d_npol, d_ndim;
  if (p->detect_state >= 2) { d_ndim = p->detect_ndim; d_npol = 4 / d_ndim; }
  else if (p->detect_state == 1) { d_ndim = 1; d_npol = 2; }
  else { d_ndim = 1; d_npol = 1; }
  const uint64_t d_span = ndat_out * d_ndim;
  std::vector<float> det_local;
  float* det = detected;
  if (!det) { det_local.resize(uint64_t(out_nchan) * d_npol * d_span); det = det_local.data(); }
  orc_detect(p->detect_state, d_ndim, volt.data(), v_span, out_nchan, p->npol, ndat_out, det, d_span);
  std::vector<float>().swap(volt);

  uint64_t ndat_folded = 0;
  if (p->nbin) {
    std::vector<unsigned> binplan(ndat_out);
    ndat_folded = orc_fold_plan_weighted(phi, pps, p->nbin, 0, ndat_out, ndat_per_weight ? weights.data() : nullptr,
                                         nweights, ndat_per_weight, weight_idat, binplan.data(), hits, nullptr);
    if (ndat_folded == UINT64_MAX) return ndat_folded;
    orc_fold(det, d_span, out_nchan, d_npol, d_ndim, 0, ndat_out, binplan.data(), p->nbin, profile);
  }
  return ndat_folded;
}

// P threads, thread t takes blocks t, t+P, ... ; each block = parts_per_block parts.
// phi[b], pps[b] per block.  profile/hits must be zeroed by the caller.
void orc_pipe_run(const orc_pipe* p, const uint8_t* raw, uint64_t nblock, uint64_t parts_per_block,
                  const double* phi, const double* pps, unsigned nthread, float* profile, unsigned* hits,
                  uint64_t* ndat_total) {
  unsigned out_nchan = p->use_filterbank ? p->fb.nchan : p->conv.nchan;
  unsigned per = (p->detect_state >= 2) ? 4 : (p->detect_state == 1 ? 2 : 1);
  const uint64_t nfloat = uint64_t(out_nchan) * per * p->nbin;
  if (nthread < 1) nthread = 1;
  std::vector<std::vector<float>> profs(nthread, std::vector<float>(nfloat, 0.f));
  std::vector<std::vector<unsigned>> hts(nthread, std::vector<unsigned>(p->nbin, 0u));
  std::vector<uint64_t> folded(nthread, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nthread; t++)
    th.emplace_back([&, t]() {
      for (uint64_t b = t; b < nblock; b += nthread) {
        const uint64_t n = orc_pipe_block(p, raw, b * parts_per_block, parts_per_block, phi ? phi[b] : 0.0,
                                          pps ? pps[b] : 0.0, profs[t].data(), hts[t].data(), nullptr);
        folded[t] = (n == UINT64_MAX || folded[t] == UINT64_MAX) ? UINT64_MAX : folded[t] + n;
      }
    });
  for (auto& t : th) t.join();
  double il = 0; uint64_t nt = 0;
  for (unsigned t = 0; t < nthread; t++)
    orc_phaseseries_combine(profile, hits, &il, &nt, profs[t].data(), hts[t].data(), 0,
                            folded[t], nfloat, p->nbin);
  if (ndat_total) {
    *ndat_total = nt;
    for (unsigned t = 0; t < nthread; t++) if (folded[t] == UINT64_MAX) *ndat_total = UINT64_MAX;
  }
}

}  // extern "C"
