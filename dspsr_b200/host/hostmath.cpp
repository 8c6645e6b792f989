// hostmath.cpp -- host-side (no device) pieces of the path that the reference computes on the
// CPU even in its GPU build, exported through the C ABI so that any host language drives the
// engines with the same numbers:
//   * dsp::BitTable 8-bit lookup table            (Kernel/Classes/BitTable.C:121-218)
//   * dsp::Dedispersion prepare / build / match   (Signal/General/Dedispersion.C:167-331,383-556;
//                                                  Response.C:132-181,259-311,649-700;
//                                                  optimize_fft.c:63-127)
//   * TEMPO polyco phase / frequency              (Pulsar::Predictor as used by Fold.C:943-958)
// Inside a real dspsr these come from dspsr/PSRCHIVE themselves (the engines receive the
// finished Response, FilterbankCUDA.cu:134-165); they are here for stand-alone use of the
// library (bench.py, the Python handles, the C++ demo driver).
// Compile with -ffp-contract=off: double results must not depend on FMA contraction.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200dsp.h"

namespace {

// ---- BitTable -------------------------------------------------------------------------------
// optimal threshold spacing of an n-bit uniform quantiser in units of sigma
// (JenetAnderson98::get_optimal_spacing in PSRCHIVE; Jenet & Anderson 1998, Table 3)
double optimal_spacing(unsigned nbit) {
  static const double table[9] = {0, 0, 0.9674, 0.5605, 0.3188, 0.1789, 0.09925, 0.05445, 0.02957};
  return nbit <= 8 ? table[nbit] : 0.0;
}

double normal_cdf(double x) { return 0.5 * (1.0 + std::erf(x / std::sqrt(2.0))); }

class BitTable {
 public:
  BitTable(unsigned nbit, bool twos_complement) : nbit_(nbit), twos_(twos_complement), scale_(1.0) {}
  // BitTable::generate_unique_values (BitTable.C:165-218)
  void unique_values(float* values) {
    const unsigned n = 1u << nbit_;
    const double out_spacing = 1.0 / double(n);
    const double out_middle = double(n - 1) / 2.0;
    const unsigned in_middle = n / 2;
    const double in_spacing = optimal_spacing(nbit_);
    const unsigned in_offset = twos_ ? n / 2 : 0;
    double cumulative_probability = 0.0, variance = 0.0;
    for (unsigned i = 0; i < n; i++) {
      const double output = (double(i) - out_middle) * out_spacing;
      values[(i + in_offset) % n] = output;
      if (i < in_middle) {
        const double threshold = double(int(i + 1) - int(in_middle)) * in_spacing;
        const double cumulative = normal_cdf(threshold);
        const double interval = cumulative - cumulative_probability;
        cumulative_probability = cumulative;
        variance += output * output * interval;
      }
    }
    variance *= 2.0;
    scale_ = 1.0 / std::sqrt(variance);
    for (unsigned i = 0; i < n; i++) values[i] *= scale_;
    scale_ *= out_spacing;
  }
  double scale() const { return scale_; }

 private:
  unsigned nbit_;
  bool twos_;
  double scale_;
};

// ---- Dedispersion ---------------------------------------------------------------------------
inline double sqr(double x) { return x * x; }

class Dedispersion {
 public:
  explicit Dedispersion(const b200_dedispersion& p) : p_(p) {}

  static constexpr double dm_dispersion = 2.41e-4;            // Dedispersion.C:28
  static constexpr double smearing_buffer = 0.1;              // Dedispersion.C:30
  static constexpr unsigned smearing_threshold = 16 * 1024 * 1024;   // Dedispersion.C:214

  double delay_time(double f1, double f2) const {             // :348-356
    const double dispersion = p_.dispersion_measure / dm_dispersion;
    return dispersion * (1.0 / sqr(f1) - 1.0 / sqr(f2));
  }
  double smearing_time(int half, unsigned skip) const {       // :383-430
    const double abs_bw = std::fabs(p_.bandwidth);
    double ch_abs_bw = abs_bw / double(p_.nchan);
    double lower_ch_cfreq = p_.centre_frequency - (abs_bw - ch_abs_bw) / 2.0;
    for (unsigned i = 0; i < skip; i++) lower_ch_cfreq += ch_abs_bw;
    if (half) {
      ch_abs_bw /= 2.0;
      lower_ch_cfreq += double(half) * ch_abs_bw;
    }
    return delay_time(lower_ch_cfreq - std::fabs(0.5 * ch_abs_bw), lower_ch_cfreq + std::fabs(0.5 * ch_abs_bw));
  }
  unsigned smearing_samples(int half, unsigned skip) const {  // :432-475
    double tsmear = smearing_time(half, skip);
    const double sampling_rate = std::fabs(p_.bandwidth) / double(p_.nchan) * 1e6;
    tsmear *= (1.0 + smearing_buffer);
    return unsigned(std::ceil(tsmear * sampling_rate));
  }
  static unsigned minimum_ndat(unsigned pos, unsigned neg) {  // Response.C:259-275
    const double tot = pos + neg;
    if (tot == 0) return 0;
    unsigned min = unsigned(std::pow(2.0, std::ceil(std::log(tot) / std::log(2.0))));
    while (min <= tot) min *= 2;
    return min;
  }
  static int64_t optimal_fft_length(uint64_t nbad, uint64_t nfft_max) {   // optimize_fft.c:63-127
    if (!nbad) return -1;
    uint64_t nfft_min = (uint64_t)std::pow(2.0, std::ceil(std::log((double)nbad) / std::log(2.0)));
    if (nfft_max && nfft_max < nfft_min) return -1;
    uint64_t nfft = nfft_min;
    double timescale = (double)nfft * std::log((double)nfft) / (double)(nfft - nbad);
    while (nfft_max == 0 || nfft * 2 < nfft_max) {
      const double prev = timescale;
      nfft *= 2;
      timescale = (double)nfft * std::log((double)nfft) / (double)(nfft - nbad);
      if (timescale > prev) {
        nfft /= 2;
        break;
      }
    }
    return (int64_t)nfft;
  }

  // Dedispersion::prepare (:216-248) + ndat choice (:296-308, Response.C:282-311)
  int prepare(b200_dedispersion* out) const {
    const unsigned threshold = smearing_threshold / p_.nchan;
    unsigned skip = 0, neg;
    while ((neg = smearing_samples(-1, skip)) > threshold) {
      if (++skip == p_.nchan) return B200_ERR_INVALID;
    }
    const unsigned pos = smearing_samples(1, skip);
    unsigned ndat;
    if (p_.frequency_resolution) {
      ndat = p_.frequency_resolution;
      if (ndat < minimum_ndat(pos, neg)) return B200_ERR_INVALID;
    } else {
      const int64_t n = optimal_fft_length(uint64_t(pos) + neg, 0);
      if (n < 0) return B200_ERR_INVALID;
      ndat = unsigned(n);
    }
    out->impulse_pos = pos;
    out->impulse_neg = neg;
    out->ndat = ndat;
    return B200_OK;
  }

  static void swap_halves(float* buffer, uint64_t nfloat, unsigned divisions) {   // Response::doswap :649-700
    const uint64_t half = nfloat / (2 * divisions);
    for (unsigned d = 0; d < divisions; d++) {
      float* a = buffer + uint64_t(d) * 2 * half;
      float* b = a + half;
      for (uint64_t i = 0; i < half; i++) std::swap(a[i], b[i]);
    }
  }

  // The same response restricted to channels [chan0, chan0+nloc): what one rank of a channel-sharded
  // run uploads (SURVEY 8e i; channels are independent, Convolution.C:389-391).  Only where the band
  // re-ordering of Response::match stays inside a channel: nchan == input_nchan (Convolution), no input swap.
  int build_channels(unsigned chan0, unsigned nloc, float* H) const {
    if (p_.nchan != p_.input_nchan || p_.input_swap || chan0 + nloc > p_.nchan) return B200_ERR_UNSUPPORTED;
    const unsigned ndat = p_.ndat, nchan = p_.nchan;
    const double bw = p_.bandwidth, cf = p_.centre_frequency;
    const double sign = bw / std::fabs(bw);
    const double chanwidth = bw / double(nchan);
    const double binwidth = chanwidth / double(ndat);
    double lower_cfreq = cf - 0.5 * bw;
    if (!p_.input_dc_centred) lower_cfreq += 0.5 * chanwidth;
    const double dispersion_per_MHz = 1e6 * p_.dispersion_measure / dm_dispersion;
    std::complex<float>* phasors = reinterpret_cast<std::complex<float>*>(H);
    for (unsigned ichan = chan0; ichan < chan0 + nloc; ichan++) {
      const double chan_cfreq = lower_cfreq + double(ichan) * chanwidth;
      const double coeff = -sign * 2 * M_PI * dispersion_per_MHz / sqr(chan_cfreq);
      std::complex<float>* row = phasors + uint64_t(ichan - chan0) * ndat;
      for (unsigned ipt = 0; ipt < ndat; ipt++) {
        const double freq = double(ipt) * binwidth - 0.5 * chanwidth;
        const float phase = coeff * sqr(freq) / (chan_cfreq + freq);
        row[ipt] = std::polar(float(1.0), phase);
      }
    }
    if (chan0 == 0) phasors[0] = 0;
    if (p_.input_dual_sideband)               // per input channel = per response row here (Response.C:165-171)
      swap_halves(H, uint64_t(ndat) * nloc * 2, nloc);
    if (chan0 == 0) H[0] = H[1] = 0.0f;
    return B200_OK;
  }

  // Dedispersion::build (:291-331,478-556) + Response::match (:132-181) + DC zap (:278,323)
  void build(float* H) const {
    const unsigned ndat = p_.ndat, nchan = p_.nchan;
    const double bw = p_.bandwidth, cf = p_.centre_frequency;
    const double sign = bw / std::fabs(bw);
    const double chanwidth = bw / double(nchan);
    const double binwidth = chanwidth / double(ndat);
    double lower_cfreq = cf - 0.5 * bw;
    if (!p_.input_dc_centred) lower_cfreq += 0.5 * chanwidth;
    const double dispersion_per_MHz = 1e6 * p_.dispersion_measure / dm_dispersion;
    std::complex<float>* phasors = reinterpret_cast<std::complex<float>*>(H);
    for (unsigned ichan = 0; ichan < nchan; ichan++) {
      const double chan_cfreq = lower_cfreq + double(ichan) * chanwidth;
      const double coeff = -sign * 2 * M_PI * dispersion_per_MHz / sqr(chan_cfreq);
      for (unsigned ipt = 0; ipt < ndat; ipt++) {
        const double freq = double(ipt) * binwidth - 0.5 * chanwidth;
        const double delay_phase = -2.0 * M_PI * freq * 0.0;
        const float phase = coeff * sqr(freq) / (chan_cfreq + freq) + delay_phase;   // stored as float (:311,545)
        phasors[uint64_t(ichan) * ndat + ipt] = std::polar(float(1.0), phase);      // :320
      }
    }
    phasors[0] = 0;
    const uint64_t nfloat = uint64_t(ndat) * nchan * 2;
    if (p_.input_nchan == 1) {
      if (p_.input_dual_sideband) swap_halves(H, nfloat, 1);
    } else {
      if (p_.input_dual_sideband) swap_halves(H, nfloat, p_.input_nchan);
      if (p_.input_swap) swap_halves(H, nfloat, 1);
    }
    H[0] = H[1] = 0.0f;
  }

 private:
  b200_dedispersion p_;
};

// ---- polyco ---------------------------------------------------------------------------------
double fortran_double(std::string t) {
  for (auto& c : t)
    if (c == 'D' || c == 'd') c = 'e';
  return std::strtod(t.c_str(), nullptr);
}

double minutes_since_tmid(const b200_polyco* pc, int day, int sec, double frac) {
  const double dsec = double(day - pc->tmid_day) * 86400.0 + (double(sec) - pc->tmid_sec) + frac;
  return dsec / 60.0;
}

}  // namespace

extern "C" {

int b200_bittable8(int twos_complement, float* lut256, double* scale) {
  if (!lut256) return B200_ERR_INVALID;
  BitTable t(8, twos_complement != 0);
  t.unique_values(lut256);
  if (scale) *scale = t.scale();
  return B200_OK;
}

int b200_dedispersion_prepare(b200_dedispersion* d) {
  if (!d || d->nchan == 0 || d->input_nchan == 0 || d->bandwidth == 0.0) return B200_ERR_INVALID;
  return Dedispersion(*d).prepare(d);
}

int b200_dedispersion_build(const b200_dedispersion* d, float* h_response) {
  if (!d || !h_response || d->ndat == 0) return B200_ERR_INVALID;
  Dedispersion(*d).build(h_response);
  return B200_OK;
}

int b200_dedispersion_build_channels(const b200_dedispersion* d, unsigned first_chan, unsigned nchan_local,
                                     float* h_response) {
  if (!d || !h_response || d->ndat == 0 || nchan_local == 0) return B200_ERR_INVALID;
  return Dedispersion(*d).build_channels(first_chan, nchan_local, h_response);
}

// ---- two-bit excision tables ---------------------------------------------------------------------
// ExcisionUnpacker::set_limits (ExcisionUnpacker.C:95-158) + TwoBitLookup::lookup_build (TwoBitLookup.C:63-98)
// with the Jenet & Anderson (1998) section 6 levels (PSRCHIVE's JenetAnderson98::set_Phi, restated):
//   alpha = ierf(Phi); lo^2 = 1 - (2 alpha/sqrt pi) exp(-alpha^2)/Phi; hi^2 = 1 + (2 alpha/sqrt pi) exp(-alpha^2)/(1-Phi)
static double inverse_erf(double y) {
  const double a = 0.147;
  const double ln1 = std::log(1.0 - y * y);
  const double t = 2.0 / (M_PI * a) + 0.5 * ln1;
  double x = std::copysign(std::sqrt(std::sqrt(t * t - ln1 / a) - t), y);
  for (int i = 0; i < 4; i++) x -= (std::erf(x) - y) / (2.0 / std::sqrt(M_PI) * std::exp(-x * x));
  return x;
}

int b200_twobit_prepare(double threshold, float cutoff_sigma, int table_type, unsigned npol,
                        unsigned ndat_per_weight, b200_twobit_desc* d) {
  if (!d || npol < 1 || npol > 2 || table_type < 0 || table_type > 2 || ndat_per_weight < 4 ||
      ndat_per_weight > 512 || ndat_per_weight % 128 != 0)
    return B200_ERR_INVALID;
  d->table_type = table_type;
  d->npol = npol;
  d->ndat_per_weight = ndat_per_weight;
  if (cutoff_sigma == 0.0) {
    d->nlow_min = 0;
    d->nlow_max = ndat_per_weight;
  } else {
    const double mean_Phi = std::erf(threshold / std::sqrt(2.0));
    const double var_Phi = mean_Phi * (1.0 - mean_Phi);
    float fsample = ndat_per_weight;
    float nlo_mean = fsample * mean_Phi;
    float nlo_variance = fsample * var_Phi;
    float nlo_sigma = sqrt(nlo_variance);
    d->nlow_max = unsigned(nlo_mean + (cutoff_sigma * nlo_sigma));
    if (d->nlow_max >= ndat_per_weight) d->nlow_max = ndat_per_weight - 1;
    if (cutoff_sigma * nlo_sigma >= nlo_mean + 1.0) d->nlow_min = 1;
    else d->nlow_min = unsigned(nlo_mean - (cutoff_sigma * nlo_sigma));
  }
  const double root_pi = std::sqrt(M_PI);
  for (unsigned nlo = d->nlow_min; nlo <= d->nlow_max; nlo++) {
    unsigned use_nlow = nlo == 0 ? 1 : nlo;
    // TwoBitLookup.C:83-84 means to clamp the all-low row as well but tests the member `nlow` (always 0 there)
    // instead of the loop variable, so with cutoff_sigma = 0 the reference evaluates Phi = 1 -> ierf(1) = inf ->
    // NaN levels, and any all-low window poisons its whole FFT part.  Deliberate deviation: clamp as intended.
    if (nlo == ndat_per_weight) use_nlow = ndat_per_weight - 1;
    float p_in = (float)use_nlow / (float)ndat_per_weight;
    const double Phi = p_in;
    const double alpha = inverse_erf(Phi);
    const double expon = std::exp(-alpha * alpha);
    d->lo[nlo - d->nlow_min] = float(std::sqrt(1.0 - (2.0 * alpha / root_pi) * (expon / Phi)));
    d->hi[nlo - d->nlow_min] = float(std::sqrt(1.0 + (2.0 * alpha / root_pi) * (expon / (1.0 - Phi))));
  }
  return B200_OK;
}

// ---- TimeDivide (seconds mode) ------------------------------------------------------------------
int b200_time_divide_init(b200_time_divide* td, double division_seconds) {
  if (!td || !(division_seconds > 0)) return B200_ERR_INVALID;
  td->division_seconds = division_seconds;
  td->lower = td->upper = td->current_end = 0.0;
  td->is_valid = 0;
  td->division = 0;
  return B200_OK;
}

int b200_time_divide_set_bounds(b200_time_divide* td, double input_start, double rate, uint64_t input_ndat,
                                b200_time_bounds* o) {
  if (!td || !o || !(rate > 0)) return B200_ERR_INVALID;
  const double input_end = input_start + double(input_ndat) / rate;
  // TimeDivide.C:147-160: where to start
  double divide_start = input_start;
  if (td->is_valid) divide_start = std::max(td->current_end, input_start);
  o->new_division = o->end_reached = o->in_next = 0;
  if (input_end < td->lower || divide_start + 0.5 / rate > td->upper) {       // :166-196
    o->new_division = 1;
    const double t = divide_start + 0.55 / rate;                              // set_boundaries(:349-425)
    const double ds = std::max(0.0, t);
    td->division = uint64_t(ds / td->division_seconds);
    td->lower = double(td->division) * td->division_seconds;
    td->upper = double(td->division + 1) * td->division_seconds;
  }
  divide_start = std::max(td->lower, divide_start);
  o->division = td->division;
  // :210-231: how far into the block to start
  double start_sample = std::rint((divide_start - input_start) * rate);
  if (start_sample < 0) start_sample = 0;
  o->idat_start = uint64_t(start_sample);
  o->ndat = 0;
  if (o->idat_start >= input_ndat) {
    td->is_valid = o->is_valid = 0;
    return B200_OK;
  }
  // :233-300: how far into the block to end
  double divide_end = std::min(input_end, td->upper);
  uint64_t idat_end = uint64_t(std::rint((divide_end - input_start) * rate));
  if (idat_end <= o->idat_start) {
    // A boundary exactly half-way between two samples leaves the old division with nothing to fold; the
    // reference throws InvalidState here (TimeDivide.C:258-287).  Start the next division instead.
    o->new_division = 1;
    td->division += 1;
    td->lower = double(td->division) * td->division_seconds;
    td->upper = double(td->division + 1) * td->division_seconds;
    divide_end = std::min(input_end, td->upper);
    idat_end = uint64_t(std::rint((divide_end - input_start) * rate));
    o->division = td->division;
    if (idat_end <= o->idat_start) return B200_ERR_INVALID;
  }
  if (idat_end > input_ndat) idat_end = input_ndat;
  else if (idat_end < input_ndat) o->in_next = 1;
  o->ndat = idat_end - o->idat_start;
  // :302-317: has the end of the division been reached?
  if ((td->upper - divide_end) * rate < 0.5) o->end_reached = 1;
  td->is_valid = o->is_valid = 1;
  td->current_end = input_start + double(idat_end) / rate;
  return B200_OK;
}

int64_t b200_optimal_fft_length(uint64_t nbadperfft, uint64_t nfft_max) {
  return Dedispersion::optimal_fft_length(nbadperfft, nfft_max);
}

int b200_polyco_parse(const char* text, b200_polyco* pc) {
  if (!text || !pc) return B200_ERR_INVALID;
  std::vector<std::string> tok;
  std::string cur;
  for (const char* p = text;; p++) {
    if (*p == 0 || *p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') {
      if (!cur.empty()) tok.push_back(cur);
      cur.clear();
      if (*p == 0) break;
    } else
      cur.push_back(*p);
  }
  if (tok.size() < 13) return B200_ERR_INVALID;
  auto split = [](const std::string& s, double* ipart, double* fpart) {
    size_t dot = s.find('.');
    *ipart = std::strtod(s.substr(0, dot).c_str(), nullptr);
    std::string f = "0" + (dot == std::string::npos ? std::string(".0") : s.substr(dot));
    *fpart = std::strtod(f.c_str(), nullptr);
    if (!s.empty() && s[0] == '-') *fpart = -*fpart;
  };
  double ip, fp;
  split(tok[3], &ip, &fp);
  pc->tmid_day = int(ip);
  pc->tmid_sec = fp * 86400.0;
  pc->dm = std::strtod(tok[4].c_str(), nullptr);
  split(tok[7], &pc->rphase_int, &pc->rphase_frac);
  pc->f0 = std::strtod(tok[8].c_str(), nullptr);
  pc->span_min = std::strtod(tok[10].c_str(), nullptr);
  pc->ncoef = std::atoi(tok[11].c_str());
  pc->obsfreq = std::strtod(tok[12].c_str(), nullptr);
  if (pc->ncoef < 1 || pc->ncoef > 32 || tok.size() < size_t(13 + pc->ncoef)) return B200_ERR_INVALID;
  for (int i = 0; i < pc->ncoef; i++) pc->coef[i] = fortran_double(tok[13 + i]);
  return B200_OK;
}

double b200_polyco_phase(const b200_polyco* pc, int day, int sec, double frac, double* turns) {
  const double dt = minutes_since_tmid(pc, day, sec, frac);
  double poly = 0.0, poweroft = 1.0;
  for (int i = 0; i < pc->ncoef; i++) {
    poly += pc->coef[i] * poweroft;
    poweroft *= dt;
  }
  const double spin = dt * 60.0 * pc->f0;
  const double spin_int = std::floor(spin), poly_int = std::floor(poly);
  const double f = (spin - spin_int) + (poly - poly_int) + pc->rphase_frac;
  const double fi = std::floor(f);
  if (turns) *turns = pc->rphase_int + spin_int + poly_int + fi;
  return f - fi;
}

double b200_polyco_frequency(const b200_polyco* pc, int day, int sec, double frac) {
  const double dt = minutes_since_tmid(pc, day, sec, frac);
  double dpoly = 0.0, poweroft = 1.0;
  for (int i = 1; i < pc->ncoef; i++) {
    dpoly += double(i) * pc->coef[i] * poweroft;
    poweroft *= dt;
  }
  return pc->f0 + dpoly / 60.0;
}

}  // extern "C"
