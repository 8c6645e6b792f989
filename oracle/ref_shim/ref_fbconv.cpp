// oracle/ref_shim/ref_fbconv.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" doors to the overlap-save loops of
//   dsp::Filterbank::filterbank       Signal/General/Filterbank.C:563-660   
//   dsp::Convolution::transformation  Signal/General/Convolution.C:389-458  (row a11)
// compiled FROM THE REFERENCE'S OWN TEXT: oracle/ref.mk cuts the two loop nests out of the files where they lie into
// oracle/_ref/gen/*.inc (build products, git-ignored) and this harness supplies the local variables around them.
// The response multiply inside the loops is the reference's own dsp::Response::operate (Response.C, compiled in
// place into this library); the three FFT calls (FTransform::Plan::frc1d / fcc1d / bcc1d = FFTW, a third-party
// dependency that is not in the tree) go to the oracle's restatement of the FFTW conventions (oracle/orc_fft.cpp).
// What this pins: every pointer step, the order of parts / channels / polarisations, which spectrum slice each
// response channel multiplies, the nfilt_pos discard and the copy-out.
#include <stdint.h>
#include <string.h>

#include <complex>
#include <iostream>
#include <vector>

#include "dsp/Response.h"

extern "C" {
void orc_fft_fcc1d(unsigned n, float* out, const float* in);
void orc_fft_bcc1d(unsigned n, float* out, const float* in);
void orc_fft_frc1d(unsigned n, float* out, const float* in);
}

#ifndef DEBUG
#define DEBUG(x)
#endif

using std::cerr;
using std::endl;

namespace {

namespace Signal {
enum State { Nyquist, Analytic };
}

struct Plan {
  void frc1d(unsigned n, float* out, const float* in) { orc_fft_frc1d(n, out, in); }
  void fcc1d(unsigned n, float* out, const float* in) { orc_fft_fcc1d(n, out, in); }
  void bcc1d(unsigned n, float* out, const float* in) { orc_fft_bcc1d(n, out, in); }
};
struct Apodization {
  void operate(float*, float*) {}
};
struct Passband {
  void integrate(float*, unsigned, unsigned) {}
  void integrate(float*, float*, unsigned) {}
};
struct InSeries {
  const float* base;
  uint64_t span;
  unsigned nchan, npol, ndim;
  Signal::State state;
  unsigned get_nchan() const { return nchan; }
  unsigned get_npol() const { return npol; }
  unsigned get_ndim() const { return ndim; }
  Signal::State get_state() const { return state; }
  const float* get_datptr(unsigned ichan, unsigned ipol) const { return base + (uint64_t(ichan) * npol + ipol) * span; }
};
struct OutSeries {
  float* base;
  uint64_t span;
  unsigned npol;
  float* get_datptr(unsigned ichan, unsigned ipol) { return base + (uint64_t(ichan) * npol + ipol) * span; }
};

// the reference's Response holding H[nchan][ndat] complex (one polarisation: the same filter for both)
void fill_response(dsp::Response& r, const float* H, unsigned nchan, unsigned ndat) {
  r.resize(1, nchan, ndat, 2);
  for (unsigned c = 0; c < nchan; c++) memcpy(r.get_datptr(c, 0), H + uint64_t(c) * ndat * 2, sizeof(float) * ndat * 2);
}

}  // namespace

extern "C" {

// in: FPT planes [input_nchan][npol] of in_span floats; out: planes [input_nchan*nchan_subband][npol] of out_span
// floats (complex); H: [input_nchan*nchan_subband][freq_res] complex in the order Response::match leaves it, or null
int ref_filterbank(const float* in, uint64_t in_span, unsigned input_nchan, unsigned npol_, int real_input,
                   unsigned nchan_subband, unsigned freq_res, unsigned nfilt_pos, unsigned nfilt_neg, unsigned nsamp_fft,
                   unsigned nsamp_step, uint64_t npart, const float* H, float* out, uint64_t out_span) {
  try {
    InSeries in_obj = {in, in_span, input_nchan, npol_, real_input ? 1u : 2u, real_input ? Signal::Nyquist : Signal::Analytic};
    OutSeries out_obj = {out, out_span, npol_};
    const InSeries* input = &in_obj;
    OutSeries* output = &out_obj;
    dsp::Response resp;
    const dsp::Response* response = 0;
    if (H) {
      fill_response(resp, H, input_nchan * nchan_subband, freq_res);
      response = &resp;
    }
    Plan plan;
    Plan* forward = &plan;
    Plan* backward = &plan;
    Apodization* apodization = 0;
    Passband* passband = 0;
    const bool matrix_convolution = false;
    const unsigned nfilt_tot = nfilt_pos + nfilt_neg;
    // Filterbank.C:480-530: scratch and the counters of the loop
    unsigned bigfftsize = nchan_subband * freq_res * 2;
    if (input->get_state() == Signal::Nyquist) bigfftsize += 256;
    std::vector<float> scratch(bigfftsize + 2 * freq_res + 16);
    float* c_spectrum[2];
    c_spectrum[0] = &scratch[0];
    c_spectrum[1] = c_spectrum[0];
    float* c_time = c_spectrum[1] + bigfftsize;
    float* windowed_time_domain = 0;
    unsigned cross_pol = 1;
    const unsigned long in_step = nsamp_step * input->get_ndim();
    const unsigned nkeep = freq_res - nfilt_tot;
    const unsigned long out_step = nkeep * 2;
    unsigned ipt, ipol, jpol, ichan;
    uint64_t ipart;
    const unsigned npol = input->get_npol();
    uint64_t in_offset, out_offset;
    float* time_dom_ptr = NULL;
    float* freq_dom_ptr = NULL;
    uint64_t* data_into = NULL;
    uint64_t* data_from = NULL;
#include "gen/filterbank_loop.inc"
    (void)windowed_time_domain;
  } catch (Error& e) {
    std::fprintf(stderr, "ref_filterbank: %s: %s\n", e.function.c_str(), e.message.c_str());
    return -1;
  }
  return 0;
}

// in / out: FPT planes [nchan][npol]; H: [nchan][n_fft] complex (always present: Convolution requires a response)
int ref_convolution(const float* in, uint64_t in_span, unsigned nchan_, unsigned npol_, int real_input, unsigned n_fft,
                    unsigned nfilt_pos, unsigned nsamp_fft, unsigned nsamp_step, uint64_t npart, const float* H, float* out,
                    uint64_t out_span) {
  try {
    InSeries in_obj = {in, in_span, nchan_, npol_, real_input ? 1u : 2u, real_input ? Signal::Nyquist : Signal::Analytic};
    OutSeries out_obj = {out, out_span, npol_};
    const InSeries* input = &in_obj;
    OutSeries* output = &out_obj;
    dsp::Response resp;
    fill_response(resp, H, nchan_, n_fft);
    const dsp::Response* response = &resp;
    Plan plan;
    Plan* forward = &plan;
    Plan* backward = &plan;
    Apodization* apodization = 0;
    Passband* passband = 0;
    const bool matrix_convolution = false;
    // Convolution.C:340-387
    Signal::State state = input->get_state();
    const unsigned npol = input->get_npol();
    const unsigned nchan = input->get_nchan();
    const unsigned ndim = input->get_ndim();
    std::vector<float> scratch(size_t(n_fft) * 4 + 16);
    float* spectrum[2];
    spectrum[0] = &scratch[0];
    spectrum[1] = spectrum[0];
    float* complex_time = spectrum[1] + n_fft * 2;
    if (state == Signal::Nyquist) complex_time += 4;
    const unsigned nbytes_step = nsamp_step * ndim * sizeof(float);
    const unsigned cross_pol = matrix_convolution ? 2 : 1;
    float* ptr = 0;
    unsigned jpol = 0;
    uint64_t offset;
    const uint64_t step = nsamp_step * ndim;
#include "gen/convolution_loop.inc"
  } catch (Error& e) {
    std::fprintf(stderr, "ref_convolution: %s: %s\n", e.function.c_str(), e.message.c_str());
    return -1;
  }
  return 0;
}

}  // extern "C"
