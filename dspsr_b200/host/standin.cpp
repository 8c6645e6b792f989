// standin.cpp -- the engine-dispatch branches of the reference operators, restated for the
// stand-in classes of dsp/standin.h so that the shims can be exercised without PSRCHIVE.
// Only the code that runs WHEN AN ENGINE IS SET is here; there is no CPU path.
#include <cmath>
#include <cstring>

#include "dsp/standin.h"

namespace dsp {

// ---- Filterbank (Signal/General/Filterbank.C) -------------------------------------------------
void Filterbank::set_engine(Engine* e) { engine = e; }        // Filterbank.C:36-39

void Filterbank::prepare() {                                  // make_preparations, :55-263
  if (nchan < input->get_nchan())
    throw Error(InvalidState, "dsp::Filterbank::make_preparations", "output nchan=%d < input nchan=%d", nchan,
                input->get_nchan());
  if (nchan % input->get_nchan() != 0)
    throw Error(InvalidState, "dsp::Filterbank::make_preparations",
                "output nchan=%d not a multiple of input nchan=%d", nchan, input->get_nchan());
  nchan_subband = nchan / input->get_nchan();                 // :68
  nfilt_pos = nfilt_neg = 0;
  if (response) {
    if (response->get_nchan() != nchan)
      throw Error(InvalidState, "dsp::Filterbank::make_preparations", "response nchan=%d != output nchan=%d",
                  response->get_nchan(), nchan);
    nfilt_pos = response->get_impulse_pos();                  // :90-93
    nfilt_neg = response->get_impulse_neg();
    freq_res = response->get_ndat();
  }
  if (freq_res == 0) throw Error(InvalidState, "dsp::Filterbank::make_preparations", "Response.ndat = 0");
  const unsigned n_fft = nchan_subband * freq_res;            // :107
  const unsigned nfilt_tot = nfilt_pos + nfilt_neg;
  if (input->get_state() == Signal::Nyquist) {                // :139-148
    nsamp_fft = 2 * n_fft;
    nsamp_overlap = 2 * nfilt_tot * nchan_subband;
  } else if (input->get_state() == Signal::Analytic) {
    nsamp_fft = n_fft;
    nsamp_overlap = nfilt_tot * nchan_subband;
  } else
    throw Error(InvalidState, "dsp::Filterbank::make_preparations", "invalid input data state");
  nsamp_step = nsamp_fft - nsamp_overlap;                     // :155
  // prepare_output (:265-379): the attributes the downstream operators read
  output->copy_configuration(input);
  output->set_nchan(nchan);
  output->set_ndim(2);
  output->set_state(Signal::Analytic);
  output->rescale(double(n_fft) * double(freq_res));          // :124-125,328
  output->set_rate(input->get_rate() * double(freq_res) / double(nsamp_fft));   // :338-339
  if (!engine) throw Error(InvalidState, "dsp::Filterbank::make_preparations", "stand-in has no CPU path: set an engine");
  engine->setup(this);                                        // :219-225
  prepared = true;
}

void Filterbank::operate() {                                  // transformation :432-475 + filterbank :477-553
  if (!prepared) prepare();
  const uint64_t ndat = input->get_ndat();
  uint64_t npart = 0;
  if (ndat > nsamp_overlap) npart = (ndat - nsamp_overlap) / nsamp_step;   // :401-402
  const unsigned nkeep = freq_res - (nfilt_pos + nfilt_neg);               // :409
  output->resize(npart * nkeep);
  if (!npart) return;
  const uint64_t in_step = uint64_t(nsamp_step) * input->get_ndim();       // :517
  const uint64_t out_step = uint64_t(nkeep) * 2;                           // :523
  engine->set_scratch(0);                                                   // :549 (the B200 engine owns its scratch)
  engine->perform(input, output, npart, in_step, out_step);                // :550
}

// ---- Convolution (Signal/General/Convolution.C) -----------------------------------------------
void Convolution::set_engine(Engine* e) { engine = e; }

void Convolution::prepare() {                                 // :105-221
  if (!response) throw Error(InvalidState, "dsp::Convolution::prepare", "no frequency response");
  if (response->get_ndat() < 2) throw Error(InvalidState, "dsp::Convolution::prepare", "invalid response size");
  if (response->get_nchan() != input->get_nchan())
    throw Error(InvalidState, "dsp::Convolution::prepare", "invalid response nsub=%d != nchan=%d",
                response->get_nchan(), input->get_nchan());
  n_fft = response->get_ndat();
  nfilt_pos = response->get_impulse_pos();
  nfilt_neg = response->get_impulse_neg();
  const unsigned nfilt_tot = nfilt_pos + nfilt_neg;
  if (input->get_state() == Signal::Nyquist) { nsamp_fft = n_fft * 2; nsamp_overlap = nfilt_tot * 2; }
  else if (input->get_state() == Signal::Analytic) { nsamp_fft = n_fft; nsamp_overlap = nfilt_tot; }
  else throw Error(InvalidState, "dsp::Convolution::prepare", "Cannot transform this Signal::State");
  if (nsamp_fft < nsamp_overlap)
    throw Error(InvalidState, "dsp::Convolution::prepare", "error nfft=%d < nfilt=%d", nsamp_fft, nsamp_overlap);
  nsamp_step = nsamp_fft - nsamp_overlap;
  output->copy_configuration(input);
  output->set_state(Signal::Analytic);
  output->set_ndim(2);
  if (input->get_state() == Signal::Nyquist) output->set_rate(0.5 * input->get_rate());
  output->rescale(double(nsamp_fft) * double(n_fft));         // :303-305
  if (!engine) throw Error(InvalidState, "dsp::Convolution::prepare", "stand-in has no CPU path: set an engine");
  engine->prepare(this);                                      // :202-209
  prepared = true;
}

void Convolution::operate() {                                 // :338-365
  if (!prepared) prepare();
  const uint64_t ndat = input->get_ndat();
  npart = 0;
  if (ndat >= nsamp_fft) npart = (ndat - nsamp_overlap) / nsamp_step;      // :236-238
  uint64_t output_ndat = npart * nsamp_step;
  if (input->get_state() == Signal::Nyquist) output_ndat /= 2;             // :290-292
  output->resize(output_ndat);
  if (!npart) return;
  engine->set_scratch(0);
  engine->perform(input, output, unsigned(npart));
}

// ---- Detection (Signal/General/Detection.C) ---------------------------------------------------
void Detection::set_engine(Engine* e) { engine = e; }

void Detection::operate() {                                   // :74-147
  if (!engine) throw Error(InvalidState, "dsp::Detection::transformation", "stand-in has no CPU path: set an engine");
  const bool inplace = (input.get() == output.get());
  unsigned output_ndim = 1, output_npol = input->get_npol();  // resize_output :153-205
  if (state == Signal::Stokes || state == Signal::Coherence) {
    if (input->get_npol() != 2 || input->get_state() != Signal::Analytic)
      throw Error(InvalidState, "dsp::Detection::polarimetry",
                  "Cannot detect polarization when ndim != 2 or state != Analytic");
    output_ndim = ndim;
    output_npol = 4 / ndim;
  } else if (state == Signal::PPQQ) output_npol = 2;
  else if (state == Signal::Intensity) output_npol = 1;
  if (!inplace) {
    output->copy_configuration(input);
    output->set_npol(output_npol);
    output->set_ndim(output_ndim);
    output->resize(input->get_ndat());
  }
  if (state == Signal::Coherence || state == Signal::Stokes) engine->polarimetry(ndim, input, output);   // :327-334
  else engine->square_law(input, output);                                                                 // :223-229
  output->set_state(state);
}

// ---- Fold (Signal/Pulsar/Fold.C) --------------------------------------------------------------
void Fold::set_engine(Engine* e) {
  engine = e;
  if (engine) engine->set_parent(this);
}

void Fold::Engine::setup() {                                  // Fold.C:973-1011
  if (!parent) throw Error(InvalidState, "dsp::Fold::Engine::setup", "no parent");
  const TimeSeries* in = parent->get_input();
  nchan = in->get_nchan();
  npol = in->get_npol();
  ndim = in->get_ndim();
  input = in->get_datptr(0, 0);
  input_span = unsigned(in->get_nfloat_span());
  PhaseSeries* out = get_profiles();
  output = out->get_datptr(0, 0);
  output_span = unsigned(out->get_nfloat_span());
  hits = out->get_hits();
  hits_nchan = out->get_hits_nchan();
  zeroed_samples = false;
}

void Fold::operate() {                                        // transformation :510-604 + fold :626-829
  if (!engine) throw Error(InvalidState, "dsp::Fold::fold", "stand-in has no CPU path: set an engine");
  if (input->get_ndat() == 0) return;
  if (!folding_nbin) throw Error(InvalidState, "dsp::Fold::fold", "nbin not set");
  idat_start = 0;                                             // set_limits :961-965
  ndat_fold = input->get_ndat();
  const uint64_t idat_end = idat_start + ndat_fold;
  unsigned* hits = output->get_hits();
  engine->set_nbin(folding_nbin);                             // :728
  engine->set_ndat(idat_end - idat_start, idat_start);        // :729
  uint64_t ndat_folded = 0;
  if (engine->use_set_bins) {                                 // :730-740
    ndat_folded = engine->set_bins(phi, phase_per_sample, idat_end - idat_start, idat_start);
    for (unsigned ibin = 0; ibin < folding_nbin; ibin++) hits[ibin] += unsigned(engine->get_bin_hits(int(ibin)));
  } else {
    double p = phi;                                           // :744-788
    for (uint64_t idat = idat_start; idat < idat_end; idat++) {
      p -= floor(p);
      double double_ibin = p * double(folding_nbin);
      unsigned ibin = unsigned(double_ibin);
      p += phase_per_sample;
      engine->set_bin(idat, double_ibin, phase_per_sample * double(folding_nbin));
      hits[ibin]++;
      ndat_folded++;
    }
  }
  output->integration_length += double(ndat_folded) / input->get_rate();   // :792-803
  output->ndat_total += ndat_fold;
  engine->fold();                                             // :817-829
}

PhaseSeries* Fold::get_result() {                             // :123-135
  if (engine) engine->synch(output);
  return output;
}

}  // namespace dsp
