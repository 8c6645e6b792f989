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

// ---- PhaseSeries (Signal/Pulsar/PhaseSeries.C) -------------------------------------------------
bool PhaseSeries::mixable(const Observation& obs, unsigned nbin, int64_t istart, int64_t fold_ndat) {   // :336-418
  MJD obsStart = obs.get_start_time() + double(istart) / obs.get_rate();
  MJD obsEnd = fold_ndat == 0 ? obs.get_end_time() : obsStart + double(fold_ndat) / obs.get_rate();
  if (integration_length == 0.0) {
    // the integration is currently empty: adopt the observation, size and zero the bins (:355-395)
    const uint64_t backup_ndat_total = ndat_total;
    const PhaseSeries* like = dynamic_cast<const PhaseSeries*>(&obs);
    TimeSeries::copy_configuration(&obs);
    if (like) end_time = like->end_time;
    resize_bins(nbin);
    zero();
    end_time = obsEnd;
    start_time = obsStart;
    ndat_total = backup_ndat_total;
    return true;
  }
  if (obs.get_nchan() != nchan || obs.get_npol() != npol || obs.get_ndim() != ndim || obs.get_state() != state)
    return false;                                             // Observation::combinable
  if (get_nbin() != nbin) return false;
  if (end_time < obsEnd) end_time = obsEnd;                   // :408-409
  if (obsStart < start_time) start_time = obsStart;
  return true;
}

void PhaseSeries::combine(const PhaseSeries* prof) {          // :442-480
  if (!prof || prof->get_nbin() == 0) return;
  if (!integration_length) {                                  // *this = *prof
    internal_match(prof);
    copy_configuration(prof);
    memory->do_copy(buffer, prof->buffer, size_t(internal_get_size()));
    return;
  }
  // mixable(*prof, nbin) with fold_ndat = 0 uses the observation's end time: a PhaseSeries ends at end_time
  if (prof->get_nchan() != nchan || prof->get_npol() != npol || prof->get_ndim() != ndim || prof->get_nbin() != get_nbin())
    throw Error(InvalidParam, "PhaseSeries::combine", "PhaseSeries !mixable");
  if (end_time < prof->end_time) end_time = prof->end_time;
  if (prof->start_time < start_time) start_time = prof->start_time;
  const size_t n = size_t(span) * nchan * npol;               // TimeSeries::operator +=
  for (size_t i = 0; i < n; i++) buffer[i] += prof->buffer[i];
  for (size_t i = 0; i < hits.size(); i++) hits[i] += prof->hits[i];
  integration_length += prof->integration_length;
  ndat_total += prof->ndat_total;
}

// ---- Fold (Signal/Pulsar/Fold.C) --------------------------------------------------------------
void Fold::set_engine(Engine* e) {
  engine = e;
  if (engine) engine->set_parent(this);
}

PhaseSeries* Fold::get_output() const {                       // :88-94
  if (engine) return engine->get_profiles();
  return output;
}

void Fold::reset() {                                          // :137-148
  if (engine) engine->zero();
  if (output) output->zero();
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
  zeroed_samples = in->get_zeroed_data();
}

void Fold::operate() {                                        // transformation :510-604 + fold :626-829
  if (!engine) throw Error(InvalidState, "dsp::Fold::fold", "stand-in has no CPU path: set an engine");
  if (input->get_ndat() == 0) return;
  if (!folding_nbin) throw Error(InvalidState, "dsp::Fold::fold", "nbin not set");
  PhaseSeries* use = get_output();                            // :536 -- the engine's PhaseSeries
  idat_start = 0;                                             // set_limits :961-965
  ndat_fold = input->get_ndat();
  if (!use->mixable(*input, folding_nbin, int64_t(idat_start), int64_t(ndat_fold)))   // prepare_output :495-504
    throw Error(InvalidParam, "dsp::Fold::prepare_output", "input and output are not mixable");
  const uint64_t idat_end = idat_start + ndat_fold;
  unsigned* hits = get_output()->get_hits();                  // :721
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
  PhaseSeries* result = get_output();                         // :800
  result->integration_length += double(ndat_folded) / input->get_rate();   // :792-803
  result->ndat_total += ndat_fold;
  if (result->get_nbin() != folding_nbin)
    throw Error(InvalidParam, "dsp::Fold::fold", "folding_nbin != output->nbin (%d != %d)", folding_nbin, result->get_nbin());
  engine->fold();                                             // :817-829
}

PhaseSeries* Fold::get_result() {                             // :123-135
  if (engine) engine->synch(output);
  return output;
}

// ---- Unpacker device hooks ----------------------------------------------------------------------
void MeerKATUnpacker::set_engine(Engine* e) { engine = e; }
bool MeerKATUnpacker::get_device_supported(Memory* m) const { return engine && engine->get_device_supported(m); }
void MeerKATUnpacker::set_device(Memory* m) {                 // MeerKATUnpacker.C:120-143
  if (engine) { engine->set_device(m); engine->setup(); }
}
void MeerKATUnpacker::operate() {                             // unpack(), engine branch :186-206
  if (!engine) throw Error(InvalidState, "dsp::MeerKATUnpacker::unpack", "stand-in has no CPU path: set an engine");
  prepare_output(2);
  if (input->get_ndat() == 0) return;
  const unsigned sample_swap = input->get_machine() == "MKBFRo" ? 2 : 1;
  engine->unpack(float(table_scale), input, output, sample_swap);
}

void UWBUnpacker::set_engine(Engine* e) { engine = e; }
bool UWBUnpacker::get_device_supported(Memory* m) const { return engine && engine->get_device_supported(m); }
void UWBUnpacker::set_device(Memory* m) {                     // UWBUnpacker.C:113-139
  if (engine) { engine->set_device(m); engine->setup(); }
}
void UWBUnpacker::operate() {                                 // unpack(), engine branch :150-168
  if (!engine) throw Error(InvalidState, "dsp::UWBUnpacker::unpack", "stand-in has no CPU path: set an engine");
  prepare_output(2);
  engine->unpack(input, output);
}

}  // namespace dsp
