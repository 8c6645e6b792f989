// B200Engines.cpp -- see B200Engines.h.  No CUDA here: everything goes through the C ABI.
#include "B200Engines.h"

namespace B200 {

void check(int status, const char* where) {
  if (status != B200_OK) throw Error(FailedCall, where, "%s", b200_last_error());
}

b200_context* context_for(void* cuda_stream, int device) {
  // one context per (device, stream) = per dspsr pipeline thread (SingleThread.C:237-244); cached so that
  // the four engines of a thread share it
  static std::vector<std::pair<std::pair<int, void*>, b200_context*> > cache;
  for (size_t i = 0; i < cache.size(); i++)
    if (cache[i].first.first == device && cache[i].first.second == cuda_stream) return cache[i].second;
  b200_context* ctx = 0;
  check(b200_context_create(device, cuda_stream, &ctx), "B200::context_for");
  cache.push_back(std::make_pair(std::make_pair(device, cuda_stream), ctx));
  return ctx;
}

void* DeviceMemory::do_allocate(size_t nbytes) {
  void* p = 0;
  check(b200_malloc(ctx, nbytes, &p), "B200::DeviceMemory::do_allocate");
  return p;
}
void DeviceMemory::do_free(void* p) { b200_free(ctx, p); }
void DeviceMemory::do_zero(void* p, size_t nbytes) { check(b200_memset(ctx, p, 0, nbytes), "B200::DeviceMemory::do_zero"); }
void DeviceMemory::do_copy(void* to, const void* from, size_t nbytes) {
  check(b200_memcpy_d2d(ctx, to, from, nbytes), "B200::DeviceMemory::do_copy");
}

// ---- Unpacker device hooks ----------------------------------------------------------------------
// Both engines receive device pointers: the BitSeries was loaded with File::load_bytes_device and the unpacked
// TimeSeries lives in DeviceMemory (SingleThread.C:249-275).
void MeerKATUnpackerEngine::unpack(float scale, const dsp::BitSeries* input, dsp::TimeSeries* output, unsigned sample_swap) {
  b200_unpack_desc d;
  d.format = B200_FMT_MEERKAT8;
  d.nchan = input->get_nchan();
  d.npol = input->get_npol();
  d.ndim = 2;
  d.scale = scale;
  d.sample_swap = sample_swap;
  d.twobit = 0;
  check(b200_unpack(ctx, &d, input->get_rawptr(), input->get_ndat(), output->get_datptr(0, 0), output->get_nfloat_span()),
        "B200::MeerKATUnpackerEngine::unpack");
}
bool MeerKATUnpackerEngine::get_device_supported(dsp::Memory* memory) const { return dynamic_cast<DeviceMemory*>(memory) != 0; }
void MeerKATUnpackerEngine::set_device(dsp::Memory* memory) {
  DeviceMemory* m = dynamic_cast<DeviceMemory*>(memory);
  if (!m) throw Error(InvalidState, "B200::MeerKATUnpackerEngine::set_device", "not a B200::DeviceMemory");
  ctx = m->get_context();
}

void UWBUnpackerEngine::unpack(const dsp::BitSeries* input, dsp::TimeSeries* output) {
  b200_unpack_desc d;
  d.format = B200_FMT_UWB16;
  d.nchan = 1;
  d.npol = input->get_npol();
  d.ndim = 2;
  d.scale = 0;
  d.sample_swap = 1;
  d.twobit = 0;
  check(b200_unpack(ctx, &d, input->get_rawptr(), input->get_ndat(), output->get_datptr(0, 0), output->get_nfloat_span()),
        "B200::UWBUnpackerEngine::unpack");
}
bool UWBUnpackerEngine::get_device_supported(dsp::Memory* memory) const { return dynamic_cast<DeviceMemory*>(memory) != 0; }
void UWBUnpackerEngine::set_device(dsp::Memory* memory) {
  DeviceMemory* m = dynamic_cast<DeviceMemory*>(memory);
  if (!m) throw Error(InvalidState, "B200::UWBUnpackerEngine::set_device", "not a B200::DeviceMemory");
  ctx = m->get_context();
}

// ---- Filterbank ---------------------------------------------------------------------------------
FilterbankEngine::~FilterbankEngine() { b200_fb_plan_destroy(plan); }

void FilterbankEngine::setup(dsp::Filterbank* filterbank) {
  // what CUDA::FilterbankEngine::setup reads (FilterbankCUDA.cu:60-170)
  b200_fb_desc d;
  d.input_real = filterbank->get_input()->get_state() == Signal::Nyquist;
  d.input_nchan = filterbank->get_input()->get_nchan();
  d.npol = filterbank->get_input()->get_npol();
  d.nchan_subband = filterbank->get_nchan_subband();
  d.freq_res = filterbank->get_freq_res();
  d.nfilt_pos = d.nfilt_neg = 0;
  d.h_response = 0;
  d.max_npart = 0;
  if (filterbank->has_response()) {
    const dsp::Response* r = filterbank->get_response();
    if (r->get_ndim() != 2) throw Error(InvalidState, "B200::FilterbankEngine::setup", "matrix responses are not supported");
    d.nfilt_pos = r->get_impulse_pos();
    d.nfilt_neg = r->get_impulse_neg();
    d.h_response = r->get_datptr(0, 0);
  }
  filterbank->set_passband(NULL);   // the engine does not integrate the bandpass (FilterbankCUDA.cu:73-77)
  if (plan) b200_fb_plan_destroy(plan);
  plan = 0;
  check(b200_fb_plan_create(ctx, &d, &plan), "B200::FilterbankEngine::setup");
}

void FilterbankEngine::perform(const dsp::TimeSeries* in, dsp::TimeSeries* out, uint64_t npart,
                               const uint64_t in_step, const uint64_t out_step) {
  check(b200_fb_perform(plan, in->get_datptr(0, 0), in->get_nfloat_span(), out->get_datptr(0, 0),
                        out->get_nfloat_span(), npart, in_step, out_step),
        "B200::FilterbankEngine::perform");
}

void FilterbankEngine::finish() { check(b200_context_synchronize(ctx), "B200::FilterbankEngine::finish"); }

// ---- Convolution --------------------------------------------------------------------------------
ConvolutionEngine::~ConvolutionEngine() { b200_fb_plan_destroy(plan); }

void ConvolutionEngine::prepare(dsp::Convolution* convolution) {
  const dsp::Response* r = convolution->get_response();
  b200_fb_desc d;
  d.input_real = convolution->get_input()->get_state() == Signal::Nyquist;
  d.input_nchan = convolution->get_input()->get_nchan();
  d.npol = convolution->get_input()->get_npol();
  d.nchan_subband = 1;
  d.freq_res = r->get_ndat();
  d.nfilt_pos = r->get_impulse_pos();
  d.nfilt_neg = r->get_impulse_neg();
  d.h_response = r->get_datptr(0, 0);
  d.max_npart = 0;
  if (plan) b200_fb_plan_destroy(plan);
  plan = 0;
  check(b200_fb_plan_create(ctx, &d, &plan), "B200::ConvolutionEngine::prepare");
  b200_fb_info info;
  check(b200_fb_plan_info(plan, &info), "B200::ConvolutionEngine::prepare");
  nsamp_step = info.nsamp_step;
  nkeep = info.nkeep;
  ndim = d.input_real ? 1 : 2;
  if (info.nsamp_fft != convolution->get_minimum_samples())
    throw Error(InvalidState, "B200::ConvolutionEngine::prepare", "nsamp_fft mismatch %u != %u", info.nsamp_fft,
                convolution->get_minimum_samples());
}

void ConvolutionEngine::perform(const dsp::TimeSeries* in, dsp::TimeSeries* out, unsigned npart) {
  check(b200_fb_perform(plan, in->get_datptr(0, 0), in->get_nfloat_span(), out->get_datptr(0, 0),
                        out->get_nfloat_span(), npart, uint64_t(nsamp_step) * ndim, uint64_t(nkeep) * 2),
        "B200::ConvolutionEngine::perform");
}

// ---- Detection ----------------------------------------------------------------------------------
void DetectionEngine::polarimetry(unsigned ndim, const dsp::TimeSeries* in, dsp::TimeSeries* out) {
  // the caller fixes the output state before the call in the out-of-place case (Detection.C:109-110)
  const int state = out->get_state() == Signal::Stokes ? B200_STOKES : B200_COHERENCE;
  check(b200_detect(ctx, state, ndim, in->get_datptr(0, 0), in->get_nfloat_span(), in->get_nchan(), in->get_npol(),
                    in->get_ndat(), out->get_datptr(0, 0), out->get_nfloat_span()),
        "B200::DetectionEngine::polarimetry");
}

void DetectionEngine::square_law(const dsp::TimeSeries* in, dsp::TimeSeries* out) {
  const int state = out->get_npol() == 1 ? B200_INTENSITY : B200_PPQQ;
  check(b200_detect(ctx, state, 1, in->get_datptr(0, 0), in->get_nfloat_span(), in->get_nchan(), in->get_npol(),
                    in->get_ndat(), out->get_datptr(0, 0), out->get_nfloat_span()),
        "B200::DetectionEngine::square_law");
}

// ---- Fold ---------------------------------------------------------------------------------------
FoldEngine::FoldEngine(b200_context* c) : ctx(c), handle(0), nbin(0), ndat_folded(0) {
  use_set_bins = true;   // Fold::fold then calls set_bins once per block instead of set_bin per sample
  // as CUDA::FoldEngine (FoldCUDA.cu:43-44): the accumulating PhaseSeries belongs to the engine and lives on the device
  device_profiles = new dsp::PhaseSeries;
  device_profiles->set_memory(new DeviceMemory(ctx));
  synchronized = true;   // no data on either side yet (FoldCUDA.cu:53-54)
}
FoldEngine::~FoldEngine() { b200_fold_destroy(handle); }

void FoldEngine::set_nbin(unsigned n) { nbin = n; }

void FoldEngine::ensure() {
  setup();   // fills nchan, npol, ndim, input, input_span, output, output_span from the parent (Fold.C:973-1011);
             // the input pointer may change from block to block
  if (handle) return;
  check(b200_fold_create(ctx, nchan, npol, ndim, nbin, &handle), "B200::FoldEngine");
}

void FoldEngine::set_ndat(uint64_t, uint64_t) {}

void FoldEngine::set_bin(uint64_t, double, double) {
  throw Error(InvalidState, "B200::FoldEngine::set_bin", "this engine plans bins itself (use_set_bins)");
}

uint64_t FoldEngine::set_bins(double phi, double phase_per_sample, uint64_t ndat, uint64_t idat0) {
  ensure();
  check(b200_fold_set_bins(handle, phi, phase_per_sample, ndat, idat0, &ndat_folded), "B200::FoldEngine::set_bins");
  last_hits.resize(nbin);
  check(b200_fold_get_bin_hits(handle, last_hits.data()), "B200::FoldEngine::set_bins");
  synchronized = false;
  return ndat_folded;
}

uint64_t FoldEngine::get_bin_hits(int ibin) { return last_hits[ibin]; }

// Never parent->get_output(): the reference's Fold::get_output() IS engine->get_profiles() (Fold.C:88-94)
dsp::PhaseSeries* FoldEngine::get_profiles() { return device_profiles; }

void FoldEngine::fold() {
  // `output` / `output_span` are the engine-owned PhaseSeries' device buffer and plane span (Fold::Engine::setup)
  check(b200_fold_fold_into(handle, input, input_span, output, output_span), "B200::FoldEngine::fold");
}

void FoldEngine::synch(dsp::PhaseSeries* out) {
  if (synchronized) return;   // idempotent, as FoldCUDA.cu:132-147
  // TransferPhaseSeriesCUDA (TransferPhaseSeriesCUDA.C:25-83): match shape, copy attributes (incl. the host-side hits),
  // then one device-to-host copy of the whole buffer
  out->internal_match(device_profiles);
  out->copy_configuration(device_profiles);
  check(b200_memcpy_d2h(ctx, out->internal_get_buffer(), device_profiles->internal_get_buffer(),
                        device_profiles->internal_get_size()),
        "B200::FoldEngine::synch");
  synchronized = true;
}

void FoldEngine::zero() {
  // CUDA::FoldEngine::zero (FoldCUDA.cu): the profiles are zeroed through their Memory (DeviceMemory::do_zero)
  device_profiles->zero();
  if (handle) check(b200_fold_zero(handle), "B200::FoldEngine::zero");
  synchronized = false;
}

}  // namespace B200
