// dsp/standin.h -- MINIMAL header-compatible stand-ins for the dspsr / PSRCHIVE types the engine
// shims touch.  PSRCHIVE and dspsr's own headers cannot be present in this repository (the
// reference needs PSRCHIVE to configure, configure.ac:73-77), so the shims in this directory are
// compiled against these few classes, which expose EXACTLY the accessors the shims use with the
// reference's names and meanings.  In a real dspsr tree delete this header and include the real
// ones; the shim sources do not change.
//
// Mirrored declarations (reference file:line):
//   Reference::Able / To            PSRCHIVE Util/units/Reference*.h (intrusive ref count)
//   Error                           PSRCHIVE Util/units/Error.h
//   dsp::Observation / TimeSeries   Kernel/Classes/dsp/Observation.h, TimeSeries.h, DataSeries.C:246-259
//   dsp::Response                   Signal/General/dsp/Response.h:59-77
//   dsp::Filterbank(+Engine)        Signal/General/dsp/Filterbank.h, FilterbankEngine.h:15-44, Filterbank.C
//   dsp::Convolution(+Engine)       Signal/General/dsp/Convolution.h:158-167, Convolution.C
//   dsp::Detection(+Engine)         Signal/General/dsp/Detection.h:98-106, Detection.C
//   dsp::Fold(+Engine), PhaseSeries Signal/Pulsar/dsp/Fold.h:249-312, Fold.C, PhaseSeries.C
#ifndef B200_DSP_STANDIN_H
#define B200_DSP_STANDIN_H

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

enum ErrorCode { Undefined, InvalidState, InvalidParam, FailedCall };

class Error {
 public:
  Error(ErrorCode c, const char* func, const char* fmt = 0, ...) : code(c), function(func ? func : "") {
    if (fmt) {
      char buf[1024];
      va_list ap;
      va_start(ap, fmt);
      vsnprintf(buf, sizeof(buf), fmt, ap);
      va_end(ap);
      message = buf;
    }
  }
  Error& operator+=(const char* ctx) { function = std::string(ctx) + " <- " + function; return *this; }
  const std::string& get_message() const { return message; }
  ErrorCode code;
  std::string function, message;
};

namespace Reference {
class Able {
 public:
  Able() : refs(0) {}
  virtual ~Able() {}
  mutable int refs;
};
template <class T> class To {
 public:
  To(T* p = 0) : ptr_(0) { set(p); }
  To(const To& o) : ptr_(0) { set(o.ptr_); }
  ~To() { set(0); }
  To& operator=(T* p) { set(p); return *this; }
  To& operator=(const To& o) { set(o.ptr_); return *this; }
  T* operator->() const { return ptr_; }
  T* get() const { return ptr_; }
  T* ptr() const { return ptr_; }
  operator T*() const { return ptr_; }
 private:
  void set(T* p) {
    if (p) p->refs++;
    if (ptr_ && --ptr_->refs == 0) delete ptr_;
    ptr_ = p;
  }
  T* ptr_;
};
}  // namespace Reference

namespace Signal {
enum State { Nyquist, Analytic, Intensity, PPQQ, Coherence, Stokes };
}

namespace dsp {

class Observation : public Reference::Able {
 public:
  Observation() : state(Signal::Nyquist), nchan(1), npol(1), ndim(1), ndat(0), rate(0), scale(1) {}
  Signal::State get_state() const { return state; }
  void set_state(Signal::State s) { state = s; }
  unsigned get_nchan() const { return nchan; }
  unsigned get_npol() const { return npol; }
  unsigned get_ndim() const { return ndim; }
  uint64_t get_ndat() const { return ndat; }
  double get_rate() const { return rate; }
  void set_nchan(unsigned n) { nchan = n; }
  void set_npol(unsigned n) { npol = n; }
  void set_ndim(unsigned n) { ndim = n; }
  void set_rate(double r) { rate = r; }
  void rescale(double f) { scale *= f; }
  double get_scale() const { return scale; }
 protected:
  Signal::State state;
  unsigned nchan, npol, ndim;
  uint64_t ndat;
  double rate, scale;
};

// Device memory manager (dsp::Memory / CUDA::DeviceMemory, Kernel/Classes/MemoryCUDA.C)
class Memory : public Reference::Able {
 public:
  virtual void* do_allocate(size_t nbytes) = 0;
  virtual void do_free(void*) = 0;
  virtual bool on_host() const = 0;
};

// FPT-ordered time series: plane(ichan,ipol) = base + (ichan*npol+ipol)*span floats
class TimeSeries : public Observation {
 public:
  TimeSeries() : buffer(0), span(0), capacity(0) {}
  ~TimeSeries() { if (buffer && memory) memory->do_free(buffer); }
  void set_memory(Memory* m) { memory = m; }
  void resize(uint64_t n) {
    ndat = n;
    span = (n * ndim + 1) / 2 * 2;   // even spans keep every plane 8-byte aligned
    size_t need = size_t(span) * nchan * npol * sizeof(float);
    if (need > capacity) {
      if (buffer) memory->do_free(buffer);
      buffer = static_cast<float*>(memory->do_allocate(need ? need : 8));
      capacity = need;
    }
  }
  float* get_datptr(unsigned ichan, unsigned ipol) { return buffer + (uint64_t(ichan) * npol + ipol) * span; }
  const float* get_datptr(unsigned ichan, unsigned ipol) const { return buffer + (uint64_t(ichan) * npol + ipol) * span; }
  uint64_t get_nfloat_span() const { return span; }
  void copy_configuration(const Observation* o) {
    state = o->get_state(); nchan = o->get_nchan(); npol = o->get_npol(); ndim = o->get_ndim();
    rate = o->get_rate(); scale = o->get_scale();
  }
 protected:
  Reference::To<Memory> memory;
  float* buffer;
  uint64_t span;
  size_t capacity;
};

// Frequency response (host buffer of nchan*ndat complex floats)
class Response : public Reference::Able {
 public:
  Response() : nchan(1), ndat(0), ndim(2), impulse_pos(0), impulse_neg(0) {}
  unsigned get_nchan() const { return nchan; }
  unsigned get_ndat() const { return ndat; }
  unsigned get_ndim() const { return ndim; }
  unsigned get_impulse_pos() const { return impulse_pos; }
  unsigned get_impulse_neg() const { return impulse_neg; }
  const float* get_datptr(unsigned, unsigned) const { return data.data(); }
  void configure(unsigned _nchan, unsigned _ndat, unsigned pos, unsigned neg) {
    nchan = _nchan; ndat = _ndat; impulse_pos = pos; impulse_neg = neg;
    data.assign(size_t(nchan) * ndat * 2, 0.f);
  }
  float* writable() { return data.data(); }
 protected:
  unsigned nchan, ndat, ndim, impulse_pos, impulse_neg;
  std::vector<float> data;
};

// ---------------------------------------------------------------------------------------------
class Filterbank : public Reference::Able {
 public:
  class Engine;
  Filterbank() : nchan(0), freq_res(0), nchan_subband(0), nfilt_pos(0), nfilt_neg(0), nsamp_fft(0),
                 nsamp_overlap(0), nsamp_step(0), prepared(false), passband_set(true) {}
  void set_input(const TimeSeries* i) { input = const_cast<TimeSeries*>(i); }
  void set_output(TimeSeries* o) { output = o; }
  const TimeSeries* get_input() const { return input; }
  void set_nchan(unsigned n) { nchan = n; }
  unsigned get_nchan() const { return nchan; }
  unsigned get_freq_res() const { return freq_res; }
  unsigned get_nchan_subband() const { return nchan_subband; }
  void set_response(Response* r) { response = r; }
  bool has_response() const { return response; }
  const Response* get_response() const { return response; }
  void set_passband(Response* p) { passband_set = (p != 0); }
  void set_engine(Engine* e);
  void prepare();          // Filterbank::make_preparations (Filterbank.C:55-263), engine branch
  void operate();          // Filterbank::transformation + filterbank() engine branch (:432-553)
  unsigned get_minimum_samples() const { return nsamp_fft; }
  unsigned get_nsamp_step() const { return nsamp_step; }
  unsigned get_nsamp_overlap() const { return nsamp_overlap; }
 protected:
  Reference::To<TimeSeries> input, output;
  Reference::To<Response> response;
  Reference::To<Engine> engine;
  unsigned nchan, freq_res, nchan_subband, nfilt_pos, nfilt_neg, nsamp_fft, nsamp_overlap, nsamp_step;
  bool prepared, passband_set;
};

class Filterbank::Engine : public Reference::Able {   // FilterbankEngine.h:15-44
 public:
  Engine() { scratch = output = 0; output_span = 0; }
  virtual void setup(Filterbank*) = 0;
  virtual void set_scratch(float*) = 0;
  virtual void perform(const dsp::TimeSeries* in, dsp::TimeSeries* out, uint64_t npart, const uint64_t in_step,
                       const uint64_t out_step) = 0;
  virtual void finish() {}
 protected:
  float* scratch;
  float* output;
  unsigned output_span;
};

// ---------------------------------------------------------------------------------------------
class Convolution : public Reference::Able {
 public:
  class Engine;
  Convolution() : n_fft(0), nfilt_pos(0), nfilt_neg(0), nsamp_fft(0), nsamp_overlap(0), nsamp_step(0), npart(0),
                  prepared(false) {}
  void set_input(const TimeSeries* i) { input = const_cast<TimeSeries*>(i); }
  void set_output(TimeSeries* o) { output = o; }
  const TimeSeries* get_input() const { return input; }
  void set_response(Response* r) { response = r; }
  const Response* get_response() const { return response; }
  bool has_response() const { return response; }
  unsigned get_minimum_samples() const { return nsamp_fft; }
  unsigned get_minimum_samples_lost() const { return nsamp_overlap; }
  void set_engine(Engine* e);
  void prepare();          // Convolution::prepare (Convolution.C:105-221), engine branch
  void operate();          // Convolution::transformation (:338-365)
 protected:
  Reference::To<TimeSeries> input, output;
  Reference::To<Response> response;
  Reference::To<Engine> engine;
  unsigned n_fft, nfilt_pos, nfilt_neg, nsamp_fft, nsamp_overlap, nsamp_step;
  uint64_t npart;
  bool prepared;
};

class Convolution::Engine : public Reference::Able {  // Convolution.h:158-167
 public:
  virtual void set_scratch(void*) = 0;
  virtual void prepare(dsp::Convolution* convolution) = 0;
  virtual void perform(const TimeSeries* in, TimeSeries* out, unsigned npart) = 0;
};

// ---------------------------------------------------------------------------------------------
class Detection : public Reference::Able {
 public:
  class Engine;
  Detection() : state(Signal::Intensity), ndim(1) {}
  void set_input(const TimeSeries* i) { input = const_cast<TimeSeries*>(i); }
  void set_output(TimeSeries* o) { output = o; }
  void set_output_state(Signal::State s) { state = s; }
  void set_output_ndim(unsigned n) { ndim = n; }
  void set_engine(Engine* e);
  void operate();          // Detection::transformation (Detection.C:74-147), engine branch
 protected:
  Reference::To<TimeSeries> input, output;
  Reference::To<Engine> engine;
  Signal::State state;
  unsigned ndim;
};

class Detection::Engine : public Reference::Able {    // Detection.h:98-106
 public:
  virtual void polarimetry(unsigned ndim, const TimeSeries* in, TimeSeries* out) = 0;
  virtual void square_law(const dsp::TimeSeries* input, dsp::TimeSeries* output) = 0;
};

// ---------------------------------------------------------------------------------------------
class PhaseSeries : public TimeSeries {
 public:
  PhaseSeries() : integration_length(0), ndat_total(0) {}
  void resize_bins(unsigned nbin) { resize(nbin); hits.assign(nbin, 0u); }
  unsigned get_nbin() const { return unsigned(ndat); }
  unsigned* get_hits() { return hits.data(); }
  unsigned get_hits_nchan() const { return 1; }
  double integration_length;
  uint64_t ndat_total;
 protected:
  std::vector<unsigned> hits;
};

class Fold : public Reference::Able {
 public:
  class Engine;
  Fold() : folding_nbin(0), idat_start(0), ndat_fold(0), phi(0), phase_per_sample(0) {}
  void set_input(const TimeSeries* i) { input = const_cast<TimeSeries*>(i); }
  const TimeSeries* get_input() const { return input; }
  void set_output(PhaseSeries* o) { output = o; }
  void set_nbin(unsigned n) { folding_nbin = n; }
  // the phase of the midpoint of the first sample and the phase advance per sample, i.e. the
  // results of get_phi / get_pfold that Fold::fold computes from the predictor (Fold.C:650-657,720)
  void set_phase(double _phi, double _pps) { phi = _phi; phase_per_sample = _pps; }
  void set_engine(Engine* e);
  void operate();          // Fold::transformation + fold() engine branch (Fold.C:510-604,724-829)
  PhaseSeries* get_result();  // Fold::get_result -> engine->synch (Fold.C:123-135)
  PhaseSeries* get_output() { return output; }
 protected:
  Reference::To<TimeSeries> input;
  Reference::To<PhaseSeries> output;
  Reference::To<Engine> engine;
  unsigned folding_nbin;
  uint64_t idat_start, ndat_fold;
  double phi, phase_per_sample;
};

class Fold::Engine : public Reference::Able {         // Fold.h:249-312
 public:
  Engine() : use_set_bins(false), output(0), output_span(0), input(0), input_span(0), hits(0), hits_nchan(0),
             zeroed_samples(false), ndat_fold(0), idat_start(0), nchan(0), npol(0), ndim(0), parent(0),
             synchronized(false) {}
  void set_parent(Fold* f) { parent = f; }
  virtual void set_nbin(unsigned nbin) = 0;
  virtual void set_bin(uint64_t idat, double ibin, double bins_per_samp) = 0;
  virtual uint64_t set_bins(double phi, double phase_per_sample, uint64_t _ndat, uint64_t idat_start) = 0;
  bool use_set_bins;
  virtual uint64_t get_bin_hits(int ibin) = 0;
  virtual uint64_t get_ndat_folded() const = 0;
  virtual PhaseSeries* get_profiles() = 0;
  virtual void fold() = 0;
  virtual void synch(PhaseSeries*) = 0;
  virtual void zero() = 0;
  virtual void set_ndat(uint64_t, uint64_t) {}
 protected:
  float* output;
  unsigned output_span;
  const float* input;
  unsigned input_span;
  unsigned* hits;
  unsigned hits_nchan;
  bool zeroed_samples;
  unsigned ndat_fold;
  uint64_t idat_start;
  unsigned nchan, npol, ndim;
  void setup();            // Fold::Engine::setup (Fold.C:973-1011)
  Fold* parent;
  bool synchronized;
  friend class Fold;
};

}  // namespace dsp
#endif
