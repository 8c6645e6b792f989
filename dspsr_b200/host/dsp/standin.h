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
//   dsp::BitSeries, MeerKATUnpacker / UWBUnpacker (+Engine)  Kernel/Formats/kat/dsp/MeerKATUnpacker.h:72-85,
//                                   Kernel/Formats/uwb/dsp/UWBUnpacker.h (nested Engine), MeerKATUnpacker.C:196-206
//   MJD                             PSRCHIVE Util/genutil/MJD.h (day / second / fraction split)
#ifndef B200_DSP_STANDIN_H
#define B200_DSP_STANDIN_H

#include <cmath>
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

// split epoch as PSRCHIVE's MJD keeps it: integer day, integer second of day, fractional second
class MJD {
 public:
  MJD(int d = 0, int s = 0, double f = 0.0) : days(d), secs(s), fracsec(f) { settle(); }
  MJD operator+(double seconds) const {
    const double whole = std::floor(seconds);
    MJD r(days, secs, fracsec + (seconds - whole));
    long long s = (long long)r.secs + (long long)whole;
    long long dd = s >= 0 ? s / 86400 : -((-s + 86399) / 86400);
    r.days += int(dd);
    r.secs = int(s - dd * 86400);
    return r;
  }
  double operator-(const MJD& o) const { return double(days - o.days) * 86400.0 + double(secs - o.secs) + (fracsec - o.fracsec); }
  bool operator<(const MJD& o) const { return (*this - o) < 0; }
  bool operator==(const MJD& o) const { return days == o.days && secs == o.secs && fracsec == o.fracsec; }
  int intday() const { return days; }
  int get_secs() const { return secs; }
  double get_fracsec() const { return fracsec; }
 private:
  void settle() {
    while (fracsec >= 1.0) { fracsec -= 1.0; secs++; }
    while (fracsec < 0.0) { fracsec += 1.0; secs--; }
    while (secs >= 86400) { secs -= 86400; days++; }
    while (secs < 0) { secs += 86400; days--; }
  }
  int days, secs;
  double fracsec;
};

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
  MJD get_start_time() const { return start_time; }
  void set_start_time(const MJD& t) { start_time = t; }
  MJD get_end_time() const { return start_time + double(ndat) / rate; }      // Observation.C get_end_time
  const std::string& get_machine() const { return machine; }
  void set_machine(const std::string& m) { machine = m; }
 protected:
  MJD start_time;
  std::string machine;
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
  virtual void do_zero(void* ptr, size_t nbytes) = 0;                       // Memory.h / MemoryCUDA.C:70-82
  virtual void do_copy(void* to, const void* from, size_t nbytes) = 0;      // MemoryCUDA.C:90-106
  virtual bool on_host() const = 0;
};

// Raw (packed) data of one block, on the device when the unpacker runs there (Kernel/Classes/dsp/BitSeries.h)
class BitSeries : public Observation {
 public:
  BitSeries() : raw(0), nbit(8) {}
  const unsigned char* get_rawptr() const { return raw; }
  void set_rawptr(const unsigned char* p, uint64_t n) { raw = p; ndat = n; }
  unsigned get_nbit() const { return nbit; }
  void set_nbit(unsigned n) { nbit = n; }
 protected:
  const unsigned char* raw;
  unsigned nbit;
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
  virtual void copy_configuration(const Observation* o) {
    state = o->get_state(); nchan = o->get_nchan(); npol = o->get_npol(); ndim = o->get_ndim();
    rate = o->get_rate(); scale = o->get_scale(); start_time = o->get_start_time(); machine = o->get_machine();
  }
  void zero() { if (buffer) memory->do_zero(buffer, size_t(span) * nchan * npol * sizeof(float)); }   // DataSeries::zero
  unsigned char* internal_get_buffer() { return reinterpret_cast<unsigned char*>(buffer); }            // DataSeries.h:103-107
  const unsigned char* internal_get_buffer() const { return reinterpret_cast<const unsigned char*>(buffer); }
  uint64_t internal_get_size() const { return uint64_t(span) * nchan * npol * sizeof(float); }
  void internal_match(const TimeSeries* o) {                                 // TimeSeries.h:88: same shape and span
    nchan = o->nchan; npol = o->npol; ndim = o->ndim;
    resize(o->ndat);
  }
  bool get_zeroed_data() const { return false; }
  Memory* get_memory() const { return memory; }
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
class PhaseSeries : public TimeSeries {                       // Signal/Pulsar/dsp/PhaseSeries.h, PhaseSeries.C
 public:
  PhaseSeries() : integration_length(0), ndat_total(0) {}
  unsigned get_nbin() const { return unsigned(ndat); }
  unsigned* get_hits() { return hits.data(); }
  const unsigned* get_hits() const { return hits.data(); }
  unsigned get_hits_nchan() const { return 1; }
  MJD get_end_time() const { return end_time; }
  //! PhaseSeries::mixable (PhaseSeries.C:336-418): adopt the observation when empty, else widen the time span
  bool mixable(const Observation& obs, unsigned nbin, int64_t istart = 0, int64_t fold_ndat = 0);
  //! PhaseSeries::combine (PhaseSeries.C:442-480); host memory only
  void combine(const PhaseSeries* prof);
  //! PhaseSeries::zero (PhaseSeries.C:239-256)
  void zero() { integration_length = 0; ndat_total = 0; hits.assign(hits.size(), 0u); TimeSeries::zero(); }
  //! PhaseSeries::copy_configuration + copy_attributes (PhaseSeries.C:258-316): host hits travel with the attributes
  void copy_configuration(const Observation* o) {
    TimeSeries::copy_configuration(o);
    const PhaseSeries* like = dynamic_cast<const PhaseSeries*>(o);
    if (like) {
      integration_length = like->integration_length;
      ndat_total = like->ndat_total;
      end_time = like->end_time;
      hits = like->hits;
    }
  }
  double integration_length;
  uint64_t ndat_total;
  void resize_bins(unsigned nbin) { resize(nbin); hits.assign(nbin, 0u); }
 protected:
  MJD end_time;
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
  PhaseSeries* get_output() const;   // Fold.C:88-94: the ENGINE's PhaseSeries whenever an engine is set
  void reset();                      // Fold.C:137-148
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

// ---------------------------------------------------------------------------------------------
// Unpacker device hook (Kernel/Classes/dsp/Unpacker.h:57-61,100).  The two formats of the BASELINE configurations that
// have a nested Engine interface in the reference: MeerKATUnpacker::Engine (kat/dsp/MeerKATUnpacker.h:72-85) and
// UWBUnpacker::Engine (uwb/dsp/UWBUnpacker.h).  operate() is the engine branch of their unpack() (MeerKATUnpacker.C:
// 196-206, UWBUnpacker.C:163-168).
class Unpacker : public Reference::Able {
 public:
  void set_input(const BitSeries* i) { input = const_cast<BitSeries*>(i); }
  void set_output(TimeSeries* o) { output = o; }
 protected:
  void prepare_output(unsigned ndim_out) {                   // Unpacker::resize_output (Unpacker.C): FPT floats
    output->copy_configuration(input);
    output->set_ndim(ndim_out);
    output->resize(input->get_ndat());
  }
  Reference::To<BitSeries> input;
  Reference::To<TimeSeries> output;
};

class MeerKATUnpacker : public Unpacker {
 public:
  class Engine;
  MeerKATUnpacker() : table_scale(0) {}
  void set_table_scale(double s) { table_scale = s; }        // BitTable(8, TwosComplement).get_scale() (MeerKATUnpacker.C:36)
  void set_engine(Engine* e);
  bool get_device_supported(Memory* m) const;
  void set_device(Memory* m);
  void operate();
 protected:
  Reference::To<Engine> engine;
  double table_scale;
};
class MeerKATUnpacker::Engine : public Reference::Able {
 public:
  virtual void setup() = 0;
  virtual void unpack(float scale, const BitSeries* input, TimeSeries* output, unsigned sample_swap) = 0;
  virtual bool get_device_supported(Memory* memory) const = 0;
  virtual void set_device(Memory* memory) = 0;
};

class UWBUnpacker : public Unpacker {
 public:
  class Engine;
  void set_engine(Engine* e);
  bool get_device_supported(Memory* m) const;
  void set_device(Memory* m);
  void operate();
 protected:
  Reference::To<Engine> engine;
};
class UWBUnpacker::Engine : public Reference::Able {
 public:
  virtual void unpack(const BitSeries* input, TimeSeries* output) = 0;
  virtual bool get_device_supported(Memory* memory) const = 0;
  virtual void set_device(Memory* memory) = 0;
  virtual void setup() = 0;
};

}  // namespace dsp
#endif
