// B200Engines.h -- the engine shims: subclasses of the reference's own plugin interfaces that
// translate dsp::TimeSeries / Response / PhaseSeries into the POD descriptors of the C ABI
// (include/b200dsp.h) and `throw Error` on a non-zero status.  These four classes are what a
// maintainer adds to a dspsr tree; each is ~50 lines.
#ifndef B200_ENGINES_H
#define B200_ENGINES_H

#include "../../include/b200dsp.h"
#ifndef B200_REAL_DSPSR
#include "dsp/standin.h"
#else
#include "dsp/FilterbankEngine.h"
#include "dsp/Convolution.h"
#include "dsp/Detection.h"
#include "dsp/Fold.h"
#include "dsp/MemoryCUDA.h"
#include "dsp/MeerKATUnpacker.h"
#include "dsp/UWBUnpacker.h"
#endif

namespace B200 {

void check(int status, const char* where);
//! The context of the pipeline thread that owns `cuda_stream` (CUDA::DeviceMemory::get_stream()); created on first use
b200_context* context_for(void* cuda_stream, int device = 0);   // status != 0 -> throw Error(FailedCall, where, b200_last_error())

//! dsp::Memory on the context's device (stand-in for CUDA::DeviceMemory, MemoryCUDA.C)
class DeviceMemory : public dsp::Memory {
 public:
  explicit DeviceMemory(b200_context* c) : ctx(c) {}
  void* do_allocate(size_t nbytes);
  void do_free(void*);
  void do_zero(void* ptr, size_t nbytes);                     // MemoryCUDA.C:70-82 (stream ordered)
  void do_copy(void* to, const void* from, size_t nbytes);    // MemoryCUDA.C:90-106 (device to device)
  bool on_host() const { return false; }
  b200_context* get_context() const { return ctx; }
 protected:
  b200_context* ctx;
};

//! Device hook of dsp::MeerKATUnpacker (kat/dsp/MeerKATUnpacker.h:72-85); replaces CUDA::MeerKATUnpackerEngine
class MeerKATUnpackerEngine : public dsp::MeerKATUnpacker::Engine {
 public:
  explicit MeerKATUnpackerEngine(b200_context* c) : ctx(c) {}
  void setup() {}
  void unpack(float scale, const dsp::BitSeries* input, dsp::TimeSeries* output, unsigned sample_swap);
  bool get_device_supported(dsp::Memory* memory) const;
  void set_device(dsp::Memory* memory);
 protected:
  b200_context* ctx;
};

//! Device hook of dsp::UWBUnpacker (uwb/dsp/UWBUnpacker.h nested Engine); replaces CUDA::UWBUnpackerEngine
class UWBUnpackerEngine : public dsp::UWBUnpacker::Engine {
 public:
  explicit UWBUnpackerEngine(b200_context* c) : ctx(c) {}
  void setup() {}
  void unpack(const dsp::BitSeries* input, dsp::TimeSeries* output);
  bool get_device_supported(dsp::Memory* memory) const;
  void set_device(dsp::Memory* memory);
 protected:
  b200_context* ctx;
};

//! Replaces CUDA::FilterbankEngine (Signal/General/FilterbankCUDA.cu)
class FilterbankEngine : public dsp::Filterbank::Engine {
 public:
  explicit FilterbankEngine(b200_context* c) : ctx(c), plan(0) {}
  ~FilterbankEngine();
  void setup(dsp::Filterbank*);
  void set_scratch(float*) {}   // the plan owns its scratch (never allocates in perform)
  void perform(const dsp::TimeSeries* in, dsp::TimeSeries* out, uint64_t npart, const uint64_t in_step,
               const uint64_t out_step);
  void finish();
 protected:
  b200_context* ctx;
  b200_fb_plan* plan;
};

//! Replaces CUDA::ConvolutionEngine / ConvolutionEngineSpectral (ConvolutionCUDA*.cu)
class ConvolutionEngine : public dsp::Convolution::Engine {
 public:
  explicit ConvolutionEngine(b200_context* c) : ctx(c), plan(0), nsamp_step(0), nkeep(0), ndim(1) {}
  ~ConvolutionEngine();
  void set_scratch(void*) {}
  void prepare(dsp::Convolution* convolution);
  void perform(const dsp::TimeSeries* in, dsp::TimeSeries* out, unsigned npart);
 protected:
  b200_context* ctx;
  b200_fb_plan* plan;
  unsigned nsamp_step, nkeep, ndim;
};

//! Replaces CUDA::DetectionEngine (DetectionCUDA.cu); unlike it, supports ndim 1, 2 and 4
class DetectionEngine : public dsp::Detection::Engine {
 public:
  explicit DetectionEngine(b200_context* c) : ctx(c) {}
  void polarimetry(unsigned ndim, const dsp::TimeSeries* in, dsp::TimeSeries* out);
  void square_law(const dsp::TimeSeries* in, dsp::TimeSeries* out);
 protected:
  b200_context* ctx;
};

//! Replaces CUDA::FoldEngine (Signal/Pulsar/FoldCUDA.cu).  use_set_bins = true: the bin plan is
//! made by the library (exact reproduction of the host recurrence), not sample by sample.
class FoldEngine : public dsp::Fold::Engine {
 public:
  explicit FoldEngine(b200_context* c);
  ~FoldEngine();
  void set_nbin(unsigned nbin);
  void set_ndat(uint64_t ndat, uint64_t idat_start);
  void set_bin(uint64_t idat, double ibin, double bins_per_samp);
  uint64_t set_bins(double phi, double phase_per_sample, uint64_t ndat, uint64_t idat_start);
  uint64_t get_bin_hits(int ibin);
  uint64_t get_ndat_folded() const { return ndat_folded; }
  dsp::PhaseSeries* get_profiles();
  void fold();
  void synch(dsp::PhaseSeries*);
  void zero();
 protected:
  void ensure();
  b200_context* ctx;
  b200_fold* handle;
  unsigned nbin;
  uint64_t ndat_folded;
  std::vector<unsigned> last_hits;
  //! The engine-owned accumulating PhaseSeries on the device (what CUDA::FoldEngine calls d_profiles, FoldCUDA.cu:43-44):
  //! Fold::get_output() returns it, Fold::transformation sizes and zeroes it, fold() accumulates into its buffer.
  Reference::To<dsp::PhaseSeries> device_profiles;
};

}  // namespace B200
#endif
