// oracle/ref_shim/ref_cxx.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" doors into the reference's OWN C++ code for the hot path, compiled in place from /root/reference by
// oracle/ref.mk (no reference source is copied here; this file only constructs the reference's classes and calls
// their methods).  What is real reference code behind each door:
//   ref_bittable*        dsp::BitTable::generate_unique_values / generate / get_scale   (Kernel/Classes/BitTable.C)
//   ref_twobit_*         dsp::TwoBitTable (TwoBitTable.C), dsp::TwoBitLookup::lookup_build (TwoBitLookup.C),
//                        dsp::TwoBitFour::{lookup_build,nlow_build,prepare,unpack} (TwoBitFour.C, dsp/TwoBitFour.h),
//                        the body of ExcisionUnpacker::excision_unpack (dsp/excision_unpack.h), StepIterator
//   ref_dedispersion     dsp::Dedispersion::{prepare,smearing_samples,build,match} (Signal/General/Dedispersion.C),
//                        dsp::Response::{match,doswap,set_optimal_ndat,check_ndat,get_minimum_ndat} (Response.C),
//                        dsp::Shape::{resize,rotate} (Shape.C), optimal_fft_length (optimize_fft.c)
//   ref_response_operate dsp::Response::operate (Response.C:385-444)
// What is NOT reference code: the PSRCHIVE stand-ins under ref_shim/ (Reference, Error, OwnStream, Callback, Jones,
// ThreadContext, NormalDistribution, JenetAnderson98 -- the last two carry the restated third-party arithmetic) and the
// two dsp:: stand-in headers ref_shim/dsp/Observation.h and ref_shim/dsp/ExcisionUnpacker.h (plain data holders).
#include <cstring>
#include <vector>

#include "dsp/BitTable.h"
#include "dsp/TwoBitTable.h"
#include "dsp/TwoBitFour.h"
#include "dsp/StepIterator.h"
#include "dsp/excision_unpack.h"
#include "JenetAnderson98.h"

#include "dsp/Dedispersion.h"
#include "dsp/Observation.h"
#include "dsp/OptimalFFT.h"

// ---- never-called members of reference classes whose own .C files cannot be compiled here (link closure only) ----
bool dsp::Observation::verbose = false;
bool dsp::OptimalFFT::verbose = false;
dsp::OptimalFFT::OptimalFFT() : nchan(1), simultaneous(false) {}
void dsp::OptimalFFT::set_simultaneous(bool flag) { simultaneous = flag; }
void dsp::OptimalFFT::set_nchan(unsigned n) { nchan = n; }
unsigned dsp::OptimalFFT::get_nfft(unsigned) const { throw Error(InvalidState, "ref_shim", "OptimalFFT not available"); }
double dsp::OptimalFFT::compute_cost(unsigned, unsigned) const { return 0; }
std::string dsp::OptimalFFT::get_library(unsigned) { return ""; }
FTransform::Bench* dsp::OptimalFFT::new_bench() const { return 0; }

namespace {
// exposes the protected tables of the reference's unpacker
class FourProbe : public dsp::TwoBitFour {
 public:
  const float* base() const { return lookup_base; }
  const char* nlow_table() const { return nlow_lookup; }
  unsigned lo() const { return nlow_min; }
  unsigned hi() const { return nlow_max; }
};
struct RefTwoBit {
  dsp::TwoBitTable* table;
  FourProbe unpacker;
  JenetAnderson98 ja98;
  unsigned ndat_per_weight, nlow_min, nlow_max;
};
}  // namespace

extern "C" {

// type: 0 OffsetBinary, 1 TwosComplement, 2 SignMagnitude (dsp::BitTable::Type order)
double ref_bittable_unique_values(unsigned nbit, int type, float* values) {
  dsp::BitTable t(nbit, dsp::BitTable::Type(type));
  t.generate_unique_values(values);
  return t.get_scale();
}
// the full byte -> floats table (256 * 8/nbit floats) and get_scale()
double ref_bittable_generate(unsigned nbit, int type, float* table) {
  dsp::BitTable t(nbit, dsp::BitTable::Type(type));
  t.generate(table);
  return t.get_scale();
}

void* ref_twobit_create(int type, double threshold, unsigned nlow_min, unsigned nlow_max, unsigned ndat_per_weight,
                        unsigned ndim) {
  RefTwoBit* r = new RefTwoBit();
  r->table = new dsp::TwoBitTable(dsp::BitTable::Type(type));
  r->ja98.set_threshold(threshold);
  r->ndat_per_weight = ndat_per_weight;
  r->nlow_min = nlow_min;
  r->nlow_max = nlow_max;
  // TwoBitCorrection::build (TwoBitCorrection.C:117-131)
  r->unpacker.set_nlow_min(nlow_min);
  r->unpacker.set_nlow_max(nlow_max);
  r->unpacker.set_ndat(ndat_per_weight);
  r->unpacker.set_ndim(ndim);
  r->unpacker.lookup_build(r->table, &r->ja98);
  return r;
}
void ref_twobit_destroy(void* h) {
  RefTwoBit* r = static_cast<RefTwoBit*>(h);
  delete r->table;
  delete r;
}
// copies the (nlow_max - nlow_min + 1) x 1024 lookup rows and the 256-entry low-state counts
void ref_twobit_tables(void* h, float* lookup, char* nlow_lookup) {
  RefTwoBit* r = static_cast<RefTwoBit*>(h);
  std::memcpy(lookup, r->unpacker.base(), sizeof(float) * 1024 * (r->nlow_max - r->nlow_min + 1));
  std::memcpy(nlow_lookup, r->unpacker.nlow_table(), 256);
}
// ExcisionUnpacker::unpack (ExcisionUnpacker.C:173-256) for real-sampled data, one digitizer per polarisation, bytes of
// the polarisations interleaved (get_input_offset = idig, get_input_incr = npol; :259-277), driving the reference's
// excision_unpack template + TwoBitFour through TwoBitCorrection::dig_unpack's two lines (TwoBitCorrection.C:137-151).
// weights: [npol][nweights], caller-initialised; the cross-polarisation mask (WeightedTimeSeries::mask_weights) is
// left to the caller.  Returns 0, or -1 if the reference threw.
int ref_twobit_unpack(void* h, const unsigned char* raw, uint64_t ndat, unsigned npol, float* out, uint64_t span,
                      unsigned* weights, unsigned nweights) {
  RefTwoBit* r = static_cast<RefTwoBit*>(h);
  dsp::ExcisionUnpacker eu;
  eu.ndim_per_digitizer = 1;
  eu.ndat_per_weight = r->ndat_per_weight;
  eu.output_incr = 1;
  eu.nlow_min = r->nlow_min;
  eu.nlow_max = r->nlow_max;
  try {
    for (unsigned idig = 0; idig < npol; idig++) {
      StepIterator<const unsigned char> iterator(raw + idig);
      iterator.set_increment(npol);
      eu.excision_unpack(r->unpacker, iterator, out + uint64_t(idig) * span, ndat, 0,
                         weights ? weights + uint64_t(idig) * nweights : 0, nweights);
    }
  } catch (Error& e) {
    return -1;
  }
  return 0;
}

// Dedispersion::match on an Observation with the given fields; H receives nchan*ndat complex floats if non-null.
// frequency_resolution 0 = the reference's optimal choice.  Returns 0, or -1 if the reference threw (message to stderr).
int ref_dedispersion(double centre_frequency, double bandwidth, double dm, unsigned input_nchan, unsigned nchan,
                     int dual_sideband, int dc_centred, int swap, unsigned frequency_resolution, unsigned* impulse_pos,
                     unsigned* impulse_neg, unsigned* ndat, float* H, uint64_t H_floats) {
  try {
    dsp::Observation obs;
    obs.nchan = input_nchan;
    obs.centre_frequency = centre_frequency;
    obs.bandwidth = bandwidth;
    obs.dispersion_measure = dm;
    obs.dual_sideband = dual_sideband != 0;
    obs.dc_centred = dc_centred != 0;
    obs.swap = swap != 0;
    dsp::Dedispersion kernel;
    if (frequency_resolution) kernel.set_frequency_resolution(frequency_resolution);
    kernel.match(&obs, nchan);
    *impulse_pos = kernel.get_impulse_pos();
    *impulse_neg = kernel.get_impulse_neg();
    *ndat = kernel.get_ndat();
    const uint64_t n = uint64_t(kernel.get_nchan()) * kernel.get_ndat() * 2;
    if (H) {
      if (n > H_floats) return -2;
      std::memcpy(H, kernel.get_datptr(0, 0), n * sizeof(float));
    }
  } catch (Error& e) {
    std::fprintf(stderr, "ref_dedispersion: %s: %s\n", e.function.c_str(), e.message.c_str());
    return -1;
  }
  return 0;
}

// Response::operate (spectrum *= H) on npts complex points of one channel-less response
int ref_response_operate(const float* H, unsigned npts, float* spectrum) {
  try {
    std::vector<std::complex<float> > ph(npts);
    for (unsigned i = 0; i < npts; i++) ph[i] = std::complex<float>(H[2 * i], H[2 * i + 1]);
    dsp::Response resp;
    resp.set(ph);
    resp.operate(spectrum, 0u, -1);
  } catch (Error& e) {
    std::fprintf(stderr, "ref_response_operate: %s: %s\n", e.function.c_str(), e.message.c_str());
    return -1;
  }
  return 0;
}

}  // extern "C"
