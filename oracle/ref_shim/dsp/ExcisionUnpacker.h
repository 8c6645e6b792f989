// ref_shim/dsp/ExcisionUnpacker.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for Kernel/Classes/dsp/ExcisionUnpacker.h
// (whose base classes pull in the whole Transformation / TimeSeries / Input tree): the data members and getters
// that the member template of dsp/excision_unpack.h reads, so that the reference's template body compiles in place.
#ifndef REF_SHIM_DSP_EXCISIONUNPACKER_H
#define REF_SHIM_DSP_EXCISIONUNPACKER_H
#include <stdint.h>
#include "Error.h"
namespace dsp {
class ExcisionUnpacker {
 public:
  ExcisionUnpacker() : verbose(false), ndim_per_digitizer(1), ndat_per_weight(512), output_incr(1), nlow_min(0), nlow_max(0) {}
  unsigned get_ndim_per_digitizer() const { return ndim_per_digitizer; }
  unsigned get_ndat_per_weight() const { return ndat_per_weight; }
  unsigned get_output_incr() const { return output_incr; }
  template <class U, class Iter>
  void excision_unpack(U& unpack, Iter& input_data, float* output_data, uint64_t ndat, unsigned long* hist,
                       unsigned* weights, unsigned nweights);
  bool verbose;
  unsigned ndim_per_digitizer, ndat_per_weight, output_incr;
  unsigned nlow_min, nlow_max;
};
}  // namespace dsp
#endif
