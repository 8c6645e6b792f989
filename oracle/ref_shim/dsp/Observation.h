// ref_shim/dsp/Observation.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for Kernel/Classes/dsp/Observation.h (which
// needs PSRCHIVE's MJD, sky_coord, Types): just the getters Dedispersion.C / Response.C call on their input.
#ifndef REF_SHIM_DSP_OBSERVATION_H
#define REF_SHIM_DSP_OBSERVATION_H
#include "OwnStream.h"
#include "dsp/dsp.h"      // the reference's own (Kernel/Classes/dsp/dsp.h): psrdisp_compatible
namespace dsp {
class Observation : public OwnStream {
 public:
  static bool verbose;
  Observation() : nchan(1), centre_frequency(0), bandwidth(0), dispersion_measure(0), dc_centred(false),
                  dual_sideband(false), swap(false) {}
  unsigned get_nchan() const { return nchan; }
  double get_centre_frequency() const { return centre_frequency; }
  // Observation.C:420-452 (per-channel centre); only copied into Dedispersion::frequency_input by the callers here
  double get_centre_frequency(unsigned ichan) const {
    const double chanwidth = bandwidth / double(nchan);
    return centre_frequency - 0.5 * bandwidth + (double(ichan) + 0.5) * chanwidth;
  }
  double get_bandwidth() const { return bandwidth; }
  double get_dispersion_measure() const { return dispersion_measure; }
  void set_dispersion_measure(double dm) { dispersion_measure = dm; }
  bool get_dc_centred() const { return dc_centred; }
  bool get_dual_sideband() const { return dual_sideband; }
  bool get_swap() const { return swap; }
  unsigned nchan;
  double centre_frequency, bandwidth, dispersion_measure;
  bool dc_centred, dual_sideband, swap;
};
}  // namespace dsp
#endif
