// ref_fold.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" door to the loops of dsp::Fold::fold
// (Signal/Pulsar/Fold.C) compiled FROM THE REFERENCE'S OWN TEXT: oracle/ref.mk cuts the three statement blocks
//   Fold.C:687-716   weight set-up            (iweight, idat_nextweight, first bad window)
//   Fold.C:744-788   per-sample loop          (weight walk, phase recurrence, bin plan, hits, ndat_folded)
//   Fold.C:835-873   OrderFPT accumulation    (phdimp[idim] += timep[idim], zeroed-sample hit counting)
// and the bodies of dsp::WeightedTimeSeries::convolve_weights / scrunch_weights (Kernel/Classes/WeightedTimeSeries.C:584-696,
// 705-774)
// out of the files where they lie under /root/reference into oracle/_ref/gen/*.inc (git-ignored build products; no
// reference text enters this repository) and this harness supplies the member variables and the two accessors the
// blocks touch.  Fold.C as a whole needs the PSRCHIVE predictor / ephemeris class trees and cannot be compiled here.
#include <assert.h>
#include <math.h>
#include <stdint.h>

#include <iostream>

#include "Error.h"

using std::cerr;
using std::endl;

namespace {

struct TimeSeries {
  enum Order { OrderFPT, OrderTFP };
  const float* base;
  uint64_t span;
  unsigned npol;
  Order get_order() const { return OrderFPT; }
  const float* get_datptr(unsigned ichan, unsigned ipol) const { return base + (uint64_t(ichan) * npol + ipol) * span; }
};

struct PhaseSeriesOut {
  float* base;
  unsigned npol, nbin, ndim, hits_nchan;
  float* get_datptr(unsigned ichan, unsigned ipol) { return base + (uint64_t(ichan) * npol + ipol) * nbin * ndim; }
  unsigned get_hits_nchan() const { return hits_nchan; }
};

struct NoEngine {
  void set_bin(uint64_t, double, double) {}
};

}  // namespace

// Returns ndat_folded; UINT64_MAX where the reference throws (iweight >= nweights).  binplan [ndat_fold], hits [nbin]
// (+=), profile [nchan][npol][nbin][ndim] (+=), discarded (nullable) receives discarded_weights.
extern "C" uint64_t ref_fold(double phi, double phase_per_sample, unsigned folding_nbin, uint64_t idat_start, uint64_t ndat_fold,
                             const unsigned* weights, uint64_t nweights, unsigned ndatperweight, unsigned weight_idat,
                             int zeroed, const float* in_base, uint64_t in_span, unsigned nchan, unsigned npol, unsigned ndim,
                             unsigned* binplan, unsigned* hits, float* profile, unsigned* discarded) {
  const bool verbose = false;
  const uint64_t id = 0;
  const uint64_t idat_end = idat_start + ndat_fold;
  uint64_t iweight = 0, idat_nextweight = 0;
  unsigned bad_weights = 0, tot_weights = 0;
  uint64_t discarded_weights = 0;
  uint64_t ndat_folded = 0, ndat_not_folded = 0;
  const bool zeroed_samples = zeroed != 0;
  bool bad_data = false;
  const double double_nbin = double(folding_nbin);
  NoEngine* engine = 0;
  try {
#include "gen/fold_weights.inc"
  } catch (Error&) {
    return UINT64_MAX;
  }
  (void)id; (void)verbose;
#include "gen/fold_binplan.inc"
  if (discarded) *discarded = unsigned(discarded_weights);
  if (in_base && profile) {
    TimeSeries in_obj = {in_base, in_span, npol};
    PhaseSeriesOut out_obj = {profile, npol, folding_nbin, ndim, 1};
    const TimeSeries* in = &in_obj;
    PhaseSeriesOut* result = &out_obj;
    PhaseSeriesOut* output = &out_obj;
#include "gen/fold_accum.inc"
  }
  (void)ndat_not_folded; (void)bad_weights; (void)tot_weights;
  return ndat_folded;
}

// ---- WeightedTimeSeries::convolve_weights / scrunch_weights: the reference's function bodies inside a holder of the
//      members they touch ----
#ifndef DEBUG
#define DEBUG(x)
#endif
#define UI64 "%lu"
namespace {
struct WeightsHolder {
  unsigned* weights;
  uint64_t nweights_, ndat_;
  unsigned ndat_per_weight;
  uint64_t weight_idat;
  static const bool verbose = false;
  uint64_t get_ndat() const { return ndat_; }
  uint64_t get_nweights() const { return nweights_; }
  uint64_t get_nzero() const { return 0; }
  void convolve_weights(unsigned nfft, unsigned nkeep) {
#include "gen/wts_convolve.inc"
  }
  void scrunch_weights(unsigned nscrunch) {
#include "gen/wts_scrunch.inc"
  }
};
}  // namespace

extern "C" int ref_convolve_weights(unsigned* weights, uint64_t nweights_tot, unsigned ndat_per_weight, uint64_t weight_idat,
                                    uint64_t ndat, unsigned nfft, unsigned nkeep) {
  WeightsHolder h = {weights, nweights_tot, ndat, ndat_per_weight, weight_idat};
  try {
    h.convolve_weights(nfft, nkeep);
  } catch (Error&) {
    return -1;
  }
  return 0;
}

// the reference updates ndat_per_weight and weight_idat; the new weight count follows from them (get_nweights)
extern "C" void ref_scrunch_weights(unsigned* weights, uint64_t nweights_tot, unsigned* ndat_per_weight, uint64_t* weight_idat,
                                    unsigned nscrunch) {
  WeightsHolder h = {weights, nweights_tot, 0, *ndat_per_weight, *weight_idat};
  h.scrunch_weights(nscrunch);
  *ndat_per_weight = h.ndat_per_weight;
  *weight_idat = h.weight_idat;
}
