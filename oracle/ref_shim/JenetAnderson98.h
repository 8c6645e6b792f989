// ref_shim/JenetAnderson98.h -- TEST INFRASTRUCTURE ONLY.  PSRCHIVE's JenetAnderson98 (NOT in the reference tree,
// version un-pinned by configure.ac:73) restated from Jenet & Anderson (1998, PASP 110, 1467): Table 3 optimal
// spacings, Eqs. 40-45 for the dynamic output levels.  The arithmetic is the oracle's own (orc_ja98_* in
// oracle/dspsr_oracle.cpp, resolved at link time), so a pin test that passes says: the reference's BitTable /
// TwoBitLookup code, compiled from its own sources, and the oracle's restatement of that code agree bit for bit
// GIVEN the same JA98 numbers.  The JA98 numbers themselves stay "restated third party".
#ifndef REF_SHIM_JENETANDERSON98_H
#define REF_SHIM_JENETANDERSON98_H
#include <math.h>
extern "C" double orc_ja98_optimal_spacing(unsigned nbit);
extern "C" void orc_ja98_levels(double Phi, double* lo, double* hi);
class JenetAnderson98 {
 public:
  JenetAnderson98() : threshold(0.9674), Phi(0), lo(0), hi(0) {}
  static double get_optimal_spacing(unsigned bits) { return orc_ja98_optimal_spacing(bits); }
  void set_threshold(double t) { threshold = t; }
  double get_threshold() const { return threshold; }
  void set_Phi(double p) { Phi = p; orc_ja98_levels(p, &lo, &hi); }
  double get_lo() const { return lo; }
  double get_hi() const { return hi; }
  double get_mean_Phi() const { return ::erf(threshold / ::sqrt(2.0)); }
  double get_var_Phi() const { const double m = get_mean_Phi(); return m * (1.0 - m); }
 private:
  double threshold, Phi, lo, hi;
};
#endif
