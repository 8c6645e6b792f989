// ref_shim/NormalDistribution.h -- TEST INFRASTRUCTURE ONLY.  PSRCHIVE's NormalDistribution (not in the reference
// tree) restated: cumulative distribution of the unit normal.  Same expression as the oracle's orc_normal_cdf.
#ifndef REF_SHIM_NORMALDISTRIBUTION_H
#define REF_SHIM_NORMALDISTRIBUTION_H
#include <math.h>
class NormalDistribution {
 public:
  double cumulative_distribution(double x) const { return 0.5 * (1.0 + ::erf(x / ::sqrt(2.0))); }
};
#endif
