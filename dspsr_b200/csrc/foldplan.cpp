// foldplan.cpp -- host side of the fold bin plan (no device code).
//
// dsp::Fold::fold assigns phase bins with a SEQUENTIAL double-precision recurrence
// (Signal/Pulsar/Fold.C:765-768):
//     phi -= floor(phi); ibin = unsigned(phi * nbin); phi += phase_per_sample;
// A closed form phi0 + i*pps differs in the last bits and can move a sample across a bin edge,
// so the plan must reproduce the recurrence itself.  Doing it sample by sample on the host
// would cap the pipeline at a few hundred M output samples/s; instead we use the fact that
// while phi stays inside one binade [2^e, 2^(e+1)) every addition rounds to the same grid of
// spacing u = 2^(e-52), so the recurrence is an EXACT integer arithmetic progression
//     phi_t = (a0 + t*step) * u
// (step = round-to-nearest of pps/u; in the round-half-even tie case the increment is constant
// from the second addition on, which the code below checks instead of assuming).  The host
// walks binade to binade (about log2(1/pps) segments per pulse period) and the GPU expands the
// segments (fold.cu: k_expand_bins).  Compile with -ffp-contract=off.
#include <cmath>
#include <cstdint>

#include "../../include/b200dsp.h"

extern "C" void b200_phase_bins_sequential(double phi, double phase_per_sample, unsigned nbin, uint64_t ndat,
                                           unsigned* bins, double* phi_end) {
  const double double_nbin = double(nbin);
  for (uint64_t idat = 0; idat < ndat; idat++) {
    phi -= std::floor(phi);
    double double_ibin = phi * double_nbin;
    bins[idat] = unsigned(double_ibin);
    phi += phase_per_sample;
  }
  if (phi_end) *phi_end = phi;
}

extern "C" int64_t b200_phase_segments(double phi, double pps, uint64_t ndat, b200_phase_segment* seg,
                                       uint64_t max_segments, double* phi_end) {
  uint64_t i = 0;
  int64_t nseg = 0;
  const uint64_t TOP = (1ull << 53) - 1;
  while (i < ndat) {
    phi -= std::floor(phi);
    const double p0 = phi;
    uint64_t count = 1;
    uint64_t a0 = 0, step = 0;
    int sexp = 0;
    double last = p0;
    if (p0 >= 0x1p-900 && p0 < 1.0 && pps > 0.0) {
      const int e = std::ilogb(p0);
      const double top = std::ldexp(1.0, e + 1);
      sexp = e - 52;
      a0 = (uint64_t)std::ldexp(p0, -sexp);   // exact: p0 is a multiple of 2^sexp, a0 in [2^52, 2^53)
      const double q1 = p0 + pps;
      if (q1 < top) {
        const double q2 = q1 + pps;
        const double d1 = q1 - p0;              // exact (same binade)
        if (!(q2 < top)) {
          count = 2;
          step = (uint64_t)std::ldexp(d1, -sexp);
        } else {
          const double d2 = q2 - q1;
          if (d1 == d2) {
            step = (uint64_t)std::ldexp(d1, -sexp);
            count = step ? (TOP - a0) / step + 1 : ndat - i;   // step 0: phi no longer advances
          }
          // d1 != d2 (round-half-even tie on the first addition): emit p0 alone; the
          // progression starts at q1 on the next iteration.
        }
      }
    } else if (p0 > 0.0 && p0 < 1.0) {
      // sub-2^-900 phases: single sample, keep the exact value through a0 = 0 fallback below
      sexp = 0;
    }
    if (count > ndat - i) count = ndat - i;
    if (a0 == 0) {
      // not representable as integer*2^sexp with the scheme above (phi == 0 or tiny):
      // encode the double itself: a0 = mantissa, sexp = exponent
      if (p0 == 0.0) { a0 = 0; sexp = 0; }
      else {
        int ex;
        double m = std::frexp(p0, &ex);        // p0 = m * 2^ex, m in [0.5,1)
        a0 = (uint64_t)std::ldexp(m, 53);
        sexp = ex - 53;
      }
      step = 0;
      count = 1;
    }
    if ((uint64_t)nseg >= max_segments) return -1;
    seg[nseg].start = i;
    seg[nseg].count = count;
    seg[nseg].a0 = a0;
    seg[nseg].step = step;
    seg[nseg].scale_exp = sexp;
    seg[nseg].pad = 0;
    nseg++;
    last = std::ldexp(double(a0 + (count - 1) * step), sexp);
    phi = last + pps;
    i += count;
  }
  if (phi_end) *phi_end = phi;
  return nseg;
}
