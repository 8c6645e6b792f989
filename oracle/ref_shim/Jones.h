// ref_shim/Jones.h -- TEST INFRASTRUCTURE ONLY.  2x2 complex matrix with the one accessor Response::set uses.
#ifndef REF_SHIM_JONES_H
#define REF_SHIM_JONES_H
#include <complex>
template <class T>
class Jones {
 public:
  std::complex<T> j[2][2];
  std::complex<T> operator()(unsigned i, unsigned k) const { return j[i][k]; }
};
#endif
