// ref_shim/Reference.h -- TEST INFRASTRUCTURE ONLY.  Minimal stand-in for PSRCHIVE's Reference.h so that
// reference sources compile in place (oracle/ref.mk).  Reference counting is not reproduced: objects handed
// to a Reference::To live until the test process exits.
#ifndef REF_SHIM_REFERENCE_H
#define REF_SHIM_REFERENCE_H
#include <algorithm>
#include <string>
#include "Error.h"   // PSRCHIVE Reference.h pulls Error.h in (TwoBitTable.C relies on it)
namespace Reference {
class Able {
 public:
  Able() {}
  Able(const Able&) {}
  Able& operator=(const Able&) { return *this; }
  virtual ~Able() {}
};
template <class T, bool active = true>
class To {
 public:
  To(T* p = 0) : ptr_(p) {}
  To& operator=(T* p) { ptr_ = p; return *this; }
  T* operator->() const { return ptr_; }
  T& operator*() const { return *ptr_; }
  operator T*() const { return ptr_; }
  operator bool() const { return ptr_ != 0; }
  bool operator!() const { return ptr_ == 0; }
  T* get() const { return ptr_; }
  T* ptr() const { return ptr_; }
  T* release() { T* p = ptr_; ptr_ = 0; return p; }
 private:
  T* ptr_;
};
}  // namespace Reference
#endif
