// ref_shim/Callback.h -- TEST INFRASTRUCTURE ONLY.  No listeners are ever connected in the pin tests.
#ifndef REF_SHIM_CALLBACK_H
#define REF_SHIM_CALLBACK_H
template <class T>
class Callback {
 public:
  void send(const T&) {}
};
#endif
