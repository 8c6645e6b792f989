// ref_shim/ThreadContext.h -- TEST INFRASTRUCTURE ONLY (single-threaded pin tests: the lock is a no-op).
#ifndef REF_SHIM_THREADCONTEXT_H
#define REF_SHIM_THREADCONTEXT_H
class ThreadContext {
 public:
  class Lock {
   public:
    Lock(ThreadContext*) {}
  };
};
#endif
