// ref_shim/ThreadContext.h -- TEST INFRASTRUCTURE ONLY.  PSRCHIVE's mutex + condition wrapper on pthreads (the pin
// tests are single-threaded; CASPSRUnpacker's optional worker threads are never started because n_threads stays 0).
#ifndef REF_SHIM_THREADCONTEXT_H
#define REF_SHIM_THREADCONTEXT_H
#include <pthread.h>
class ThreadContext {
 public:
  ThreadContext() { pthread_mutex_init(&m, 0); pthread_cond_init(&c, 0); }
  ~ThreadContext() { pthread_mutex_destroy(&m); pthread_cond_destroy(&c); }
  void lock() { pthread_mutex_lock(&m); }
  void unlock() { pthread_mutex_unlock(&m); }
  void wait() { pthread_cond_wait(&c, &m); }
  void signal() { pthread_cond_signal(&c); }
  void broadcast() { pthread_cond_broadcast(&c); }
  class Lock {
   public:
    Lock(ThreadContext* t) : ctx(t) { if (ctx) ctx->lock(); }
    ~Lock() { if (ctx) ctx->unlock(); }
   private:
    ThreadContext* ctx;
  };
 private:
  pthread_mutex_t m;
  pthread_cond_t c;
};
#endif
