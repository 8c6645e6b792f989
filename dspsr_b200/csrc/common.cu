// common.cu -- context, error plumbing, memory helpers and twiddle tables of libb200dsp.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b200 {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what, file, line);
  return B200_ERR_CUDA;
}

int make_twiddle(TwiddleTable& t, unsigned n, cudaStream_t s) {
  t.n = n;
  std::vector<float2> h(n ? n : 1);
  for (unsigned m = 0; m < n; m++) {
    double a = -2.0 * M_PI * double(m) / double(n);
    h[m] = make_float2(float(std::cos(a)), float(std::sin(a)));
  }
  B200_CUDA(cudaMalloc(&t.tw, sizeof(float2) * h.size()));
  B200_CUDA(cudaMemcpyAsync(t.tw, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice, s));
  // compact per-stage tables of the compile-time-sized paths (EPT = 32 and 16)
  for (int ept = 32; ept >= 16; ept -= 16) {
    std::vector<float2> st(1, make_float2(1.f, 0.f));
    if (n >= (unsigned)ept && (n & (n - 1)) == 0) {
      st.assign(stage_table_size(n, ept), make_float2(1.f, 0.f));
      unsigned ns = 1, off = 0;
      while (ns < n) {
        const int R = stage_radix(n, ept, ns);
        if (ns > 1) {
          for (int mi = 0; mi < stage_nmult(R); mi++)
            for (unsigned k = 0; k < ns; k++) {
              double a = -2.0 * M_PI * double(stage_mult(R, mi)) * double(k) / (double(ns) * R);
              st[off + mi * ns + k] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
          off += stage_nmult(R) * ns;
        }
        ns *= R;
      }
    }
    float2** dst = ept == 32 ? &t.stage : &t.stage16;
    B200_CUDA(cudaMalloc(dst, sizeof(float2) * st.size()));
    B200_CUDA(cudaMemcpy(*dst, st.data(), sizeof(float2) * st.size(), cudaMemcpyHostToDevice));
  }
  B200_CUDA(cudaStreamSynchronize(s));
  return B200_OK;
}

int make_big_twiddle(BigTwiddle& t, uint64_t n, cudaStream_t s) {
  t.n = n;
  t.nlo = 2048;
  t.nhi = n > 2048 ? unsigned(n / 2048) : 1;
  std::vector<float2> lo(t.nlo), hi(t.nhi);
  for (unsigned b = 0; b < t.nlo; b++) {
    double a = -2.0 * M_PI * double(b % n) / double(n);
    lo[b] = make_float2(float(std::cos(a)), float(std::sin(a)));
  }
  for (unsigned a_ = 0; a_ < t.nhi; a_++) {
    double a = -2.0 * M_PI * double(a_) / double(t.nhi);
    hi[a_] = t.nhi == 1 ? make_float2(1.f, 0.f) : make_float2(float(std::cos(a)), float(std::sin(a)));
  }
  B200_CUDA(cudaMalloc(&t.lo, sizeof(float2) * t.nlo));
  B200_CUDA(cudaMalloc(&t.hi, sizeof(float2) * t.nhi));
  B200_CUDA(cudaMemcpyAsync(t.lo, lo.data(), sizeof(float2) * t.nlo, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(t.hi, hi.data(), sizeof(float2) * t.nhi, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaStreamSynchronize(s));
  return B200_OK;
}

void free_twiddle(TwiddleTable& t) {
  if (t.tw) cudaFree(t.tw);
  if (t.stage) cudaFree(t.stage);
  if (t.stage16) cudaFree(t.stage16);
  t.tw = t.stage = t.stage16 = nullptr;
}
void free_big_twiddle(BigTwiddle& t) {
  if (t.lo) cudaFree(t.lo);
  if (t.hi) cudaFree(t.hi);
  t.lo = t.hi = nullptr;
}

}  // namespace b200

struct b200_context : public b200::Context {
  bool own_stream;
};

extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) { return b200::g_error; }

int b200_context_create(int device, void* cuda_stream, b200_context** out) {
  B200_REQUIRE(out != nullptr, "b200_context_create: null output pointer");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    b200::set_error("b200_context_create: no CUDA device available (%s); libb200dsp has no CPU fallback",
                    cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  B200_REQUIRE(device >= 0 && device < ndev, "b200_context_create: device %d out of range (%d devices)", device, ndev);
  B200_CUDA(cudaSetDevice(device));
  b200_context* c = new b200_context();
  c->device = device;
  c->launches = 0;
  c->timing = false;
  c->timed = new std::vector<b200::TimedLaunch>();
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
  } else {
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  B200_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  B200_CUDA(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  c->d_tables = nullptr;
  B200_CUDA(cudaMalloc(&c->d_tables, 8192));
  *out = c;
  return B200_OK;
}

int b200_context_destroy(b200_context* c) {
  if (!c) return B200_OK;
  if (c->own_stream) cudaStreamDestroy(c->stream);
  for (auto& t : *c->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
  delete c->timed;
  if (c->d_tables) cudaFree(c->d_tables);
  delete c;
  return B200_OK;
}

int b200_context_synchronize(b200_context* c) {
  B200_REQUIRE(c, "null context");
  B200_CUDA(cudaStreamSynchronize(c->stream));
  return B200_OK;
}

int b200_context_set_timing(b200_context* c, int enable) {
  B200_REQUIRE(c, "null context");
  c->timing = enable != 0;
  return B200_OK;
}

int b200_context_read_timing(b200_context* c, double* ms, unsigned long long* count) {
  B200_REQUIRE(c && ms && count, "b200_context_read_timing: null argument");
  B200_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < b200::KC_COUNT; k++) { ms[k] = 0.0; count[k] = 0; }
  for (auto& t : *c->timed) {
    float e = 0.f;
    B200_CUDA(cudaEventElapsedTime(&e, t.start, t.stop));
    ms[t.kc] += e;
    count[t.kc]++;
    cudaEventDestroy(t.start);
    cudaEventDestroy(t.stop);
  }
  c->timed->clear();
  return B200_OK;
}

unsigned long long b200_context_launch_count(const b200_context* c) { return c ? c->launches : 0; }
void* b200_context_stream(const b200_context* c) { return c ? (void*)c->stream : nullptr; }

int b200_malloc(b200_context* c, uint64_t nbytes, void** d_ptr) {
  B200_REQUIRE(c && d_ptr, "b200_malloc: null argument");
  B200_CUDA(cudaSetDevice(c->device));
  B200_CUDA(cudaMalloc(d_ptr, nbytes ? nbytes : 1));
  return B200_OK;
}
int b200_free(b200_context* c, void* d_ptr) {
  (void)c;
  if (d_ptr) B200_CUDA(cudaFree(d_ptr));
  return B200_OK;
}
int b200_malloc_host(b200_context* c, uint64_t nbytes, void** h_ptr) {
  B200_REQUIRE(c && h_ptr, "b200_malloc_host: null argument");
  B200_CUDA(cudaMallocHost(h_ptr, nbytes ? nbytes : 1));
  return B200_OK;
}
int b200_free_host(b200_context* c, void* h_ptr) {
  (void)c;
  if (h_ptr) B200_CUDA(cudaFreeHost(h_ptr));
  return B200_OK;
}
int b200_memset(b200_context* c, void* d_ptr, int value, uint64_t nbytes) {
  B200_REQUIRE(c, "null context");
  B200_CUDA(cudaMemsetAsync(d_ptr, value, nbytes, c->stream));
  return B200_OK;
}
int b200_memcpy_d2d(b200_context* c, void* d_dst, const void* d_src, uint64_t nbytes) {
  B200_REQUIRE(c, "null context");
  B200_CUDA(cudaMemcpyAsync(d_dst, d_src, nbytes, cudaMemcpyDeviceToDevice, c->stream));
  return B200_OK;
}
int b200_memcpy_h2d(b200_context* c, void* d_dst, const void* h_src, uint64_t nbytes) {
  B200_REQUIRE(c, "null context");
  B200_CUDA(cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, c->stream));
  return B200_OK;
}
int b200_memcpy_d2h(b200_context* c, void* h_dst, const void* d_src, uint64_t nbytes) {
  B200_REQUIRE(c, "null context");
  B200_CUDA(cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, c->stream));
  B200_CUDA(cudaStreamSynchronize(c->stream));
  return B200_OK;
}

}  // extern "C"
