// multi.cpp -- one host process driving the pipelines of several GPUs (include/b200multi.h).
//
// The reference's MultiThread (Signal/General/MultiThread.C) gives every worker its own pthread, stream and
// engines and merges the workers' PhaseSeries on the host.  Here every worker is a std::thread bound to one
// device that executes jobs from a queue (so CUDA calls of different devices never serialise on one host thread),
// and the merge of time shards is an ncclReduce over NVLink issued for all ranks from the calling thread inside
// one ncclGroupStart/End (the single-process multi-device form of NCCL).
#include <cuda_runtime.h>
#include <nccl.h>

#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200multi.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

struct Worker {
  int device = 0;
  b200_context* ctx = nullptr;
  b200_pipeline* pipe = nullptr;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::queue<std::function<void()>> jobs;
  bool quit = false;
  unsigned pending = 0;

  void loop() {
    cudaSetDevice(device);
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return quit || !jobs.empty(); });
        if (jobs.empty()) return;
        job = std::move(jobs.front());
        jobs.pop();
      }
      job();
      {
        std::lock_guard<std::mutex> lk(mu);
        pending--;
      }
      cv.notify_all();
    }
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      jobs.push(std::move(f));
      pending++;
    }
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return pending == 0; });
  }
};

}  // namespace

struct b200_multi {
  std::vector<Worker*> w;
  // receive buffers of the time-shard reduce on device 0
  float* d_sum = nullptr;
  unsigned* d_hits = nullptr;
  uint64_t sum_floats = 0, hits_n = 0;
};

#define MCUDA(x)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) return fail(B200_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_));  \
  } while (0)
#define MNCCL(x)                                                                              \
  do {                                                                                        \
    ncclResult_t r_ = (x);                                                                    \
    if (r_ != ncclSuccess) return fail(B200_ERR_CUDA, "%s: %s", #x, ncclGetErrorString(r_));  \
  } while (0)

extern "C" {

const char* b200_multi_last_error(void) { return g_err.c_str(); }

int b200_multi_nccl_version(void) {
  int v = 0;
  ncclGetVersion(&v);
  return v;
}

int b200_multi_create(const int* devices, unsigned ndev, b200_multi** out) {
  if (!devices || !ndev || !out) return fail(B200_ERR_INVALID, "b200_multi_create: null argument");
  int have = 0;
  MCUDA(cudaGetDeviceCount(&have));
  for (unsigned i = 0; i < ndev; i++)
    if (devices[i] < 0 || devices[i] >= have) return fail(B200_ERR_INVALID, "device %d of %d", devices[i], have);
  b200_multi* m = new b200_multi();
  std::vector<ncclComm_t> comms(ndev);
  if (ndev > 1) {
    ncclResult_t r = ncclCommInitAll(comms.data(), (int)ndev, devices);
    if (r != ncclSuccess) {
      delete m;
      return fail(B200_ERR_CUDA, "ncclCommInitAll: %s", ncclGetErrorString(r));
    }
  }
  for (unsigned i = 0; i < ndev; i++) {
    Worker* k = new Worker();
    k->device = devices[i];
    k->comm = ndev > 1 ? comms[i] : nullptr;
    int rc = b200_context_create(devices[i], nullptr, &k->ctx);
    if (rc != B200_OK) {
      g_err = b200_last_error();
      delete k;
      b200_multi_destroy(m);
      return rc;
    }
    k->stream = static_cast<cudaStream_t>(b200_context_stream(k->ctx));
    k->th = std::thread([k] { k->loop(); });
    m->w.push_back(k);
  }
  *out = m;
  return B200_OK;
}

int b200_multi_destroy(b200_multi* m) {
  if (!m) return B200_OK;
  for (Worker* k : m->w) {
    k->wait();
    {
      std::lock_guard<std::mutex> lk(k->mu);
      k->quit = true;
    }
    k->cv.notify_all();
    if (k->th.joinable()) k->th.join();
    cudaSetDevice(k->device);
    if (k->comm) ncclCommDestroy(k->comm);
    if (k->ctx) b200_context_destroy(k->ctx);
    delete k;
  }
  if (!m->w.empty()) cudaSetDevice(m->w[0]->device);
  if (m->d_sum) cudaFree(m->d_sum);
  if (m->d_hits) cudaFree(m->d_hits);
  delete m;
  return B200_OK;
}

unsigned b200_multi_ndev(const b200_multi* m) { return m ? (unsigned)m->w.size() : 0; }

b200_context* b200_multi_context(b200_multi* m, unsigned i) { return (m && i < m->w.size()) ? m->w[i]->ctx : nullptr; }

int b200_multi_set_pipeline(b200_multi* m, unsigned i, b200_pipeline* pipe) {
  if (!m || i >= m->w.size() || !pipe) return fail(B200_ERR_INVALID, "b200_multi_set_pipeline: bad argument");
  m->w[i]->pipe = pipe;
  return B200_OK;
}

int b200_multi_execute_host_obs(b200_multi* m, const void* const* h_input, const uint64_t* nbytes,
                                const uint64_t* first_sample, const uint64_t* npart, const uint64_t* obs_sample) {
  if (!m || !h_input || !nbytes || !first_sample || !npart || !obs_sample)
    return fail(B200_ERR_INVALID, "b200_multi_execute_host_obs: null argument");
  const size_t n = m->w.size();
  std::vector<int> rc(n, B200_OK);
  std::vector<std::string> msg(n);
  for (size_t i = 0; i < n; i++) {
    if (!npart[i]) continue;
    Worker* k = m->w[i];
    if (!k->pipe) return fail(B200_ERR_INVALID, "device %d has no pipeline", k->device);
    k->submit([=, &rc, &msg] {
      rc[i] = b200_pipeline_execute_host_obs(k->pipe, h_input[i], nbytes[i], first_sample[i], npart[i], obs_sample[i]);
      if (rc[i] != B200_OK) msg[i] = b200_last_error();          // thread-local of the worker: carry it over
    });
  }
  for (Worker* k : m->w) k->wait();
  for (size_t i = 0; i < n; i++)
    if (rc[i] != B200_OK) return fail(rc[i], "device %d: %s", m->w[i]->device, msg[i].c_str());
  return B200_OK;
}

int b200_multi_synchronize(b200_multi* m) {
  if (!m) return fail(B200_ERR_INVALID, "null");
  for (Worker* k : m->w) {
    k->wait();
    MCUDA(cudaSetDevice(k->device));
    MCUDA(cudaStreamSynchronize(k->stream));
  }
  return B200_OK;
}

int b200_multi_reset(b200_multi* m) {
  if (!m) return fail(B200_ERR_INVALID, "null");
  for (Worker* k : m->w) {
    k->wait();
    if (!k->pipe) continue;
    MCUDA(cudaSetDevice(k->device));
    int rc = b200_pipeline_reset(k->pipe);
    if (rc != B200_OK) return fail(rc, "device %d: %s", k->device, b200_last_error());
  }
  return B200_OK;
}

int b200_multi_combine(b200_multi* m, int mode, b200_phase_series* out) {
  if (!m || !out || !out->data || !out->hits) return fail(B200_ERR_INVALID, "b200_multi_combine: null argument");
  const size_t n = m->w.size();
  for (Worker* k : m->w) {
    k->wait();
    if (!k->pipe) return fail(B200_ERR_INVALID, "device %d has no pipeline", k->device);
  }
  // attributes and shapes of every device's PhaseSeries
  std::vector<b200_phase_series> ps(n);
  for (size_t i = 0; i < n; i++) {
    memset(&ps[i], 0, sizeof ps[i]);
    MCUDA(cudaSetDevice(m->w[i]->device));
    int rc = b200_pipeline_get_phase_series(m->w[i]->pipe, &ps[i]);
    if (rc != B200_OK) return fail(rc, "device %d: %s", m->w[i]->device, b200_last_error());
  }
  float* data = out->data;
  unsigned* hits = out->hits;
  const unsigned nbin = ps[0].nbin;

  if (mode == B200_SHARD_TIME) {
    const uint64_t nfl = uint64_t(ps[0].obs.nchan) * ps[0].obs.npol * ps[0].obs.ndim * nbin;
    for (size_t i = 1; i < n; i++)
      if (ps[i].nbin != nbin || uint64_t(ps[i].obs.nchan) * ps[i].obs.npol * ps[i].obs.ndim * nbin != nfl)
        return fail(B200_ERR_INVALID, "time shards of different shapes");
    Worker* root = m->w[0];
    MCUDA(cudaSetDevice(root->device));
    if (n > 1) {
      if (m->sum_floats < nfl) {
        if (m->d_sum) cudaFree(m->d_sum);
        MCUDA(cudaMalloc(&m->d_sum, nfl * sizeof(float)));
        m->sum_floats = nfl;
      }
      if (m->hits_n < nbin) {
        if (m->d_hits) cudaFree(m->d_hits);
        MCUDA(cudaMalloc(&m->d_hits, nbin * sizeof(unsigned)));
        m->hits_n = nbin;
      }
      // PhaseSeries::combine of the data: data +=, hits += -- on the devices, summed in NCCL's fixed rank order
      MNCCL(ncclGroupStart());
      for (size_t i = 0; i < n; i++) {
        b200_fold* f = b200_pipeline_fold(m->w[i]->pipe);
        MNCCL(ncclReduce(b200_fold_device_profile(f), m->d_sum, nfl, ncclFloat, ncclSum, 0, m->w[i]->comm, m->w[i]->stream));
      }
      MNCCL(ncclGroupEnd());
      MNCCL(ncclGroupStart());
      for (size_t i = 0; i < n; i++) {
        b200_fold* f = b200_pipeline_fold(m->w[i]->pipe);
        MNCCL(ncclReduce(b200_fold_device_hits(f), m->d_hits, nbin, ncclUint32, ncclSum, 0, m->w[i]->comm, m->w[i]->stream));
      }
      MNCCL(ncclGroupEnd());
      MCUDA(cudaSetDevice(root->device));
      MCUDA(cudaMemcpyAsync(data, m->d_sum, nfl * sizeof(float), cudaMemcpyDeviceToHost, root->stream));
      MCUDA(cudaMemcpyAsync(hits, m->d_hits, nbin * sizeof(unsigned), cudaMemcpyDeviceToHost, root->stream));
      for (Worker* k : m->w) {
        MCUDA(cudaSetDevice(k->device));
        MCUDA(cudaStreamSynchronize(k->stream));
      }
    } else {
      ps[0].data = data;
      ps[0].hits = hits;
      int rc = b200_pipeline_get_phase_series(root->pipe, &ps[0]);
      if (rc != B200_OK) return fail(rc, "%s", b200_last_error());
    }
    // the attribute rules of PhaseSeries::combine, in device order (arrays are already summed: pass none)
    b200_phase_series acc;
    memset(&acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) {
      ps[i].data = nullptr;
      ps[i].hits = nullptr;
      if (ps[i].integration_length == 0.0) continue;                 // an idle device contributes nothing
      int rc = b200_phase_series_combine(&acc, &ps[i]);
      if (rc != B200_OK) return fail(rc, "PhaseSeries !mixable (device %d)", m->w[i]->device);
    }
    *out = acc;
    out->data = data;
    out->hits = hits;
    return B200_OK;
  }

  if (mode == B200_SHARD_CHANNEL) {
    // disjoint [chan][pol][bin][dim] blocks: every device copies its block straight to its channel offset of the
    // host array on its own stream (one process owns all devices: no hop through device 0 is needed)
    uint64_t off = 0;
    unsigned nchan_total = 0;
    for (size_t i = 0; i < n; i++) {
      if (ps[i].nbin != nbin || ps[i].obs.npol != ps[0].obs.npol || ps[i].obs.ndim != ps[0].obs.ndim)
        return fail(B200_ERR_INVALID, "channel shards of different shapes");
      const uint64_t nfl = uint64_t(ps[i].obs.nchan) * ps[i].obs.npol * ps[i].obs.ndim * nbin;
      MCUDA(cudaSetDevice(m->w[i]->device));
      b200_fold* f = b200_pipeline_fold(m->w[i]->pipe);
      MCUDA(cudaMemcpyAsync(data + off, b200_fold_device_profile(f), nfl * sizeof(float), cudaMemcpyDeviceToHost, m->w[i]->stream));
      if (i == 0) MCUDA(cudaMemcpyAsync(hits, b200_fold_device_hits(f), nbin * sizeof(unsigned), cudaMemcpyDeviceToHost, m->w[i]->stream));
      off += nfl;
      nchan_total += ps[i].obs.nchan;
    }
    for (Worker* k : m->w) {
      MCUDA(cudaSetDevice(k->device));
      MCUDA(cudaStreamSynchronize(k->stream));
    }
    // every shard saw the same samples: times, integration_length and ndat_total must agree
    for (size_t i = 1; i < n; i++)
      if (ps[i].ndat_total != ps[0].ndat_total || ps[i].integration_length != ps[0].integration_length)
        return fail(B200_ERR_INVALID, "channel shards folded different data (device %d)", m->w[i]->device);
    *out = ps[0];
    // the whole band: contiguous channel ranges in device order
    double lo = ps[0].obs.centre_frequency - 0.5 * ps[0].obs.bandwidth, bw = 0;
    for (size_t i = 0; i < n; i++) bw += ps[i].obs.bandwidth;
    out->obs.nchan = nchan_total;
    out->obs.bandwidth = bw;
    out->obs.centre_frequency = lo + 0.5 * bw;
    out->data = data;
    out->hits = hits;
    return B200_OK;
  }
  return fail(B200_ERR_INVALID, "unknown shard mode %d", mode);
}

}  // extern "C"
