// digifil.cu -- the tail of the digifil path: per-channel rescaling of the detected
// series and 8-bit SIGPROC digitisation.
//   b200_rescale_*          <- dsp::Rescale::transformation / compute_various (Signal/General/Rescale.C:165-412)
//   b200_sigproc_digitize8  <- dsp::SigProcDigitizer::pack, nbit 8, FPT input (Kernel/Formats/sigproc/
//                              SigProcDigitizer.C:80-160,244-300) with ChannelSort (:38-70)
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace b200 {

// one CTA per (chan,pol) plane: sums of the samples and of their float squares in double (Rescale.C:258-262)
__global__ void k_rescale_stats(const float* __restrict__ in, uint64_t span, uint64_t start, uint64_t end,
                                double* tot, double* totsq) {
  const float* p = in + blockIdx.x * span;
  double s = 0.0, s2 = 0.0;
  for (uint64_t i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float v = p[i];
    s += double(v);
    s2 += double(__fmul_rn(v, v));
  }
  __shared__ double sh[2][32];
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, off);
    s2 += __shfl_down_sync(0xffffffffu, s2, off);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const unsigned nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? sh[0][threadIdx.x] : 0.0;
    s2 = threadIdx.x < nw ? sh[1][threadIdx.x] : 0.0;
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_down_sync(0xffffffffu, s, off);
      s2 += __shfl_down_sync(0xffffffffu, s2, off);
    }
    if (threadIdx.x == 0) { tot[blockIdx.x] += s; totsq[blockIdx.x] += s2; }
  }
}

// compute_various (Rescale.C:387-412) + zeroing of the accumulators (:300-306)
__global__ void k_rescale_update(double* tot, double* totsq, float* offset, float* scale, unsigned n, double isample,
                                 int set_values) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = tot[i] / isample;
  const double meansq = totsq[i] / isample;
  const double variance = meansq - mean * mean;
  if (set_values) {
    offset[i] = float(-mean);
    scale[i] = variance == 0.0 ? 1.0f : float(1.0 / sqrt(variance));
  }
  tot[i] = 0.0;
  totsq[i] = 0.0;
}

__global__ void k_rescale_apply(const float* __restrict__ in, uint64_t in_span, float* out, uint64_t out_span,
                                uint64_t start, uint64_t end, const float* offset, const float* scale) {
  const unsigned plane = blockIdx.y;
  const float o = offset[plane], s = scale[plane];
  const float* p = in + plane * in_span;
  float* q = out + plane * out_span;
  for (uint64_t i = start + blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < end; i += uint64_t(gridDim.x) * blockDim.x)
    q[i] = __fmul_rn(__fadd_rn(p[i], o), s);       // (in + offset) * scale  (Rescale.C:366-367)
}

// FPT floats -> TPF bytes through a 32x32 shared tile: reads coalesced along time, writes along channel
__global__ void k_digitize8(const float* __restrict__ in, uint64_t span, unsigned nchan, unsigned npol, uint64_t ndat,
                            float digi_scale, float digi_mean, float xpol_offset, int flip, int swap,
                            unsigned char* out) {
  __shared__ unsigned char tile[32][33];
  const unsigned ipol = blockIdx.z;
  const unsigned c0 = blockIdx.y * 32;
  const uint64_t t0 = uint64_t(blockIdx.x) * 32;
  const float mean = digi_mean + (ipol > 1 ? xpol_offset : 0.f);
  for (unsigned cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const unsigned oc = c0 + cc;
    const uint64_t t = t0 + threadIdx.x;
    if (oc < nchan && t < ndat) {
      unsigned ic = oc;                                   // ChannelSort (SigProcDigitizer.C:56-69)
      if (swap) ic = (ic + nchan / 2) % nchan;
      if (flip) ic = nchan - ic - 1;
      const float x = in[(uint64_t(ic) * npol + ipol) * span + t];
      const double v = double(__fadd_rn(__fmul_rn(x, digi_scale), mean)) + 0.5;   // float product and sum, + 0.5 in double
      int r = __double2int_rz(v);
      r = r < 0 ? 0 : r > 255 ? 255 : r;
      tile[cc][threadIdx.x] = (unsigned char)r;
    }
  }
  __syncthreads();
  for (unsigned tt = threadIdx.y; tt < 32; tt += blockDim.y) {
    const unsigned oc = c0 + threadIdx.x;
    const uint64_t t = t0 + tt;
    if (oc < nchan && t < ndat) out[(t * npol + ipol) * nchan + oc] = tile[threadIdx.x][tt];
  }
}

}  // namespace b200

using namespace b200;

struct b200_rescale {
  Context* ctx;
  unsigned nchan, npol;
  uint64_t interval_samples, nsample, isample;
  int constant;
  double* d_tot;
  double* d_totsq;
  float* d_offset;
  float* d_scale;
};

extern "C" {

int b200_rescale_create(b200_context* cctx, unsigned nchan, unsigned npol, uint64_t interval_samples, int constant,
                        b200_rescale** out) {
  B200_REQUIRE(cctx && out && nchan && npol, "b200_rescale_create: invalid argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  b200_rescale* r = new b200_rescale();
  memset(r, 0, sizeof(*r));
  r->ctx = ctx; r->nchan = nchan; r->npol = npol; r->interval_samples = interval_samples; r->constant = constant;
  const size_t n = size_t(nchan) * npol;
  cudaError_t e = cudaMalloc(&r->d_tot, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_totsq, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_offset, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_scale, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemsetAsync(r->d_tot, 0, n * sizeof(double), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->d_totsq, 0, n * sizeof(double), ctx->stream);
  if (e != cudaSuccess) { b200_rescale_destroy(r); return cuda_fail(e, "b200_rescale_create", __FILE__, __LINE__); }
  *out = r;
  return B200_OK;
}

int b200_rescale_destroy(b200_rescale* r) {
  if (!r) return B200_OK;
  if (r->d_tot) cudaFree(r->d_tot);
  if (r->d_totsq) cudaFree(r->d_totsq);
  if (r->d_offset) cudaFree(r->d_offset);
  if (r->d_scale) cudaFree(r->d_scale);
  delete r;
  return B200_OK;
}

int b200_rescale_transform(b200_rescale* r, const float* d_in, uint64_t in_span, uint64_t ndat, float* d_out,
                           uint64_t out_span) {
  B200_REQUIRE(r && d_in && d_out, "b200_rescale_transform: null argument");
  if (ndat == 0) return B200_OK;
  Context* ctx = r->ctx;
  const unsigned nplane = r->nchan * r->npol;
  bool first_call = r->nsample == 0;                                       // Rescale.C:181-184, init() :112-123
  if (first_call) r->nsample = r->interval_samples ? r->interval_samples : ndat;
  uint64_t start = 0;
  do {                                                                     // Rescale.C:214-381
    uint64_t end = ndat;
    const uint64_t interval_end = start + r->nsample - r->isample;
    if (interval_end < end) end = interval_end;
    {
      LaunchScope ls(ctx, KC_OTHER);
      k_rescale_stats<<<nplane, 256, 0, ctx->stream>>>(d_in, in_span, start, end, r->d_tot, r->d_totsq);
    }
    r->isample += end - start;
    if (r->isample == r->nsample || first_call) {
      LaunchScope ls(ctx, KC_OTHER);
      k_rescale_update<<<(nplane + 127) / 128, 128, 0, ctx->stream>>>(r->d_tot, r->d_totsq, r->d_offset, r->d_scale,
                                                                     nplane, double(r->isample),
                                                                     (!r->constant || first_call) ? 1 : 0);
      r->isample = 0;
      first_call = false;
    }
    {
      const uint64_t n = end - start;
      dim3 grid((unsigned)std::min<uint64_t>((n + 255) / 256, 64), nplane);
      LaunchScope ls(ctx, KC_OTHER);
      k_rescale_apply<<<grid, 256, 0, ctx->stream>>>(d_in, in_span, d_out, out_span, start, end, r->d_offset, r->d_scale);
    }
    start = end;
  } while (start < ndat);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_rescale_get(b200_rescale* r, float* h_offset, float* h_scale) {
  B200_REQUIRE(r && h_offset && h_scale, "b200_rescale_get: null argument");
  const size_t n = size_t(r->nchan) * r->npol;
  B200_CUDA(cudaMemcpyAsync(h_offset, r->d_offset, n * sizeof(float), cudaMemcpyDeviceToHost, r->ctx->stream));
  B200_CUDA(cudaMemcpyAsync(h_scale, r->d_scale, n * sizeof(float), cudaMemcpyDeviceToHost, r->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(r->ctx->stream));
  return B200_OK;
}

int b200_sigproc_digitize8(b200_context* cctx, const float* d_in, uint64_t in_span, unsigned nchan, unsigned npol,
                           uint64_t ndat, float digi_scale, float digi_mean, float xpol_offset, int flip_band,
                           int swap_band, unsigned char* d_out) {
  B200_REQUIRE(cctx && d_in && d_out && nchan && npol, "b200_sigproc_digitize8: invalid argument");
  if (ndat == 0) return B200_OK;
  Context* ctx = reinterpret_cast<Context*>(cctx);
  B200_REQUIRE((ndat + 31) / 32 <= 0x7fffffffull && npol <= 65535 && (nchan + 31) / 32 <= 65535,
               "b200_sigproc_digitize8: block too large");
  dim3 grid((unsigned)((ndat + 31) / 32), (nchan + 31) / 32, npol), block(32, 8);
  LaunchScope ls(ctx, KC_OTHER);
  k_digitize8<<<grid, block, 0, ctx->stream>>>(d_in, in_span, nchan, npol, ndat, digi_scale, digi_mean, xpol_offset,
                                              flip_band, swap_band, d_out);
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // extern "C"
