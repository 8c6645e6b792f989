// weights.cu -- the per-window validity flags of a WeightedTimeSeries on the device.
//
// Two-bit excision unpackers flag windows of ndat_per_weight samples (fold.cu k_unpack_twobit); the reference then
// carries the flags through every transformation on the host:
//   WeightedTimeSeries::convolve_weights (Kernel/Classes/WeightedTimeSeries.C:582-690), called by
//     Filterbank::prepare_output (Filterbank.C:302-307) and Convolution::prepare_output (Convolution.C:312-319):
//     an overlap-save transform that contains a flagged window is flagged as a whole;
//   WeightedTimeSeries::scrunch_weights (:692-780): the flags follow the change of time resolution;
//   Fold::fold (Fold.C:687-716,746-763): samples of flagged windows are not folded and give no hit.
// Here the flags never leave the GPU.  convolve_weights is sequential in the reference (the flagging of transform i
// is applied while transform i+1 is examined, "so that it does not affect the next test"); as long as the step
// between transforms is at least one window (nkeep >= ndat_per_weight) no test ever reads a flag written by an
// earlier transform, so every transform can be tested against the ORIGINAL flags in parallel and the flagged
// ranges applied afterwards -- the same result, two small kernels.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace b200 {

// bad[i] = 1 iff transform i (samples [i*nkeep, i*nkeep + nfft) of the block) contains a flagged window;
// one warp per transform.  Index arithmetic in double, exactly as the reference's (:614-617).
__global__ void k_weights_test(const unsigned* __restrict__ w, uint64_t nweights, double weights_per_dat,
                               uint64_t weight_idat, unsigned nfft, unsigned nkeep, uint64_t nblocks,
                               unsigned* __restrict__ bad, unsigned* __restrict__ overflow) {
  const uint64_t i = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (i >= nblocks) return;
  const uint64_t wt_idat = i * nkeep + weight_idat;
  const uint64_t sw = uint64_t(double(wt_idat) * weights_per_dat);
  const uint64_t ew = uint64_t(ceil(double(wt_idat + nfft) * weights_per_dat));
  if (ew > nweights) {                     // the reference throws here (:619-623)
    if (lane == 0) atomicExch(overflow, 1u);
    return;
  }
  unsigned z = 0;
  for (uint64_t k = sw + lane; k < ew; k += 32) z |= (w[k] == 0u);
  z = __any_sync(0xffffffffu, z);
  if (lane == 0) bad[i] = z ? 1u : 0u;
}

// out[k] = 0 if w[k] == 0 or k lies in the range a flagged transform zeroes: [sw_i, ceil((i*nkeep + nkeep) * wpd))
// (:652-653 -- the end is computed WITHOUT weight_idat in the reference; kept)
__global__ void k_weights_apply(const unsigned* __restrict__ w, uint64_t nweights, double weights_per_dat,
                                uint64_t weight_idat, unsigned nkeep, unsigned ndat_per_weight, uint64_t nblocks,
                                const unsigned* __restrict__ bad, unsigned* __restrict__ out) {
  for (uint64_t k = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; k < nweights; k += uint64_t(gridDim.x) * blockDim.x) {
    unsigned v = w[k];
    if (v != 0u && nblocks) {
      const uint64_t ic = (k * ndat_per_weight) / nkeep;
      const uint64_t i0 = ic >= 2 ? ic - 2 : 0, i1 = min(nblocks, ic + 3);
      for (uint64_t i = i0; i < i1; i++) {
        if (!bad[i]) continue;
        const uint64_t start_idat = i * nkeep;
        const uint64_t sw = uint64_t(double(start_idat + weight_idat) * weights_per_dat);
        const uint64_t ze = uint64_t(ceil(double(start_idat + nkeep) * weights_per_dat));
        if (k >= sw && k < ze) v = 0u;
      }
    }
    out[k] = v;
  }
}

// scrunch_weights when one output sample spans more than one window (:741-775)
__global__ void k_weights_scrunch(const unsigned* __restrict__ w, uint64_t nweights, unsigned nscrunch, uint64_t nnew,
                                  unsigned* __restrict__ out) {
  for (uint64_t k = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; k < nnew; k += uint64_t(gridDim.x) * blockDim.x) {
    unsigned n = nscrunch;
    if ((k + 1) * nscrunch > nweights) n = unsigned(nweights % nscrunch);
    unsigned acc = 0;
    bool zero = false;
    for (unsigned j = 0; j < n; j++) {
      const unsigned v = w[k * nscrunch + j];
      if (v == 0u) { zero = true; break; }
      acc += v;
    }
    out[k] = zero ? 0u : acc / n;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_weights_convolve(b200_context* cctx, const unsigned* d_weights, uint64_t nweights, unsigned ndat_per_weight,
                          uint64_t weight_idat, uint64_t ndat, unsigned nfft, unsigned nkeep, unsigned* d_out,
                          unsigned* d_scratch) {
  B200_REQUIRE(cctx && d_weights && d_out, "b200_weights_convolve: null argument");
  B200_REQUIRE(d_out != d_weights, "b200_weights_convolve: in place is not supported (the tests read the original flags)");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  cudaStream_t st = ctx->stream;
  // the early returns of the reference (:584-609): nothing to convolve
  if (ndat_per_weight == 0 || ndat_per_weight >= nfft || ndat + nkeep < nfft) {
    B200_CUDA(cudaMemcpyAsync(d_out, d_weights, nweights * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    return B200_OK;
  }
  if (nkeep < ndat_per_weight) {
    set_error("convolve_weights: nsamp_step=%u shorter than one window of %u samples is not built", nkeep, ndat_per_weight);
    return B200_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(d_scratch, "b200_weights_convolve: scratch of (transforms + 1) words needed");
  const uint64_t nblocks = (ndat + nkeep - nfft) / nkeep;
  const double wpd = 1.0 / ndat_per_weight;
  // the reference throws when a transform reaches beyond the flags; checked on the host with the same arithmetic
  if (nblocks) {
    const uint64_t last = (nblocks - 1) * nkeep + weight_idat;
    const uint64_t ew = uint64_t(std::ceil(double(last + nfft) * wpd));
    B200_REQUIRE(ew <= nweights, "convolve_weights: end_weight=%llu > nweights=%llu (WeightedTimeSeries.C:619)",
                 (unsigned long long)ew, (unsigned long long)nweights);
  }
  unsigned* bad = d_scratch;
  unsigned* overflow = d_scratch + nblocks;
  B200_CUDA(cudaMemsetAsync(overflow, 0, sizeof(unsigned), st));
  LaunchScope ls(ctx, KC_OTHER);
  if (nblocks) {
    const unsigned threads = 256;
    const unsigned grid = unsigned((nblocks * 32 + threads - 1) / threads);
    k_weights_test<<<grid, threads, 0, st>>>(d_weights, nweights, wpd, weight_idat, nfft, nkeep, nblocks, bad, overflow);
  }
  {
    const unsigned threads = 256;
    const unsigned grid = unsigned(std::min<uint64_t>((nweights + threads - 1) / threads, uint64_t(ctx->sm_count) * 8));
    k_weights_apply<<<std::max(1u, grid), threads, 0, st>>>(d_weights, nweights, wpd, weight_idat, nkeep, ndat_per_weight,
                                                            nblocks, bad, d_out);
  }
  B200_CUDA(cudaGetLastError());
  return B200_OK;
}

int b200_weights_scrunch(b200_context* cctx, const unsigned* d_weights, uint64_t* nweights, unsigned* ndat_per_weight,
                         uint64_t* weight_idat, unsigned nscrunch, unsigned* d_out) {
  B200_REQUIRE(cctx && nweights && ndat_per_weight && weight_idat && nscrunch, "b200_weights_scrunch: null argument");
  Context* ctx = reinterpret_cast<Context*>(cctx);
  if (!*ndat_per_weight) return B200_OK;
  const double points_per_weight = double(*ndat_per_weight) / double(nscrunch);
  if (points_per_weight >= 1.0) {                      // :709-726: only the bookkeeping changes
    *ndat_per_weight = unsigned(points_per_weight);
    const bool leftover = (*weight_idat % *ndat_per_weight != 0);
    *weight_idat /= *ndat_per_weight;
    if (leftover) (*weight_idat)++;
    if (d_out && d_out != d_weights)
      B200_CUDA(cudaMemcpyAsync(d_out, d_weights, *nweights * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
  }
  B200_REQUIRE(d_weights && d_out && d_out != d_weights, "b200_weights_scrunch: distinct input and output arrays needed");
  const uint64_t nnew = *nweights / nscrunch + ((*nweights % nscrunch) ? 1 : 0);
  {
    LaunchScope ls(ctx, KC_OTHER);
    const unsigned threads = 256;
    const unsigned grid = unsigned(std::min<uint64_t>((nnew + threads - 1) / threads, uint64_t(ctx->sm_count) * 8));
    k_weights_scrunch<<<std::max(1u, grid), threads, 0, ctx->stream>>>(d_weights, *nweights, nscrunch, nnew, d_out);
  }
  B200_CUDA(cudaGetLastError());
  *nweights = nnew;
  *ndat_per_weight = 1;
  return B200_OK;
}

}  // extern "C"
