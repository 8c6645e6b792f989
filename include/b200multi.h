/* b200multi.h -- C ABI of libb200multi.so: the hot path on the GPUs of one box from ONE host process.
 *
 * Replaces: dsp::MultiThread (Signal/General/MultiThread.C:171-222 construction of one SingleThread per
 * worker, :240-271 one pthread each, :274-379 run / finish / combine of the threads' PhaseSeries) in the
 * configuration `dspsr --cuda=0,1,...`: one host thread + one CUDA stream + one pipeline per device.
 * New relative to the reference: the per-device PhaseSeries are summed on the devices with NCCL (ncclReduce
 * over NVLink) instead of being copied to the host and added there (MultiThread.C:329-342 ->
 * PhaseSeries::combine), and the attribute rules of PhaseSeries::combine are applied to the host-side
 * attributes (b200dsp.h b200_phase_series_combine).
 *
 * Sharding: B200_SHARD_TIME -- every device takes different blocks of one stream (nchan = 1 inputs;
 * combine = sum); B200_SHARD_CHANNEL -- every device owns a contiguous channel range of the same blocks
 * (combine = concatenation along the channel axis, no arithmetic, hits taken from device 0).
 */
#ifndef B200MULTI_H
#define B200MULTI_H

#include "b200dsp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_multi b200_multi;
typedef enum { B200_SHARD_TIME = 0, B200_SHARD_CHANNEL = 1 } b200_shard_mode;

/* One worker thread, context (private stream) and NCCL rank per entry of devices[]. */
int b200_multi_create(const int* devices, unsigned ndev, b200_multi** out);
int b200_multi_destroy(b200_multi* m);
unsigned b200_multi_ndev(const b200_multi* m);
/* the context of worker i: create that device's pipeline on it (b200_pipeline_create), then hand it over */
b200_context* b200_multi_context(b200_multi* m, unsigned idev);
int b200_multi_set_pipeline(b200_multi* m, unsigned idev, b200_pipeline* pipe);
/* Every worker with npart[i] > 0 runs b200_pipeline_execute_host_obs(pipe_i, h_input[i], nbytes[i],
 * first_sample[i], npart[i], obs_sample[i]) on its own thread; returns when all have queued their work
 * (first non-zero status wins; b200_multi_last_error names the device). */
int b200_multi_execute_host_obs(b200_multi* m, const void* const* h_input, const uint64_t* nbytes,
                                const uint64_t* first_sample, const uint64_t* npart, const uint64_t* obs_sample);
/* Sub-integration boundary.  TIME: ncclReduce(sum) of the float sums and of the uint32 hits onto device 0,
 * attributes merged with b200_phase_series_combine in device order.  CHANNEL: each device's block is copied to
 * its channel offset of out->data, attributes of device 0 with nchan / centre frequency / bandwidth of the whole
 * band.  out->data ([nchan_total][npol][nbin][ndim]) and out->hits ([nbin]) are caller-owned host buffers. */
int b200_multi_combine(b200_multi* m, int shard_mode, b200_phase_series* out);
/* Fold::reset on every device */
int b200_multi_reset(b200_multi* m);
int b200_multi_synchronize(b200_multi* m);
const char* b200_multi_last_error(void);
/* NCCL version the communicator runs on (diagnostics) */
int b200_multi_nccl_version(void);

#ifdef __cplusplus
}
#endif
#endif
