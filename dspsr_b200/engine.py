"""Python handles on the C-ABI engines (one class per reference engine interface).

These are thin: they own a handle, translate torch CUDA tensors into (pointer, span) pairs
exactly like the C++ shims in dspsr_b200/host/ translate dsp::TimeSeries, and raise B200Error
on a non-zero status (the shims `throw Error`).  torch is used for device memory and streams
only; every computation happens in libb200dsp.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _need_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA tensor" % name)


class Context:
    """One CUDA stream of one device (reference: one pipeline thread, SingleThread.C:237-244)."""

    def __init__(self, device=0, stream=None):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("dspsr_b200 needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(device)
        self.device = device
        self.torch_stream = stream if stream is not None else torch.cuda.current_stream(device)
        h = C.c_void_p()
        # torch's default stream has handle 0, which the C ABI reads as "create a private stream";
        # cudaStreamLegacy (0x1) names the same default stream explicitly.
        handle = self.torch_stream.cuda_stream or 1
        L.check(self.lib.b200_context_create(device, C.c_void_p(handle), C.byref(h)))
        self.h = h

    def synchronize(self):
        L.check(self.lib.b200_context_synchronize(self.h))

    def set_timing(self, enable):
        L.check(self.lib.b200_context_set_timing(self.h, int(enable)))

    def read_timing(self):
        """-> ({class: ms}, {class: launches}) since the last read; synchronises the stream."""
        ms = (C.c_double * 5)()
        n = (C.c_ulonglong * 5)()
        L.check(self.lib.b200_context_read_timing(self.h, ms, n))
        names = ["cols_fwd", "rows", "inverse", "bins", "other"]
        return {k: ms[i] for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}

    @property
    def launches(self):
        return int(self.lib.b200_context_launch_count(self.h))

    def __del__(self):
        try:
            self.lib.b200_context_destroy(self.h)
        except Exception:
            pass


def make_unpack_desc(fmt, nchan, npol, ndim, lut=None, scale=0.0, sample_swap=1):
    d = L.UnpackDesc()
    d.format, d.nchan, d.npol, d.ndim = fmt, nchan, npol, ndim
    if lut is not None:
        lut = np.ascontiguousarray(lut, np.float32)
        C.memmove(d.lut, lut.ctypes.data, 1024)
    d.scale = scale
    d.sample_swap = sample_swap
    return d


def make_twobit_desc(npol=2, threshold=0.9674, cutoff_sigma=10.0, table_type=0, ndat_per_weight=512):
    """Level table of the two-bit excision unpacker (TwoBitCorrection::build); table_type 0 = OffsetBinary."""
    d = L.TwoBitDesc()
    L.check(L.load().b200_twobit_prepare(threshold, cutoff_sigma, table_type, npol, ndat_per_weight, C.byref(d)))
    return d


def make_twobit_unpack_desc(twobit):
    """b200_unpack_desc for FMT_TWOBIT; keeps the level table alive on the returned object."""
    d = make_unpack_desc(L.FMT_TWOBIT, 1, twobit.npol, 1)
    d.twobit = C.pointer(twobit)
    d._keep = twobit
    return d


def unpack_twobit(ctx, twobit, raw, ndat):
    """Two-bit excision unpacker: raw uint8 CUDA tensor -> (float32 [1, npol, ndat], weights uint32 [ndat/512])."""
    _need_cuda(raw, "raw")
    out = torch.empty((1, twobit.npol, ndat), dtype=torch.float32, device=raw.device)
    w = torch.empty(ndat // twobit.ndat_per_weight, dtype=torch.int32, device=raw.device)
    L.check(ctx.lib.b200_unpack_twobit(ctx.h, C.byref(twobit), _ptr(raw), ndat, _ptr(out), ndat, _ptr(w)))
    return out, w


def unpack(ctx, desc, raw, ndat):
    """Unpacker device hook: raw uint8 CUDA tensor -> float32 [nchan, npol, ndat*ndim]."""
    _need_cuda(raw, "raw")
    out = torch.empty((desc.nchan, desc.npol, ndat * desc.ndim), dtype=torch.float32, device=raw.device)
    L.check(ctx.lib.b200_unpack(ctx.h, C.byref(desc), _ptr(raw), ndat, _ptr(out), ndat * desc.ndim))
    return out


def make_fb_desc(input_real, input_nchan, npol, nchan_subband, freq_res, nfilt_pos, nfilt_neg, response=None,
                 max_npart=0):
    d = L.FbDesc()
    d.input_real = int(input_real)
    d.input_nchan, d.npol, d.nchan_subband, d.freq_res = input_nchan, npol, nchan_subband, freq_res
    d.nfilt_pos, d.nfilt_neg = nfilt_pos, nfilt_neg
    d.max_npart = max_npart
    keep = None
    if response is not None:
        keep = np.ascontiguousarray(response, np.complex64)
        assert keep.size == input_nchan * nchan_subband * freq_res, "response size"
        d.h_response = keep.ctypes.data
    return d, keep


class FilterbankEngine:
    """dsp::Filterbank::Engine / dsp::Convolution::Engine (FilterbankEngine.h:15-44, Convolution.h:158-167)."""

    def __init__(self, ctx, input_real, input_nchan, npol, nchan_subband, freq_res, nfilt_pos, nfilt_neg,
                 response=None, max_npart=0):
        self.ctx = ctx
        d, keep = make_fb_desc(input_real, input_nchan, npol, nchan_subband, freq_res, nfilt_pos, nfilt_neg,
                               response, max_npart)
        h = C.c_void_p()
        L.check(ctx.lib.b200_fb_plan_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.desc = d
        info = L.FbInfo()
        L.check(ctx.lib.b200_fb_plan_info(h, C.byref(info)))
        self.info = info
        self.ndim = 1 if input_real else 2
        self.nchan = input_nchan * nchan_subband

    def npart(self, ndat):
        # Filterbank::resize_output (Filterbank.C:401-402)
        i = self.info
        return (ndat - i.nsamp_overlap) // i.nsamp_step if ndat > i.nsamp_overlap else 0

    def perform(self, x, npart=None):
        """x: [input_nchan, npol, ndat*ndim] float32 CUDA -> [nchan, npol, npart*nkeep*2] float32."""
        _need_cuda(x, "x")
        i = self.info
        ndat = x.shape[2] // self.ndim
        if npart is None:
            npart = self.npart(ndat)
        out = torch.empty((self.nchan, self.desc.npol, npart * i.nkeep * 2), dtype=torch.float32, device=x.device)
        L.check(self.ctx.lib.b200_fb_perform(self.h, _ptr(x), x.shape[2], _ptr(out), out.shape[2], npart,
                                             i.nsamp_step * self.ndim, i.nkeep * 2))
        return out

    def __del__(self):
        try:
            self.ctx.lib.b200_fb_plan_destroy(self.h)
        except Exception:
            pass


def detect(ctx, state, ndim_out, v, npol=2):
    """dsp::Detection::Engine. v: [nchan, npol, ndat*2] float32 CUDA -> [nchan, npol', ndat*ndim']."""
    _need_cuda(v, "v")
    s = L.STATE[state] if isinstance(state, str) else state
    nchan, npol, n2 = v.shape
    ndat = n2 // 2
    if s >= L.COHERENCE:
        onpol, ondim = 4 // ndim_out, ndim_out
    else:
        onpol, ondim = (npol if s == L.PPQQ else 1), 1
    out = torch.empty((nchan, onpol, ndat * ondim), dtype=torch.float32, device=v.device)
    L.check(ctx.lib.b200_detect(ctx.h, s, ondim, _ptr(v), n2, nchan, npol, ndat, _ptr(out), ndat * ondim))
    return out


class FoldEngine:
    """dsp::Fold::Engine (Fold.h:249-312): owns the accumulating device PhaseSeries."""

    def __init__(self, ctx, nchan, npol, ndim, nbin, handle=None):
        self.ctx = ctx
        self.nchan, self.npol, self.ndim, self.nbin = nchan, npol, ndim, nbin
        self.owned = handle is None
        if handle is None:
            handle = C.c_void_p()
            L.check(ctx.lib.b200_fold_create(ctx.h, nchan, npol, ndim, nbin, C.byref(handle)))
        self.h = handle

    def set_bins(self, phi, phase_per_sample, ndat, idat_start=0):
        n = C.c_uint64(0)
        L.check(self.ctx.lib.b200_fold_set_bins(self.h, phi, phase_per_sample, ndat, idat_start, C.byref(n)))
        return n.value

    def set_bins_weighted(self, phi, phase_per_sample, ndat, idat_start, weights, ndatperweight, weight_idat=0):
        """weights: int32 CUDA tensor of per-window flags (Fold.C:687-716)."""
        _need_cuda(weights, "weights")
        L.check(self.ctx.lib.b200_fold_set_bins_weighted(self.h, phi, phase_per_sample, ndat, idat_start, _ptr(weights),
                                                         weights.numel(), ndatperweight, weight_idat))

    def get_bin_hits(self):
        h = np.zeros(self.nbin, np.uint32)
        L.check(self.ctx.lib.b200_fold_get_bin_hits(self.h, h.ctypes.data_as(C.c_void_p)))
        return h

    def fold(self, x):
        _need_cuda(x, "x")
        L.check(self.ctx.lib.b200_fold_fold(self.h, _ptr(x), x.shape[2]))

    def synch(self):
        p = np.zeros((self.nchan, self.npol, self.nbin * self.ndim), np.float32)
        L.check(self.ctx.lib.b200_fold_synch(self.h, p.ctypes.data_as(C.c_void_p)))
        return p

    def hits(self):
        h = np.zeros(self.nbin, np.uint32)
        n = C.c_uint64(0)
        L.check(self.ctx.lib.b200_fold_get_hits(self.h, h.ctypes.data_as(C.c_void_p), C.byref(n)))
        return h, n.value

    def zero(self):
        L.check(self.ctx.lib.b200_fold_zero(self.h))

    def device_profile(self):
        """torch view of the device profile (for NCCL reductions at sub-integration boundaries)."""
        ptr = self.ctx.lib.b200_fold_device_profile(self.h)
        n = self.nchan * self.npol * self.nbin * self.ndim
        return _tensor_from_ptr(ptr, n, torch.float32, self.ctx.device)

    def device_hits(self):
        ptr = self.ctx.lib.b200_fold_device_hits(self.h)
        return _tensor_from_ptr(ptr, self.nbin, torch.int32, self.ctx.device)

    def __del__(self):
        try:
            if self.owned:
                self.ctx.lib.b200_fold_destroy(self.h)
        except Exception:
            pass


class _CudaArrayView:
    def __init__(self, ptr, nbytes, typestr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _tensor_from_ptr(ptr, n, dtype, device):
    typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaArrayView(ptr, n * 4, typestr, n), device="cuda:%d" % device)


class Pipeline:
    """Fused raw bytes -> PhaseSeries path (b200_pipeline_*)."""

    def __init__(self, ctx, unpack_desc, fb_desc, response_keepalive, detect_state, detect_ndim, nbin):
        self.ctx = ctx
        d = L.PipelineDesc()
        d.unpack = unpack_desc
        d.fb = fb_desc
        d.detect_state = L.STATE[detect_state] if isinstance(detect_state, str) else detect_state
        d.detect_ndim = detect_ndim
        d.nbin = nbin
        self._keep = response_keepalive
        h = C.c_void_p()
        L.check(ctx.lib.b200_pipeline_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.desc = d
        info = L.FbInfo()
        L.check(ctx.lib.b200_pipeline_info(h, C.byref(info)))
        self.info = info
        self.nchan = fb_desc.input_nchan * fb_desc.nchan_subband
        s = d.detect_state
        if s >= L.COHERENCE:
            self.dnpol, self.dndim = 4 // detect_ndim, detect_ndim
        else:
            self.dnpol, self.dndim = (fb_desc.npol if s == L.PPQQ else 1), 1
        self.nbin = nbin
        self.fold = None
        if nbin:
            fh = C.c_void_p(ctx.lib.b200_pipeline_fold(h))
            self.fold = FoldEngine(ctx, self.nchan, self.dnpol, self.dndim, nbin, handle=fh)

    def _detected(self, npart, out):
        """(tensor, pointer, span) of the detected series when the pipeline has no fold stage."""
        if self.nbin:
            return None, None, 0
        dspan = npart * self.info.nkeep * self.dndim
        if out is None:
            out = torch.empty((self.nchan, self.dnpol, dspan), dtype=torch.float32, device="cuda:%d" % self.ctx.device)
        else:
            _need_cuda(out, "out")
            assert out.shape[0] == self.nchan and out.shape[1] == self.dnpol and out.shape[2] >= dspan
            dspan = out.shape[2]
        return out, _ptr(out), dspan

    def execute(self, d_input, npart, phi=0.0, pps=0.0, first_sample=0, input_span=0, out=None):
        _need_cuda(d_input, "d_input")
        det, dptr, dspan = self._detected(npart, out)
        L.check(self.ctx.lib.b200_pipeline_execute(self.h, _ptr(d_input), input_span, first_sample, npart, phi, pps,
                                                   dptr, dspan))
        return det

    def execute_host(self, h_input, npart, phi=0.0, pps=0.0, first_sample=0, out=None):
        """h_input: numpy uint8 array or pinned torch CPU tensor of raw bytes; it must stay untouched until
        input_consumed() (the copies are asynchronous)."""
        if isinstance(h_input, torch.Tensor):
            ptr, nbytes = h_input.data_ptr(), h_input.numel() * h_input.element_size()
        else:
            ptr, nbytes = h_input.ctypes.data, h_input.nbytes
        det, dptr, dspan = self._detected(npart, out)
        L.check(self.ctx.lib.b200_pipeline_execute_host(self.h, C.c_void_p(ptr), nbytes, first_sample, npart, phi, pps,
                                                        dptr, dspan))
        return det

    # ---- observation-driven blocks: the library evaluates the predictor and keeps the PhaseSeries attributes ----
    def set_observation(self, raw_obs):
        L.check(self.ctx.lib.b200_pipeline_set_observation(self.h, C.byref(raw_obs)))

    def set_predictor(self, predictor, reference_phase=0.0):
        """predictor: hostmath.Polyco"""
        L.check(self.ctx.lib.b200_pipeline_set_predictor(self.h, C.byref(predictor.pc), reference_phase))

    def set_folding_period(self, period, reference_phase=0.0, reference_epoch=None):
        """reference_epoch: (day, sec, frac) or None (= MJD 0, the reference's default)"""
        ep = C.byref(L.Mjd(*reference_epoch)) if reference_epoch is not None else None
        L.check(self.ctx.lib.b200_pipeline_set_folding_period(self.h, period, reference_phase, ep))

    def execute_obs(self, d_input, npart, obs_sample, first_sample=0, input_span=0):
        _need_cuda(d_input, "d_input")
        L.check(self.ctx.lib.b200_pipeline_execute_obs(self.h, _ptr(d_input), input_span, first_sample, npart, obs_sample))

    def execute_host_obs(self, h_input, npart, obs_sample, first_sample=0):
        if isinstance(h_input, torch.Tensor):
            ptr, nbytes = h_input.data_ptr(), h_input.numel() * h_input.element_size()
        else:
            ptr, nbytes = h_input.ctypes.data, h_input.nbytes
        L.check(self.ctx.lib.b200_pipeline_execute_host_obs(self.h, C.c_void_p(ptr), nbytes, first_sample, npart, obs_sample))

    def phase_series(self):
        """-> phaseseries.PhaseSeries holding the accumulated sums, hits and attributes (synchronises)."""
        from . import phaseseries as P
        out = P.PhaseSeries(self.nchan, self.dnpol, self.dndim, self.nbin)
        L.check(self.ctx.lib.b200_pipeline_get_phase_series(self.h, C.byref(out.ps)))
        out._bind()
        return out

    def reset(self):
        L.check(self.ctx.lib.b200_pipeline_reset(self.h))

    # ---- streaming input with block-edge carry (InputBuffering) ----
    def stream_begin(self, max_block_samples, obs_sample0=0):
        L.check(self.ctx.lib.b200_pipeline_stream_begin(self.h, max_block_samples, obs_sample0))

    def feed(self, block, nsamples, out=None):
        """block: raw bytes of `nsamples` new samples -- a CUDA uint8 tensor, a pinned CPU tensor or a numpy array.
        Returns the number of overlap-save parts this feed completed (and `out` holds their detected samples)."""
        n = C.c_uint64(0)
        dptr, dspan = (None, 0) if out is None else (_ptr(out), out.shape[2])
        if isinstance(block, torch.Tensor) and block.is_cuda:
            L.check(self.ctx.lib.b200_pipeline_feed(self.h, _ptr(block), nsamples, dptr, dspan, C.byref(n)))
        else:
            ptr = block.data_ptr() if isinstance(block, torch.Tensor) else block.ctypes.data
            L.check(self.ctx.lib.b200_pipeline_feed_host(self.h, C.c_void_p(ptr), nsamples, dptr, dspan, C.byref(n)))
        return n.value

    def set_deterministic(self, lsb=-1.0):
        """Reproducible fixed-point accumulation of the PhaseSeries (b200_fold_set_deterministic); lsb < 0: automatic."""
        L.check(self.ctx.lib.b200_pipeline_set_deterministic(self.h, lsb))

    def reserve(self, max_npart):
        L.check(self.ctx.lib.b200_pipeline_reserve(self.h, max_npart))

    def input_consumed(self):
        L.check(self.ctx.lib.b200_pipeline_input_consumed(self.h))

    def synch(self):
        p = np.zeros((self.nchan, self.dnpol, self.nbin * self.dndim), np.float32)
        h = np.zeros(self.nbin, np.uint32)
        n = C.c_uint64(0)
        L.check(self.ctx.lib.b200_pipeline_synch(self.h, p.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p),
                                                 C.byref(n)))
        return p, h, n.value

    def zero(self):
        L.check(self.ctx.lib.b200_pipeline_zero(self.h))

    def __del__(self):
        try:
            self.ctx.lib.b200_pipeline_destroy(self.h)
        except Exception:
            pass


def weights_convolve(ctx, w, ndat_per_weight, weight_idat, ndat, nfft, nkeep):
    """WeightedTimeSeries::convolve_weights on device flags (int32 CUDA tensor) -> new tensor."""
    _need_cuda(w, "w")
    out = torch.empty_like(w)
    scratch = torch.empty(ndat // nkeep + 2, dtype=torch.int32, device=w.device)
    L.check(ctx.lib.b200_weights_convolve(ctx.h, _ptr(w), w.numel(), ndat_per_weight, weight_idat, ndat, nfft, nkeep,
                                          _ptr(out), _ptr(scratch)))
    return out


def weights_scrunch(ctx, w, ndat_per_weight, weight_idat, nscrunch):
    """WeightedTimeSeries::scrunch_weights -> (flags tensor, ndat_per_weight, weight_idat)."""
    _need_cuda(w, "w")
    n, npw, wi = C.c_uint64(w.numel()), C.c_uint(ndat_per_weight), C.c_uint64(weight_idat)
    out = torch.empty_like(w)
    L.check(ctx.lib.b200_weights_scrunch(ctx.h, _ptr(w), C.byref(n), C.byref(npw), C.byref(wi), nscrunch, _ptr(out)))
    return out[: n.value], npw.value, wi.value


# ---- host-only helpers (exact fold bin plan) -------------------------------------------------
def phase_segments(phi, pps, ndat, max_segments=1 << 16):
    lib = L.load()
    seg = (L.PhaseSegment * max_segments)()
    phi_end = C.c_double(0)
    n = lib.b200_phase_segments(phi, pps, ndat, seg, max_segments, C.byref(phi_end))
    if n < 0:
        raise L.B200Error(1, "more than %d phase segments" % max_segments)
    return [seg[i] for i in range(n)], phi_end.value


def phase_bins_sequential(phi, pps, nbin, ndat):
    lib = L.load()
    bins = np.zeros(ndat, np.uint32)
    phi_end = C.c_double(0)
    lib.b200_phase_bins_sequential(phi, pps, nbin, ndat, bins.ctypes.data_as(C.c_void_p), C.byref(phi_end))
    return bins, phi_end.value


def expand_segments_numpy(segs, nbin, ndat):
    """Host expansion of phase segments with the same exact arithmetic as k_expand_bins (tests)."""
    bins = np.zeros(ndat, np.uint32)
    for s in segs:
        t = np.arange(s.count, dtype=np.uint64)
        a = np.uint64(s.a0) + t * np.uint64(s.step)
        # k_expand_bins scales by multiplying with the power of two (exact: the product is a normal double) and keeps
        # ldexp for phases below 2^-900
        if s.scale_exp >= -1000:
            phi = a.astype(np.float64) * np.ldexp(1.0, s.scale_exp)
        else:
            phi = np.ldexp(a.astype(np.float64), s.scale_exp)
        bins[s.start:s.start + s.count] = (phi * float(nbin)).astype(np.uint32)
    return bins


class Rescale:
    """dsp::Rescale on the device (b200_rescale_*): per-channel offset/scale of a detected FPT series."""

    def __init__(self, ctx, nchan, npol, interval_samples=0, constant=False):
        self.ctx, self.nchan, self.npol = ctx, nchan, npol
        h = C.c_void_p()
        L.check(ctx.lib.b200_rescale_create(ctx.h, nchan, npol, interval_samples, int(constant), C.byref(h)))
        self.h = h

    def transform(self, x, out=None):
        _need_cuda(x, "x")
        out = torch.empty_like(x) if out is None else out
        L.check(self.ctx.lib.b200_rescale_transform(self.h, _ptr(x), x.shape[2], x.shape[2], _ptr(out), out.shape[2]))
        return out

    def offset_scale(self):
        o = np.zeros((self.nchan, self.npol), np.float32)
        s = np.zeros((self.nchan, self.npol), np.float32)
        L.check(self.ctx.lib.b200_rescale_get(self.h, o.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)))
        return o, s

    def __del__(self):
        try:
            self.ctx.lib.b200_rescale_destroy(self.h)
        except Exception:
            pass


def sigproc_digitize8(ctx, x, input_scale=1.0, scale_fac=1.0, rescale=True, bandwidth=-1.0, swap=False, out=None):
    """dsp::SigProcDigitizer::pack (8 bit): detected FPT CUDA tensor [nchan, npol, ndat] -> uint8 [ndat, npol, nchan]."""
    _need_cuda(x, "x")
    nchan, npol, ndat = x.shape
    digi_mean, digi_scale, xpol = np.float32(127.5), np.float32(np.float32(127.5) / np.float32(6)), np.float32(0)
    if not rescale:
        xpol, digi_mean, digi_scale = digi_mean, np.float32(0), np.float32(1)
    digi_scale = np.float32(np.float64(digi_scale) / (np.float64(input_scale) * np.float64(scale_fac)))
    if out is None:
        out = torch.empty((ndat, npol, nchan), dtype=torch.uint8, device=x.device)
    L.check(ctx.lib.b200_sigproc_digitize8(ctx.h, _ptr(x), ndat, nchan, npol, ndat, float(digi_scale), float(digi_mean),
                                          float(xpol), int(bandwidth > 0), int(swap), _ptr(out)))
    return out
