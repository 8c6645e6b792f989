"""ctypes wrapper of the CPU oracle (oracle/dspsr_oracle.cpp, oracle/orc_fft.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (dspsr_b200/) never imports this.
Parity status: pinned bit for bit to reference code compiled in place (oracle/ref.mk -> oracle/_ref/, checked by
tests/test_ref_pin.py) for every stage; UNPINNED only for the third-party primitives that live in PSRCHIVE (FFTW
transforms, TEMPO polyco evaluation, JA98 / normal-distribution numbers) -- see the header of dspsr_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-j4"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class DedispParams(C.Structure):
    _fields_ = [
        ("centre_frequency", C.c_double),
        ("bandwidth", C.c_double),
        ("dispersion_measure", C.c_double),
        ("doppler_shift", C.c_double),
        ("input_nchan", C.c_uint),
        ("nchan", C.c_uint),
        ("input_dual_sideband", C.c_int),
        ("input_dc_centred", C.c_int),
        ("input_swap", C.c_int),
        ("frequency_resolution", C.c_uint),
        ("impulse_pos", C.c_uint),
        ("impulse_neg", C.c_uint),
        ("ndat", C.c_uint),
        ("unsupported_channels", C.c_uint),
    ]


class FB(C.Structure):
    _fields_ = [
        ("input_real", C.c_int),
        ("input_nchan", C.c_uint),
        ("npol", C.c_uint),
        ("nchan", C.c_uint),
        ("freq_res", C.c_uint),
        ("nfilt_pos", C.c_uint),
        ("nfilt_neg", C.c_uint),
        ("nchan_subband", C.c_uint),
        ("n_fft", C.c_uint),
        ("nsamp_fft", C.c_uint),
        ("nsamp_overlap", C.c_uint),
        ("nsamp_step", C.c_uint),
        ("nkeep", C.c_uint),
    ]


class Conv(C.Structure):
    _fields_ = [
        ("input_real", C.c_int),
        ("nchan", C.c_uint),
        ("npol", C.c_uint),
        ("n_fft", C.c_uint),
        ("nfilt_pos", C.c_uint),
        ("nfilt_neg", C.c_uint),
        ("nsamp_fft", C.c_uint),
        ("nsamp_overlap", C.c_uint),
        ("nsamp_step", C.c_uint),
    ]


class Polyco(C.Structure):
    _fields_ = [
        ("tmid_day", C.c_int),
        ("tmid_sec", C.c_double),
        ("rphase_int", C.c_double),
        ("rphase_frac", C.c_double),
        ("f0", C.c_double),
        ("span_min", C.c_double),
        ("obsfreq", C.c_double),
        ("dm", C.c_double),
        ("ncoef", C.c_int),
        ("coef", C.c_double * 32),
    ]


class Pipe(C.Structure):
    _fields_ = [
        ("unpack_fmt", C.c_int),
        ("input_nchan", C.c_uint),
        ("npol", C.c_uint),
        ("ndim", C.c_uint),
        ("lut", C.c_void_p),
        ("scale", C.c_float),
        ("use_filterbank", C.c_int),
        ("fb", FB),
        ("conv", Conv),
        ("H", C.c_void_p),
        ("detect_state", C.c_int),
        ("detect_ndim", C.c_uint),
        ("nbin", C.c_uint),
        ("twobit", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_ja98_optimal_spacing.restype = C.c_double
        L.orc_ja98_optimal_spacing.argtypes = [C.c_uint]
        L.orc_normal_cdf.restype = C.c_double
        L.orc_normal_cdf.argtypes = [C.c_double]
        L.orc_bittable_unique_values.restype = C.c_double
        L.orc_bittable_unique_values.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.orc_bittable8.restype = C.c_double
        L.orc_bittable8.argtypes = [C.c_int, C.c_void_p]
        L.orc_optimal_fft_length.restype = C.c_int64
        L.orc_optimal_fft_length.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_dedisp_prepare.restype = C.c_int
        L.orc_dedisp_build.restype = C.c_int
        L.orc_fb_npart.restype = C.c_uint64
        L.orc_fb_npart.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_conv_npart.restype = C.c_uint64
        L.orc_conv_npart.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_fold_plan.restype = C.c_uint64
        L.orc_fold_plan.argtypes = [C.c_double, C.c_double, C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_polyco_parse.restype = C.c_int
        L.orc_polyco_phase.restype = C.c_double
        L.orc_polyco_phase.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.orc_polyco_frequency.restype = C.c_double
        L.orc_polyco_frequency.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- a1
def bittable8(twos_complement=True):
    lut = np.zeros(256, np.float32)
    scale = lib().orc_bittable8(int(twos_complement), _p(lut))
    return lut, scale


def bittable_values(nbit, twos_complement):
    v = np.zeros(1 << nbit, np.float32)
    scale = lib().orc_bittable_unique_values(nbit, int(twos_complement), _p(v))
    return v, scale


# ----------------------------------------------------------------------------- a2-a5
def unpack_caspsr(raw, ndat, lut):
    out = np.zeros((1, 2, ndat), np.float32)
    lib().orc_unpack_caspsr(_p(raw), C.c_uint64(ndat), _p(lut), _p(out), C.c_uint64(ndat))
    return out


def unpack_generic8(raw, ndat, nchan, npol, ndim, lut):
    out = np.zeros((nchan, npol, ndat * ndim), np.float32)
    lib().orc_unpack_generic8(_p(raw), C.c_uint64(ndat), nchan, npol, ndim, _p(lut), _p(out),
                              C.c_uint64(ndat * ndim), None)
    return out


def unpack_meerkat(raw, ndat, nchan, npol, scale, sample_swap=1):
    out = np.zeros((nchan, npol, ndat * 2), np.float32)
    lib().orc_unpack_meerkat(_p(raw), C.c_uint64(ndat), nchan, npol, C.c_float(scale), sample_swap, _p(out),
                             C.c_uint64(ndat * 2))
    return out


def unpack_uwb(raw, ndat, npol):
    out = np.zeros((1, npol, ndat * 2), np.float32)
    lib().orc_unpack_uwb(_p(raw), C.c_uint64(ndat), npol, _p(out), C.c_uint64(ndat * 2))
    return out


# ----------------------------------------------------------------------------- a6
class TwoBit:
    """TwoBitCorrection (CPSR2 convention): level lookup + excision limits, and the unpack loop."""

    def __init__(self, table_type=0, threshold=0.9674, cutoff_sigma=10.0, ndat_per_weight=512):
        L = lib()
        L.orc_twobit_create.restype = C.c_void_p
        L.orc_twobit_create.argtypes = [C.c_int, C.c_double, C.c_float, C.c_uint]
        self.h = C.c_void_p(L.orc_twobit_create(table_type, threshold, cutoff_sigma, ndat_per_weight))
        self.ndat_per_weight = ndat_per_weight
        a, b = C.c_uint(0), C.c_uint(0)
        L.orc_twobit_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.orc_twobit_info(self.h, C.byref(a), C.byref(b))
        self.nlow_min, self.nlow_max = a.value, b.value

    def levels(self, nlow):
        lo, hi = C.c_float(0), C.c_float(0)
        f = lib().orc_twobit_levels
        f.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        f(self.h, nlow, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def unpack(self, raw, ndat, npol):
        """raw uint8 -> (float32 [1, npol, ndat], weights uint32 [npol, ndat/ndat_per_weight] after mask_weights)."""
        raw = np.ascontiguousarray(raw, np.uint8)
        out = np.zeros((1, npol, ndat), np.float32)
        nw = ndat // self.ndat_per_weight
        w = np.zeros((npol, nw), np.uint32)
        f = lib().orc_unpack_twobit
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        f(self.h, _p(raw), ndat, npol, _p(out), ndat, _p(w), nw)
        return out, w

    def __del__(self):
        try:
            lib().orc_twobit_destroy.argtypes = [C.c_void_p]
            lib().orc_twobit_destroy(self.h)
        except Exception:
            pass


def ja98_levels(phi):
    lo, hi = C.c_double(0), C.c_double(0)
    f = lib().orc_ja98_levels
    f.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(phi, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


# ----------------------------------------------------------------------------- a7-a9
def optimal_fft_length(nbad, nfft_max=0):
    return lib().orc_optimal_fft_length(nbad, nfft_max)


def dedispersion(centre_frequency, bandwidth, dm, input_nchan, nchan, input_real, frequency_resolution=0,
                 dual_sideband=None, dc_centred=False, swap=False, build=True):
    """Dedispersion::prepare + build + match.  Returns (params, H[nchan, ndat] complex64 or None)."""
    d = DedispParams()
    d.centre_frequency = centre_frequency
    d.bandwidth = bandwidth
    d.dispersion_measure = dm
    d.doppler_shift = 1.0
    d.input_nchan = input_nchan
    d.nchan = nchan
    d.input_dual_sideband = int((not input_real) if dual_sideband is None else dual_sideband)
    d.input_dc_centred = int(dc_centred)
    d.input_swap = int(swap)
    d.frequency_resolution = frequency_resolution
    rc = lib().orc_dedisp_prepare(C.byref(d))
    if rc != 0:
        raise ValueError("orc_dedisp_prepare failed rc=%d" % rc)
    H = None
    if build:
        H = np.zeros((nchan, d.ndat), np.complex64)
        lib().orc_dedisp_build(C.byref(d), _p(H))
    return d, H


def response_operate(spectrum, H):
    """Response::operate (Response.C:385-444): spectrum * H in the reference's float operation order."""
    out = np.ascontiguousarray(spectrum, np.complex64).copy()
    Hc = np.ascontiguousarray(H, np.complex64)
    lib().orc_response_operate(_p(out), _p(Hc), C.c_uint64(out.size))
    return out


# ----------------------------------------------------------------------------- FFT
def fcc1d(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    lib().orc_fft_fcc1d(x.size, _p(out), _p(x))
    return out


def bcc1d(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    lib().orc_fft_bcc1d(x.size, _p(out), _p(x))
    return out


def frc1d(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size // 2 + 1, np.complex64)
    lib().orc_fft_frc1d(x.size, _p(out), _p(x))
    return out


# ----------------------------------------------------------------------------- a10/a11
def fb_sizes(input_real, input_nchan, npol, nchan, freq_res, nfilt_pos, nfilt_neg):
    f = FB()
    f.input_real = int(input_real)
    f.input_nchan, f.npol, f.nchan, f.freq_res = input_nchan, npol, nchan, freq_res
    f.nfilt_pos, f.nfilt_neg = nfilt_pos, nfilt_neg
    lib().orc_fb_sizes(C.byref(f))
    return f


def filterbank(f, x, H):
    """x: [input_nchan, npol, ndat*ndim] float32 -> [nchan, npol, npart*nkeep] complex64."""
    x = np.ascontiguousarray(x, np.float32)
    ndim = 1 if f.input_real else 2
    ndat = x.shape[2] // ndim
    npart = lib().orc_fb_npart(C.byref(f), ndat)
    out = np.zeros((f.nchan, f.npol, npart * f.nkeep), np.complex64)
    if npart:
        Hp = _p(np.ascontiguousarray(H, np.complex64)) if H is not None else None
        lib().orc_filterbank_parts(C.byref(f), _p(x), C.c_uint64(x.shape[2]), Hp, _p(out),
                                   C.c_uint64(2 * npart * f.nkeep), C.c_uint64(0), C.c_uint64(npart))
    return out


def conv_sizes(input_real, nchan, npol, n_fft, nfilt_pos, nfilt_neg):
    c = Conv()
    c.input_real = int(input_real)
    c.nchan, c.npol, c.n_fft, c.nfilt_pos, c.nfilt_neg = nchan, npol, n_fft, nfilt_pos, nfilt_neg
    lib().orc_conv_sizes(C.byref(c))
    return c


def convolution(c, x, H):
    x = np.ascontiguousarray(x, np.float32)
    ndim = 1 if c.input_real else 2
    ndat = x.shape[2] // ndim
    npart = lib().orc_conv_npart(C.byref(c), ndat)
    nkeep = c.n_fft - c.nfilt_pos - c.nfilt_neg
    out = np.zeros((c.nchan, c.npol, npart * nkeep), np.complex64)
    if npart:
        Hh = np.ascontiguousarray(H, np.complex64)
        lib().orc_convolution_parts(C.byref(c), _p(x), C.c_uint64(x.shape[2]), _p(Hh), _p(out),
                                    C.c_uint64(2 * npart * nkeep), C.c_uint64(0), C.c_uint64(npart))
    return out


# ----------------------------------------------------------------------------- a12
STATE = {"Intensity": 0, "PPQQ": 1, "Coherence": 2, "Stokes": 3}


def detect_shape(state, ndim_out):
    s = STATE[state] if isinstance(state, str) else state
    if s >= 2:
        return 4 // ndim_out, ndim_out
    return (2, 1) if s == 1 else (1, 1)


def detect(state, ndim_out, v):
    """v: [nchan, 2, ndat] complex64 -> [nchan, npol', ndat*ndim'] float32."""
    s = STATE[state] if isinstance(state, str) else state
    v = np.ascontiguousarray(v, np.complex64)
    nchan, npol, ndat = v.shape
    onpol, ondim = detect_shape(s, ndim_out)
    out = np.zeros((nchan, onpol, ndat * ondim), np.float32)
    lib().orc_detect(s, ondim, _p(v), C.c_uint64(2 * ndat), nchan, npol, C.c_uint64(ndat), _p(out),
                     C.c_uint64(ndat * ondim))
    return out


# ----------------------------------------------------------------------------- a13
def fold_plan(phi, pps, nbin, ndat):
    binplan = np.zeros(ndat, np.uint32)
    hits = np.zeros(nbin, np.uint32)
    phi_out = C.c_double(0)
    n = lib().orc_fold_plan(phi, pps, nbin, ndat, _p(binplan), _p(hits), C.byref(phi_out))
    return binplan, hits, n, phi_out.value


def fold(x, ndim, binplan, nbin, profile=None, idat_start=0):
    """x: [nchan, npol, ndat*ndim] float32; returns profile [nchan, npol, nbin*ndim] (+=)."""
    x = np.ascontiguousarray(x, np.float32)
    nchan, npol, n = x.shape
    if profile is None:
        profile = np.zeros((nchan, npol, nbin * ndim), np.float32)
    lib().orc_fold(_p(x), C.c_uint64(n), nchan, npol, ndim, C.c_uint64(idat_start), C.c_uint64(binplan.size),
                   _p(binplan), nbin, _p(profile))
    return profile


# ----------------------------------------------------------------------------- a16
def polyco_parse(text):
    pc = Polyco()
    rc = lib().orc_polyco_parse(text.encode(), C.byref(pc))
    if rc != 0:
        raise ValueError("orc_polyco_parse rc=%d" % rc)
    return pc


def polyco_phase(pc, day, sec, frac):
    turns = C.c_double(0)
    f = lib().orc_polyco_phase(C.byref(pc), day, sec, frac, C.byref(turns))
    return f, turns.value


def polyco_frequency(pc, day, sec, frac):
    return lib().orc_polyco_frequency(C.byref(pc), day, sec, frac)


# ----------------------------------------------------------------------------- whole path
def make_pipe(unpack_fmt, input_nchan, npol, ndim, lut, scale, fb, conv, H, detect_state, detect_ndim, nbin,
              twobit=None):
    """unpack_fmt: 0 CASPSR, 1 generic 8-bit, 2 MeerKAT, 3 UWB, 5 two-bit (pass `twobit`, a TwoBit)."""
    p = Pipe()
    p.unpack_fmt = unpack_fmt
    p.input_nchan, p.npol, p.ndim = input_nchan, npol, ndim
    p._keep = (lut, H, twobit)
    p.twobit = twobit.h if twobit is not None else None
    p.lut = lut.ctypes.data if lut is not None else None
    p.scale = scale
    p.use_filterbank = int(fb is not None)
    if fb is not None:
        p.fb = fb
    if conv is not None:
        p.conv = conv
    p.H = H.ctypes.data if H is not None else None
    p.detect_state = STATE[detect_state] if isinstance(detect_state, str) else detect_state
    p.detect_ndim = detect_ndim
    p.nbin = nbin
    return p


def pipe_profile_shape(p):
    out_nchan = p.fb.nchan if p.use_filterbank else p.conv.nchan
    onpol, ondim = detect_shape(p.detect_state, p.detect_ndim)
    return out_nchan, onpol, p.nbin * ondim


def pipe_run(p, raw, nblock, parts_per_block, phi, pps, nthread=1, with_total=False):
    """-> (profile, hits[, ndat_total]): nblock Fold calls of parts_per_block parts each, `dspsr -t nthread` style."""
    shape = pipe_profile_shape(p)
    profile = np.zeros(shape, np.float32)
    hits = np.zeros(p.nbin, np.uint32)
    phi = np.ascontiguousarray(phi, np.float64)
    pps = np.ascontiguousarray(pps, np.float64)
    ntot = C.c_uint64(0)
    lib().orc_pipe_run(C.byref(p), _p(raw), C.c_uint64(nblock), C.c_uint64(parts_per_block), _p(phi), _p(pps),
                       C.c_uint(nthread), _p(profile), _p(hits), C.byref(ntot))
    if ntot.value == 2 ** 64 - 1:
        raise RuntimeError("oracle: the reference would throw here (weights exhausted, Fold.C:699 / WeightedTimeSeries.C:626)")
    if with_total:
        return profile, hits, ntot.value
    return profile, hits


def pipe_blocks(p, raw, blocks):
    """Fold calls of arbitrary extent: blocks = [(ipart0, npart, phi, pps), ...] -> (profile, hits, ndat_folded)."""
    shape = pipe_profile_shape(p)
    profile = np.zeros(shape, np.float32)
    hits = np.zeros(p.nbin, np.uint32)
    f = lib().orc_pipe_block
    f.restype = C.c_uint64
    nfold = 0
    for ipart0, npart, phi, pps in blocks:
        n = f(C.byref(p), _p(raw), C.c_uint64(ipart0), C.c_uint64(npart), C.c_double(phi), C.c_double(pps),
              _p(profile), _p(hits), None)
        if n == 2 ** 64 - 1:
            raise RuntimeError("oracle: the reference would throw here")
        nfold += n
    return profile, hits, nfold


def pipe_block_detected(p, raw, ipart0, npart):
    """One block of the path without fold: the detected series [out_nchan, out_npol, npart*nkeep*out_ndim]."""
    out_nchan = p.fb.nchan if p.use_filterbank else p.conv.nchan
    nkeep = p.fb.nkeep if p.use_filterbank else p.conv.n_fft - p.conv.nfilt_pos - p.conv.nfilt_neg
    onpol, ondim = detect_shape(p.detect_state, p.detect_ndim)
    det = np.zeros((out_nchan, onpol, npart * nkeep * ondim), np.float32)
    f = lib().orc_pipe_block
    f.restype = C.c_uint64
    f(C.byref(p), _p(raw), C.c_uint64(ipart0), C.c_uint64(npart), C.c_double(0), C.c_double(0), None, None, _p(det))
    return det


def convolve_weights(weights, ndat_per_weight, weight_idat, ndat, nfft, nkeep):
    """WeightedTimeSeries::convolve_weights (WeightedTimeSeries.C:582-690) on a copy of `weights`."""
    w = np.ascontiguousarray(weights, np.uint32).copy()
    rc = lib().orc_convolve_weights(_p(w), C.c_uint64(w.size), C.c_uint(ndat_per_weight), C.c_uint64(weight_idat),
                                    C.c_uint64(ndat), C.c_uint(nfft), C.c_uint(nkeep))
    if rc != 0:
        raise RuntimeError("convolve_weights: end_weight > nweights")
    return w


def scrunch_weights(weights, ndat_per_weight, weight_idat, nscrunch):
    """WeightedTimeSeries::scrunch_weights (:692-780): -> (weights, ndat_per_weight, weight_idat)."""
    w = np.ascontiguousarray(weights, np.uint32).copy()
    n, npw, wi = C.c_uint64(w.size), C.c_uint(ndat_per_weight), C.c_uint64(weight_idat)
    lib().orc_scrunch_weights(_p(w), C.byref(n), C.byref(npw), C.byref(wi), C.c_uint(nscrunch))
    return w[: n.value], npw.value, wi.value


def fold_plan_weighted(phi, pps, nbin, idat_start, ndat, weights, ndat_per_weight, weight_idat):
    """Fold.C:687-788 with a weighted input: -> (binplan with nbin for skipped samples, hits, ndat_folded)."""
    binplan = np.zeros(ndat, np.uint32)
    hits = np.zeros(nbin, np.uint32)
    w = np.ascontiguousarray(weights, np.uint32)
    f = lib().orc_fold_plan_weighted
    f.restype = C.c_uint64
    n = f(C.c_double(phi), C.c_double(pps), C.c_uint(nbin), C.c_uint64(idat_start), C.c_uint64(ndat), _p(w),
          C.c_uint64(w.size), C.c_uint(ndat_per_weight), C.c_uint64(weight_idat), _p(binplan), _p(hits), None)
    if n == 2 ** 64 - 1:
        raise RuntimeError("fold: iweight >= nweights (Fold.C:699)")
    return binplan, hits, n


# ----------------------------------------------------------------------------- a15
class TimeDivide:
    """dsp::TimeDivide for divisions in seconds, restated line by line from Signal/Pulsar/TimeDivide.C
    (set_bounds :132-330, set_boundaries :349-425) with times in seconds since the observation start."""

    def __init__(self, division_seconds):
        self.division_seconds = division_seconds
        self.lower = self.upper = self.current_end = 0.0
        self.is_valid = False
        self.division = 0

    def set_bounds(self, input_start, rate, input_ndat):
        input_end = input_start + input_ndat / rate
        divide_start = input_start
        if self.is_valid:
            divide_start = max(self.current_end, input_start)
        new_division = end_reached = in_next = False
        if input_end < self.lower or divide_start + 0.5 / rate > self.upper:
            new_division = True
            seconds = max(0.0, divide_start + 0.55 / rate)
            self.division = int(seconds / self.division_seconds)
            self.lower = float(self.division) * self.division_seconds
            self.upper = float(self.division + 1) * self.division_seconds
        divide_start = max(self.lower, divide_start)
        idat_start = int(max(0.0, np.rint((divide_start - input_start) * rate)))
        if idat_start >= input_ndat:
            self.is_valid = False
            return dict(is_valid=False, new_division=new_division, end_reached=False, in_next=False,
                        idat_start=idat_start, ndat=0, division=self.division)
        divide_end = min(input_end, self.upper)
        idat_end = int(np.rint((divide_end - input_start) * rate))
        assert idat_end > idat_start
        if idat_end > input_ndat:
            idat_end = input_ndat
        elif idat_end < input_ndat:
            in_next = True
        if (self.upper - divide_end) * rate < 0.5:
            end_reached = True
        self.is_valid = True
        self.current_end = input_start + idat_end / rate
        return dict(is_valid=True, new_division=new_division, end_reached=end_reached, in_next=in_next,
                    idat_start=idat_start, ndat=idat_end - idat_start, division=self.division)


# ----------------------------------------------------------------------------- f1 (digifil tail)
class Rescale:
    """dsp::Rescale (Signal/General/Rescale.C:165-385, compute_various :387-412), FPT order, default mode
    (not exact, no decay): statistics accumulate over `interval_samples` (0: the first block's length);
    on the first call and at every interval end offset = -mean, scale = 1/sqrt(variance)."""

    def __init__(self, interval_samples=0, constant=False):
        self.interval_samples = interval_samples
        self.constant = constant
        self.nsample = 0
        self.isample = 0
        self.offset = self.scale = None

    def transform(self, x):
        """x: [nchan, npol, ndat] float32 -> same shape."""
        x = np.ascontiguousarray(x, np.float32)
        nchan, npol, ndat = x.shape
        first_call = self.nsample == 0
        if first_call:
            self.nsample = self.interval_samples if self.interval_samples else ndat
            self.tot = np.zeros((nchan, npol), np.float64)
            self.totsq = np.zeros((nchan, npol), np.float64)
            self.offset = np.zeros((nchan, npol), np.float32)
            self.scale = np.ones((nchan, npol), np.float32)
        out = np.empty_like(x)
        start = 0
        while True:
            end = min(ndat, start + self.nsample - self.isample)
            seg = x[:, :, start:end]
            # sequential double accumulation of float samples and float squares (Rescale.C:258-262)
            self.tot += np.cumsum(seg.astype(np.float64), axis=2)[:, :, -1] if end > start else 0.0
            self.totsq += np.cumsum((seg * seg).astype(np.float64), axis=2)[:, :, -1] if end > start else 0.0
            self.isample += end - start
            if self.isample == self.nsample or first_call:
                mean = self.tot / self.isample
                var = self.totsq / self.isample - mean * mean
                if not self.constant or first_call:
                    self.offset = (-mean).astype(np.float32)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        self.scale = np.where(var == 0.0, 1.0, 1.0 / np.sqrt(var)).astype(np.float32)
                self.isample = 0
                first_call = False
                self.tot[:] = 0
                self.totsq[:] = 0
            out[:, :, start:end] = (seg + self.offset[:, :, None]) * self.scale[:, :, None]
            start = end
            if end >= ndat:
                break
        return out


def sigproc_channel_sort(nchan, bandwidth, swap=False, nsub_swap=0):
    """ChannelSort (Kernel/Formats/sigproc/SigProcDigitizer.C:38-70): output channel -> input channel."""
    m = np.arange(nchan)
    if nsub_swap > 1:
        if swap:
            m = (m + nchan // 2) % nchan
        sub = nchan // nsub_swap
        m = (m // sub) * sub + ((m % sub) + sub // 2) % sub
    elif swap:
        m = (m + nchan // 2) % nchan
    if bandwidth > 0:
        m = nchan - m - 1
    return m


def sigproc_digitize(x, nbit=8, input_scale=1.0, scale_fac=1.0, rescale=True, bandwidth=-1.0, swap=False):
    """SigProcDigitizer::pack, FPT input, nbit 8 (SigProcDigitizer.C:80-160,244-300): TPF bytes
    [ndat][npol][nchan] = clip(int(x*digi_scale + mean + 0.5), 0, 255)."""
    assert nbit == 8
    x = np.ascontiguousarray(x, np.float32)
    nchan, npol, ndat = x.shape
    digi_mean, digi_sigma = np.float32(127.5), np.float32(6)
    digi_scale = np.float32(digi_mean / digi_sigma)
    xpol_offset = np.float32(0)
    if not rescale:
        xpol_offset, digi_mean, digi_scale = digi_mean, np.float32(0), np.float32(1)
    digi_scale = np.float32(np.float64(digi_scale) / (np.float64(input_scale) * np.float64(scale_fac)))
    chan = sigproc_channel_sort(nchan, bandwidth, swap)
    out = np.zeros((ndat, npol, nchan), np.uint8)
    for ipol in range(npol):
        mean = np.float32(digi_mean + (xpol_offset if ipol > 1 else 0))
        v = (x[chan, ipol, :] * digi_scale + mean).astype(np.float64) + 0.5  # float product and sum, + 0.5 in double
        r = np.clip(np.trunc(v), 0, 255).astype(np.uint8)
        out[:, ipol, :] = r.T
    return out
