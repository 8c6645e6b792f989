"""Host-side (no GPU) helpers bound from libb200dsp.so: BitTable, Dedispersion, polyco predictor.
Product code -- independent of oracle/ (which restates the same reference functions for checking)."""
import ctypes as C
import math

import numpy as np

from . import _lib as L


def bittable8(twos_complement=True):
    """dsp::BitTable(8): (256 float32 values, get_scale())  [BitTable.C:121-218]."""
    lut = np.zeros(256, np.float32)
    scale = C.c_double(0)
    L.check(L.load().b200_bittable8(int(twos_complement), lut.ctypes.data_as(C.c_void_p), C.byref(scale)))
    return lut, scale.value


def dedispersion(centre_frequency, bandwidth, dm, input_nchan, nchan, input_real, frequency_resolution=0,
                 dual_sideband=None, dc_centred=False, swap=False, build=True):
    """dsp::Dedispersion::prepare/build/match.  Returns (params, H[nchan, ndat] complex64 | None)."""
    d = L.Dedispersion()
    d.centre_frequency, d.bandwidth, d.dispersion_measure = centre_frequency, bandwidth, dm
    d.input_nchan, d.nchan = input_nchan, nchan
    d.input_dual_sideband = int((not input_real) if dual_sideband is None else dual_sideband)
    d.input_dc_centred, d.input_swap = int(dc_centred), int(swap)
    d.frequency_resolution = frequency_resolution
    L.check(L.load().b200_dedispersion_prepare(C.byref(d)))
    H = None
    if build:
        H = np.zeros((nchan, d.ndat), np.complex64)
        L.check(L.load().b200_dedispersion_build(C.byref(d), H.ctypes.data_as(C.c_void_p)))
    return d, H


def dedispersion_channels(d, first_chan, nchan_local):
    """Rows [first_chan, first_chan+nchan_local) of the matched response of a prepared `d` (channel shard)."""
    H = np.zeros((nchan_local, d.ndat), np.complex64)
    L.check(L.load().b200_dedispersion_build_channels(C.byref(d), first_chan, nchan_local, H.ctypes.data_as(C.c_void_p)))
    return H


class Polyco:
    """TEMPO polyco predictor (Pulsar::Predictor::phase / frequency as used by Fold.C:943-958)."""

    def __init__(self, text):
        self.pc = L.Polyco()
        L.check(L.load().b200_polyco_parse(text.encode(), C.byref(self.pc)))

    def phase(self, mjd):
        day, sec, frac = mjd
        return L.load().b200_polyco_phase(C.byref(self.pc), day, sec, frac, None)

    def frequency(self, mjd):
        day, sec, frac = mjd
        return L.load().b200_polyco_frequency(C.byref(self.pc), day, sec, frac)


def utc_to_mjd(utc):
    """'YYYY-MM-DD-hh:mm:ss' -> (day, sec, frac) split MJD (ASCIIObservation UTC_START)."""
    y, m, d, hms = utc.split("-")
    hh, mm, ss = hms.split(":")
    y, m, d = int(y), int(m), int(d)
    a = (14 - m) // 12
    yy = y + 4800 - a
    mo = m + 12 * a - 3
    jdn = d + (153 * mo + 2) // 5 + 365 * yy + yy // 4 - yy // 100 + yy // 400 - 32045
    return (jdn - 2400001, int(hh) * 3600 + int(mm) * 60 + int(ss), 0.0)


def mjd_add(mjd, seconds):
    day, sec, frac = mjd
    frac += seconds
    whole = math.floor(frac)
    frac -= whole
    sec += int(whole)
    day += sec // 86400
    sec %= 86400
    return (day, sec, frac)


def fold_phase(predictor, start_mjd, idat_start, rate, reference_phase=0.0):
    """dsp::Fold::fold's phase set-up (Fold.C:650-657,718-720): returns (phi, phase_per_sample)."""
    t0 = mjd_add(start_mjd, (idat_start + 0.5) / rate)
    phi = predictor.phase(t0) - reference_phase
    pfold = 1.0 / predictor.frequency(t0)
    return phi, (1.0 / rate) / pfold


def optimal_fft_length(nbadperfft, nfft_max=0):
    """optimal_fft_length (Signal/General/optimize_fft.c:63-127)."""
    f = L.load().b200_optimal_fft_length
    f.restype = C.c_int64
    f.argtypes = [C.c_uint64, C.c_uint64]
    return f(nbadperfft, nfft_max)
