"""Host-side PhaseSeries (b200_phase_series_*): the folded accumulator's attributes, PhaseSeries::mixable /
combine, Archiver::set normalisation and the self-describing sub-integration dump.  Thin ctypes wrapper: the
arithmetic lives in libb200dsp.so (dspsr_b200/host/phaseseries.cpp)."""
import ctypes as C

import numpy as np

from . import _lib as L


def observation(nchan, npol, ndim, rate, start_mjd, ndat=0, centre_frequency=0.0, bandwidth=0.0, scale=1.0, dm=0.0,
                state=L.COHERENCE, nbit=8, source="J0835-4510", telescope="PKS", machine="CASPSR", mode="PSR",
                receiver="", format="", rm=0.0, swap=0, nsub_swap=0, dc_centred=0, basis=0, type=0):
    o = L.Observation()
    o.telescope, o.receiver, o.source = telescope.encode(), receiver.encode(), source.encode()
    o.mode, o.machine, o.format = mode.encode(), machine.encode(), format.encode()
    o.centre_frequency, o.bandwidth, o.rate, o.scale = centre_frequency, bandwidth, rate, scale
    o.dispersion_measure, o.rotation_measure = dm, rm
    o.nchan, o.npol, o.ndim, o.nbit = nchan, npol, ndim, nbit
    o.state, o.type, o.basis, o.swap, o.nsub_swap, o.dc_centred = state, type, basis, swap, nsub_swap, dc_centred
    o.start_time = L.Mjd(*start_mjd)
    o.ndat = ndat
    return o


def combinable(a, b):
    """Observation::combinable -> (bool, reason)."""
    buf = C.create_string_buffer(2048)
    ok = L.load().b200_observation_combinable(C.byref(a), C.byref(b), buf, 2048)
    return bool(ok), buf.value.decode()


class PhaseSeries:
    """A host PhaseSeries: numpy arrays + the POD the library's rules operate on."""

    def __init__(self, nchan, npol, ndim, nbin, folding_period=0.0, reference_phase=0.0):
        self.data = np.zeros((nchan, npol, nbin * ndim), np.float32)
        self.hits = np.zeros(nbin, np.uint32)
        self.ps = L.PhaseSeries()
        self.ps.obs.nchan, self.ps.obs.npol, self.ps.obs.ndim = nchan, npol, ndim
        self.ps.nbin, self.ps.hits_nchan = nbin, 1
        self.ps.folding_period, self.ps.reference_phase = folding_period, reference_phase
        self._bind()

    def _bind(self):
        self.ps.data = self.data.ctypes.data
        self.ps.hits = self.hits.ctypes.data

    def mixable(self, obs, nbin, istart=0, fold_ndat=0):
        return bool(L.load().b200_phase_series_mixable(C.byref(self.ps), C.byref(obs), nbin, istart, fold_ndat))

    def folded(self, ndat_folded, ndat_fold):
        L.check(L.load().b200_phase_series_folded(C.byref(self.ps), ndat_folded, ndat_fold))

    def combine(self, other):
        L.check(L.load().b200_phase_series_combine(C.byref(self.ps), C.byref(other.ps)))

    def normalise(self):
        """Archiver::set -> (profiles [nchan, npol, ndim, nbin], weights [nchan, npol, ndim], corrupted)."""
        o = self.ps.obs
        prof = np.zeros((o.nchan, o.npol, o.ndim, self.ps.nbin), np.float32)
        w = np.zeros((o.nchan, o.npol, o.ndim), np.float32)
        bad = C.c_uint(0)
        L.check(L.load().b200_phase_series_normalise(C.byref(self.ps), prof.ctypes.data, w.ctypes.data, C.byref(bad)))
        return prof, w, bad.value

    def unload(self, path):
        L.check(L.load().b200_phase_series_unload(C.byref(self.ps), str(path).encode()))

    @property
    def integration_length(self):
        return self.ps.integration_length

    @property
    def ndat_total(self):
        return self.ps.ndat_total

    @property
    def start_time(self):
        t = self.ps.obs.start_time
        return (t.day, t.sec, t.frac)

    @property
    def end_time(self):
        t = self.ps.end_time
        return (t.day, t.sec, t.frac)


def load(path):
    """-> (L.PhaseSeries header, profiles, weights, hits, raw sums) of a dump written by PhaseSeries.unload."""
    lib = L.load()
    ps = L.PhaseSeries()
    L.check(lib.b200_phase_series_load(str(path).encode(), C.byref(ps), None, None, None, None))
    o = ps.obs
    prof = np.zeros((o.nchan, o.npol, o.ndim, ps.nbin), np.float32)
    w = np.zeros((o.nchan, o.npol, o.ndim), np.float32)
    hits = np.zeros((ps.hits_nchan, ps.nbin), np.uint32)
    raw = np.zeros((o.nchan, o.npol, ps.nbin * o.ndim), np.float32)
    L.check(lib.b200_phase_series_load(str(path).encode(), C.byref(ps), prof.ctypes.data, w.ctypes.data,
                                       hits.ctypes.data, raw.ctypes.data))
    return ps, prof, w, hits, raw
