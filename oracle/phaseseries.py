"""ORACLE (test infrastructure only): numpy restatement of the PhaseSeries rules of the reference.

  mixable / combine      Signal/Pulsar/PhaseSeries.C:336-418,442-480
  combinable             Kernel/Classes/Observation.C:139-310
  normalise              dsp::Archiver::set, Signal/Pulsar/Archiver.C:773-895
  fold bookkeeping       Signal/Pulsar/Fold.C:789-802

Written against the reference text, independently of dspsr_b200/host/phaseseries.cpp (plain Python dicts and
Fractions for the times instead of split MJDs).  Parity unpinned: PhaseSeries.C / Archiver.C need PSRCHIVE
(Pulsar::Profile, MJD, Reference::Able) and cannot be compiled here; the tests compare the product with this
restatement on seeded cases.  Only tests/ may import this module."""
from fractions import Fraction

import numpy as np

EPS = 0.000001
STRINGS = ("telescope", "receiver", "source", "machine", "format")
NUMS_EXACT = ("nchan", "npol", "ndim", "nbit", "type", "state", "basis", "rate", "swap", "nsub_swap", "dc_centred")


def combinable(a, b):
    """dicts of Observation attributes -> bool (Observation.C:139-310)."""
    ok = all(a[k] == b[k] for k in STRINGS)
    if abs(a["centre_frequency"] - b["centre_frequency"]) > EPS:
        ok = False
    elif abs(a["bandwidth"] - b["bandwidth"]) > EPS:
        ok = False
    ok = ok and all(a[k] == b[k] for k in NUMS_EXACT)
    if abs(a["scale"] - b["scale"]) > EPS * abs(a["scale"]):
        ok = False
    if a["mode"] != b["mode"] and not (a["mode"][:5] == b["mode"][:5] == "2-bit"):
        ok = False
    if abs(a["dm"] - b["dm"]) > EPS or abs(a["rm"] - b["rm"]) > EPS:
        ok = False
    return ok


class PS:
    """PhaseSeries state: attrs (dict), start/end (Fraction seconds since an arbitrary epoch), arrays."""

    def __init__(self, nchan, npol, ndim, nbin):
        self.attrs = None
        self.nbin = nbin
        self.data = np.zeros((nchan, npol, nbin * ndim), np.float32)
        self.hits = np.zeros(nbin, np.uint32)
        self.integration_length = 0.0
        self.ndat_total = 0
        self.start = self.end = None

    def mixable(self, attrs, obs_start, obs_ndat, nbin, istart=0, fold_ndat=0):
        rate = attrs["rate"]
        s = obs_start + Fraction(istart) / Fraction(rate)
        e = (obs_start + Fraction(obs_ndat) / Fraction(rate)) if fold_ndat == 0 else s + Fraction(fold_ndat) / Fraction(rate)
        if self.integration_length == 0.0:
            keep = self.ndat_total
            self.attrs = dict(attrs)
            self.start, self.end = s, e
            self.nbin = nbin
            self.data[:] = 0
            self.hits[:] = 0
            self.ndat_total = keep
            return True
        if not combinable(self.attrs, attrs) or self.nbin != nbin:
            return False
        self.end = max(self.end, e)
        self.start = min(self.start, s)
        return True

    def folded(self, ndat_folded, ndat_fold):
        self.integration_length += float(ndat_folded) / self.attrs["rate"]
        self.ndat_total += ndat_fold

    def combine(self, other):
        if other is None or other.nbin == 0:
            return
        if not self.integration_length:
            self.attrs = dict(other.attrs)
            self.start, self.end, self.nbin = other.start, other.end, other.nbin
            self.data[:] = other.data
            self.hits[:] = other.hits
            self.integration_length, self.ndat_total = other.integration_length, other.ndat_total
            return
        if not combinable(self.attrs, other.attrs) or self.nbin != other.nbin:
            raise ValueError("PhaseSeries !mixable")
        self.end = max(self.end, other.end)
        self.start = min(self.start, other.start)
        self.data += other.data
        self.hits += other.hits
        self.integration_length += other.integration_length
        self.ndat_total += other.ndat_total


def normalise(data, hits, scale, ndim):
    """Archiver::set for every (chan, pol, dim): data [nchan, npol, nbin*ndim] -> (profiles [nchan, npol, ndim, nbin],
    weights [nchan, npol, ndim])."""
    nchan, npol, _ = data.shape
    nbin = hits.size
    out = np.zeros((nchan, npol, ndim, nbin), np.float32)
    w = np.ones((nchan, npol, ndim), np.float32)
    for c in range(nchan):
        for p in range(npol):
            for d in range(ndim):
                frm = data[c, p, d::ndim]
                into = np.zeros(nbin, np.float32)
                hit = hits != 0
                finite = np.isfinite(frm)
                good = hit & finite
                into[good] = (frm[good].astype(np.float64) / (scale * hits[good].astype(np.float64))).astype(np.float32)
                if np.any(hit & ~finite):
                    into[:] = 0
                    w[c, p, d] = 0
                if np.any(~hit):
                    cnt = int(hit.sum()) or 1
                    mean = float(np.sum(into[hit].astype(np.float64))) / cnt
                    into[~hit] = mean
                out[c, p, d] = into
    return out, w
