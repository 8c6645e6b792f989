"""Host-side PhaseSeries rules: the library (dspsr_b200/host/phaseseries.cpp) against the numpy
restatement of PhaseSeries.C / Observation.C / Archiver.C in oracle/phaseseries.py.  No GPU needed."""
from fractions import Fraction

import numpy as np
import pytest

from dspsr_b200 import _lib as L
from dspsr_b200 import phaseseries as P

START = (55299, 7545, 0.0)


def _attrs(o):
    return dict(telescope=o.telescope.decode(), receiver=o.receiver.decode(), source=o.source.decode(),
                machine=o.machine.decode(), format=o.format.decode(), mode=o.mode.decode(),
                centre_frequency=o.centre_frequency, bandwidth=o.bandwidth, nchan=o.nchan, npol=o.npol, ndim=o.ndim,
                nbit=o.nbit, type=o.type, state=o.state, basis=o.basis, rate=o.rate, swap=o.swap, nsub_swap=o.nsub_swap,
                dc_centred=o.dc_centred, scale=o.scale, dm=o.dispersion_measure, rm=o.rotation_measure)


def _secs(mjd):
    return Fraction(mjd[0] - START[0]) * 86400 + Fraction(mjd[1] - START[1]) + Fraction(mjd[2])


def _obs(**kw):
    base = dict(nchan=4, npol=1, ndim=4, rate=195312.5, start_mjd=START, ndat=100000, centre_frequency=1382.0,
                bandwidth=-400.0, scale=2097152.0 * 8192.0, dm=67.99)
    base.update(kw)
    return P.observation(**base)


def test_combinable_follows_observation_rules():
    from oracle import phaseseries as OP
    a = _obs()
    cases = [dict(), dict(centre_frequency=1382.0 + 5e-7), dict(centre_frequency=1383.0), dict(bandwidth=-399.0),
             dict(nchan=8), dict(rate=195312.0), dict(scale=2097152.0 * 8192.0 * (1 + 5e-7)),
             dict(scale=2097152.0 * 8192.0 * 1.001), dict(dm=68.0), dict(source="J0437-4715"), dict(mode="2-bit,foo"),
             dict(state=L.STOKES), dict(swap=1), dict(machine="UWB")]
    seen = set()
    for kw in cases:
        b = _obs(**kw)
        got, why = P.combinable(a, b)
        want = OP.combinable(_attrs(a), _attrs(b))
        assert got == want, (kw, why)
        assert got or why.startswith("\n\tdifferent ")
        seen.add(got)
    assert seen == {True, False}
    # two 2-bit modes may differ in their tail (Observation.C:267-279)
    x, y = _obs(mode="2-bit,mean=1"), _obs(mode="2-bit,mean=2")
    assert P.combinable(x, y)[0] and OP.combinable(_attrs(x), _attrs(y))


def test_mixable_fold_bookkeeping_and_combine_match_restatement():
    """Two 'threads' fold disjoint blocks of one observation (dspsr -t 2 / two GPUs), then PhaseSeries::combine:
    integration_length, ndat_total, start/end bounds, data and hits."""
    from oracle import phaseseries as OP
    rng = np.random.default_rng(5)
    nchan, npol, ndim, nbin = 4, 1, 4, 64
    obs = _obs()
    blocks = [(0, 30000), (30000, 25000), (55000, 45000)]
    owner = [0, 1, 0]
    mine = [P.PhaseSeries(nchan, npol, ndim, nbin) for _ in range(2)]
    ref = [OP.PS(nchan, npol, ndim, nbin) for _ in range(2)]
    for (istart, n), t in zip(blocks, owner):
        assert mine[t].mixable(obs, nbin, istart, n)
        assert ref[t].mixable(_attrs(obs), _secs(START), obs.ndat, nbin, istart, n)
        d = rng.standard_normal(mine[t].data.shape).astype(np.float32)
        h = rng.integers(0, 50, nbin).astype(np.uint32)
        mine[t].data += d
        mine[t].hits += h
        ref[t].data += d
        ref[t].hits += h
        mine[t].folded(n - 7, n)
        ref[t].folded(n - 7, n)
    for t in range(2):
        assert mine[t].integration_length == ref[t].integration_length and mine[t].ndat_total == ref[t].ndat_total
        assert _secs(mine[t].start_time) == pytest.approx(float(ref[t].start), abs=1e-9)
        assert _secs(mine[t].end_time) == pytest.approx(float(ref[t].end), abs=1e-9)
    # a different nbin or an uncombinable observation is refused once data are in
    assert not mine[0].mixable(obs, nbin * 2, 0, 10)
    assert not mine[0].mixable(_obs(dm=10.0), nbin, 0, 10)
    total, rtotal = P.PhaseSeries(nchan, npol, ndim, nbin), OP.PS(nchan, npol, ndim, nbin)
    for t in range(2):                       # first combine copies ("this is empty"), second adds
        total.combine(mine[t])
        rtotal.combine(ref[t])
    assert np.array_equal(total.data, rtotal.data) and np.array_equal(total.hits, rtotal.hits)
    assert total.integration_length == rtotal.integration_length == pytest.approx((100000 - 21) / obs.rate)
    assert total.ndat_total == rtotal.ndat_total == 100000
    assert float(_secs(total.start_time)) == pytest.approx(0.0, abs=1e-9)
    assert float(_secs(total.end_time)) == pytest.approx(100000 / obs.rate, abs=1e-9)
    bad = P.PhaseSeries(nchan, npol, ndim, nbin)
    assert bad.mixable(_obs(source="other"), nbin, 0, 10)
    bad.folded(10, 10)
    with pytest.raises(L.B200Error):
        total.combine(bad)


def test_normalise_and_dump_round_trip(tmp_path):
    """Archiver::set: amplitude / (scale * hits); bins without hits take the mean of the hit bins; a non-finite
    amplitude zeroes the profile and its weight.  The dump holds exactly what normalise returns."""
    from oracle import phaseseries as OP
    rng = np.random.default_rng(6)
    nchan, npol, ndim, nbin = 3, 2, 2, 32
    ps = P.PhaseSeries(nchan, npol, ndim, nbin, folding_period=0.0893)
    obs = _obs(nchan=nchan, npol=npol, ndim=ndim, scale=1234.5)
    assert ps.mixable(obs, nbin, 0, 5000)
    ps.data += rng.standard_normal(ps.data.shape).astype(np.float32) * 1e4
    ps.hits += rng.integers(1, 200, nbin).astype(np.uint32)
    ps.hits[[3, 17]] = 0
    ps.data[1, 0, 5 * ndim + 1] = np.inf
    ps.folded(5000, 5000)
    prof, w, bad = ps.normalise()
    rprof, rw = OP.normalise(ps.data, ps.hits, 1234.5, ndim)
    assert bad == 1 and w[1, 0, 1] == 0 and np.all(prof[1, 0, 1] == 0) and w.sum() == w.size - 1
    assert np.array_equal(w, rw)
    hit = ps.hits != 0
    assert np.array_equal(prof[..., hit], rprof[..., hit])
    assert np.allclose(prof[..., ~hit], rprof[..., ~hit], rtol=1e-6, atol=0)
    assert prof[0, 0, 0, 3] == pytest.approx(prof[0, 0, 0, hit].astype(np.float64).mean(), rel=1e-6)
    path = tmp_path / "subint.b200ps"
    ps.unload(path)
    hdr, fprof, fw, fhits, fraw = P.load(path)
    assert (hdr.nbin, hdr.obs.nchan, hdr.obs.npol, hdr.obs.ndim) == (nbin, nchan, npol, ndim)
    assert hdr.integration_length == ps.integration_length and hdr.ndat_total == 5000
    assert hdr.obs.scale == 1234.5 and hdr.folding_period == 0.0893 and hdr.obs.source == b"J0835-4510"
    assert (hdr.obs.start_time.day, hdr.obs.start_time.sec) == START[:2]
    assert np.array_equal(fprof, prof) and np.array_equal(fw, w) and np.array_equal(fhits[0], ps.hits)
    assert np.array_equal(fraw, np.nan_to_num(ps.data, posinf=np.inf)) or np.array_equal(fraw[np.isfinite(fraw)], ps.data[np.isfinite(ps.data)])
    raw = open(path, "rb").read(4096)
    assert raw.startswith(b"HDR_VERSION") and b"FILE_TYPE            B200_PHASESERIES" in raw and raw[-1] == 0
