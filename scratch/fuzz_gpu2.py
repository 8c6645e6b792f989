"""Randomised parity sweep, part 2: convolution pipelines (generic 8-bit complex / real, multi-channel) and the
stand-alone fold engine with arbitrary periods.   python scratch/fuzz_gpu2.py [ncases] [seed]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import oracle as O, synth
from dspsr_b200 import _lib as L, engine as E

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
ctx = E.Context(0)
lut, _ = O.bittable8()
bad = 0


def sparse_err(a, b):
    nz = b != 0
    rms = np.sqrt(np.mean(b[nz].astype(np.float64) ** 2)) if nz.any() else 1.0
    return float(np.max(np.abs(a.astype(np.float64) - b)) / rms)


for case in range(ncases):
    kind = int(rng.integers(0, 3))
    if kind < 2:
        # convolution pipeline: generic 8-bit, ndim 2 (complex) or 1 (real), nchan channels
        ndim = 2 if kind == 0 else 1
        nchan = int(rng.integers(1, 5))
        F = 1 << int(rng.integers(4, 17))
        nf = int(rng.integers(1, max(2, F // 4)))
        npos = int(rng.integers(0, nf + 1)); nneg = nf - npos
        npart = int(rng.integers(1, 4)); nblock = int(rng.integers(1, 3))
        state, dndim = [("Coherence", 4), ("Stokes", 2), ("PPQQ", 1), ("Intensity", 1)][int(rng.integers(0, 4))]
        nbin = int(2 ** rng.integers(2, 11))
        c = O.conv_sizes(ndim == 1, nchan, 2, F, npos, nneg)
        if ndim == 1 and c.nsamp_step % 2: continue
        ndat = nblock * npart * c.nsamp_step + c.nsamp_overlap
        raw = rng.integers(0, 256, size=ndat * nchan * 2 * ndim, dtype=np.uint8)
        H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
        nkeep = F - npos - nneg if ndim == 2 else (F - npos - nneg)
        pps = 1.0 / (rng.uniform(0.02, 3.0) * c.nsamp_step * npart + 2.5)
        phis = [float(rng.uniform(0, 1)) for _ in range(nblock)]
        op = O.make_pipe(L.FMT_GENERIC8, nchan, 2, ndim, lut, 0.0, None, c, H, state, dndim, nbin)
        ref, ref_hits = O.pipe_run(op, raw, nblock, npart, phis, [pps] * nblock, nthread=1)
        ud = E.make_unpack_desc(L.FMT_GENERIC8, nchan, 2, ndim, lut)
        fd, keep = E.make_fb_desc(ndim == 1, nchan, 2, 1, F, npos, nneg, H, int(rng.integers(0, 3)))
        pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, nbin)
        d_raw = torch.from_numpy(raw).cuda()
        step = pipe.info.nsamp_step
        for b in range(nblock):
            pipe.execute(d_raw, npart, phis[b], pps, first_sample=b * npart * step)
        prof, hits, ntot = pipe.synch()
        err = sparse_err(prof, ref)
        ok = np.array_equal(hits, ref_hits) and err <= 2e-5
        desc = dict(kind="conv", ndim=ndim, nchan=nchan, F=F, npos=npos, nneg=nneg, npart=npart, nblock=nblock, state=state, dndim=dndim, nbin=nbin)
        del pipe
    else:
        # stand-alone fold engine
        nchan = int(rng.integers(1, 6)); ndim = int([1, 2, 4][rng.integers(0, 3)]); npol = int([1, 2, 4][rng.integers(0, 3)])
        nbin = int(2 ** rng.integers(1, 11)); ndat = int(rng.integers(100, 60000))
        x = (rng.standard_normal((nchan, npol, ndat * ndim)) + 1.0).astype(np.float32)
        phi = float(rng.uniform(0, 1)); pps = 1.0 / rng.uniform(1.5, 5000.0)
        fe = E.FoldEngine(ctx, nchan, npol, ndim, nbin)
        fe.set_bins(phi, pps, ndat, 0)
        fe.fold(torch.from_numpy(x).cuda())
        bp, hh, _, _ = O.fold_plan(phi, pps, nbin, ndat)
        ref = O.fold(x, ndim, bp, nbin)
        out = fe.synch(); hits, ntot = fe.hits()
        err = sparse_err(out, ref)
        ok = np.array_equal(hits, hh) and ntot == ndat and err <= 2e-5
        desc = dict(kind="fold", nchan=nchan, ndim=ndim, npol=npol, nbin=nbin, ndat=ndat, period=1 / pps)
    if not ok:
        bad += 1
        print("FAIL", desc, "err %.2e" % err)
print("fuzz2 done: %d cases, %d failures" % (ncases, bad))
