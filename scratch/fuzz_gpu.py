"""Randomised GPU-vs-oracle parity sweep over pipeline shapes (run by hand on the GPU box):
   python scratch/fuzz_gpu.py [ncases] [seed]"""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import oracle as O, synth
from dspsr_b200 import _lib as L, engine as E

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = E.Context(0)
lut, _ = O.bittable8()
bad = 0
for case in range(ncases):
    lgF = int(rng.integers(1, 14))                       # freq_res 2 .. 8192
    lgC = int(rng.integers(0, max(1, min(9, 22 - lgF - 1))))
    F, C = 1 << lgF, 1 << lgC
    if C * F < 16 or C * F > (1 << 21):
        continue
    nf = int(rng.integers(0, max(1, F // 3)))
    npos = int(rng.integers(0, nf + 1)); nneg = nf - npos
    if F == 1: npos = nneg = 0
    if npos + nneg >= F: continue
    npart = int(rng.integers(1, 6))
    nblock = int(rng.integers(1, 3))
    state, dndim = [("Coherence", 4), ("Coherence", 2), ("Coherence", 1), ("Stokes", 4), ("Stokes", 2), ("PPQQ", 1), ("Intensity", 1)][int(rng.integers(0, 7))]
    nbin = int(2 ** rng.integers(3, 11))
    max_npart = int(rng.integers(0, 4))
    f = O.fb_sizes(1, 1, 2, C, F, npos, nneg)
    if f.nsamp_step % 4: continue
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=1000 + case)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C, F))).astype(np.complex64)
    pps = 1.0 / (rng.uniform(0.05, 3.0) * f.nkeep * npart + 3.0)
    phis = [float(rng.uniform(0, 1)) for _ in range(nblock)]
    op = O.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, state, dndim, nbin)
    ref, ref_hits = O.pipe_run(op, raw, nblock, npart, phis, [pps] * nblock, nthread=1)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H, max_npart)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, nbin)
    d_raw = torch.from_numpy(raw).cuda()
    for b in range(nblock):
        pipe.execute(d_raw, npart, phis[b], pps, first_sample=b * npart * f.nsamp_step)
    prof, hits, ntot = pipe.synch()
    nz = ref != 0
    rms = np.sqrt(np.mean(ref[nz].astype(np.float64) ** 2)) if nz.any() else 1.0   # RMS over filled bins
    err = float(np.max(np.abs(prof.astype(np.float64) - ref)) / rms)
    ok = np.array_equal(hits, ref_hits) and err <= 1e-5
    if not ok:
        bad += 1
        print("FAIL", dict(C=C, F=F, npos=npos, nneg=nneg, npart=npart, nblock=nblock, state=state, dndim=dndim, nbin=nbin, max_npart=max_npart), "err %.2e hits %s" % (err, np.array_equal(hits, ref_hits)))
    del pipe
print("fuzz done: %d cases, %d failures" % (ncases, bad))
