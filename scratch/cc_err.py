"""prints the parity error of the 65536-point convolution + fold pipeline for a few shapes (B200_LIB selects the library)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import test_gpu_parity as T
import synth, oracle
from dspsr_b200 import engine as E, _lib as L
ctx = E.Context(0)
F, npos, nneg = 65536, 2536, 2543
for (nchan, npart, state, dndim, nbin) in [(5, 7, "Stokes", 4, 37), (5, 7, "Coherence", 4, 37), (5, 7, "Stokes", 4, 1024), (2, 2, "Stokes", 4, 37)]:
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    rng = np.random.default_rng(77)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    nsamp = npart * c.nsamp_step + c.nsamp_overlap
    ndat = (nsamp + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=78)
    _, scale = oracle.bittable8()
    err = T._pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart, state, dndim, nbin, scale=np.float32(scale))
    print(os.environ.get("B200_LIB", "default")[-20:], nchan, npart, state, nbin, "err %.3g" % err)
