#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/ from the CPU oracle.

The reference (demorest/dspsr) holds no golden vectors or known-answer tests for this path
 and cannot be built here, so the fixtures are produced by the restated
oracle on small seeded inputs and frozen: CPU tests check that the oracle still reproduces
them, GPU tests check the CUDA path against them.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
import synth  # noqa: E402


def filterbank_case():
    # miniature of BASELINE configs[0]: CASPSR 8-bit real dual-pol, -F 8:D, Coherence, fold 32 bins
    lut, scale = O.bittable8()
    C = 8
    d, H = O.dedispersion(1382.0, -400.0, 0.002, 1, C, True)
    f = O.fb_sizes(1, 1, 2, C, d.ndat, d.impulse_pos, d.impulse_neg)
    nblock, npart, nbin = 2, 3, 32
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=0xD5B5 + 1)
    x = O.unpack_caspsr(raw, ndat, lut)
    volt = O.filterbank(f, x, H)
    det = O.detect("Coherence", 4, volt)
    phi = [0.2, 0.7]
    pps = [1.0 / 41.3, 1.0 / 41.3000001]
    p = O.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    prof, hits = O.pipe_run(p, raw, nblock, npart, phi, pps, 1)
    np.savez_compressed(os.path.join(HERE, "filterbank_mini.npz"), raw=raw, lut=lut, scale=scale, H=H,
                        freq_res=d.ndat, nfilt_pos=d.impulse_pos, nfilt_neg=d.impulse_neg, nchan=C, nbin=nbin,
                        nblock=nblock, npart=npart, phi=phi, pps=pps, volt=volt, det=det, profile=prof, hits=hits)


def convolution_case():
    # miniature of BASELINE configs[2]: 4 complex input channels, 8-bit MeerKAT heaps, convolution
    _, scale = O.bittable8()
    scale = float(np.float32(scale))
    nchan, npol = 4, 2
    d, H = O.dedispersion(1284.0, 16.0, 0.5, nchan, nchan, False)
    c = O.conv_sizes(False, nchan, npol, d.ndat, d.impulse_pos, d.impulse_neg)
    npart = 2
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, npol, seed=0xD5B5 + 3)
    x = O.unpack_meerkat(raw, ndat, nchan, npol, scale)
    volt = O.convolution(c, x, H)
    det = O.detect("Stokes", 4, volt)
    binplan, hits, _, _ = O.fold_plan(0.4, 1.0 / 57.7, 16, volt.shape[2])
    prof = O.fold(det, 4, binplan, 16)
    np.savez_compressed(os.path.join(HERE, "convolution_mini.npz"), raw=raw, scale=scale, H=H, freq_res=d.ndat,
                        nfilt_pos=d.impulse_pos, nfilt_neg=d.impulse_neg, nchan=nchan, npart=npart, ndat=ndat,
                        volt=volt, profile=prof, hits=hits, phi=0.4, pps=1.0 / 57.7, nbin=16)


def plan_case():
    # fold bin plans: (phi, pps, nbin, ndat) -> CRC-like digest of the bins + hits
    cases = [(0.3, 7.161113589011971e-06, 1024, 200000), (0.999999, 0.013, 1000, 5000), (0.0, 0.5, 64, 1000),
             (-0.3, 3.3e-4, 128, 50000), (0.5, 1.0 / 3.0, 7, 999)]
    out = {}
    for i, (phi, pps, nbin, ndat) in enumerate(cases):
        bins, hits, _, phi_end = O.fold_plan(phi, pps, nbin, ndat)
        out["case%d" % i] = np.array([phi, pps, nbin, ndat, phi_end])
        out["hits%d" % i] = hits
        out["digest%d" % i] = np.array([np.bitwise_xor.reduce(bins * np.arange(1, ndat + 1, dtype=np.uint32)),
                                        int(bins.astype(np.uint64).sum())], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "fold_plan.npz"), **out)


if __name__ == "__main__":
    filterbank_case()
    convolution_case()
    plan_case()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
