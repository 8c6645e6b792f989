"""GPU parity tests: every C-ABI engine against the CPU oracle on the same seeded inputs.
Tolerances follow BASELINE.json north_star: unpack bit-exact; voltages and folded profiles
<= 1e-5 relative (normalised to RMS); hits / ndat_total exact."""
import ctypes as C

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _torch():
    import torch
    return torch


def _E():
    from dspsr_b200 import engine as E
    return E


def _L():
    from dspsr_b200 import _lib as L
    return L


# ------------------------------------------------------------------------------------ unpack
def test_unpack_caspsr_bitexact(ctx, oracle):
    torch, E, L = _torch(), _E(), _L()
    ndat = 1 << 16
    raw = synth.caspsr_bytes(ndat)
    lut, _ = oracle.bittable8()
    ref = oracle.unpack_caspsr(raw, ndat, lut)
    d = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    out = E.unpack(ctx, d, torch.from_numpy(raw).cuda(), ndat).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("nchan,npol,ndim", [(1, 2, 1), (4, 2, 2), (3, 1, 2), (16, 2, 1)])
def test_unpack_generic8_bitexact(ctx, oracle, nchan, npol, ndim):
    torch, E, L = _torch(), _E(), _L()
    ndat = 5000
    raw = synth.generic8_bytes(ndat, nchan, npol, ndim)
    lut, _ = oracle.bittable8()
    ref = oracle.unpack_generic8(raw, ndat, nchan, npol, ndim, lut)
    d = E.make_unpack_desc(L.FMT_GENERIC8, nchan, npol, ndim, lut)
    out = E.unpack(ctx, d, torch.from_numpy(raw).cuda(), ndat).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("swap", [1, 2])
def test_unpack_meerkat_bitexact(ctx, oracle, swap):
    torch, E, L = _torch(), _E(), _L()
    ndat, nchan, npol = 1024, 8, 2
    raw = synth.meerkat_bytes(ndat, nchan, npol)
    _, scale = oracle.bittable8()
    scale = float(np.float32(scale))
    ref = oracle.unpack_meerkat(raw, ndat, nchan, npol, scale, swap)
    d = E.make_unpack_desc(L.FMT_MEERKAT8, nchan, npol, 2, None, scale, swap)
    out = E.unpack(ctx, d, torch.from_numpy(raw).cuda(), ndat).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("npol", [1, 2])
def test_unpack_uwb_bitexact(ctx, oracle, npol):
    torch, E, L = _torch(), _E(), _L()
    ndat = 8192
    raw = synth.uwb_bytes(ndat, npol)
    ref = oracle.unpack_uwb(raw.view(np.int16), ndat, npol)
    d = E.make_unpack_desc(L.FMT_UWB16, 1, npol, 2)
    out = E.unpack(ctx, d, torch.from_numpy(raw).cuda(), ndat).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


# ------------------------------------------------------------------------------------ filterbank
def _fb_case(ctx, oracle, input_real, input_nchan, npol, C, F, npos, nneg, npart, with_response=True, seed=1,
             max_npart=0):
    torch, E = _torch(), _E()
    rng = np.random.default_rng(seed)
    nchan = input_nchan * C
    f = oracle.fb_sizes(input_real, input_nchan, npol, nchan, F, npos, nneg)
    ndim = 1 if input_real else 2
    ndat = npart * f.nsamp_step + f.nsamp_overlap
    x = rng.standard_normal((input_nchan, npol, ndat * ndim)).astype(np.float32)
    H = None
    if with_response:
        ph = rng.uniform(-np.pi, np.pi, (nchan, F))
        H = np.exp(1j * ph).astype(np.complex64)
        H[0, 0] = 0
    ref = oracle.filterbank(f, x, H)
    eng = E.FilterbankEngine(ctx, input_real, input_nchan, npol, C, F, npos, nneg, H, max_npart)
    assert eng.info.nsamp_step == f.nsamp_step and eng.info.nkeep == f.nkeep
    out = eng.perform(torch.from_numpy(x).cuda()).cpu().numpy().view(np.complex64)
    assert out.shape == ref.shape
    return synth.relerr(out.view(np.float32), ref.view(np.float32)), eng


@pytest.mark.parametrize("case", [
    # (input_real, input_nchan, npol, C, F, nfilt_pos, nfilt_neg, npart)
    (1, 1, 2, 16, 64, 5, 6, 5),          # single-pass forward (Nc = 1024)
    (1, 1, 2, 64, 512, 30, 31, 3),       # two-pass forward (Nc = 32768)
    (1, 1, 1, 8, 16, 2, 1, 7),           # single pol
    (0, 1, 2, 32, 128, 9, 10, 4),        # complex input, single-pass
    (0, 3, 2, 8, 256, 11, 13, 3),        # complex multi-channel input
    (0, 1, 2, 128, 256, 20, 21, 3),      # complex two-pass (Nc = 32768)
    (1, 1, 2, 4096, 8, 1, 1, 6),         # cfg2 shape: 4096 x 8
    (1, 1, 2, 2, 8192, 457, 459, 2),     # F = 8192 inverse (cfg1's channel transform), Nc = 16384
    (1, 1, 2, 64, 4, 1, 0, 9),           # tiny freq_res
    (1, 1, 2, 64, 2, 0, 1, 9),
])
def test_filterbank_voltages(ctx, oracle, case):
    err, _ = _fb_case(ctx, oracle, *case)
    assert err <= TOL, err


def test_filterbank_freq_res_1_no_response(ctx, oracle):
    # freq_res == 1: spectrum copied straight to the output channels (Filterbank.C:621-631)
    err, _ = _fb_case(ctx, oracle, 1, 1, 2, 64, 1, 0, 0, 20, with_response=False)
    assert err <= TOL, err
    err, _ = _fb_case(ctx, oracle, 0, 2, 2, 32, 1, 0, 0, 20, with_response=False)
    assert err <= TOL, err


def test_filterbank_batching_independent(ctx, oracle):
    # more parts than the internal batch: results must not depend on the batch size
    e1, _ = _fb_case(ctx, oracle, 1, 1, 2, 16, 64, 5, 6, 11, max_npart=3)
    e2, _ = _fb_case(ctx, oracle, 1, 1, 2, 16, 64, 5, 6, 11, max_npart=16)
    assert e1 <= TOL and e2 <= TOL


def test_filterbank_cfg1_shape(ctx, oracle):
    # BASELINE configs[0]: 256 x 8192, M = 457 + 459, r2c 4,194,304 (3 parts)
    err, eng = _fb_case(ctx, oracle, 1, 1, 2, 256, 8192, 457, 459, 3)
    assert (eng.info.fft_rows, eng.info.fft_cols) == (2048, 1024)
    assert err <= TOL, err


# ------------------------------------------------------------------------------------ convolution
@pytest.mark.parametrize("case", [
    (1, 1, 2, 1, 1024, 40, 41, 4),       # real, in-shared-memory transform
    (0, 4, 2, 1, 4096, 100, 101, 3),     # complex multi-channel (cfg3-like, small)
    (0, 1, 2, 1, 16384, 500, 501, 3),    # two-pass inverse (convolution path)
    (1, 1, 2, 1, 32768, 900, 901, 2),    # real + two-pass inverse
    (0, 2, 2, 1, 65536, 2536, 2543, 2),  # cfg3's per-channel transform (one-kernel path, voltage sink)
    (0, 3, 2, 1, 32768, 1000, 1001, 9),  # one-kernel path, several tiles per group
    (0, 1, 2, 1, 131072, 4000, 4001, 2), # one-kernel path, 8192-point rows
    (0, 1, 2, 1, 262144, 8000, 8001, 2), # long-transform kernels (512 x 512), float input, voltage sink
    (0, 2, 2, 1, 1 << 20, 30000, 30001, 2),  # long-transform kernels (1024 x 1024), two channels
])
def test_convolution_voltages(ctx, oracle, case):
    torch, E = _torch(), _E()
    input_real, nchan, npol, _, F, npos, nneg, npart = case
    rng = np.random.default_rng(3)
    c = oracle.conv_sizes(input_real, nchan, npol, F, npos, nneg)
    ndim = 1 if input_real else 2
    ndat = npart * c.nsamp_step + c.nsamp_overlap
    x = rng.standard_normal((nchan, npol, ndat * ndim)).astype(np.float32)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    ref = oracle.convolution(c, x, H)
    eng = E.FilterbankEngine(ctx, input_real, nchan, npol, 1, F, npos, nneg, H)
    out = eng.perform(torch.from_numpy(x).cuda()).cpu().numpy().view(np.complex64)
    assert out.shape == ref.shape
    err = synth.relerr(out.view(np.float32), ref.view(np.float32))
    assert err <= TOL, err


# ------------------------------------------------------------------------------------ detection
@pytest.mark.parametrize("state,ndim", [("Intensity", 1), ("PPQQ", 1), ("Coherence", 1), ("Coherence", 2),
                                        ("Coherence", 4), ("Stokes", 1), ("Stokes", 2), ("Stokes", 4)])
def test_detection_bitexact(ctx, oracle, state, ndim):
    torch, E = _torch(), _E()
    rng = np.random.default_rng(5)
    nchan, ndat = 5, 3001
    v = (rng.standard_normal((nchan, 2, ndat)) + 1j * rng.standard_normal((nchan, 2, ndat))).astype(np.complex64)
    ref = oracle.detect(state, ndim, v)
    out = E.detect(ctx, state, ndim, torch.from_numpy(v.view(np.float32)).cuda()).cpu().numpy()
    assert out.shape == ref.shape
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_detection_inplace_ndim2(ctx, oracle):
    torch, E, L = _torch(), _E(), _L()
    rng = np.random.default_rng(6)
    nchan, ndat = 3, 1000
    v = (rng.standard_normal((nchan, 2, ndat)) + 1j * rng.standard_normal((nchan, 2, ndat))).astype(np.complex64)
    ref = oracle.detect("Coherence", 2, v)
    t = torch.from_numpy(v.view(np.float32).copy()).cuda()
    import ctypes as C
    L.check(ctx.lib.b200_detect(ctx.h, L.COHERENCE, 2, C.c_void_p(t.data_ptr()), 2 * ndat, nchan, 2, ndat,
                                C.c_void_p(t.data_ptr()), 2 * ndat))
    assert np.array_equal(t.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_detection_errors(ctx):
    torch, E, L = _torch(), _E(), _L()
    t = torch.zeros((1, 1, 64), device="cuda")
    with pytest.raises(L.B200Error):
        E.detect(ctx, "Coherence", 2, t)      # npol != 2 (Detection.C:476-489)


# ------------------------------------------------------------------------------------ fold
@pytest.mark.parametrize("ndim,npol", [(1, 1), (1, 4), (2, 2), (4, 1)])
def test_fold_engine(ctx, oracle, ndim, npol):
    torch, E = _torch(), _E()
    rng = np.random.default_rng(7)
    nchan, nbin, ndat = 6, 128, 40000
    x = rng.standard_normal((nchan, npol, ndat * ndim)).astype(np.float32) + 3.0
    phi, pps = 0.37, 1.0 / 777.7
    fe = E.FoldEngine(ctx, nchan, npol, ndim, nbin)
    n1 = fe.set_bins(phi, pps, ndat // 2, 0)
    h1 = fe.get_bin_hits()
    xt = torch.from_numpy(x).cuda()
    fe.fold(xt)
    bp1, hh1, _, phi_mid = oracle.fold_plan(phi, pps, nbin, ndat // 2)
    assert n1 == ndat // 2 and np.array_equal(h1, hh1)
    prof = oracle.fold(x, ndim, bp1, nbin)
    # second call with idat_start (Subint-style slice) and its own phase
    phi2 = 0.91
    fe.set_bins(phi2, pps, ndat - ndat // 2, ndat // 2)
    fe.fold(xt)
    bp2, hh2, _, _ = oracle.fold_plan(phi2, pps, nbin, ndat - ndat // 2)
    prof = oracle.fold(x, ndim, bp2, nbin, profile=prof, idat_start=ndat // 2)
    out = fe.synch()
    hits, ntot = fe.hits()
    assert np.array_equal(hits, hh1 + hh2) and ntot == ndat and hits.sum() == ndat
    assert synth.relerr(out, prof) <= TOL
    fe.zero()
    assert not fe.synch().any() and not fe.hits()[0].any()


# ------------------------------------------------------------------------------------ fused path
def _pipeline_case(ctx, oracle, C, F, npos, nneg, npart, state, dndim, nbin, nblock=2, pps=None, from_host=False):
    torch, E, L = _torch(), _E(), _L()
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C, F, npos, nneg)
    ndat = nblock * npart * f.nsamp_step + f.nsamp_overlap
    ndat = (ndat + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=11)
    rng = np.random.default_rng(12)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C, F))).astype(np.complex64)
    H[0, 0] = 0
    if pps is None:
        pps = 1.0 / (0.37 * f.nkeep * npart)
    phis = [0.123 + 0.31 * b for b in range(nblock)]
    ppss = [pps * (1 + 1e-6 * b) for b in range(nblock)]
    op = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, state, dndim, nbin)
    ref, ref_hits = oracle.pipe_run(op, raw, nblock, npart, phis, ppss, nthread=1)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, nbin)
    d_raw = torch.from_numpy(raw).cuda()
    for b in range(nblock):
        first = b * npart * f.nsamp_step
        if from_host:
            nbytes = (npart * f.nsamp_step + f.nsamp_overlap) * 2
            pipe.execute_host(raw[first * 2: first * 2 + nbytes], npart, phis[b], ppss[b], 0)
        else:
            pipe.execute(d_raw, npart, phis[b], ppss[b], first_sample=first)
    prof, hits, ntot = pipe.synch()
    assert np.array_equal(hits, ref_hits)
    assert ntot == nblock * npart * f.nkeep == hits.sum()
    return synth.relerr(prof, ref)


@pytest.mark.parametrize("state,dndim", [("Coherence", 4), ("Coherence", 2), ("Coherence", 1), ("Stokes", 4),
                                         ("PPQQ", 1), ("Intensity", 1)])
def test_pipeline_small(ctx, oracle, state, dndim):
    err = _pipeline_case(ctx, oracle, 16, 256, 20, 21, 5, state, dndim, 64)
    assert err <= TOL, err


def test_pipeline_many_channels_per_cta(ctx, oracle):
    err = _pipeline_case(ctx, oracle, 256, 16, 2, 3, 8, "Coherence", 4, 32)
    assert err <= TOL, err


def test_pipeline_from_host(ctx, oracle):
    err = _pipeline_case(ctx, oracle, 16, 256, 20, 21, 5, "Coherence", 4, 64, from_host=True)
    assert err <= TOL, err


def test_pipeline_cfg1(ctx, oracle):
    # BASELINE configs[0] at full shape: 256 x 8192, 1024 bins, Coherence, 2 blocks x 2 parts
    err = _pipeline_case(ctx, oracle, 256, 8192, 457, 459, 2, "Coherence", 4, 1024, nblock=2, pps=7.18e-6)
    assert err <= TOL, err


# ------------------------------------------------------------------------------------ C++ engine shims
def test_cpp_engine_shims_demo(ctx, oracle, tmp_path):
    """host/b200_demo drives B200::FilterbankEngine / DetectionEngine / FoldEngine through the
    stand-in dsp::Filterbank / Detection / Fold operators (the reference's engine interfaces)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "dspsr_b200", "host", "b200_demo")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    lut, _ = oracle.bittable8()
    C, nbin, npart = 16, 64, 4
    d, H = oracle.dedispersion(1382.0, -400.0, 0.05, 1, C, True)
    f = oracle.fb_sizes(1, 1, 2, C, d.ndat, d.impulse_pos, d.impulse_neg)
    ndat = npart * f.nsamp_step + f.nsamp_overlap
    ndat = (ndat + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=31)
    phi, pps = 0.31, 1.0 / (0.45 * npart * f.nkeep)
    (tmp_path / "raw.bin").write_bytes(raw.tobytes())
    (tmp_path / "resp.c64").write_bytes(H.tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(tmp_path / "raw.bin"), str(tmp_path / "resp.c64"), str(C), str(d.ndat),
                        str(d.impulse_pos), str(d.impulse_neg), str(nbin), repr(phi), repr(pps), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    buf = np.fromfile(out, dtype=np.uint8)
    nprof = C * 4 * nbin
    prof = buf[: nprof * 4].view(np.float32).reshape(C, 1, nbin * 4)
    hits = buf[nprof * 4:nprof * 4 + nbin * 4].view(np.uint32)
    tail = buf[nprof * 4 + nbin * 4:]
    integration_length, ndat_total = tail[:8].view(np.float64)[0], int(tail[8:16].view(np.uint64)[0])
    # the demo hands the operators the whole stream at once: npart' = (ndat - overlap) / step parts
    npart_all = (ndat - f.nsamp_overlap) // f.nsamp_step
    p = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    ref, ref_hits = oracle.pipe_run(p, raw, 1, npart_all, [phi], [pps], 1)
    assert np.array_equal(hits, ref_hits)
    assert synth.relerr(prof, ref) <= TOL
    # Fold.C:792-803 bookkeeping kept on the engine-owned PhaseSeries and carried over by synch()
    assert ndat_total == npart_all * f.nkeep == int(hits.sum())
    rate_out = 800e6 * d.ndat / f.nsamp_fft
    assert abs(integration_length - ndat_total / rate_out) <= 1e-12 * integration_length


def test_cpp_engine_shims_demo_meerkat_convolution(ctx, oracle, tmp_path):
    """BASELINE configs[2] wiring in C++: B200::MeerKATUnpackerEngine (Unpacker device hook) -> ConvolutionEngine ->
    DetectionEngine -> FoldEngine through the stand-in operators, whose Fold::get_output() is the ENGINE's PhaseSeries
    as in the reference (Fold.C:88-94).  The demo folds the block, resets (engine->zero()), then folds it twice: the
    result must be exactly two accumulations of the oracle's single fold."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "dspsr_b200", "host", "b200_demo")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    nchan, F, npos, nneg, npart, nbin = 6, 4096, 150, 170, 3, 128
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=77)
    rng = np.random.default_rng(78)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    H[:, 0] = 0
    phi, pps = 0.12, 1.0 / 977.3
    (tmp_path / "raw.bin").write_bytes(raw.tobytes())
    (tmp_path / "resp.c64").write_bytes(H.tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, "--meerkat", str(tmp_path / "raw.bin"), str(tmp_path / "resp.c64"), str(nchan), str(F),
                        str(npos), str(nneg), str(nbin), repr(phi), repr(pps), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    buf = np.fromfile(out, dtype=np.uint8)
    nprof = nchan * 4 * nbin
    prof = buf[: nprof * 4].view(np.float32).reshape(nchan, 1, nbin * 4)
    hits = buf[nprof * 4:nprof * 4 + nbin * 4].view(np.uint32)
    tail = buf[nprof * 4 + nbin * 4:]
    ndat_total = int(tail[8:16].view(np.uint64)[0])
    _, scale = oracle.bittable8()
    npart_all = (ndat - c.nsamp_overlap) // c.nsamp_step
    p = oracle.make_pipe(2, nchan, 2, 2, None, np.float32(scale), None, c, H, "Coherence", 4, nbin)
    ref, ref_hits = oracle.pipe_run(p, raw, 1, npart_all, [phi], [pps], 1)
    assert np.array_equal(hits, 2 * ref_hits)
    assert ndat_total == 2 * npart_all * (c.n_fft - c.nfilt_pos - c.nfilt_neg)
    assert synth.relerr(prof, 2.0 * ref) <= TOL


# ------------------------------------------------------------------------------------ two-bit excision (a6)
@pytest.mark.parametrize("npol", [1, 2])
def test_unpack_twobit_bitexact(ctx, oracle, npol):
    """TwoBitCorrection (CPSR2 convention) on the device: floats and weights identical to the CPU loops,
    including all-zero windows, windows outside the nlow limits and windows at other input powers."""
    torch, E = _torch(), _E()
    ndat = 512 * 40
    raw = synth.twobit_bytes(ndat, npol, seed=5).reshape(-1, npol).copy()
    loud = synth.twobit_bytes(ndat, npol, seed=6, sigma=1.6).reshape(-1, npol)
    quiet = synth.twobit_bytes(ndat, npol, seed=7, sigma=0.7).reshape(-1, npol)
    raw[128 * 10:128 * 14] = loud[128 * 10:128 * 14]        # nlow well below the mean: other table rows
    raw[128 * 20:128 * 24] = quiet[128 * 20:128 * 24]
    raw[128 * 3:128 * 4, 0] = 0x00                         # all-zero bytes
    raw[128 * 5:128 * 6, npol - 1] = 0xFF                  # nlow = 0
    raw[128 * 7:128 * 8, 0] = 0x99                         # nlow = 512
    raw = raw.reshape(-1)
    t = oracle.TwoBit()
    ref, wref = t.unpack(raw, ndat, npol)
    tb = E.make_twobit_desc(npol=npol)
    assert (tb.nlow_min, tb.nlow_max) == (t.nlow_min, t.nlow_max)
    out, w = E.unpack_twobit(ctx, tb, torch.from_numpy(raw).cuda(), ndat)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(w.cpu().numpy().astype(np.uint32), wref[0])
    assert wref[0].sum() < len(wref[0])                     # some windows really were excised
    # through the generic Unpacker hook as well
    ud = E.make_twobit_unpack_desc(tb)
    out2 = E.unpack(ctx, ud, torch.from_numpy(raw).cuda(), ndat)
    assert np.array_equal(out2.cpu().numpy(), ref)


def test_pipeline_cfg2_twobit_filterbank_intensity(ctx, oracle):
    """BASELINE configs[1] (digifil-style): 2-bit dual-pol real input, -F 4096:D coherent filterbank
    (freq_res 8, nfilt 1+1), Intensity detection, no fold -- detected float series vs the oracle chain."""
    torch, E, L = _torch(), _E(), _L()
    C, F, npos, nneg, npart = 4096, 8, 1, 1, 5
    d, H = oracle.dedispersion(1400.0, 128.0, 50.0, 1, C, True)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == (F, npos, nneg)
    f = oracle.fb_sizes(1, 1, 2, C, F, npos, nneg)
    ndat = npart * f.nsamp_step + f.nsamp_overlap
    ndat = (ndat + 511) // 512 * 512
    raw = synth.twobit_bytes(ndat, 2, seed=21)
    t = oracle.TwoBit()
    x, _ = t.unpack(raw, ndat, 2)
    volt = oracle.filterbank(f, x[:, :, : npart * f.nsamp_step + f.nsamp_overlap], H)
    ref = oracle.detect("Intensity", 1, volt)
    tb = E.make_twobit_desc(npol=2)
    ud = E.make_twobit_unpack_desc(tb)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Intensity", 1, 0)
    det = pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 0.0, first_sample=0)
    det = det.cpu().numpy()
    assert det.shape == ref.shape
    assert synth.relerr(det, ref) <= TOL


# ------------------------------------------------------------------------------------ other BASELINE configs, end to end
def _pipe_generic(ctx, oracle, fmt, input_nchan, npol, ndim, raw, ndat_unpacked, fbs, convs, H, C, F, npos, nneg,
                  npart, state, dndim, nbin, lut=None, scale=0.0, swap=1, nblock=1, max_npart=0):
    """raw bytes of `fmt` -> pipeline (fold) on the GPU vs the oracle pipeline."""
    torch, E = _torch(), _E()
    nkeep = F - npos - nneg
    pps = 1.0 / (0.41 * nkeep * npart)
    phis = [0.2 + 0.17 * b for b in range(nblock)]
    ppss = [pps] * nblock
    op = oracle.make_pipe(fmt, input_nchan, npol, ndim, lut, scale, fbs, convs, H, state, dndim, nbin)
    ref, ref_hits = oracle.pipe_run(op, raw, nblock, npart, phis, ppss, nthread=1)
    ud = E.make_unpack_desc(fmt, input_nchan, npol, ndim, lut, scale, swap)
    fd, keep = E.make_fb_desc(ndim == 1, input_nchan, npol, C, F, npos, nneg, H, max_npart)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, nbin)
    d_raw = torch.from_numpy(raw).cuda()
    step = pipe.info.nsamp_step
    for b in range(nblock):
        pipe.execute(d_raw, npart, phis[b], ppss[b], first_sample=b * npart * step)
    prof, hits, ntot = pipe.synch()
    assert np.array_equal(hits, ref_hits) and ntot == nblock * npart * nkeep
    return synth.relerr(prof, ref)


def test_pipeline_cfg3_meerkat_convolution_fold(ctx, oracle):
    """BASELINE configs[2] on a channel shard: MeerKAT 8-bit complex, 8 of the 1024 input channels,
    Convolution with the 65536-point response (M = 2536 + 2543), Coherence, fold 1024 bins."""
    L = _L()
    nchan, F, npos, nneg, npart = 8, 65536, 2536, 2543, 2
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    rng = np.random.default_rng(31)
    raw = rng.integers(-60, 60, size=ndat * nchan * 2 * 2, dtype=np.int8).view(np.uint8)
    _, scale = oracle.bittable8()
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    err = _pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                        "Coherence", 4, 1024, scale=np.float32(scale))
    assert err <= TOL, err


@pytest.mark.parametrize("fmt,nchan,npart,nblock,state,dndim,nbin", [
    ("meerkat", 5, 7, 1, "Stokes", 4, 257),       # 35 tiles: several per cluster, the channel changes inside a cluster's range
    ("meerkat", 3, 2, 2, "PPQQ", 1, 1024),        # two blocks into one PhaseSeries
    ("meerkat", 4, 19, 1, "Intensity", 1, 64),    # 76 tiles: five per group, the software pipeline in steady state
    ("uwb", 1, 3, 1, "Intensity", 1, 64),         # 16-bit blocks of 2048 samples
    ("generic8", 2, 3, 1, "Coherence", 2, 128),   # TFP bytes through the 8-bit table, Coherence with ndim 2
])
def test_cluster_convolution_kernel(ctx, oracle, fmt, nchan, npart, nblock, state, dndim, nbin):
    """clusterconv.cu: 65536-point convolutions folded on the fly run as ONE kernel on groups of 16 co-resident CTAs
    (exchanges through L2-resident matrices, counter barriers; -DCC_GROUP=0: hardware clusters + distributed shared
    memory).  Every source format it unpacks itself, every detection state, tile counts that do not divide into the
    groups, several tiles per group, several blocks into one PhaseSeries -- against the oracle pipeline."""
    L = _L()
    F, npos, nneg = 65536, 2536, 2543
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    rng = np.random.default_rng(77)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    nsamp = nblock * npart * c.nsamp_step + c.nsamp_overlap
    if fmt == "meerkat":
        ndat = (nsamp + 255) // 256 * 256
        raw = synth.meerkat_bytes(ndat, nchan, 2, seed=78)
        _, scale = oracle.bittable8()
        err = _pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin, scale=np.float32(scale), nblock=nblock)
    elif fmt == "uwb":
        ndat = (nsamp + 2047) // 2048 * 2048
        raw = synth.uwb_bytes(ndat, 2, seed=79)
        err = _pipe_generic(ctx, oracle, L.FMT_UWB16, 1, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin, nblock=nblock)
    else:
        ndat = nsamp
        raw = synth.generic8_bytes(ndat, nchan, 2, 2, seed=80)
        lut, _ = oracle.bittable8()
        err = _pipe_generic(ctx, oracle, L.FMT_GENERIC8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin, lut=lut, nblock=nblock)
    assert err <= TOL, err


@pytest.mark.parametrize("F,npos,nneg,nchan,npart", [(16384, 700, 650, 3, 30), (32768, 1200, 1300, 2, 21),
                                                       (131072, 5000, 5100, 2, 3), (131072, 0, 9000, 1, 41)])
def test_one_kernel_convolution_other_lengths(ctx, oracle, F, npos, nneg, nchan, npart):
    """clusterconv.cu is instantiated for N = 16 Q, Q = 1024 ... 8192: every length, with tile counts that leave one
    (131072: 18 groups) or several tiles per group, MeerKAT input, Coherence, against the oracle pipeline."""
    L = _L()
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=90 + nchan)
    _, scale = oracle.bittable8()
    rng = np.random.default_rng(F)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    err = _pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                        "Coherence", 4, 512, scale=np.float32(scale))
    assert err <= TOL, err


@pytest.mark.parametrize("state,dndim,F", [("Stokes", 4, 65536), ("Intensity", 1, 16384), ("Coherence", 2, 32768)])
def test_one_kernel_convolution_detected_series(ctx, oracle, state, dndim, F):
    """The detected-series sink (no fold: digifil with coherent dedispersion) of the one-kernel convolution path: UWB /
    MeerKAT bytes in, detected planes out, against the oracle chain."""
    torch, E, L = _torch(), _E(), _L()
    nchan, npart = 2, 5
    npos, nneg = F // 30, F // 29
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=95)
    _, scale = oracle.bittable8()
    rng = np.random.default_rng(96)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    op = oracle.make_pipe(L.FMT_MEERKAT8, nchan, 2, 2, None, np.float32(scale), None, c, H, state, dndim, 0)
    ref = oracle.pipe_block_detected(op, raw, 0, npart)
    ud = E.make_unpack_desc(L.FMT_MEERKAT8, nchan, 2, 2, None, np.float32(scale), 1)
    fd, keep = E.make_fb_desc(False, nchan, 2, 1, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, 0)
    det = pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 0.0, first_sample=0).cpu().numpy()
    assert det.shape == ref.shape
    assert np.array_equal(det, ref) or synth.relerr(det, ref) <= TOL


def test_pipeline_4096_input_channels_grid_limit(ctx, oracle):
    """4096 input channels x 2 polarisations = 8192 (channel, pol) blocks per part; with the default
    16 parts per launch the generic kernels' grid.y would be 131072 (> 65535).  The plan caps the batch; 20 parts
    force several launches.  MeerKAT heaps, 64-point convolution."""
    L = _L()
    nchan, F, npos, nneg, npart = 4096, 64, 5, 6, 20
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=91)
    _, scale = oracle.bittable8()
    rng = np.random.default_rng(92)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    err = _pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                        "Coherence", 4, 64, scale=np.float32(scale))
    assert err <= TOL, err


@pytest.mark.parametrize("fmt,F,nchan,npart,state,dndim,nbin", [
    ("meerkat", 1 << 18, 2, 3, "Stokes", 4, 257),       # 512 x 512
    ("uwb", 1 << 19, 1, 2, "Intensity", 1, 256),        # 1024 x 512
    ("generic8", 1 << 20, 2, 2, "Coherence", 2, 1024),  # 1024 x 1024, two channels
    ("meerkat", 1 << 21, 1, 2, "PPQQ", 1, 1024),        # 2048 x 1024
    ("meerkat", 1 << 22, 1, 1, "Stokes", 4, 4096),      # 2048 x 2048 (cfg4's shape, another source and state)
    ("generic8", 1 << 23, 1, 1, "Coherence", 4, 2048),  # 2048 x 4096: above 2^22 only these kernels exist
    ("uwb", 1 << 24, 1, 1, "Intensity", 1, 4096),       # 2048 x 8192, the longest transform
])
def test_long_convolution_kernels(ctx, oracle, fmt, F, nchan, npart, state, dndim, nbin):
    """longconv.cu: convolutions of more than 131072 points (N = P Q, P = 512 ... 2048, Q = 512 ...
    8192) run as three c2-core kernels with both polarisations of a bin side by side in the spectrum scratch.  Every
    source format, detection state and factorisation against the oracle pipeline (cfg4, 2048 x 2048, has its own test;
    tests/test_gpu_variants.py isolates each of the three kernels between the generic ones)."""
    L = _L()
    npos, nneg = F // 9, F // 11
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    rng = np.random.default_rng(F % 1000 + nchan)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    nsamp = npart * c.nsamp_step + c.nsamp_overlap
    if fmt == "meerkat":
        ndat = (nsamp + 255) // 256 * 256
        raw = synth.meerkat_bytes(ndat, nchan, 2, seed=178)
        _, scale = oracle.bittable8()
        err = _pipe_generic(ctx, oracle, L.FMT_MEERKAT8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin, scale=np.float32(scale))
    elif fmt == "uwb":
        ndat = (nsamp + 2047) // 2048 * 2048
        raw = synth.uwb_bytes(ndat, 2, seed=179)
        err = _pipe_generic(ctx, oracle, L.FMT_UWB16, 1, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin)
    else:
        ndat = nsamp
        raw = synth.generic8_bytes(ndat, nchan, 2, 2, seed=180)
        lut, _ = oracle.bittable8()
        err = _pipe_generic(ctx, oracle, L.FMT_GENERIC8, nchan, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                            state, dndim, nbin, lut=lut)
    assert err <= TOL, err


def test_long_convolution_detected_series(ctx, oracle):
    """The detected-series sink of the long-transform kernels (digifil with coherent dedispersion of one wide channel)."""
    torch, E, L = _torch(), _E(), _L()
    nchan, npart, F, state, dndim = 1, 3, 1 << 18, "Stokes", 4
    npos, nneg = F // 30, F // 29
    c = oracle.conv_sizes(0, nchan, 2, F, npos, nneg)
    ndat = (npart * c.nsamp_step + c.nsamp_overlap + 255) // 256 * 256
    raw = synth.meerkat_bytes(ndat, nchan, 2, seed=195)
    _, scale = oracle.bittable8()
    rng = np.random.default_rng(196)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
    op = oracle.make_pipe(L.FMT_MEERKAT8, nchan, 2, 2, None, np.float32(scale), None, c, H, state, dndim, 0)
    ref = oracle.pipe_block_detected(op, raw, 0, npart)
    ud = E.make_unpack_desc(L.FMT_MEERKAT8, nchan, 2, 2, None, np.float32(scale), 1)
    fd, keep = E.make_fb_desc(False, nchan, 2, 1, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, 0)
    det = pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 0.0, first_sample=0).cpu().numpy()
    assert det.shape == ref.shape
    assert np.array_equal(det, ref) or synth.relerr(det, ref) <= TOL


def test_pipeline_cfg4_single_channel_4m_convolution(ctx, oracle):
    """BASELINE configs[3]: one 400 MHz channel at 12.5 GHz, DM 1500, 2^22-point overlap-save
    (M = 534848 + 588748), generic 8-bit complex input, Coherence, fold."""
    L = _L()
    F, npos, nneg, npart = 1 << 22, 534848, 588748, 2
    c = oracle.conv_sizes(0, 1, 2, F, npos, nneg)
    ndat = npart * c.nsamp_step + c.nsamp_overlap
    rng = np.random.default_rng(41)
    raw = rng.integers(0, 256, size=ndat * 2 * 2, dtype=np.uint8)
    lut, _ = oracle.bittable8()
    d, H = oracle.dedispersion(12500.0, 400.0, 1500.0, 1, 1, False, frequency_resolution=F)
    assert (d.impulse_pos, d.impulse_neg) == (npos, nneg)
    err = _pipe_generic(ctx, oracle, L.FMT_GENERIC8, 1, 2, 2, raw, ndat, None, c, H, 1, F, npos, nneg, npart,
                        "Coherence", 4, 1024, lut=lut)
    assert err <= TOL, err


@pytest.mark.parametrize("sb,F,npos,nneg", [(0, 16384, 886, 890), (6, 2048, 98, 98), (25, 64, 6, 6)])
def test_pipeline_cfg5_uwl_subband(ctx, oracle, sb, F, npos, nneg):
    """BASELINE configs[4]: UWL-like 128 MHz sub-bands, 16-bit complex dual-pol, -F 128:D, fold.
    Sub-band 0 (freq_res 16384) takes the unfused tail (voltages -> detect -> fold)."""
    L = _L()
    C, npart = 128, 2
    d, H = oracle.dedispersion(768.0 + 128.0 * sb, 128.0, 67.99, 1, C, False)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == (F, npos, nneg)
    f = oracle.fb_sizes(0, 1, 2, C, F, npos, nneg)
    ndat = (npart * f.nsamp_step + f.nsamp_overlap + 2047) // 2048 * 2048
    rng = np.random.default_rng(50 + sb)
    raw = (rng.normal(0, 2000, size=ndat * 2 * 2).astype(np.int16) ^ np.int16(-32768)).view(np.uint8)
    err = _pipe_generic(ctx, oracle, L.FMT_UWB16, 1, 2, 2, raw, ndat, f, None, H, C, F, npos, nneg, npart,
                        "Coherence", 4, 1024)
    assert err <= TOL, err


# ------------------------------------------------------------------------------------ sub-integrations (a15)
def test_subint_folder_cuts_at_division_boundaries(ctx, oracle):
    """dsp::Subint<Fold>: a detected stream folded block by block with -L seconds; every unloaded
    sub-integration equals the oracle fold of exactly the samples of that division, sliced per block with
    Fold::fold's phase set-up, and the hit totals add up to the stream length."""
    torch, E = _torch(), _E()
    from dspsr_b200 import hostmath as HM
    import workloads as W
    from dspsr_b200.subint import SubintFolder
    nchan, npol, ndim, nbin = 3, 1, 4, 256
    rate = 1.5625e6 / 8
    Ldiv = 0.05                                   # 9765.625 samples per division
    start = HM.utc_to_mjd("2010-04-13-02:05:45")
    pred = HM.Polyco(W.polyco_text())
    opc = oracle.polyco_parse(W.polyco_text())
    rng = np.random.default_rng(77)
    blocks = [int(n) for n in rng.integers(3000, 9000, 9)]
    total = sum(blocks)
    x = (rng.standard_normal((nchan, npol, total * ndim)) + 2.0).astype(np.float32)
    got = {}

    def unload(division, prof, hits, ntot, partial):
        assert division not in got
        got[division] = (prof.copy(), hits.copy(), ntot, partial)

    fe = E.FoldEngine(ctx, nchan, npol, ndim, nbin)
    sf = SubintFolder(fe, pred, start, Ldiv, unload)
    # oracle: same cutting rule (restated TimeDivide), CPU fold per slice
    ot = oracle.TimeDivide(Ldiv)
    want = {}
    t0 = 0
    for n in blocks:
        blk = x[:, :, t0 * ndim:(t0 + n) * ndim]
        sf.fold_block(torch.from_numpy(np.ascontiguousarray(blk)).cuda(), t0 / rate, rate)
        more = True
        while more:
            o = ot.set_bounds(t0 / rate, rate, n)
            more = o["in_next"]
            if not o["is_valid"]:
                continue
            tb = HM.mjd_add(start, t0 / rate)
            ts = HM.mjd_add(tb, (o["idat_start"] + 0.5) / rate)
            phi = oracle.polyco_phase(opc, *ts)[0]
            pps = (1.0 / rate) / (1.0 / oracle.polyco_frequency(opc, *ts))      # Fold.C:718-720
            bp, hh, _, _ = oracle.fold_plan(phi, pps, nbin, o["ndat"])
            prof, hits = want.get(o["division"], (None, np.zeros(nbin, np.uint32)))
            prof = oracle.fold(np.ascontiguousarray(blk), ndim, bp, nbin, profile=prof, idat_start=o["idat_start"])
            want[o["division"]] = (prof, hits + hh)
        t0 += n
    sf.finish()
    assert sorted(got) == sorted(want) and len(got) >= 5
    ntot = 0
    for div in want:
        prof, hits, n, partial = got[div]
        assert np.array_equal(hits, want[div][1]) and n == hits.sum()
        assert synth.relerr(prof, want[div][0]) <= TOL
        ntot += n
    assert ntot == total
    assert got[min(got)][3] and got[max(got)][3]          # first and last divisions are flagged partial
    assert not any(got[d][3] for d in sorted(got)[1:-1])


# ------------------------------------------------------------------------------------ digifil tail (f1)
def test_rescale_and_sigproc_digitizer(ctx, oracle):
    """dsp::Rescale over several blocks (interval not a multiple of the block length) and the 8-bit SIGPROC
    digitiser: floats within 2e-6 of the CPU restatement (the device sums in parallel), bytes identical except
    where the value sits within 1e-3 of a rounding boundary; channel order flipped for positive bandwidth."""
    torch, E = _torch(), _E()
    rng = np.random.default_rng(91)
    nchan, npol = 96, 1
    blocks = [5000, 7000, 3000, 9000]
    interval = 8192
    gain = rng.uniform(0.5, 20.0, (nchan, npol, 1)).astype(np.float32)
    base = rng.uniform(1.0, 50.0, (nchan, npol, 1)).astype(np.float32)
    ro = oracle.Rescale(interval_samples=interval)
    rg = E.Rescale(ctx, nchan, npol, interval_samples=interval)
    for i, n in enumerate(blocks):
        x = (rng.standard_normal((nchan, npol, n)).astype(np.float32) ** 2 * gain + base * (1 + 0.1 * i)).astype(np.float32)
        want = ro.transform(x)
        got = rg.transform(torch.from_numpy(x).cuda())
        assert np.max(np.abs(got.cpu().numpy() - want)) <= 2e-5 * max(1.0, np.max(np.abs(want)))
        o, s = rg.offset_scale()
        assert np.allclose(o, ro.offset, rtol=1e-6) and np.allclose(s, ro.scale, rtol=1e-6)
        for bw in (-64.0, 64.0):
            bo = oracle.sigproc_digitize(want, bandwidth=bw)
            bg = E.sigproc_digitize8(ctx, torch.from_numpy(want).cuda(), bandwidth=bw).cpu().numpy()
            assert bg.shape == bo.shape == (n, npol, nchan)
            assert np.array_equal(bg, bo)
    assert 100 < bo.mean() < 155      # rescaled noise sits around 127.5


# ------------------------------------------------------------------------------------ fast-path K3 epilogues
@pytest.mark.parametrize("state,dndim,nbin", [("Stokes", 2, 128), ("Intensity", 1, 64), ("PPQQ", 1, 256),
                                               ("Coherence", 1, 0), ("Stokes", 4, 0), ("Intensity", 1, 0)])
def test_fast_k3_all_epilogues(ctx, oracle, state, dndim, nbin):
    """freq_res 8192 with two polarisations takes the second-generation K3 (fastpath.cu) whatever the
    forward factorisation: every detection state / layout, folded (run-time state variant) and unfolded."""
    torch, E, L = _torch(), _E(), _L()
    C, F, npos, nneg, npart = 4, 8192, 457, 459, 3
    if nbin:
        err = _pipeline_case(ctx, oracle, C, F, npos, nneg, npart, state, dndim, nbin, nblock=1)
        assert err <= TOL, err
        return
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C, F, npos, nneg)
    ndat = (npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=61)
    H = np.exp(1j * np.random.default_rng(62).uniform(-np.pi, np.pi, (C, F))).astype(np.complex64)
    x = oracle.unpack_caspsr(raw, ndat, lut)
    volt = oracle.filterbank(f, x[:, :, : npart * f.nsamp_step + f.nsamp_overlap], H)
    ref = oracle.detect(state, dndim, volt)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, 0)
    det = pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 0.0, first_sample=0).cpu().numpy()
    assert det.shape == ref.shape
    assert np.array_equal(det, ref) or synth.relerr(det, ref) <= TOL


@pytest.mark.parametrize("C,F,npos,nneg,state,dndim", [(64, 4, 1, 0, "Stokes", 4), (32, 16, 2, 3, "Coherence", 2),
                                                         (128, 2, 0, 1, "PPQQ", 1), (4096, 8, 1, 1, "Intensity", 1)])
def test_short_channel_transforms_detected_series(ctx, oracle, C, F, npos, nneg, state, dndim):
    """k_chan_inv_small: freq_res 2 ... 16 (cfg2: 8) with the detected-series sink -- one thread per channel, several
    parts per thread -- for every detection state, with a part count that does not divide into the part groups."""
    torch, E, L = _torch(), _E(), _L()
    npart = 37
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C, F, npos, nneg)
    ndat = (npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=71)
    H = np.exp(1j * np.random.default_rng(72).uniform(-np.pi, np.pi, (C, F))).astype(np.complex64)
    x = oracle.unpack_caspsr(raw, ndat, lut)
    volt = oracle.filterbank(f, x[:, :, : npart * f.nsamp_step + f.nsamp_overlap], H)
    ref = oracle.detect(state, dndim, volt)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, state, dndim, 0)
    det = pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 0.0, first_sample=0).cpu().numpy()
    assert det.shape == ref.shape
    assert np.array_equal(det, ref) or synth.relerr(det, ref) <= TOL


@pytest.mark.parametrize("C,F,npos,nneg", [(32, 256, 20, 21), (16, 512, 30, 31), (8, 1024, 60, 61), (4, 2048, 98, 98),
                                           (2, 4096, 200, 201)])
def test_fast_k3_every_planned_length(ctx, oracle, C, F, npos, nneg):
    """The second-generation K3 is instantiated for every per-channel length the c2 core plans (256 ... 8192),
    with 8192/F channels sharing a CTA: fused fold (Coherence) and fused detection (Stokes, ndim 2) per length."""
    err = _pipeline_case(ctx, oracle, C, F, npos, nneg, 3, "Coherence", 4, 128, nblock=2)
    assert err <= TOL, err
    err = _pipeline_case(ctx, oracle, C, F, npos, nneg, 2, "Stokes", 2, 64, nblock=1)
    assert err <= TOL, err


def test_pipeline_cfg1_bench_scale(ctx, oracle):
    """BASELINE configs[0] at the size bench.py times (32 overlap-save parts = 119 M samples per polarisation,
    239 MB of raw bytes, 4 blocks of 8 parts with their own fold phase): whole folded profile against the
    oracle, exact hits, and the device-resident and host-fed paths agree bit for bit in hits and to 1e-6
    in the profile (they differ only in internal batch size)."""
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import hostmath as HM
    C, F, npos, nneg, nbin = 256, 8192, 457, 459, 1024
    nblock, npart = 4, 8
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C, F, npos, nneg)
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=99)
    d, H = HM.dedispersion(1382.0, -400.0, 67.99, 1, C, True)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == (F, npos, nneg)
    pps = 7.18e-6
    phis = [(0.1 + b * npart * f.nkeep * pps) % 1.0 for b in range(nblock)]
    op = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    ref, ref_hits = oracle.pipe_run(op, raw, nblock, npart, phis, [pps] * nblock, nthread=4)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
    d_raw = torch.from_numpy(raw).cuda()
    for b in range(nblock):
        pipe.execute(d_raw, npart, phis[b], pps, first_sample=b * npart * f.nsamp_step)
    prof, hits, ntot = pipe.synch()
    assert np.array_equal(hits, ref_hits) and ntot == nblock * npart * f.nkeep == hits.sum()
    err = synth.relerr(prof, ref)
    assert err <= TOL, err
    # host-fed path (chunked copies, smaller internal batches)
    pipe.zero()
    for b in range(nblock):
        lo = b * npart * f.nsamp_step * 2
        nbytes = (npart * f.nsamp_step + f.nsamp_overlap) * 2
        pipe.execute_host(raw[lo: lo + nbytes], npart, phis[b], pps, 0)
    prof2, hits2, ntot2 = pipe.synch()
    assert np.array_equal(hits2, hits) and ntot2 == ntot
    assert synth.relerr(prof2, prof) <= 1e-6


@pytest.mark.parametrize("C,F,npos,nneg,nbin,period", [(128, 128, 19, 9, 8, 23.0), (16, 256, 20, 21, 8, 11.0),
                                                       (4, 8192, 457, 459, 16, 37.0), (64, 32, 0, 4, 4, 5.5)])
def test_pipeline_fold_short_period(ctx, oracle, C, F, npos, nneg, nbin, period):
    """Pulse periods of a few samples with few bins: the phase wraps many times inside the samples one warp
    walks, so equal (channel, bin) keys are NOT contiguous across lanes (regression test for a shuffle-scan
    reduction that double counted in this regime; found by scratch/fuzz_gpu.py)."""
    err = _pipeline_case(ctx, oracle, C, F, npos, nneg, 2, "Stokes", 2, nbin, nblock=2, pps=1.0 / period)
    assert err <= TOL, err


# ------------------------------------------------------------------------------------ library-side PhaseSeries (a14, f3)
def test_pipeline_execute_obs_keeps_phase_series_attributes(ctx, oracle, tmp_path):
    """b200_pipeline_execute_obs: the library evaluates the polyco at the midpoint of each block's first output sample
    (Fold.C:650-657) and maintains the PhaseSeries attributes like Fold::transformation / Fold::fold (mixable,
    integration_length, ndat_total, start/end, scale of Filterbank.C:124-125); the result is dumped normalised
    (Archiver::set) and read back.  Oracle: the same blocks with the oracle's own predictor."""
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import hostmath as HM
    from dspsr_b200 import phaseseries as P
    import workloads as W
    C_, F, npos, nneg, npart, nblock, nbin = 16, 256, 20, 21, 3, 3, 128
    cfg = dict(W.CFG1, nchan=C_)
    S = W.sizes(cfg, F, npos, nneg)
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C_, F, npos, nneg)
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=61)
    rng = np.random.default_rng(62)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C_, F))).astype(np.complex64)
    start = W.utc_to_mjd(cfg["utc_start"])
    opc = oracle.polyco_parse(W.polyco_text())
    ph = [W.block_phase(S, start, b * npart * S["step"], lambda m: oracle.polyco_phase(opc, *m)[0],
                        lambda m: oracle.polyco_frequency(opc, *m)) for b in range(nblock)]
    op = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    ref, ref_hits = oracle.pipe_run(op, raw, nblock, npart, [p[0] for p in ph], [p[1] for p in ph], nthread=1)

    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C_, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
    robs = P.observation(1, 2, 1, S["rate_in"], start, ndat=ndat, centre_frequency=cfg["freq"], bandwidth=cfg["bw"],
                         dm=cfg["dm"], state=17, nbit=8)
    pipe.set_observation(robs)
    pipe.set_predictor(HM.Polyco(W.polyco_text()))
    d_raw = torch.from_numpy(raw).cuda()
    order = [1, 0, 2]                                          # blocks need not arrive in time order
    for b in order:
        first = b * npart * S["step"]
        pipe.execute_obs(d_raw, npart, obs_sample=first, first_sample=first)
    ps = pipe.phase_series()
    assert np.array_equal(ps.hits, ref_hits)
    # the Vela period (89 ms) is much longer than these few blocks: compare on the bins that were hit
    hit = ref_hits > 0
    assert synth.relerr(ps.data.reshape(C_, 1, nbin, 4)[:, :, hit], ref.reshape(C_, 1, nbin, 4)[:, :, hit]) <= TOL
    assert np.all(ps.data.reshape(C_, 1, nbin, 4)[:, :, ~hit] == 0)
    n_out = nblock * npart * S["nkeep"]
    assert ps.ndat_total == n_out
    assert ps.integration_length == pytest.approx(n_out / S["rate_out"], rel=1e-12)
    o = ps.ps.obs
    assert (o.nchan, o.npol, o.ndim, o.state) == (C_, 1, 4, L.COHERENCE)
    assert o.rate == pytest.approx(S["rate_out"], rel=1e-15) and o.scale == float(C_ * F) * F
    t0 = W.mjd_add(start, npos / S["rate_out"])
    t1 = W.mjd_add(t0, (nblock * npart * S["step"]) / S["rate_in"] - (npart * S["step"]) / S["rate_in"] + npart * S["nkeep"] / S["rate_out"])
    assert ps.start_time[:2] == t0[:2] and ps.start_time[2] == pytest.approx(t0[2], abs=1e-9)
    assert ps.end_time[:2] == t1[:2] and ps.end_time[2] == pytest.approx(t1[2], abs=1e-9)
    # unload: normalised by scale * hits
    path = tmp_path / "cfg1mini.b200ps"
    ps.unload(path)
    hdr, prof, w, hits, rawsum = P.load(path)
    assert np.array_equal(hits[0], ref_hits) and np.all(w == 1)
    want = ref.reshape(C_, 1, nbin, 4).transpose(0, 1, 3, 2)[..., hit] / (o.scale * ref_hits[hit].astype(np.float64))
    assert synth.relerr(prof[..., hit], want) <= TOL
    # bins without hits take the mean of the others (Archiver.C:869-889)
    assert np.allclose(prof[..., ~hit], prof[..., hit].astype(np.float64).mean(axis=-1, keepdims=True), rtol=1e-5)
    # a block of another observation is refused, and reset clears the integration
    pipe.set_observation(P.observation(1, 2, 1, S["rate_in"], start, ndat=ndat, centre_frequency=1400.0,
                                       bandwidth=cfg["bw"], dm=cfg["dm"], state=17, nbit=8))
    with pytest.raises(L.B200Error):
        pipe.execute_obs(d_raw, npart, obs_sample=0, first_sample=0)
    pipe.reset()
    assert pipe.phase_series().integration_length == 0.0


# ------------------------------------------------------------------------------------ WeightedTimeSeries flags (f4)
@pytest.mark.parametrize("nfft,nkeep,npw,widat,nscr", [(2048, 1408, 512, 0, 128), (4096, 3000, 512, 100, 512),
                                                         (65536, 49152, 512, 0, 8192), (1024, 512, 512, 7, 2)])
def test_weights_convolve_and_scrunch_match_reference_loops(ctx, oracle, nfft, nkeep, npw, widat, nscr):
    """b200_weights_convolve / b200_weights_scrunch against the literal sequential loops of
    WeightedTimeSeries.C:582-690,692-780 (oracle), random flags at several densities incl. none and all."""
    torch, E = _torch(), _E()
    rng = np.random.default_rng(nfft + nkeep)
    nblocks = 37
    ndat = nblocks * nkeep + (nfft - nkeep)
    nw = (ndat + widat + npw - 1) // npw + 1
    for density in (0.0, 0.002, 0.05, 0.5, 1.0):
        w = (rng.random(nw) >= density).astype(np.uint32) * rng.integers(1, 4, nw).astype(np.uint32)
        want = oracle.convolve_weights(w, npw, widat, ndat, nfft, nkeep)
        got = E.weights_convolve(ctx, torch.from_numpy(w.astype(np.int32)).cuda(), npw, widat, ndat, nfft, nkeep)
        assert np.array_equal(got.cpu().numpy().astype(np.uint32), want), density
        w2, npw2, wi2 = oracle.scrunch_weights(want, npw, widat, nscr)
        g2, gnpw, gwi = E.weights_scrunch(ctx, got, npw, widat, nscr)
        assert (gnpw, gwi) == (npw2, wi2)
        assert np.array_equal(g2.cpu().numpy().astype(np.uint32)[: w2.size], w2)


def test_fold_engine_skips_flagged_windows(ctx, oracle):
    """b200_fold_set_bins_weighted: bins, hits and ndat_folded of Fold.C:687-788 with a weighted input."""
    torch, E = _torch(), _E()
    rng = np.random.default_rng(71)
    nchan, npol, ndim, nbin, ndat, npw, widat, start = 3, 2, 2, 64, 20000, 128, 37, 11
    w = (rng.random((start + ndat + widat) // npw + 1) > 0.2).astype(np.uint32)
    x = rng.standard_normal((nchan, npol, (start + ndat) * ndim)).astype(np.float32)
    phi, pps = 0.37, 1.0 / 713.3
    bp, hits, nfold = oracle.fold_plan_weighted(phi, pps, nbin, start, ndat, w, npw, widat)
    ref = oracle.fold(x, ndim, bp, nbin, idat_start=start)
    fe = E.FoldEngine(ctx, nchan, npol, ndim, nbin)
    fe.set_bins_weighted(phi, pps, ndat, start, torch.from_numpy(w.astype(np.int32)).cuda(), npw, widat)
    fe.fold(torch.from_numpy(x).cuda())
    h, ntot = fe.hits()
    assert np.array_equal(h, hits) and int(h.sum()) == nfold < ndat and ntot == ndat
    assert synth.relerr(fe.synch(), ref) <= TOL


@pytest.mark.parametrize("C_,F,npos,nneg", [(64, 16, 2, 2), (128, 256, 20, 20)])
def test_pipeline_twobit_fold_with_excised_windows(ctx, oracle, C_, F, npos, nneg):
    """Two-bit input with windows that the excision unpacker flags (all-zero bytes in one polarisation, a run of
    saturated samples in the other): the flags are convolved with the overlap-save transforms, scrunched to the
    filterbank's time resolution and the fold skips the flagged samples -- profile, hits, ndat_total and
    integration_length as the reference's CPU path (oracle pipeline with a WeightedTimeSeries)."""
    torch, E, L = _torch(), _E(), _L()
    npart, nbin, nblock = 9, 32, 2
    f = oracle.fb_sizes(1, 1, 2, C_, F, npos, nneg)
    assert f.nsamp_step % 512 == 0                  # blocks start on a window boundary
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 511) // 512 * 512
    raw = synth.twobit_bytes(ndat, 2, seed=81).copy()
    nwin = ndat // 512
    rng = np.random.default_rng(82)
    # few enough that some transforms stay clean (a transform holds nsamp_fft / 512 windows)
    for wbad in rng.choice(nwin, max(2, nwin * 512 // (12 * f.nsamp_fft)), replace=False):
        if wbad % 2:
            raw[wbad * 256: wbad * 256 + 256: 2] = 0            # polarisation 0: all-zero bytes
        else:
            raw[wbad * 256 + 1: wbad * 256 + 256: 2] = 0xFF     # polarisation 1: every sample in the top state
    t = oracle.TwoBit()
    _, w = t.unpack(raw, ndat, 2)
    assert 0 < (w[0] == 0).sum() < nwin
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C_, F))).astype(np.complex64)
    pps = 1.0 / (0.41 * f.nkeep * npart)
    phis = [0.2 + 0.17 * b for b in range(nblock)]
    op = oracle.make_pipe(5, 1, 2, 1, None, 0.0, f, None, H, "Coherence", 4, nbin, twobit=t)
    ref, ref_hits, ref_nfold = oracle.pipe_run(op, raw, nblock, npart, phis, [pps] * nblock, nthread=1, with_total=True)
    assert 0 < ref_nfold < nblock * npart * f.nkeep
    tb = E.make_twobit_desc(npol=2)
    ud = E.make_twobit_unpack_desc(tb)
    fd, keep = E.make_fb_desc(1, 1, 2, C_, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
    pipe.reserve(npart)
    d_raw = torch.from_numpy(raw).cuda()
    for b in range(nblock):
        pipe.execute(d_raw, npart, phis[b], pps, first_sample=b * npart * f.nsamp_step)
    prof, hits, ntot = pipe.synch()
    assert np.array_equal(hits, ref_hits) and int(hits.sum()) == ref_nfold
    assert ntot == nblock * npart * f.nkeep
    assert synth.relerr(prof, ref) <= TOL


# ------------------------------------------------------------------------------------ streaming input with carry (f2)
@pytest.mark.parametrize("from_host", [False, True])
def test_stream_feed_carries_block_edges(ctx, oracle, from_host):
    """b200_pipeline_stream_begin / feed: blocks of ragged length; each feed processes the whole parts of
    [carried tail | new block] as ONE Fold call (phase evaluated by the library for the first output sample of that
    call) and carries the rest (InputBuffering::set_next_start).  Oracle: the same partition of the stream."""
    import math
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import phaseseries as P
    import workloads as W
    C_, F, npos, nneg, nbin = 16, 256, 20, 21, 64
    cfg = dict(W.CFG1, nchan=C_)
    S = W.sizes(cfg, F, npos, nneg)
    step, overlap = S["step"], S["overlap"]
    rng = np.random.default_rng(91)
    feeds = [int(x) // 4 * 4 for x in rng.integers(step // 3, 4 * step, 9)] + [4, 0, 5 * step]
    ndat = sum(feeds)
    raw = synth.caspsr_bytes(ndat, seed=92)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C_, F))).astype(np.complex64)
    lut, _ = oracle.bittable8()
    start = W.utc_to_mjd(cfg["utc_start"])
    period = 0.37 * 3 * S["nkeep"] / S["rate_out"]

    def phase(m):
        return math.fmod((m[0] - start[0]) * 86400.0 + (m[1] - start[1]) + (m[2] - start[2]), period) / period
    # the partition the feeds imply
    blocks, have, done = [], 0, 0
    for n in feeds:
        have += n
        npart = (have - done * step - overlap) // step if have - done * step > overlap else 0
        if npart:
            phi, pps = W.block_phase(S, start, done * step, phase, lambda m: 1.0 / period)
            blocks.append((done, npart, phi, pps))
            done += npart
    assert len(blocks) >= 8 and done == (ndat - overlap) // step
    f = oracle.fb_sizes(1, 1, 2, C_, F, npos, nneg)
    op = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    ref, ref_hits, _ = oracle.pipe_blocks(op, raw, blocks)

    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
    fd, keep = E.make_fb_desc(1, 1, 2, C_, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
    pipe.set_observation(P.observation(1, 2, 1, S["rate_in"], start, ndat=ndat, centre_frequency=cfg["freq"],
                                       bandwidth=cfg["bw"], dm=cfg["dm"], state=17))
    pipe.set_folding_period(period, reference_epoch=start)
    pipe.stream_begin(max(feeds))
    pos, got_parts = 0, []
    h_raw = torch.from_numpy(raw).pin_memory()
    d_raw = torch.from_numpy(raw).cuda()
    for n in feeds:
        src = h_raw[2 * pos: 2 * (pos + n)] if from_host else d_raw[2 * pos: 2 * (pos + n)]
        if n == 0:
            src = h_raw[:4] if from_host else d_raw[:4]
        got_parts.append(pipe.feed(src, n))
        pos += n
    assert [g for g in got_parts if g] == [b[1] for b in blocks]
    ps = pipe.phase_series()
    assert np.array_equal(ps.hits, ref_hits)
    assert synth.relerr(ps.data, ref) <= TOL
    assert ps.ndat_total == done * S["nkeep"]


def test_stream_feed_twobit_filterbank_detected_and_fil(ctx, oracle, tmp_path):
    """cfg2's chain as a stream: 2-bit blocks -> coherent filterbank -> Intensity, carried across block edges; the
    concatenated detected series equals the oracle's over the whole stream; then Rescale + 8-bit digitiser and a
    SIGPROC file whose payload is the oracle's digitised bytes."""
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import phaseseries as P
    C_, F, npos, nneg = 64, 8, 1, 1
    f = oracle.fb_sizes(1, 1, 2, C_, F, npos, nneg)
    rng = np.random.default_rng(95)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C_, F))).astype(np.complex64)
    feeds = [int(x) // 512 * 512 for x in rng.integers(600, 6 * f.nsamp_step, 8)]
    ndat = sum(feeds)
    raw = synth.twobit_bytes(ndat, 2, seed=96)
    t = oracle.TwoBit()
    op = oracle.make_pipe(5, 1, 2, 1, None, 0.0, f, None, H, "Intensity", 1, 0, twobit=t)
    nparts = (ndat - f.nsamp_overlap) // f.nsamp_step
    ref = oracle.pipe_block_detected(op, raw, 0, nparts)
    tb = E.make_twobit_desc(npol=2)
    ud = E.make_twobit_unpack_desc(tb)
    fd, keep = E.make_fb_desc(1, 1, 2, C_, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "Intensity", 1, 0)
    pipe.stream_begin(max(feeds))
    d_raw = torch.from_numpy(raw).cuda()
    pos, chunks = 0, []
    maxparts = max(feeds) // f.nsamp_step + 2
    out = torch.empty((C_, 1, maxparts * f.nkeep), dtype=torch.float32, device="cuda")
    for n in feeds:
        k = pipe.feed(d_raw[pos // 2: (pos + n) // 2], n, out=out)
        if k:
            chunks.append(out[:, :, : k * f.nkeep].clone())
        pos += n
    det = torch.cat(chunks, dim=2)
    assert det.shape[2] == nparts * f.nkeep
    assert synth.relerr(det.cpu().numpy(), ref) <= TOL
    # digifil tail: Rescale over the whole series, 8-bit digitiser, .fil
    r = E.Rescale(ctx, C_, 1, interval_samples=0)
    scaled = r.transform(det.contiguous())
    fil = E.sigproc_digitize8(ctx, scaled, bandwidth=128.0)
    orr = oracle.Rescale(interval_samples=0)
    want = oracle.sigproc_digitize(orr.transform(ref), nbit=8, bandwidth=128.0)
    got = fil.cpu().numpy()
    assert got.shape == want.shape and np.mean(got != want) < 2e-3         # float sums differ in the last bits
    obs = P.observation(C_, 1, 1, 128e6 * 2 * F / f.nsamp_fft, (55299, 7545, 0.0), centre_frequency=1400.0,
                        bandwidth=128.0, state=L.INTENSITY, telescope="PKS", machine="CPSR2")
    h = L.SigprocHeader()
    L.check(ctx.lib.b200_sigproc_header_from_observation(C.byref(obs), 8, C.byref(h)))
    path = tmp_path / "cfg2mini.fil"
    L.check(ctx.lib.b200_sigproc_file_write(str(path).encode(), C.byref(h), got.ctypes.data, got.size, 0))
    blob = open(path, "rb").read()
    assert blob.endswith(got.tobytes()) and b"HEADER_END" in blob and h.nchans == C_ and h.foff == -2.0


# ------------------------------------------------------------------------------------ science known answer
@pytest.mark.parametrize("F,dm", [(8192, 1.0), (65536, 6.0)])
def test_injected_dispersed_pulsar_folds_to_a_spike(ctx, oracle, F, dm):
    """A known answer that does not come from the oracle's arithmetic: a train of one-sample pulses (period 1777.25
    samples) is dispersed with the ANALYTIC cold-plasma chirp evaluated on the full-length frequency grid (what the
    interstellar medium does), quantised to 8-bit complex samples and sent through unpack -> coherent dedispersion ->
    detection -> fold on the GPU with the reference's Dedispersion response.  The folded profile must be a spike in
    the phase bin the construction predicts, holding nearly all of the pulsed power."""
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import hostmath as HM
    bw, cf, nbin, npart = 16.0, 1400.0, 256, 6
    d, H = HM.dedispersion(cf, bw, dm, 1, 1, False, frequency_resolution=F)
    npos, nneg = d.impulse_pos, d.impulse_neg
    step, overlap = F - npos - nneg, npos + nneg
    n = npart * step + overlap
    period = 1777.25
    t0 = npos + 40.0
    pulse = np.zeros(n, np.complex128)
    k = 0
    while t0 + k * period < n - nneg - 2:
        pulse[int(round(t0 + k * period))] = 1.0
        k += 1
    assert k >= 20
    kk = np.arange(n)
    f = np.where(kk < n // 2, kk / n, kk / n - 1.0) * bw                     # baseband frequency of FFT bin (MHz)
    phase = -2 * np.pi * (1e6 * dm / 2.41e-4) / cf ** 2 * f * f / (cf + f)   # Dedispersion.C:534-545, bw > 0
    disp = np.fft.ifft(np.fft.fft(pulse) * np.exp(-1j * phase))              # the ISM applies the inverse filter
    amp = 100.0 / np.abs(np.concatenate([disp.real, disp.imag])).max()
    raw = np.zeros((n, 1, 2, 2), np.int8)                                    # generic TFP bytes [t][chan][pol][re,im]
    raw[:, 0, 0, 0] = np.clip(np.rint(disp.real * amp - 0.5), -128, 127)     # the table adds 0.5 (BitTable.C:176)
    raw[:, 0, 0, 1] = np.clip(np.rint(disp.imag * amp - 0.5), -128, 127)
    raw = raw.reshape(-1).view(np.uint8)
    lut, _ = HM.bittable8()
    ud = E.make_unpack_desc(L.FMT_GENERIC8, 1, 2, 2, lut)
    fd, keep = E.make_fb_desc(False, 1, 2, 1, F, npos, nneg, H)
    pipe = E.Pipeline(ctx, ud, fd, keep, "PPQQ", 1, nbin)
    # output sample m is input sample m + nfilt_pos; fold with the train's own period, phase 0 at output sample 0
    pipe.execute(torch.from_numpy(raw).cuda(), npart, 0.0, 1.0 / period, first_sample=0)
    prof, hits, ntot = pipe.synch()
    assert ntot == npart * step and hits.sum() == ntot
    pp = prof[0, 0] / np.maximum(hits, 1)
    want_bin = int(((t0 - npos) / period % 1.0) * nbin)
    assert int(pp.argmax()) == want_bin
    on = pp[want_bin] + pp[(want_bin + 1) % nbin] + pp[(want_bin - 1) % nbin]
    off = np.median(pp)
    assert (on - 3 * off) / max(pp.sum() - nbin * off, 1e-30) > 0.9
    # recovered pulse energy: each unit pulse was scaled by amp and by the table's gain; the unnormalised FFT pair
    # multiplies amplitudes by F (Convolution.C:303-305), so a dedispersed pulse carries (amp * g * F)^2
    g = float(lut[129] - lut[128])                                           # table units per count
    per_pulse = (amp * g * F) ** 2
    pulses_kept = sum(1 for i in range(k) if npos <= int(round(t0 + i * period)) < npos + npart * step)
    got = (prof[0, 0, want_bin] + prof[0, 0, (want_bin + 1) % nbin] + prof[0, 0, (want_bin - 1) % nbin])
    assert got == pytest.approx(pulses_kept * per_pulse, rel=0.08)


# ------------------------------------------------------------------------------------ reproducible fold (hard part 3)
def test_deterministic_fold_is_bit_reproducible(ctx, oracle):
    """b200_pipeline_set_deterministic: fixed-point accumulation makes the folded profile bit-identical from run to run
    (cfg1 shape: 148 persistent CTAs add their run sums in whatever order they finish), still within the parity
    tolerance of the oracle; the same for the convolution path and the stand-alone fold engine."""
    torch, E, L = _torch(), _E(), _L()
    from dspsr_b200 import hostmath as HM
    C_, npart, nbin = 256, 3, 1024
    d, H = HM.dedispersion(1382.0, -400.0, 67.99, 1, C_, True)
    lut, _ = HM.bittable8()
    f = oracle.fb_sizes(1, 1, 2, C_, d.ndat, d.impulse_pos, d.impulse_neg)
    ndat = (npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=111)
    phi, pps = 0.25, 1.0 / (0.37 * npart * f.nkeep)
    d_raw = torch.from_numpy(raw).cuda()

    def run(det):
        ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
        fd, keep = E.make_fb_desc(1, 1, 2, C_, d.ndat, d.impulse_pos, d.impulse_neg, H)
        pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
        if det:
            pipe.set_deterministic()
        out = []
        for _ in range(3):
            pipe.zero()
            pipe.execute(d_raw, npart, phi, pps)
            out.append(pipe.synch()[0].copy())
        return out
    a = run(True)
    assert np.array_equal(a[0], a[1]) and np.array_equal(a[0], a[2])
    lo, _ = oracle.bittable8()
    _, Ho = oracle.dedispersion(1382.0, -400.0, 67.99, 1, C_, True)
    op = oracle.make_pipe(0, 1, 2, 1, lo, 0.0, f, None, Ho, "Coherence", 4, nbin)
    ref, _ = oracle.pipe_run(op, raw, 1, npart, [phi], [pps], 1)
    assert synth.relerr(a[0], ref) <= TOL
    b = run(False)
    assert synth.relerr(b[0], ref) <= TOL
    # stand-alone fold engine
    rng = np.random.default_rng(112)
    x = torch.from_numpy(rng.standard_normal((8, 2, 50000 * 2)).astype(np.float32) * 1e3).cuda()
    res = []
    for _ in range(2):
        fe = E.FoldEngine(ctx, 8, 2, 2, 64)
        L.check(ctx.lib.b200_fold_set_deterministic(fe.h, 2.0 ** -12))
        fe.set_bins(0.1, 1.0 / 977.0, 50000)
        fe.fold(x)
        res.append(fe.synch())
    assert np.array_equal(res[0], res[1])
    bp, hh, _, _ = oracle.fold_plan(0.1, 1.0 / 977.0, 64, 50000)
    assert synth.relerr(res[0], oracle.fold(x.cpu().numpy(), 2, bp, 64)) <= TOL
