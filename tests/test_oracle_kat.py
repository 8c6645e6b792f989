"""CPU tests of the oracle itself: known-answer tests derived from the mathematics of the path
(the reference holds no golden vectors -- SURVEY.md 4, 8c -- so these pin the restatement)."""
import numpy as np
import pytest
import scipy.fft

import synth


def test_fft_conventions_against_numpy_and_pocketfft(oracle):
    rng = np.random.default_rng(0)
    for n in (8, 64, 1024, 8192, 1 << 16, 1 << 18):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        ref = np.fft.fft(x.astype(np.complex128))
        rms = np.sqrt(np.mean(np.abs(ref) ** 2))
        assert np.abs(oracle.fcc1d(x) - ref).max() / rms < 2e-6            # forward: exp(-i), unnormalised
        refb = np.fft.ifft(x.astype(np.complex128)) * n
        assert np.abs(oracle.bcc1d(x) - refb).max() / rms < 2e-6           # backward: exp(+i), no 1/N
        # an independent single-precision FFT (pocketfft) agrees to float accuracy
        assert np.abs(oracle.fcc1d(x) - scipy.fft.fft(x)).max() / rms < 3e-6
    for n in (16, 4096, 1 << 17):
        x = rng.standard_normal(n).astype(np.float32)
        ref = np.fft.rfft(x.astype(np.float64))
        rms = np.sqrt(np.mean(np.abs(ref) ** 2))
        y = oracle.frc1d(x)
        assert y.size == n // 2 + 1                                        # FFTW r2c layout
        assert np.abs(y - ref).max() / rms < 2e-6


def test_bittable8_properties(oracle):
    lut, scale = oracle.bittable8(True)
    # two's complement: byte 0 -> +0.5 step, byte 255 -> -0.5 step, byte 128 -> most negative
    assert lut[0] > 0 and lut[255] < 0 and lut[0] == -lut[255]
    assert lut[128] == lut.min() and lut[127] == lut.max()
    d = np.diff(lut[:128].astype(np.float64))
    assert np.allclose(d, scale, rtol=2e-5)      # float32 table: spacing = get_scale() to float accuracy
    assert scale == pytest.approx(0.02957, rel=2e-3)   # ~ JA98 optimal 8-bit spacing
    lut_ob, _ = oracle.bittable8(False)
    assert np.array_equal(lut_ob, np.roll(lut, 128))   # offset binary = rotated table


def test_unpack_layouts(oracle):
    lut, scale = oracle.bittable8()
    # CASPSR: byte 8*(i/4) + 4*p + i%4
    ndat = 64
    raw = np.arange(2 * ndat, dtype=np.uint8)
    out = oracle.unpack_caspsr(raw, ndat, lut)
    for p in range(2):
        for i in (0, 1, 5, 63):
            assert out[0, p, i] == lut[raw[8 * (i // 4) + 4 * p + i % 4]]
    # generic TFP: byte i*(nchan*npol*ndim) + ndim*(npol*c+p) + d
    nchan, npol, ndim, ndat = 3, 2, 2, 10
    raw = np.random.default_rng(1).integers(0, 256, ndat * nchan * npol * ndim, dtype=np.uint8)
    out = oracle.unpack_generic8(raw, ndat, nchan, npol, ndim, lut)
    for (c, p, i, d) in [(0, 0, 0, 0), (2, 1, 9, 1), (1, 0, 4, 1)]:
        assert out[c, p, i * ndim + d] == lut[raw[i * nchan * npol * ndim + ndim * (npol * c + p) + d]]
    # MeerKAT: [heap][pol][chan][256 x (re, im)] int8, (x + 0.5) * scale
    nchan, npol, ndat = 4, 2, 512
    raw = synth.meerkat_bytes(ndat, nchan, npol, seed=2)
    s = float(np.float32(scale))
    out = oracle.unpack_meerkat(raw, ndat, nchan, npol, s)
    r8 = raw.view(np.int8)
    for (c, p, i) in [(0, 0, 0), (3, 1, 300), (2, 0, 511)]:
        w = ((i // 256 * npol + p) * nchan + c) * 256 + i % 256
        assert out[c, p, 2 * i] == np.float32((np.float32(r8[2 * w]) + 0.5) * s)
        assert out[c, p, 2 * i + 1] == np.float32((np.float32(r8[2 * w + 1]) + 0.5) * s)
    # UWB: blocks of 2048 samples per pol, int16 offset binary, no scale
    raw = synth.uwb_bytes(4096, 2, seed=3)
    out = oracle.unpack_uwb(raw.view(np.int16), 4096, 2)
    r16 = (raw.view(np.uint16) ^ np.uint16(0x8000)).view(np.int16)
    for (p, i, d) in [(0, 0, 0), (1, 2049, 1), (0, 4095, 0)]:
        assert out[0, p, 2 * i + d] == float(r16[(i // 2048 * 2 + p) * 4096 + 2 * (i % 2048) + d])


@pytest.mark.parametrize("args,expect", [
    # SURVEY Appendix B (computed independently with the reference's formulas)
    ((1382, -400, 67.99, 1, 256, True), (8192, 457, 459)),
    ((1400, 128, 50, 1, 4096, True), (8, 1, 1)),
    ((1284, 856, 500, 1024, 1024, False), (65536, 2536, 2543)),
    ((768, 128, 67.99, 1, 128, False), (16384, 886, 890)),
    ((1536, 128, 67.99, 1, 128, False), (2048, 98, 98)),
    ((2304, 128, 67.99, 1, 128, False), (512, 28, 28)),
    ((3968, 128, 67.99, 1, 128, False), (64, 6, 6)),
])
def test_dedispersion_sizes(oracle, args, expect):
    d, _ = oracle.dedispersion(*args, build=False)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == expect


def test_dedispersion_override_and_threshold(oracle):
    d, _ = oracle.dedispersion(12500, 400, 1500, 1, 1, False, frequency_resolution=4194304, build=False)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == (4194304, 534848, 588748)
    with pytest.raises(ValueError):     # 400 MHz at L band, DM 1500: above the 16 Mi-sample threshold
        oracle.dedispersion(1400, 400, 1500, 1, 1, False, build=False)
    with pytest.raises(ValueError):     # -x smaller than the minimum ndat (Response::check_ndat)
        oracle.dedispersion(1382, -400, 67.99, 1, 256, True, frequency_resolution=512, build=False)


def test_optimal_fft_length(oracle):
    assert oracle.optimal_fft_length(916) == 8192
    assert oracle.optimal_fft_length(2) == 8
    assert oracle.optimal_fft_length(1776) == 16384
    assert oracle.optimal_fft_length(0) == -1


def test_chirp_is_unit_modulus_and_matches_formula(oracle):
    d, H = oracle.dedispersion(1382, -400, 67.99, 1, 256, True)
    assert H[0, 0] == 0
    a = np.abs(H.ravel()[1:])
    assert np.abs(a - 1).max() < 2e-7
    # independent double-precision evaluation of the phase (Dedispersion.C:534-545)
    c, j = 17, 1234
    bw, cf, nchan, F = -400.0, 1382.0, 256, d.ndat
    chanwidth = bw / nchan
    fc = cf - bw / 2 + chanwidth / 2 + c * chanwidth
    f = j * chanwidth / F - chanwidth / 2
    phase = np.float32(-(-1.0) * 2 * np.pi * (1e6 * 67.99 / 2.41e-4) / fc ** 2 * f * f / (fc + f))
    assert H[c, j] == np.complex64(complex(np.cos(phase), np.sin(phase))) or \
        abs(H[c, j] - complex(np.cos(np.float64(phase)), np.sin(np.float64(phase)))) < 2e-7


def test_complex_input_swaps_halves(oracle):
    # Response::match: single-channel complex input swaps the whole band (Response.C:138-148)
    _, Hr = oracle.dedispersion(1400, 64, 1, 1, 8, True, frequency_resolution=256, dual_sideband=False)
    _, Hc = oracle.dedispersion(1400, 64, 1, 1, 8, False, frequency_resolution=256)
    flat_r, flat_c = Hr.ravel().copy(), Hc.ravel()
    n = flat_r.size
    swapped = np.concatenate([flat_r[n // 2:], flat_r[:n // 2]])
    # DC zap happens before (bin 0 of the unswapped band) and after the swap (bin 0 again)
    swapped[0] = 0
    assert np.array_equal(swapped, flat_c)


def test_filterbank_pure_tone_lands_in_one_channel(oracle):
    # a real tone at the centre of output channel c0 -> all power in channel c0 (no response)
    C, F = 16, 32
    f = oracle.fb_sizes(True, 1, 1, C, F, 0, 0)
    n = 4 * f.nsamp_step
    c0 = 5
    k = c0 * F + F // 2                    # bin of the 2*C*F-point real FFT... tone frequency k/(2CF)
    t = np.arange(n)
    x = np.cos(2 * np.pi * k * t / (2 * C * F)).astype(np.float32)[None, None, :]
    y = oracle.filterbank(f, x, None)
    p = (np.abs(y[:, 0, :]) ** 2).sum(axis=1)
    assert p.argmax() == c0 and p[c0] / p.sum() > 0.999


def test_filterbank_impulse_reproduces_chirp(oracle):
    # an impulse at sample 0 of a part: spectrum = 1 -> per-channel inverse FFT of H
    C, F, npos, nneg = 4, 64, 5, 6
    f = oracle.fb_sizes(False, 1, 1, C, F, npos, nneg)
    rng = np.random.default_rng(4)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C, F))).astype(np.complex64)
    x = np.zeros((1, 1, 2 * f.nsamp_fft), np.float32)
    x[0, 0, 0] = 1.0
    y = oracle.filterbank(f, x, H)
    for c in range(C):
        ref = np.fft.ifft(H[c].astype(np.complex128)) * F
        assert np.abs(y[c, 0, :f.nkeep] - ref[npos:npos + f.nkeep]).max() < 1e-5


def test_convolution_matches_numpy_overlap_save(oracle):
    # independent float64 re-derivation of the overlap-save bookkeeping (Convolution.C:389-458)
    F, npos, nneg, npart = 512, 20, 23, 4
    rng = np.random.default_rng(6)
    c = oracle.conv_sizes(False, 2, 1, F, npos, nneg)
    n = npart * c.nsamp_step + c.nsamp_overlap
    x = (rng.standard_normal((2, 1, n)) + 1j * rng.standard_normal((2, 1, n))).astype(np.complex64)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (2, F))).astype(np.complex64)
    y = oracle.convolution(c, x.view(np.float32), H)
    nkeep = F - npos - nneg
    for ch in range(2):
        for r in range(npart):
            seg = x[ch, 0, r * c.nsamp_step: r * c.nsamp_step + F].astype(np.complex128)
            ref = (np.fft.ifft(np.fft.fft(seg) * H[ch]) * F)[npos:npos + nkeep]
            got = y[ch, 0, r * nkeep:(r + 1) * nkeep]
            assert np.abs(got - ref).max() / np.sqrt(np.mean(np.abs(ref) ** 2)) < 2e-6


def test_convolution_undoes_dispersion(oracle):
    # disperse a narrow pulse with the analytic chirp evaluated on the full-length grid; the
    # oracle's Dedispersion response must bring it back to one sample at the right place
    F, bw, cf, dm = 8192, 16.0, 1400.0, 1.0
    d, H = oracle.dedispersion(cf, bw, dm, 1, 1, False, frequency_resolution=F)
    c = oracle.conv_sizes(False, 1, 1, F, d.impulse_pos, d.impulse_neg)
    npart = 3
    n = npart * c.nsamp_step + c.nsamp_overlap
    pulse = np.zeros(n, np.complex128)
    t0 = c.nsamp_step + 1000 + d.impulse_pos
    pulse[t0] = 1.0
    k = np.arange(n)
    f = np.where(k < n // 2, k / n, k / n - 1.0) * bw          # baseband frequency of FFT bin k (MHz)
    disp_per_mhz = 1e6 * dm / 2.41e-4
    phase = -2 * np.pi * disp_per_mhz / cf ** 2 * f * f / (cf + f)   # Dedispersion.C:534-545 (bw > 0)
    disp = np.fft.ifft(np.fft.fft(pulse) * np.exp(-1j * phase))      # the ISM applies the inverse
    x = np.ascontiguousarray(disp.astype(np.complex64)).view(np.float32)[None, None, :]
    y = oracle.convolution(c, x, H)[0, 0]
    peak = np.abs(y).argmax()
    assert peak == t0 - d.impulse_pos          # output sample m is input sample m + nfilt_pos
    assert np.abs(y[peak]) ** 2 / (np.abs(y) ** 2).sum() > 0.9


def test_detection_identities(oracle):
    rng = np.random.default_rng(5)
    v = (rng.standard_normal((2, 2, 100)) + 1j * rng.standard_normal((2, 2, 100))).astype(np.complex64)
    coh = oracle.detect("Coherence", 4, v).reshape(2, 100, 4)
    sto = oracle.detect("Stokes", 4, v).reshape(2, 100, 4)
    inten = oracle.detect("Intensity", 1, v)[:, 0]
    ppqq = oracle.detect("PPQQ", 1, v)
    assert np.array_equal(coh[..., 0], ppqq[:, 0]) and np.array_equal(coh[..., 1], ppqq[:, 1])
    assert np.array_equal(sto[..., 0], coh[..., 0] + coh[..., 1])      # I = pp + qq
    assert np.array_equal(sto[..., 1], coh[..., 0] - coh[..., 1])      # Q = pp - qq
    assert np.array_equal(sto[..., 2], 2 * coh[..., 2]) and np.array_equal(sto[..., 3], 2 * coh[..., 3])
    assert np.array_equal(inten, sto[..., 0])
    # layouts (Detection::get_result_pointers): ndim 2 -> (pp,qq) plane and (Re,Im) plane; ndim 1 -> 4 planes
    c2 = oracle.detect("Coherence", 2, v)
    c1 = oracle.detect("Coherence", 1, v)
    assert np.array_equal(c2[:, 0].reshape(2, 100, 2), coh[..., :2]) and np.array_equal(c2[:, 1].reshape(2, 100, 2), coh[..., 2:])
    for q in range(4):
        assert np.array_equal(c1[:, q], coh[..., q])
    p, q = v[:, 0], v[:, 1]
    assert np.allclose(coh[..., 2] + 1j * coh[..., 3], np.conj(p) * q, rtol=1e-5, atol=1e-6)


def test_fold_boxcar_and_hits(oracle):
    nbin, ndat = 64, 64000
    period = 1000.0                                   # samples
    phi0, pps = 0.125, 1.0 / period
    binplan, hits, nfold, phi_end = oracle.fold_plan(phi0, pps, nbin, ndat)
    assert nfold == ndat and hits.sum() == ndat
    assert hits.min() >= 64 * 15 and hits.max() <= 64 * 16      # 15.6 samples per bin and turn, 64 turns
    phase = (phi0 + pps * np.arange(ndat)) % 1.0
    x = ((phase >= 0.25) & (phase < 0.30)).astype(np.float32)[None, None, :]
    prof = oracle.fold(x, 1, binplan, nbin)[0, 0]
    on = np.flatnonzero(prof)
    assert on.min() >= int(0.25 * nbin) - 1 and on.max() <= int(0.30 * nbin) + 1
    assert prof.sum() == x.sum()


def test_polyco_fixture(oracle):
    import workloads as W
    pc = oracle.polyco_parse(W.polyco_text())
    assert pc.tmid_day == 55299 and pc.ncoef == 15 and pc.f0 == pytest.approx(11.1946499395)
    assert pc.tmid_sec == pytest.approx(0.1041666666 * 86400, abs=1e-3)
    # at TMID the phase is RPHASE + c0 and the frequency F0 + c1/60
    sec = int(pc.tmid_sec)
    ph, turns = oracle.polyco_phase(pc, 55299, sec, pc.tmid_sec - sec)
    assert turns == pytest.approx(3616377136.0, abs=1) and abs(ph - 0.814839) < 1e-5
    f = oracle.polyco_frequency(pc, 55299, sec, pc.tmid_sec - sec)
    assert f == pytest.approx(pc.f0 + pc.coef[1] / 60.0, rel=1e-12)
    # phase advances by ~f per second
    ph2, t2 = oracle.polyco_phase(pc, 55299, sec + 10, pc.tmid_sec - sec)
    assert (t2 + ph2) - (turns + ph) == pytest.approx(10 * f, abs=1e-4)


def test_pipeline_threads_equal_serial(oracle):
    # dspsr -t P (MultiThread): P pipelines over time blocks + combine == one pipeline, up to
    # float reassociation of the final sum
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, 8, 64, 5, 6)
    nblock, npart = 4, 3
    ndat = (nblock * npart * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=9)
    H = np.exp(1j * np.random.default_rng(9).uniform(-3, 3, (8, 64))).astype(np.complex64)
    p = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, 32)
    phi = [0.1 * b for b in range(nblock)]
    pps = [1 / 97.0] * nblock
    a, ha = oracle.pipe_run(p, raw, nblock, npart, phi, pps, nthread=1)
    b, hb = oracle.pipe_run(p, raw, nblock, npart, phi, pps, nthread=3)
    assert np.array_equal(ha, hb) and ha.sum() == nblock * npart * f.nkeep
    assert synth.relerr(b, a) < 1e-6


# ----------------------------------------------------------------------------- a6 two-bit excision
def test_twobit_levels_and_limits(oracle):
    """JA98 section 6 levels are the conditional rms of a unit Gaussian below / above the threshold, so the
    digitised power equals the undigitised power (Phi lo^2 + (1-Phi) hi^2 = 1); checked against numerical
    integration.  ExcisionUnpacker::set_limits for the defaults (512 samples, threshold 0.9674, cutoff
    10 sigma) gives [234, 447]."""
    for phi in (0.2, 0.5, 0.6667, 0.9):
        lo, hi = oracle.ja98_levels(phi)
        assert phi * lo * lo + (1 - phi) * hi * hi == pytest.approx(1.0, abs=1e-12)
        assert 0 < lo < 1 < hi
    import math
    from scipy import integrate, stats
    u = 0.9674
    lo, hi = oracle.ja98_levels(math.erf(u / math.sqrt(2)))
    p_lo = integrate.quad(lambda x: x * x * stats.norm.pdf(x), -u, u)[0] / (stats.norm.cdf(u) - stats.norm.cdf(-u))
    p_hi = 2 * integrate.quad(lambda x: x * x * stats.norm.pdf(x), u, 12)[0] / (2 * stats.norm.cdf(-u))
    assert lo == pytest.approx(math.sqrt(p_lo), rel=1e-9) and hi == pytest.approx(math.sqrt(p_hi), rel=1e-9)
    t = oracle.TwoBit()
    assert (t.nlow_min, t.nlow_max) == (234, 447)
    t0 = oracle.TwoBit(cutoff_sigma=0.0)
    assert (t0.nlow_min, t0.nlow_max) == (0, 512)
    # rows are ordered: more low states = more input power below threshold = smaller sigma estimate
    lo_a, hi_a = t.levels(300)
    lo_b, hi_b = t.levels(400)
    assert lo_a < lo_b and hi_a < hi_b


def test_twobit_unpack_semantics(oracle):
    """Window statistics select the level row; all-zero windows and windows whose nlow is outside the
    limits are zeroed and flagged (excision_unpack.h:79-97), per digitizer; weights are masked over pols."""
    ndat, npol = 512 * 6, 2
    raw = synth.twobit_bytes(ndat, npol, seed=3).reshape(-1, npol).copy()
    raw[128 * 1:128 * 2, 0] = 0x00          # pol 0, window 1: all-zero bytes -> bad
    raw[128 * 3:128 * 4, 1] = 0xFF          # pol 1, window 3: every sample +hi -> nlow = 0 < nlow_min
    raw[128 * 4:128 * 5, 0] = 0x66          # pol 0, window 4: every sample low -> nlow = 512 > nlow_max
    t = oracle.TwoBit()
    out, w = t.unpack(raw.reshape(-1), ndat, npol)
    assert np.array_equal(w[0], [1, 0, 1, 0, 0, 1]) and np.array_equal(w[0], w[1])
    assert not out[0, 0, 512:1024].any() and out[0, 1, 512:1024].any()       # only the bad digitizer is zeroed
    assert not out[0, 1, 1536:2048].any() and out[0, 0, 1536:2048].any()
    assert not out[0, 0, 2048:2560].any()
    # a good window: four distinct values -hi, -lo, lo, hi of the row selected by its own nlow
    win = out[0, 0, :512]
    codes = np.array([(b >> s) & 3 for b in raw[:128, 0] for s in (6, 4, 2, 0)])
    nlow = int(np.sum((codes == 1) | (codes == 2)))
    lo, hi = t.levels(nlow)
    want = np.array([-hi, -lo, lo, hi], np.float32)[codes]
    assert np.array_equal(win, want)
    # unit variance on average at nominal power
    assert np.var(out[0, 0, 2560:]) == pytest.approx(1.0, rel=0.15)


# ----------------------------------------------------------------------------- f1 digifil tail
def test_rescale_and_digitizer_known_answers(oracle):
    """dsp::Rescale: the first block is normalised with its own statistics (zero mean, unit variance per
    channel), later blocks with the statistics of the last completed interval; SigProcDigitizer 8 bit:
    0 -> 127.5 + 0.5 truncated = 128, +-6 sigma span the byte range, positive bandwidth flips the channel order."""
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((5, 1, 4000)) * np.arange(1, 6)[:, None, None] + 10).astype(np.float32)
    r = oracle.Rescale()
    y = r.transform(x)
    assert np.allclose(y.mean(axis=2), 0, atol=1e-5) and np.allclose(y.var(axis=2), 1, rtol=1e-4)
    y2 = r.transform(x + 1.0)                       # interval = first block: statistics of THIS block apply at its end
    assert np.allclose(y2.mean(axis=2), 0, atol=1e-5)
    r3 = oracle.Rescale(interval_samples=6000)
    a = r3.transform(x)                             # first call: own statistics
    b = r3.transform(x + 1.0)                       # samples 4000..5999 complete the interval inside this block
    assert np.allclose(a.mean(axis=2), 0, atol=1e-5)
    assert np.allclose(b[:, :, :2000].mean(axis=2), 1.0 / np.arange(1, 6)[:, None], rtol=0.1)   # old offsets still in force
    z = np.zeros((4, 1, 3), np.float32)
    z[:, 0, 1] = [6, -6, 3, -3]
    z[:, 0, 2] = [100, -100, 0, 0]
    d = oracle.sigproc_digitize(z, bandwidth=-1.0)
    assert d.shape == (3, 1, 4) and np.array_equal(d[0, 0], [128, 128, 128, 128])
    assert np.array_equal(d[1, 0], [255, 0, 191, 64]) and np.array_equal(d[2, 0], [255, 0, 128, 128])
    assert np.array_equal(oracle.sigproc_digitize(z, bandwidth=1.0)[1, 0], [64, 191, 0, 255])
    assert np.array_equal(oracle.sigproc_channel_sort(8, -1.0, swap=True), [4, 5, 6, 7, 0, 1, 2, 3])
