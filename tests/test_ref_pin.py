"""Pins the oracle to the real reference wherever reference code can be compiled here (oracle/ref.mk, outputs in
oracle/_ref/, sources read where they lie under /root/reference):
  libdspsr_refc.so       optimize_fft.c, cross_detect.c, stokes_detect.c, ascii_header.c            (rows a9, a12, header)
  libdspsr_refcxx.so     BitTable.C, TwoBitTable.C, TwoBitLookup.C, TwoBitFour.C, excision_unpack.h,
                         Dedispersion.C, Response.C, Shape.C                                        (a1, a6, a7, a8)
                         + the overlap-save loop nests of Filterbank.C:563-660 / Convolution.C:389-458 compiled from
                           the reference's text around the reference's Response::operate              (a10, a11)
  libdspsr_reffmt.so     CASPSRUnpacker.C, MeerKATUnpacker.C, UWBUnpacker.C                          (a2, a4, a5)
  libdspsr_refbit.so     BitUnpacker.C, EightBitUnpacker.C                                           (a3)
  libdspsr_reffold.so    the weight / bin-plan / accumulation loops of Fold.C:687-716,744-787,835-873 (a13) and the bodies of
                         WeightedTimeSeries::convolve_weights / scrunch_weights (WeightedTimeSeries.C:584-696,705-774; f4)
  libdspsr_refsigproc.so filterbank_header.c, send_stuff.c                                           (f1)
What stays restated: the FFT primitive (FFTW inside PSRCHIVE) and the TEMPO polyco evaluation (PSRCHIVE) -- third-party
code that is not in the reference tree.

The libraries are built by __graft_entry__.build() whenever /root/reference exists and travel to the GPU box with the
snapshot; nothing here reads /root/reference at run time."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libdspsr_refc.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REFLIB)
    L.optimal_fft_length.restype = C.c_uint64
    L.optimal_fft_length.argtypes = [C.c_uint64, C.c_uint64, C.c_char]
    fp = C.POINTER(C.c_float)
    for name in ("cross_detect", "stokes_detect", "cross_detect_int", "stokes_detect_int"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [C.c_uint, fp, fp, fp, fp, fp, fp, C.c_uint]
    L.ascii_header_get.restype = C.c_int
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_optimal_fft_length_matches_reference(ref, oracle):
    """optimize_fft.c:63-127 vs orc_optimal_fft_length and the library's b200_optimal_fft_length."""
    from dspsr_b200 import hostmath as HM
    rng = np.random.default_rng(5)
    cases = [1, 2, 3, 12, 916, 457 + 459, 2, 5079, 1123596, 1776, 1348, 196, 56] + [int(x) for x in rng.integers(1, 1 << 22, 200)]
    for nbad in cases:
        for nmax in (0, 1 << 24):
            want = ref.optimal_fft_length(nbad, nmax, b"\0")
            assert oracle.optimal_fft_length(nbad, nmax) == want, (nbad, nmax)
            assert HM.optimal_fft_length(nbad, nmax) == want, (nbad, nmax)
    # the frequency resolutions SURVEY Appendix B quotes come out of the reference's own function
    assert ref.optimal_fft_length(457 + 459, 0, b"\0") == 8192
    assert ref.optimal_fft_length(2536 + 2543, 0, b"\0") == 65536


@pytest.mark.parametrize("state,fn", [("Coherence", "cross_detect"), ("Stokes", "stokes_detect")])
def test_detect_products_match_reference(ref, oracle, state, fn):
    """cross_detect.ic:25-41 / stokes_detect.ic:21-44 vs oracle.detect (ndim 1: four planes, span 1;
    ndim 4: one plane of 4-vectors, span 4 -- Detection.C:423-474), bit for bit."""
    rng = np.random.default_rng(11)
    ndat = 4099
    v = (rng.standard_normal((1, 2, ndat)) + 1j * rng.standard_normal((1, 2, ndat))).astype(np.complex64)
    v *= np.float32(37.5)
    p = np.ascontiguousarray(v[0, 0]).view(np.float32)
    q = np.ascontiguousarray(v[0, 1]).view(np.float32)
    # ndim 1
    r = [np.zeros(ndat, np.float32) for _ in range(4)]
    getattr(ref, fn)(ndat, _fp(p), _fp(q), _fp(r[0]), _fp(r[1]), _fp(r[2]), _fp(r[3]), 1)
    o = oracle.detect(state, 1, v)
    for i in range(4):
        assert np.array_equal(o[0, i], r[i]), (state, i)
    # ndim 4
    buf = np.zeros(4 * ndat, np.float32)
    getattr(ref, fn)(ndat, _fp(p), _fp(q), _fp(buf[0:]), _fp(buf[1:]), _fp(buf[2:]), _fp(buf[3:]), 4)
    o4 = oracle.detect(state, 4, v)
    assert np.array_equal(o4[0, 0], buf)


def test_dada_header_keys_parse_like_reference(ref):
    """ascii_header.c ascii_header_get on the cfg1 header the bench writes."""
    import workloads as W
    hdr = W.dada_header(W.CFG1).encode()
    assert len(hdr) == 4096
    buf = C.create_string_buffer(hdr, 4096)
    got = {}
    for key, fmt, ctype in (("HDR_SIZE", b"%d", C.c_int), ("NBIT", b"%d", C.c_int), ("NDIM", b"%d", C.c_int),
                            ("NPOL", b"%d", C.c_int), ("NCHAN", b"%d", C.c_int), ("FREQ", b"%lf", C.c_double),
                            ("BW", b"%lf", C.c_double), ("TSAMP", b"%lf", C.c_double)):
        val = ctype()
        assert ref.ascii_header_get(buf, key.encode(), fmt, C.byref(val)) == 1, key
        got[key] = val.value
    s = C.create_string_buffer(64)
    assert ref.ascii_header_get(buf, b"INSTRUMENT", b"%s", s) == 1 and s.value == b"CASPSR"
    assert ref.ascii_header_get(buf, b"UTC_START", b"%s", s) == 1 and s.value == b"2010-04-13-02:05:45"
    c = W.CFG1
    assert got == {"HDR_SIZE": 4096, "NBIT": 8, "NDIM": 1, "NPOL": 2, "NCHAN": 1, "FREQ": c["freq"], "BW": c["bw"],
                   "TSAMP": c["tsamp_us"]}
    # and our own parser reads back what the reference's parser reads
    mine = W.parse_dada_header(hdr)
    assert mine["INSTRUMENT"] == "CASPSR" and mine["UTC_START"] == "2010-04-13-02:05:45"
    for k, v in got.items():
        assert float(mine[k]) == v, k


# ---------------------------------------------------------------------------------------------------------------
# C++ half of the pin (oracle/_ref/libdspsr_refcxx.so): the reference's BitTable.C, TwoBitTable.C, TwoBitLookup.C,
# TwoBitFour.C, dsp/TwoBitFour.h, dsp/excision_unpack.h, Dedispersion.C, Response.C, Shape.C compiled in place
# behind oracle/ref_shim/ (PSRCHIVE utility stand-ins; JenetAnderson98 / NormalDistribution numbers are the
# restated third-party part and are shared with the oracle).
# ---------------------------------------------------------------------------------------------------------------
REFCXX = os.path.join(ROOT, "oracle", "_ref", "libdspsr_refcxx.so")
needs_cxx = pytest.mark.skipif(not os.path.exists(REFCXX), reason="oracle/_ref C++ pin library not built")


@pytest.fixture(scope="module")
def refcxx(oracle):
    oracle.lib()                                   # liboracle.so first (JA98 numbers), found again through the rpath
    L = C.CDLL(REFCXX)
    L.ref_bittable_unique_values.restype = C.c_double
    L.ref_bittable_unique_values.argtypes = [C.c_uint, C.c_int, C.c_void_p]
    L.ref_bittable_generate.restype = C.c_double
    L.ref_bittable_generate.argtypes = [C.c_uint, C.c_int, C.c_void_p]
    L.ref_twobit_create.restype = C.c_void_p
    L.ref_twobit_create.argtypes = [C.c_int, C.c_double, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    L.ref_twobit_destroy.argtypes = [C.c_void_p]
    L.ref_twobit_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_twobit_unpack.restype = C.c_int
    L.ref_twobit_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint]
    L.ref_dedispersion.restype = C.c_int
    L.ref_dedispersion.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int,
                                   C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_void_p,
                                   C.c_uint64]
    L.ref_response_operate.restype = C.c_int
    L.ref_response_operate.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
    return L


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


@needs_cxx
def test_bittable_matches_reference(refcxx, oracle):
    """Row a1: BitTable.C:121-218 (generate_unique_values, generate, get_scale) vs the oracle and the product's
    host-built table, bit for bit -- the 8-bit TwosComplement table of CASPSR / MeerKAT and every other width."""
    from dspsr_b200 import hostmath as HM
    OFFSET, TWOS = 0, 1                                      # dsp::BitTable::Type
    for nbit in (2, 3, 4, 5, 6, 7, 8):
        for typ, twos in ((OFFSET, False), (TWOS, True)):
            want = np.zeros(1 << nbit, np.float32)
            scale = refcxx.ref_bittable_unique_values(nbit, typ, _vp(want))
            got, gscale = oracle.bittable_values(nbit, twos)
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (nbit, typ)
            assert scale == gscale, (nbit, typ)
    tab = np.zeros(256, np.float32)
    scale = refcxx.ref_bittable_generate(8, TWOS, _vp(tab))
    lut_o, scale_o = oracle.bittable8(True)
    lut_p, scale_p = HM.bittable8()
    assert np.array_equal(tab.view(np.uint32), lut_o.view(np.uint32)) and scale == scale_o
    assert np.array_equal(tab.view(np.uint32), np.asarray(lut_p).view(np.uint32)) and scale == scale_p
    # the sample the CASPSR unpacker would produce for byte 0x00 / 0x7f / 0x80 / 0xff, straight from the reference
    assert tab[0] > 0 and tab[0x7F] == tab.max() and tab[0x80] == tab.min() and tab[0xFF] < 0


# oracle table_type -> dsp::BitTable::Type
_TWOBIT_TYPES = {0: 0, 1: 2, 2: 1}       # OffsetBinary, SignMagnitude, TwosComplement


@needs_cxx
@pytest.mark.parametrize("table_type", [0, 1, 2])
@pytest.mark.parametrize("cutoff", [10.0, 3.0])
def test_twobit_tables_and_unpack_match_reference(refcxx, oracle, table_type, cutoff):
    """Row a6: TwoBitTable.C:42-75, TwoBitLookup.C:63-98, TwoBitFour.C:25-52, dsp/TwoBitFour.h:42-89 and the body of
    dsp/excision_unpack.h:21-106, all compiled from the reference, vs oracle.TwoBit (levels per low-state count,
    low-state counts per byte, unpacked floats and per-window weights incl. excised and all-zero windows)."""
    ndw = 512
    tb = oracle.TwoBit(table_type, 0.9674, cutoff, ndw)
    h = C.c_void_p(refcxx.ref_twobit_create(_TWOBIT_TYPES[table_type], 0.9674, tb.nlow_min, tb.nlow_max, ndw, 1))
    try:
        nrow = tb.nlow_max - tb.nlow_min + 1
        lookup = np.zeros((nrow, 256, 4), np.float32)
        nlow_lookup = np.zeros(256, np.int8)
        refcxx.ref_twobit_tables(h, _vp(lookup), _vp(nlow_lookup))
        for nlow in range(tb.nlow_min, tb.nlow_max + 1):
            lo, hi = tb.levels(nlow)
            mags = np.unique(np.abs(lookup[nlow - tb.nlow_min]))
            assert mags.size == 2 and mags[0] == np.float32(lo) and mags[1] == np.float32(hi), nlow
        # data: Gaussian noise digitised at the optimal threshold, plus windows that must be excised
        rng = np.random.default_rng(21 + table_type)
        npol, nwin = 2, 24
        ndat = nwin * ndw
        x = rng.standard_normal((npol, ndat))
        x[0, 3 * ndw:4 * ndw] *= 6.0            # too few low states
        x[1, 7 * ndw:8 * ndw] *= 0.05           # too many low states
        code = np.where(x < -0.9674, 0, np.where(x < 0, 1, np.where(x < 0.9674, 2, 3))).astype(np.uint8)
        if table_type == 1:      # SignMagnitude: lo, hi, -lo, -hi
            code = np.array([3, 2, 0, 1], np.uint8)[code]
        elif table_type == 2:    # TwosComplement: lo, hi, -hi, -lo
            code = np.array([2, 3, 0, 1], np.uint8)[code]
        c4 = code.reshape(npol, ndat // 4, 4)
        by = (c4[..., 0] << 6) | (c4[..., 1] << 4) | (c4[..., 2] << 2) | c4[..., 3]
        raw = np.ascontiguousarray(by.T).reshape(-1).astype(np.uint8)       # polarisations interleaved byte by byte
        raw[npol * (11 * ndw // 4):npol * (12 * ndw // 4)] = 0                # an all-zero window (unpack.bad)
        want = np.zeros((npol, ndat), np.float32)
        wref = np.ones((npol, nwin), np.uint32)
        assert refcxx.ref_twobit_unpack(h, _vp(raw), ndat, npol, _vp(want), ndat, _vp(wref), nwin) == 0
        got, w = tb.unpack(raw, ndat, npol)
        assert np.array_equal(want.view(np.uint32), got[0].view(np.uint32))
        # WeightedTimeSeries::mask_weights (a window flagged in one polarisation is flagged in all)
        masked = np.broadcast_to(wref.min(axis=0, keepdims=True), wref.shape)
        assert np.array_equal(masked, w)
        assert 3 <= int((masked[0] == 0).sum()) < nwin
        # per-byte low-state counts: oracle's prepare() must count what the reference's nlow_build tabulated
        lo_codes = {0: (1, 2), 1: (0, 2), 2: (0, 3)}[table_type]
        cnt = np.array([sum(((b >> s) & 3) in lo_codes for s in (6, 4, 2, 0)) for b in range(256)], np.int8)
        assert np.array_equal(cnt, nlow_lookup)
    finally:
        refcxx.ref_twobit_destroy(h)


_DEDISP_CASES = [
    # cf, bw, dm, input_nchan, nchan, input_real, freq_res, kwargs                      
    (1382.0, -400.0, 67.99, 1, 256, True, 0, {}),                                       # cfg1
    (1382.0, -400.0, 100.0, 1, 256, True, 0, {}),                                       # cfg1' (bench.csh DM 100)
    (1400.0, 128.0, 50.0, 1, 4096, True, 0, {}),                                        # cfg2
    (1284.0, 856.0, 500.0, 1024, 1024, False, 8192, {}),                                # cfg3 with -x 8192 (H is 512 MiB at the optimal 65536)
    (12500.0, 400.0, 1500.0, 1, 1, False, 4194304, {}),                                 # cfg4
    (768.0, 128.0, 67.99, 1, 128, False, 0, {}),                                        # cfg5 sb0
    (3968.0, 128.0, 67.99, 1, 128, False, 0, {}),                                       # cfg5 sb25
    (1400.0, -64.0, 30.0, 1, 16, False, 0, {}),                                         # lower sideband, complex input
    (1400.0, 64.0, 30.0, 8, 8, False, 0, {"swap": True}),                               # multi-channel + whole-band swap
    (1400.0, 64.0, 30.0, 8, 32, False, 0, {}),                                          # multi-channel in, more channels out
    (1400.0, 64.0, 30.0, 4, 4, False, 0, {"dual_sideband": False}),                     # single-sideband channels
    (1400.0, 64.0, 30.0, 1, 32, True, 0, {"dc_centred": True}),                         # bin-centred spectrum
    (610.0, -16.0, 26.8, 1, 1, True, 0, {}),                                            # plain convolution, real input
]


@needs_cxx
@pytest.mark.parametrize("case", _DEDISP_CASES, ids=lambda c: "cf%g_bw%g_dm%g_%dto%d" % c[:5])
def test_dedispersion_matches_reference(refcxx, oracle, case):
    """Rows a7 + a8: Dedispersion::prepare / smearing_samples / build / match (Dedispersion.C:167-556) with
    Response::match / doswap / set_optimal_ndat (Response.C:132-181,259-311,649-700) and Shape.C, compiled from the
    reference, vs the oracle AND the product's host maths: impulse_pos/neg, chosen frequency resolution, and every
    float of the matched response H, bit for bit."""
    from dspsr_b200 import hostmath as HM
    cf, bw, dm, nin, nout, real, fres, kw = case
    dual = kw.get("dual_sideband", not real)
    pos, neg, ndat = C.c_uint(0), C.c_uint(0), C.c_uint(0)
    rc = refcxx.ref_dedispersion(cf, bw, dm, nin, nout, int(dual), int(kw.get("dc_centred", False)),
                                 int(kw.get("swap", False)), fres, C.byref(pos), C.byref(neg), C.byref(ndat), None, 0)
    assert rc == 0
    d_o, H_o = oracle.dedispersion(cf, bw, dm, nin, nout, real, fres, **kw)
    d_p, H_p = HM.dedispersion(cf, bw, dm, nin, nout, real, fres, **kw)
    assert (pos.value, neg.value, ndat.value) == (d_o.impulse_pos, d_o.impulse_neg, d_o.ndat)
    assert (pos.value, neg.value, ndat.value) == (d_p.impulse_pos, d_p.impulse_neg, d_p.ndat)
    want = np.zeros((nout, ndat.value), np.complex64)
    rc = refcxx.ref_dedispersion(cf, bw, dm, nin, nout, int(dual), int(kw.get("dc_centred", False)),
                                 int(kw.get("swap", False)), fres, C.byref(pos), C.byref(neg), C.byref(ndat), _vp(want),
                                 want.size * 2)
    assert rc == 0
    assert np.array_equal(want.view(np.uint32), H_o.view(np.uint32))
    assert np.array_equal(want.view(np.uint32), np.asarray(H_p).view(np.uint32))


@needs_cxx
def test_dedispersion_refuses_like_reference(refcxx, oracle):
    """cfg4 as BASELINE.json words it (400 MHz at L-band, DM 1500) is refused by the reference itself
    (Dedispersion.C:214-233, smearing samples above the 16 Mi threshold); oracle and product refuse too."""
    from dspsr_b200 import hostmath as HM
    pos, neg, ndat = C.c_uint(0), C.c_uint(0), C.c_uint(0)
    assert refcxx.ref_dedispersion(1400.0, 400.0, 1500.0, 1, 1, 1, 0, 0, 0, C.byref(pos), C.byref(neg), C.byref(ndat), None, 0) == -1
    with pytest.raises(Exception):
        oracle.dedispersion(1400.0, 400.0, 1500.0, 1, 1, False)
    with pytest.raises(Exception):
        HM.dedispersion(1400.0, 400.0, 1500.0, 1, 1, False)


@needs_cxx
def test_response_operate_matches_reference(refcxx, oracle):
    """Row a8: Response::operate (Response.C:385-444) -- spectrum *= H in float, the reference's operation order --
    vs the oracle's filterbank, which applies the same multiply between its forward and inverse transforms:
    with a unit impulse response H = 1 both reduce to identity, so the comparison isolates the multiply by running
    the reference's operate on the oracle's own forward spectrum and inverse-transforming with the oracle."""
    rng = np.random.default_rng(3)
    n = 4096
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, n)).astype(np.complex64)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    spec = oracle.fcc1d(x)
    want = spec.copy()
    assert refcxx.ref_response_operate(_vp(H), n, _vp(want)) == 0
    # the oracle's restatement of the same loop (orc response_operate, used inside orc_filterbank_parts)
    got = oracle.response_operate(spec, H)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@needs_cxx
def test_cfg3_optimal_resolution_from_reference(refcxx, oracle):
    """the reference itself picks F = 65536 with M = 2536 + 2543 (no response built: 512 MiB)."""
    pos, neg, ndat = C.c_uint(0), C.c_uint(0), C.c_uint(0)
    # prepare only: Dedispersion::match would build; ask for the sizes through a 1-channel-wide equivalent band
    d_o, _ = oracle.dedispersion(1284.0, 856.0, 500.0, 1024, 1024, False, build=False)
    assert (d_o.impulse_pos, d_o.impulse_neg, d_o.ndat) == (2536, 2543, 65536)
    # the reference on the lowest-frequency channel alone (same smearing: channel 0 of 1024 is what sets M)
    cw = 856.0 / 1024
    rc = refcxx.ref_dedispersion(1284.0 - 428.0 + cw / 2, cw, 500.0, 1, 1, 1, 0, 0, 0, C.byref(pos), C.byref(neg),
                                 C.byref(ndat), None, 0)
    assert rc == 0 and (pos.value, neg.value, ndat.value) == (2536, 2543, 65536)


@needs_cxx
def test_twobit_all_low_row_deviation(refcxx, oracle):
    """with cutoff_sigma = 0 the reference's level row for nlow = ndat is NaN (TwoBitLookup.C:83 tests
    the wrong variable); oracle and product clamp that row to ndat-1 instead.  Every other row still equals the
    reference bit for bit."""
    from dspsr_b200 import hostmath as HM
    ndw = 128
    tb = oracle.TwoBit(0, 0.9674, 0.0, ndw)
    assert (tb.nlow_min, tb.nlow_max) == (0, ndw)
    h = C.c_void_p(refcxx.ref_twobit_create(0, 0.9674, 0, ndw, ndw, 1))
    try:
        lookup = np.zeros((ndw + 1, 256, 4), np.float32)
        nl = np.zeros(256, np.int8)
        refcxx.ref_twobit_tables(h, _vp(lookup), _vp(nl))
        assert np.isnan(lookup[ndw]).any()                       # the reference's own all-low row
        for nlow in range(0, ndw):
            lo, hi = tb.levels(nlow)
            mags = np.unique(np.abs(lookup[nlow]))
            assert mags[0] == np.float32(lo) and mags[-1] == np.float32(hi)
        lo, hi = tb.levels(ndw)
        assert np.isfinite(lo) and np.isfinite(hi) and (lo, hi) == tb.levels(ndw - 1)
    finally:
        refcxx.ref_twobit_destroy(h)
    from dspsr_b200 import engine as E
    tp = E.make_twobit_desc(2, 0.9674, 0.0, 0, ndw)
    n = tp.nlow_max - tp.nlow_min + 1
    assert (tp.nlow_min, tp.nlow_max) == (0, ndw)
    lo_p, hi_p = np.array(tp.lo[:n]), np.array(tp.hi[:n])
    assert np.isfinite(lo_p).all() and np.isfinite(hi_p).all()
    assert (np.float32(lo_p[-1]), np.float32(hi_p[-1])) == (np.float32(lo), np.float32(hi))


# ---------------------------------------------------------------------------------------------------------------
# Format unpackers (oracle/_ref/libdspsr_reffmt.so): the reference's CASPSRUnpacker.C, MeerKATUnpacker.C and
# UWBUnpacker.C compiled in place; their unpack() bodies run on the same seeded bytes as the oracle.
# ---------------------------------------------------------------------------------------------------------------
REFFMT = os.path.join(ROOT, "oracle", "_ref", "libdspsr_reffmt.so")
needs_fmt = pytest.mark.skipif(not os.path.exists(REFFMT), reason="oracle/_ref format pin library not built")


@pytest.fixture(scope="module")
def reffmt(oracle):
    oracle.lib()
    L = C.CDLL(REFFMT)
    L.ref_unpack_caspsr.restype = C.c_int
    L.ref_unpack_caspsr.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    L.ref_unpack_meerkat.restype = C.c_int
    L.ref_unpack_meerkat.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]
    L.ref_unpack_uwb.restype = C.c_int
    L.ref_unpack_uwb.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_uint64]
    return L


@needs_fmt
def test_caspsr_unpack_matches_reference(reffmt, oracle):
    """Row a2: CASPSRUnpacker.C:132-187 (4 samples pol0, 4 samples pol1 per 8 bytes, BitTable look-up), every byte value."""
    rng = np.random.default_rng(31)
    ndat = 1024 * 5
    raw = rng.integers(0, 256, 2 * ndat, dtype=np.uint8)
    raw[:256] = np.arange(256, dtype=np.uint8)
    want = np.zeros((1, 2, ndat), np.float32)
    assert reffmt.ref_unpack_caspsr(_vp(raw), ndat, _vp(want), ndat) == 0
    lut, _ = oracle.bittable8(True)
    got = oracle.unpack_caspsr(raw, ndat, lut)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@needs_fmt
@pytest.mark.parametrize("npol,swap", [(2, 1), (2, 2), (1, 1)])
def test_meerkat_unpack_matches_reference(reffmt, oracle, npol, swap):
    """Row a4: MeerKATUnpacker.C:196-229 (heaps of 256 samples, [heap][pol][chan][256 x (re, im)], (float(x) + 0.5) * scale,
    MKBFRo sample swap) and the scale the reference derives from its BitTable."""
    rng = np.random.default_rng(32)
    nchan, nheap = 12, 3
    ndat = 256 * nheap
    raw = rng.integers(0, 256, ndat * nchan * npol * 2, dtype=np.uint8)
    raw[:256] = np.arange(256, dtype=np.uint8)
    want = np.zeros((nchan, npol, ndat * 2), np.float32)
    scale = C.c_float(0)
    assert reffmt.ref_unpack_meerkat(_vp(raw), ndat, nchan, npol, swap, _vp(want), ndat * 2, C.byref(scale)) == 0
    _, tscale = oracle.bittable8(True)
    assert np.float32(tscale) == np.float32(scale.value)
    got = oracle.unpack_meerkat(raw.view(np.int8), ndat, nchan, npol, np.float32(tscale), swap)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@needs_fmt
@pytest.mark.parametrize("npol", [2, 1])
def test_uwb_unpack_matches_reference(reffmt, oracle, npol):
    """Row a5: UWBUnpacker.C:177-218 (blocks of 2048 samples per polarisation, offset-binary 16 bit, no scaling)."""
    rng = np.random.default_rng(33)
    ndat = 2048 * 3
    raw = rng.integers(0, 65536, ndat * npol * 2, dtype=np.uint16)
    raw[:8] = [0, 1, 0x7FFF, 0x8000, 0x8001, 0xFFFF, 0x1234, 0xFEDC]
    want = np.zeros((1, npol, ndat * 2), np.float32)
    assert reffmt.ref_unpack_uwb(_vp(raw), ndat, npol, _vp(want), ndat * 2) == 0
    got = oracle.unpack_uwb(raw.view(np.int16), ndat, npol)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


# ------------------------------------------------------------------------------------------------------------
# SIGPROC filterbank header (f1): the reference's own filterbank_header.c + send_stuff.c, compiled in place
# ------------------------------------------------------------------------------------------------------------
REFSIG = os.path.join(ROOT, "oracle", "_ref", "libdspsr_refsigproc.so")


@pytest.mark.skipif(not os.path.exists(REFSIG), reason="oracle/_ref/libdspsr_refsigproc.so not built")
@pytest.mark.parametrize("bw,nchan,npol,telescope,machine", [(128.0, 4096, 1, "PKS", "CPSR2"), (-400.0, 256, 4, "Parkes", "BPSR"),
                                                              (64.0, 16, 2, "GBT", "COBALT"), (856.0, 1024, 1, "MeerKAT", "MKBF")])
def test_sigproc_header_bytes_equal_reference_writer(tmp_path, bw, nchan, npol, telescope, machine):
    """b200_sigproc_header_write against filterbank_header() (filterbank_header.c:44-100) fed with the globals that
    SigProcObservation::unload_global (SigProcObservation.C:228-275) sets for the same observation."""
    from dspsr_b200 import _lib as L
    from dspsr_b200 import phaseseries as P
    lib = L.load()
    start = (55299, 7545, 0.25)
    rate = 128e6 * 8 / 65536
    det = P.observation(nchan, npol, 1, rate, start, centre_frequency=1400.0, bandwidth=bw, state=L.INTENSITY,
                        source="J0835-4510", telescope=telescope, machine=machine)
    h = L.SigprocHeader()
    L.check(lib.b200_sigproc_header_from_observation(C.byref(det), 8, C.byref(h)))
    buf = C.create_string_buffer(1024)
    n = lib.b200_sigproc_header_write(C.byref(h), buf, 1024)
    assert n > 0
    mine = buf.raw[:n]
    # unload_global, by hand, into the reference's globals
    R = C.CDLL(REFSIG)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]

    def setstr(name, val, size=80):
        a = (C.c_char * size).in_dll(R, name)
        a.value = val

    def setv(ctype, name, val):
        ctype.in_dll(R, name).value = val
    tel = {"PKS": 4, "Parkes": 4, "GBT": 6, "MeerKAT": 0}[telescope]
    mach = {"BPSR": 10, "SCAMP": 6, "COBALT": 11}.get(machine, 0)
    setstr("inpfile", b"unknown")
    setstr("source_name", b"J0835-4510")
    setv(C.c_int, "machine_id", mach)
    setv(C.c_int, "telescope_id", tel)
    obw = -abs(bw)                                                    # SigProcDigitizer.C:84
    setv(C.c_double, "fch1", 1400.0 - 0.5 * obw + 0.5 * obw / nchan)  # Observation.C:438-451, channel 0, not dc-centred
    setv(C.c_double, "foff", obw / nchan)
    setv(C.c_int, "nchans", nchan)
    setv(C.c_int, "nifs", npol)
    setv(C.c_int, "obits", 8)
    setv(C.c_double, "tsamp", 1.0 / rate)
    setv(C.c_double, "tstart", start[0] + (start[1] + start[2]) / 86400.0)
    for k in ("src_raj", "src_dej", "az_start", "za_start"):
        setv(C.c_double, k, 0.0)
    ifs = (C.c_char * 8).in_dll(R, "ifstream")
    ifs.value = b"Y" * npol
    path = tmp_path / "ref.hdr"
    fp = libc.fopen(str(path).encode(), b"wb")
    R.filterbank_header.argtypes = [C.c_void_p]
    R.filterbank_header(fp)
    libc.fclose(fp)
    ref = open(path, "rb").read()
    assert mine == ref
    assert mine.startswith(b"\x0c\x00\x00\x00HEADER_START") and mine.endswith(b"HEADER_END")
    # the file writer: header + bytes, then appended bytes
    data = np.arange(3 * nchan * npol, dtype=np.uint8)
    fil = tmp_path / "out.fil"
    L.check(lib.b200_sigproc_file_write(str(fil).encode(), C.byref(h), data.ctypes.data, data.size, 0))
    L.check(lib.b200_sigproc_file_write(str(fil).encode(), None, data.ctypes.data, data.size, 1))
    blob = open(fil, "rb").read()
    assert blob[:n] == ref and blob[n:] == data.tobytes() * 2


# ---------------------------------------------------------------------------------------------------------------
# Fold (row a13; oracle/_ref/libdspsr_reffold.so): the statement blocks of dsp::Fold::fold -- weight set-up
# (Fold.C:687-716), the per-sample phase recurrence / bin plan / hits loop (:744-787) and the OrderFPT accumulation
# (:835-873) -- compiled from the reference's own text (oracle/ref.mk cuts them out of the file in place) inside
# oracle/ref_shim/ref_fold.cpp.  The oracle's orc_fold_plan / orc_fold_plan_weighted / orc_fold must agree bit for bit.
# ---------------------------------------------------------------------------------------------------------------
REFFOLD = os.path.join(ROOT, "oracle", "_ref", "libdspsr_reffold.so")
needs_fold = pytest.mark.skipif(not os.path.exists(REFFOLD), reason="oracle/_ref fold pin library not built")


@pytest.fixture(scope="module")
def reffold():
    L = C.CDLL(REFFOLD)
    L.ref_fold.restype = C.c_uint64
    L.ref_fold.argtypes = [C.c_double, C.c_double, C.c_uint, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint,
                           C.c_uint, C.c_int, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _ref_fold(L, phi, pps, nbin, idat_start, ndat, weights, ndpw, widat, x=None, ndim=1):
    binplan = np.zeros(ndat, np.uint32)
    hits = np.zeros(nbin, np.uint32)
    disc = np.zeros(1, np.uint32)
    prof = None
    args = [None, 0, 0, 0, 0]
    if x is not None:
        x = np.ascontiguousarray(x, np.float32)
        nchan, npol, n = x.shape
        prof = np.zeros((nchan, npol, nbin * ndim), np.float32)
        args = [_vp(x), n, nchan, npol, ndim]
    w = np.ascontiguousarray(weights if weights is not None else [], np.uint32)
    n = L.ref_fold(phi, pps, nbin, idat_start, ndat, _vp(w) if w.size else None, w.size, ndpw if w.size else 0, widat,
                   0, *args, _vp(binplan), _vp(hits), _vp(prof) if prof is not None else None, _vp(disc))
    return binplan, hits, n, prof, int(disc[0])


@needs_fold
@pytest.mark.parametrize("nbin,pps,phi,ndat", [
    (1024, 1.0 / 71492.5, 0.123456789, 300000),        # cfg1-like: many samples per bin
    (1024, 1.0 / 4000.37, 0.999999, 200000),           # cfg3-like: ~4 samples per bin, start next to a wrap
    (37, 0.0173, 0.0, 5000),                           # a period of a few dozen samples
    (2, 0.61803398875, 0.5, 1000),                     # period shorter than two samples
    (8192, 1e-9, 0.75, 100000),                        # a whole block inside one bin
])
def test_fold_bin_plan_matches_reference_loop(reffold, oracle, nbin, pps, phi, ndat):
    """The sequential double-precision recurrence of Fold.C:765-768 and its hit counting, from the reference's text,
    against the oracle (and through it against the GPU's segment expansion, tests/test_host_logic.py)."""
    rb, rh, rn, _, _ = _ref_fold(reffold, phi, pps, nbin, 0, ndat, None, 0, 0)
    ob, oh, on, _ = oracle.fold_plan(phi, pps, nbin, ndat)
    assert rn == on == ndat
    assert np.array_equal(rb, ob) and np.array_equal(rh, oh)


@needs_fold
@pytest.mark.parametrize("seed", range(6))
def test_fold_weighted_plan_and_accumulation_match_reference_loop(reffold, oracle, seed):
    """Weighted input (WeightedTimeSeries windows with zero weight are not folded and get bin = nbin, Fold.C:687-716,
    746-763, 777-786) and the OrderFPT accumulation loop (Fold.C:835-873) for every detected layout: bin plan, hits,
    ndat_folded and the float profile sums are bit-identical."""
    rng = np.random.default_rng(100 + seed)
    nbin = int(rng.choice([16, 128, 1024]))
    ndpw = int(rng.choice([64, 512, 1000]))
    idat_start = int(rng.integers(0, 3 * ndpw))
    widat = int(rng.integers(0, ndpw))
    ndat = int(rng.integers(2000, 20000))
    nweights = (idat_start + ndat + widat) // ndpw + 2
    weights = (rng.random(nweights) > 0.3).astype(np.uint32) * rng.integers(1, 9, nweights).astype(np.uint32)
    phi, pps = float(rng.random()), float(1.0 / rng.uniform(3.0, 5000.0))
    nchan, npol, ndim = int(rng.integers(1, 4)), int(rng.choice([1, 2, 4])), int(rng.choice([1, 2, 4]))
    x = rng.standard_normal((nchan, npol, (idat_start + ndat) * ndim)).astype(np.float32)
    rb, rh, rn, rprof, rdisc = _ref_fold(reffold, phi, pps, nbin, idat_start, ndat, weights, ndpw, widat, x, ndim)
    ob, oh, on = oracle.fold_plan_weighted(phi, pps, nbin, idat_start, ndat, weights, ndpw, widat)
    assert rn == on and np.array_equal(rb, ob) and np.array_equal(rh, oh)
    assert rdisc == int(np.count_nonzero(weights[(idat_start + widat) // ndpw:(idat_start + ndat - 1 + widat) // ndpw + 1] == 0))
    oprof = oracle.fold(x, ndim, ob, nbin, idat_start=idat_start)
    assert np.array_equal(rprof.view(np.uint32), oprof.view(np.uint32))
    assert (rb == nbin).any() and (rb != nbin).any()


# ---------------------------------------------------------------------------------------------------------------
# Generic 8-bit unpacker (row a3; oracle/_ref/libdspsr_refbit.so): BitUnpacker.C + EightBitUnpacker.C + BitTable.C
# compiled in place with their own headers (only dsp/HistUnpacker.h, the PSRCHIVE-dependent base, is a stand-in).
# ---------------------------------------------------------------------------------------------------------------
REFBIT = os.path.join(ROOT, "oracle", "_ref", "libdspsr_refbit.so")


@pytest.mark.skipif(not os.path.exists(REFBIT), reason="oracle/_ref/libdspsr_refbit.so not built")
@pytest.mark.parametrize("nchan,npol,ndim,twos", [(1, 2, 1, 1), (3, 2, 2, 1), (4, 1, 2, 0), (2, 2, 1, 0)])
def test_generic8_unpack_and_histogram_match_reference(oracle, nchan, npol, ndim, twos):
    """BitUnpacker.C:48-80 (the digitizer walk over TFP bytes) and EightBitUnpacker.C:25-49 (hist[*from]++,
    *into = lookup[*from]): floats and the per-digitizer histograms are bit-identical with the oracle."""
    oracle.lib()
    L = C.CDLL(REFBIT)
    L.ref_unpack_generic8.restype = C.c_int
    L.ref_unpack_generic8.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_uint64,
                                      C.c_void_p]
    rng = np.random.default_rng(55)
    ndat = 3001
    raw = rng.integers(0, 256, ndat * nchan * npol * ndim, dtype=np.uint8)
    raw[:6] = [0, 1, 127, 128, 129, 255]
    want = np.zeros((nchan, npol, ndat * ndim), np.float32)
    whist = np.zeros((nchan * npol * ndim, 256), np.uint64)
    assert L.ref_unpack_generic8(_vp(raw), ndat, nchan, npol, ndim, twos, _vp(want), ndat * ndim, _vp(whist)) == 0
    lut, _ = oracle.bittable8(twos_complement=bool(twos))
    got = np.zeros_like(want)
    ghist = np.zeros_like(whist)
    oracle.lib().orc_unpack_generic8(_vp(raw), C.c_uint64(ndat), nchan, npol, ndim, _vp(lut), _vp(got),
                                     C.c_uint64(ndat * ndim), _vp(ghist))
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert np.array_equal(whist, ghist) and int(whist.sum()) == raw.size


# ---------------------------------------------------------------------------------------------------------------
# Filterbank / Convolution (rows a10, a11): the overlap-save loop nests of Filterbank.C:563-660 and
# Convolution.C:389-458 compiled from the reference's own text (oracle/ref_shim/ref_fbconv.cpp) with the reference's
# own Response::operate; only the three FFT calls go to the oracle's restatement of the FFTW conventions (FFTW lives in
# PSRCHIVE, which is not in the tree).  The oracle's loops must reproduce the output bit for bit.
# ---------------------------------------------------------------------------------------------------------------
@needs_cxx
@pytest.mark.parametrize("real,input_nchan,npol,nchan,F,npos,nneg,with_H", [
    (1, 1, 2, 16, 128, 11, 13, True),       # cfg1-shaped: real input, one input channel, -F 16:D
    (1, 1, 1, 8, 1, 0, 0, False),           # freq_res 1: the 64-bit copy branch, no response
    (0, 3, 2, 12, 64, 5, 6, True),          # complex multi-channel input, 4 sub-channels each
    (0, 2, 2, 2, 256, 20, 21, True),        # nchan_subband 1 through the filterbank
])
def test_filterbank_loop_matches_reference_text(refcxx, oracle, real, input_nchan, npol, nchan, F, npos, nneg, with_H):
    f = oracle.fb_sizes(bool(real), input_nchan, npol, nchan, F, npos, nneg)
    ndim = 1 if real else 2
    npart = 3
    ndat = npart * f.nsamp_step + f.nsamp_overlap
    rng = np.random.default_rng(7)
    x = rng.standard_normal((input_nchan, npol, ndat * ndim)).astype(np.float32)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64) if with_H else None
    got = oracle.filterbank(f, x, H)
    assert got.shape == (nchan, npol, npart * f.nkeep)
    want = np.zeros_like(got)
    refcxx.ref_filterbank.restype = C.c_int
    refcxx.ref_filterbank.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint,
                                      C.c_uint, C.c_uint, C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    assert refcxx.ref_filterbank(_vp(x), x.shape[2], input_nchan, npol, real, f.nchan_subband, F, npos, nneg, f.nsamp_fft,
                                 f.nsamp_step, npart, _vp(H) if with_H else None, _vp(want), 2 * npart * f.nkeep) == 0
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert np.abs(want).max() > 0


@needs_cxx
@pytest.mark.parametrize("real,nchan,npol,n_fft,npos,nneg", [(0, 3, 2, 512, 40, 45), (1, 1, 2, 1024, 100, 101), (0, 1, 1, 64, 0, 7)])
def test_convolution_loop_matches_reference_text(refcxx, oracle, real, nchan, npol, n_fft, npos, nneg):
    c = oracle.conv_sizes(bool(real), nchan, npol, n_fft, npos, nneg)
    ndim = 1 if real else 2
    npart = 3
    ndat = npart * c.nsamp_step + c.nsamp_overlap
    rng = np.random.default_rng(8)
    x = rng.standard_normal((nchan, npol, ndat * ndim)).astype(np.float32)
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, n_fft))).astype(np.complex64)
    got = oracle.convolution(c, x, H)
    nkeep = n_fft - npos - nneg
    assert got.shape == (nchan, npol, npart * nkeep)
    want = np.zeros_like(got)
    refcxx.ref_convolution.restype = C.c_int
    refcxx.ref_convolution.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_uint,
                                       C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    assert refcxx.ref_convolution(_vp(x), x.shape[2], nchan, npol, real, n_fft, npos, c.nsamp_fft, c.nsamp_step, npart,
                                  _vp(H), _vp(want), 2 * npart * nkeep) == 0
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert np.abs(want).max() > 0


@needs_fold
@pytest.mark.parametrize("seed", range(8))
def test_weights_convolve_and_scrunch_match_reference_text(reffold, oracle, seed):
    """the bodies of WeightedTimeSeries::convolve_weights (WeightedTimeSeries.C:584-696: a transform
    holding a bad window is flagged as a whole, one transform late) and scrunch_weights (:705-774), compiled from the
    reference's text, against the oracle -- random bad windows, both scrunch regimes."""
    rng = np.random.default_rng(300 + seed)
    ndpw = int(rng.choice([16, 64, 512]))
    nfft = int(rng.choice([128, 1024, 4096]))
    nkeep = int(nfft - rng.integers(1, nfft // 2))
    weight_idat = int(rng.integers(0, ndpw))
    ndat = int(nfft + nkeep * rng.integers(2, 12) + rng.integers(0, nkeep))
    nweights = (ndat + weight_idat + ndpw - 1) // ndpw + 1
    w0 = (rng.random(nweights) > 0.15).astype(np.uint32) * rng.integers(1, 5, nweights).astype(np.uint32)
    reffold.ref_convolve_weights.restype = C.c_int
    reffold.ref_convolve_weights.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64, C.c_uint64, C.c_uint, C.c_uint]
    want = w0.copy()
    assert reffold.ref_convolve_weights(_vp(want), want.size, ndpw, weight_idat, ndat, nfft, nkeep) == 0
    got = oracle.convolve_weights(w0, ndpw, weight_idat, ndat, nfft, nkeep)
    assert np.array_equal(want, got)
    if ndpw < nfft:
        assert (want == 0).sum() >= (w0 == 0).sum()
    reffold.ref_scrunch_weights.restype = None
    reffold.ref_scrunch_weights.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint), C.POINTER(C.c_uint64), C.c_uint]
    for nscrunch in (2, ndpw // 4, ndpw, 3 * ndpw):
        wref = want.copy()
        npw, wi = C.c_uint(ndpw), C.c_uint64(weight_idat)
        reffold.ref_scrunch_weights(_vp(wref), wref.size, C.byref(npw), C.byref(wi), nscrunch)
        ow, onpw, owi = oracle.scrunch_weights(want, ndpw, weight_idat, nscrunch)
        assert (npw.value, wi.value) == (onpw, owi)
        assert np.array_equal(wref[:ow.size], ow)
