"""CPU tests of the product's host-side logic (no GPU, no compute kernels): the C-ABI library loads
and exports what include/b200dsp.h declares, the exact fold bin plan, the host math against the
oracle, the FFT core index algebra (host emulation), sharding helpers, and a world_size-2 gloo run
of the sub-integration combine."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from dspsr_b200 import _lib as L
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(lib, name), "libb200dsp.so does not export %s" % name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.b200_version() >= 100


def test_multi_library_loads_and_exports_every_declared_symbol():
    from dspsr_b200 import multi as M
    lib = M.load()
    hdr = open(os.path.join(ROOT, "include", "b200multi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200_multi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 11
    for name in sorted(declared):
        assert hasattr(lib, name), "libb200multi.so does not export %s" % name
    assert declared == set(M.SIGNATURES), declared ^ set(M.SIGNATURES)
    assert lib.b200_multi_nccl_version() >= 22000


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dspsr_b200 import _lib as L
    h = C.c_void_p()
    rc = L.load().b200_context_create(0, None, C.byref(h))
    assert rc != 0 and b"no CUDA device" in L.load().b200_last_error()
    from dspsr_b200 import engine as E
    with pytest.raises(RuntimeError):
        E.Context(0)


def test_product_never_touches_the_oracle():
    # the product (dspsr_b200/, include/) must not import, link or name anything under oracle/
    for base, _, files in os.walk(os.path.join(ROOT, "dspsr_b200")):
        if "_build" in base or "__pycache__" in base:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(base, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), fn
                assert "liboracle" not in txt and "orc_" not in txt, fn


def test_phase_segments_equal_sequential_recurrence(oracle):
    from dspsr_b200 import engine as E
    rng = np.random.default_rng(0)
    cases = [(0.3, 7.2e-6, 300000), (0.999999, 0.013, 5000), (0.0, 0.5, 1000), (0.7, 1e-9, 100000),
             (0.25, 0.25, 100), (0.1, 0.7, 1000), (-0.3, 3.3e-4, 50000), (5.75, 1e-3, 10000)]
    for _ in range(30):
        cases.append((rng.random(), 10 ** rng.uniform(-7, -1.3), int(rng.integers(1, 100000))))
    for _ in range(20):   # few significant bits: round-half-even ties in the recurrence
        cases.append((rng.integers(0, 1 << 20) / float(1 << 20),
                      rng.integers(1, 1 << 12) / float(1 << (12 + rng.integers(1, 40))), 20000))
    for phi, pps, n in cases:
        for nbin in (1024, 1000):
            ref, phi_end_ref = E.phase_bins_sequential(phi, pps, nbin, n)
            orc, _, _, phi_end_orc = oracle.fold_plan(phi, pps, nbin, n)
            segs, phi_end = E.phase_segments(phi, pps, n, 1 << 18)
            got = E.expand_segments_numpy(segs, nbin, n)
            assert np.array_equal(ref, orc) and np.array_equal(got, ref), (phi, pps, n, nbin)
            assert phi_end == phi_end_ref == phi_end_orc
            assert sum(s.count for s in segs) == n
            for s in segs:   # the device scales by a multiplication with 2^scale_exp: the same double as ldexp
                a = np.uint64(s.a0) + np.arange(min(s.count, 64), dtype=np.uint64) * np.uint64(s.step)
                if s.scale_exp >= -1000:
                    assert np.array_equal(a.astype(np.float64) * np.ldexp(1.0, s.scale_exp), np.ldexp(a.astype(np.float64), s.scale_exp))
    segs, _ = E.phase_segments(0.3, 7.16e-6, 466000)     # a cfg1 block of 64 parts: a handful of segments
    assert len(segs) < 100
    assert E.phase_segments(0.3, 1e-3, 0)[0] == []        # empty input


def test_hostmath_bitexact_with_oracle(oracle):
    from dspsr_b200 import hostmath as HM
    import workloads as W
    for twos in (True, False):
        a, sa = HM.bittable8(twos)
        b, sb = oracle.bittable8(twos)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and sa == sb
    for args in [(1382, -400, 67.99, 1, 256, True), (1400, 128, 50, 1, 4096, True), (1284, 16, 0.5, 4, 4, False),
                 (1400, 64, 1, 1, 8, False)]:
        d, H = HM.dedispersion(*args)
        do, Ho = oracle.dedispersion(*args)
        assert (d.ndat, d.impulse_pos, d.impulse_neg) == (do.ndat, do.impulse_pos, do.impulse_neg)
        assert np.array_equal(H.view(np.uint32), Ho.view(np.uint32))
    with pytest.raises(Exception):
        HM.dedispersion(1400, 400, 1500, 1, 1, False, build=False)
    p = HM.Polyco(W.polyco_text())
    po = oracle.polyco_parse(W.polyco_text())
    start = HM.utc_to_mjd("2010-04-13-02:05:45")
    assert start == (55299, 7545, 0.0)
    for dt in (0.0, 1.2345678, 600.000001, 3599.5):
        mjd = HM.mjd_add(start, dt)
        assert p.phase(mjd) == oracle.polyco_phase(po, *mjd)[0]
        assert p.frequency(mjd) == oracle.polyco_frequency(po, *mjd)
    phi, pps = HM.fold_phase(p, start, 0, 1.5625e6)
    assert 0 <= phi < 1 and pps == pytest.approx(7.16e-6, rel=1e-2)


def test_fft_core_index_algebra_on_the_host():
    """Builds csrc/host_fft_emul.cu with nvcc (host code only) and runs it: the very same
    __host__ __device__ Stockham stage / twiddle / shared-memory map code the kernels use."""
    exe = os.path.join(ROOT, "dspsr_b200", "_build", "host_fft_emul")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "dspsr_b200", "csrc"), "../_build/host_fft_emul"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]


def test_sharding_helpers():
    from dspsr_b200 import sharding as S
    for total, world in [(64, 8), (10, 4), (3, 8), (1024, 3)]:
        parts = [S.shard_parts(total, world, r) for r in range(world)]
        assert sum(n for _, n in parts) == total
        pos = 0
        for first, n in parts:
            assert first == pos
            pos += n
        assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    lo, hi = S.part_byte_range(4, 2, 1000, 100, 2)
    assert (lo, hi) == (8000, 12200)


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    import synth
    from dspsr_b200 import sharding as S
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lut, _ = O.bittable8()
    f = O.fb_sizes(1, 1, 2, 8, 64, 5, 6)
    total_parts, nbin = 7, 32            # ragged: 4 + 3 parts
    ndat = (total_parts * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=21)
    H = np.exp(1j * np.random.default_rng(21).uniform(-3, 3, (8, 64))).astype(np.complex64)
    pipe = O.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    first, n = S.shard_parts(total_parts, world, rank)
    lo, hi = S.part_byte_range(first, n, f.nsamp_step, f.nsamp_overlap, 2)
    assert hi <= raw.size
    # every rank folds its own super-block with the phase of ITS first output sample; the CPU
    # oracle stands in for the GPU here -- the test is about sharding + combine
    pps = 1.0 / 61.7
    phi = (0.15 + first * f.nkeep * pps) % 1.0
    prof = np.zeros((8, 1, nbin * 4), np.float32)
    hits = np.zeros(nbin, np.uint32)
    O.lib().orc_pipe_block(C.byref(pipe), raw.ctypes.data_as(C.c_void_p), C.c_uint64(first), C.c_uint64(n),
                           C.c_double(phi), C.c_double(pps), prof.ctypes.data_as(C.c_void_p),
                           hits.ctypes.data_as(C.c_void_p), None)
    tp = torch.from_numpy(prof)
    th = torch.from_numpy(hits.astype(np.int32))
    il, nt = S.combine_time_sharded(tp, th, n * f.nkeep / 1e6, n * f.nkeep)
    # channel-sharded gather of ragged shards
    c0, cn = S.shard_channels(5, world, rank)
    local = torch.arange(c0, c0 + cn, dtype=torch.float32).reshape(cn, 1).repeat(1, 3)
    gathered = S.gather_channel_sharded(local, 5)
    if rank == 0:
        np.savez(tmp, prof=tp.numpy(), hits=th.numpy(), nt=nt, il=il, gathered=gathered.numpy())
    dist.destroy_process_group()


def test_gloo_world2_time_sharded_combine(oracle, tmp_path):
    import socket
    import torch.multiprocessing as mp
    import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    tmp = str(tmp_path / "out.npz")
    mp.spawn(_gloo_worker, args=(2, port, tmp), nprocs=2, join=True)
    g = np.load(tmp)
    # single-process answer: the two super-blocks folded one after the other
    lut, _ = oracle.bittable8()
    f = oracle.fb_sizes(1, 1, 2, 8, 64, 5, 6)
    total_parts, nbin = 7, 32
    ndat = (total_parts * f.nsamp_step + f.nsamp_overlap + 3) // 4 * 4
    raw = synth.caspsr_bytes(ndat, seed=21)
    H = np.exp(1j * np.random.default_rng(21).uniform(-3, 3, (8, 64))).astype(np.complex64)
    pipe = oracle.make_pipe(0, 1, 2, 1, lut, 0.0, f, None, H, "Coherence", 4, nbin)
    pps = 1.0 / 61.7
    prof = np.zeros((8, 1, nbin * 4), np.float32)
    hits = np.zeros(nbin, np.uint32)
    for first, n in ((0, 4), (4, 3)):
        phi = (0.15 + first * f.nkeep * pps) % 1.0
        oracle.lib().orc_pipe_block(C.byref(pipe), raw.ctypes.data_as(C.c_void_p), C.c_uint64(first), C.c_uint64(n),
                                    C.c_double(phi), C.c_double(pps), prof.ctypes.data_as(C.c_void_p),
                                    hits.ctypes.data_as(C.c_void_p), None)
    assert np.array_equal(g["hits"].astype(np.uint32), hits) and int(g["nt"]) == total_parts * f.nkeep
    assert synth.relerr(g["prof"], prof) < 1e-6
    assert np.array_equal(g["gathered"][:, 0], np.arange(5, dtype=np.float32))


def test_twobit_tables_bitexact_with_oracle(oracle):
    """b200_twobit_prepare (product) and the oracle's TwoBitCorrection::build restatement produce the same
    limits and bit-identical float levels for every table row."""
    from dspsr_b200 import engine as E
    for cutoff in (10.0, 3.0, 0.0):
        t = oracle.TwoBit(cutoff_sigma=cutoff)
        d = E.make_twobit_desc(npol=2, cutoff_sigma=cutoff)
        assert (d.nlow_min, d.nlow_max) == (t.nlow_min, t.nlow_max)
        for nlow in range(max(t.nlow_min, 1), min(t.nlow_max, 511) + 1):
            lo, hi = t.levels(nlow)
            assert (d.lo[nlow - d.nlow_min], d.hi[nlow - d.nlow_min]) == (lo, hi), nlow


def test_time_divide_matches_oracle_and_partitions_the_stream(oracle):
    """b200_time_divide_set_bounds (product) vs the oracle's TimeDivide restatement on block sequences whose
    boundaries do not line up with the divisions: identical decisions; every sample lands in exactly one
    slice; slices of one division cover [k L, (k+1) L) of the stream."""
    import ctypes as C
    from dspsr_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(9)
    for rate, Ldiv in ((1.5625e6, 0.01), (1e4, 0.37), (128e6 / 128, 10.0 / 4099)):
        td = L.TimeDivide()
        L.check(lib.b200_time_divide_init(C.byref(td), Ldiv))
        ot = oracle.TimeDivide(Ldiv)
        t0 = 0
        covered = []
        for blk in range(40):
            ndat = int(rng.integers(1000, 30000))
            start = t0 / rate
            more = True
            guard = 0
            while more:
                b = L.TimeBounds()
                L.check(lib.b200_time_divide_set_bounds(C.byref(td), start, rate, ndat, C.byref(b)))
                o = ot.set_bounds(start, rate, ndat)
                got = dict(is_valid=bool(b.is_valid), new_division=bool(b.new_division), end_reached=bool(b.end_reached),
                           in_next=bool(b.in_next), idat_start=b.idat_start, ndat=b.ndat, division=b.division)
                assert got == o, (blk, got, o)
                more = o["in_next"]
                if o["is_valid"]:
                    covered.append((t0 + o["idat_start"], t0 + o["idat_start"] + o["ndat"], o["division"], o["end_reached"]))
                guard += 1
                assert guard < 100
            t0 += ndat
        # contiguous, non-overlapping cover of [0, t0)
        assert covered[0][0] == 0 and covered[-1][1] == t0
        for a, b2 in zip(covered, covered[1:]):
            assert a[1] == b2[0]
        # a division's slices end within half a sample of its boundary
        for s0, s1, div, end in covered:
            assert s0 >= int(np.floor(div * Ldiv * rate - 0.51)) and s1 <= int(np.ceil((div + 1) * Ldiv * rate + 0.51))
            if end:
                assert abs(s1 - (div + 1) * Ldiv * rate) <= 0.5 + 1e-6


def test_time_divide_half_sample_boundary_does_not_throw():
    """A division boundary exactly half-way between samples makes the reference throw (TimeDivide.C:258-287);
    the product starts the next division instead and still partitions the stream."""
    import ctypes as C
    from dspsr_b200 import _lib as L
    lib = L.load()
    rate, Ldiv = 1e6, 10.0 / 4096          # 2441.40625 samples per division: boundaries at x.5 occur
    td = L.TimeDivide()
    L.check(lib.b200_time_divide_init(C.byref(td), Ldiv))
    rng = np.random.default_rng(9)
    t0, prev_end = 0, 0
    for blk in range(200):
        ndat = int(rng.integers(1000, 30000))
        more = True
        while more:
            b = L.TimeBounds()
            L.check(lib.b200_time_divide_set_bounds(C.byref(td), t0 / rate, rate, ndat, C.byref(b)))
            more = bool(b.in_next)
            if b.is_valid:
                assert t0 + b.idat_start == prev_end and b.ndat > 0
                prev_end = t0 + b.idat_start + b.ndat
        t0 += ndat
    assert prev_end == t0


def test_bind_cpu_affinity_is_harmless_without_gpu():
    """sharding.bind_cpu_affinity: no GPU / no NVML here -> returns 0 and leaves the affinity mask alone."""
    import os
    from dspsr_b200 import sharding
    before = os.sched_getaffinity(0)
    n = sharding.bind_cpu_affinity(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before


def test_tile_image_spectrum_layout_spec():
    """Executable statement of the order of Z between K2 and K3 on the cfg1 fast path (fastpath.cu
    zi_pos / k_tile_response): a K2 tile (8 mirror row pairs) leaves the SM as the verbatim image of K2's shared memory,
    bin k = r + P*k2 -> Z'[tile][g/4][k2/4][which][((g%4)*4 + k2%4) ^ h(k2/4)], h(m) = ((m>>3) ^ (m<<2)) & 15.
    The map is a bijection onto [0, Nc); a K2 tile owns one contiguous 128 KiB region and each half-CTA group one
    contiguous 64 KiB half of it; the four k2 of one K3 channel and row are one aligned 32-byte group (permuted inside
    by h & 3); the layout is bank-conflict free for K2's three shared-memory access patterns."""
    P, Q = 2048, 1024
    r = np.arange(P, dtype=np.int64)[:, None]
    k2 = np.arange(Q, dtype=np.int64)[None, :]
    m = np.where(r > P // 2, P - r, r)
    tile = np.where(r == P // 2, 0, m >> 3)
    which = np.where(r >= P // 2, 1, 0)
    g = np.where(r == P // 2, 0, m & 7)

    def h(mm):
        return ((mm >> 3) ^ (mm << 2)) & 15

    def zi(g, which, k2):
        return (g >> 2) * 8192 + (k2 >> 2) * 32 + which * 16 + ((((g & 3) << 2) | (k2 & 3)) ^ h(k2 >> 2))

    pos = tile * (Q * 16) + zi(g, which, k2)
    flat = pos.ravel()
    assert flat.min() == 0 and flat.max() == P * Q - 1 and np.unique(flat).size == P * Q
    # a K2 tile (rows 8t..8t+7 and their mirrors) = one region of 16 Q elements; pairs 0-3 / 4-7 = its two halves
    t = 5
    lo = np.arange(8 * t, 8 * t + 8)
    assert np.array_equal(np.sort(pos[np.concatenate([lo, P - lo])].ravel()), np.arange(t * Q * 16, (t + 1) * Q * 16))
    assert np.array_equal(np.sort(pos[np.concatenate([lo[:4], P - lo[:4]])].ravel()),
                          np.arange(t * Q * 16, t * Q * 16 + Q * 8))
    # K3: thread (pol, j0) of channel csub holds bins f = j0 + 256 e, e < 32: rows j0 + 256 i (i < 8), k2 = 4 csub + q
    for csub, j0 in ((77, 201), (8, 0), (255, 255), (130, 1)):
        hq = h(np.int64(csub)) & 3
        for i in range(8):
            row = j0 + 256 * i
            offs = pos[row, 4 * csub: 4 * csub + 4]
            base = offs.min()
            assert base % 4 == 0 and np.array_equal(offs, base + (np.arange(4) ^ hq))    # one aligned 32-byte group
    # the four rows g % 4 of a group fill one 128-byte line (16 float2)
    line = pos[8 * t: 8 * t + 4, 4 * 77: 4 * 77 + 4].ravel()
    assert line.max() - line.min() == 15 and line.min() % 16 == 0
    # shared-memory bank pairs (64-bit accesses: 16 lanes per wavefront, 16 pairs of banks)
    j = np.arange(16)
    for gg in range(8):
        for rr in range(32):                                   # scatter of stage 0: lane j writes element 32 j + r
            assert np.unique(zi(np.int64(gg), 0, 32 * j + rr) & 15).size == 16
            assert np.unique(zi(np.int64(gg), 1, 32 * (j + 16) + rr) & 15).size == 16
        for e in range(32):                                    # gather / natural store: lane j holds element j + 32 e
            assert np.unique(zi(np.int64(gg), 0, j + 32 * e) & 15).size == 16
            assert np.unique(zi(np.int64(gg), 0, j + 16 + 32 * e) & 15).size == 16
    lane = np.arange(16)                                       # split walk: 4 consecutive k2 x 4 pairs per half warp
    for m0 in range(16):
        for it in range(8):
            kk = (lane & 3) + 4 * (m0 + 16 * it)
            gl = (lane >> 2) & 3
            assert np.unique(zi(gl, 0, kk) & 15).size == 16
            assert np.unique(zi(gl, 1, Q - 1 - kk) & 15).size == 16
            # the mirror element Q-1-k2 shares the swizzled low part: slot = (255 - m) * 32 + same low bits
            assert np.array_equal(zi(gl, 0, Q - 1 - kk) - (255 - (kk >> 2)) * 32, zi(gl, 0, kk) - (kk >> 2) * 32)


def test_long_transform_factorisation_spec():
    """Executable statement of the arithmetic of the long-transform kernels (longconv.cu k_bc_cols_fwd / k_bc_rows /
    k_bc_cols_inv): N = P Q, n = Q n1 + n2, k = k1 + P k2.  Thread j of a column (T = P/16 threads)
    or of a row (NT = Q/16) holds elements j + T e, e < 16, and builds its twiddles from ONE table value per thread times a
    16-entry table per tile; the row pass reads the response from the transposed copy Ht[k1][k2] = H[k1 + P k2] and runs
    its inverse transform as conj(FFT(conj .)); the scratch holds both polarisations of a bin side by side.  The chain
    must equal the plain overlap-save convolution N * ifft(fft(x) * H) of both polarisations."""
    rng = np.random.default_rng(5)
    P, Q = 64, 32
    N, T, NT = P * Q, P // 16, Q // 16
    x = (rng.standard_normal((2, N)) + 1j * rng.standard_normal((2, N)))
    H = np.exp(1j * rng.uniform(-np.pi, np.pi, N))
    W = lambda m: np.exp(-2j * np.pi * (np.asarray(m) % N) / N)          # the two-level table of W_N (big_twiddle)

    # K1: columns n2; register e of thread j = row k1 = j + T e; twiddle W_N^(n2 k1) = W_N^(n2 j) * W_N^(n2 T e)
    A = np.zeros((P, Q, 2), dtype=complex)                               # [k1][n2][pol]: one float4 per bin
    for n2 in range(Q):
        col = np.fft.fft(x[:, n2::Q], axis=1)                            # over n1, both polarisations
        for j in range(T):
            wbase = W(n2 * j)
            for e in range(16):
                k1 = j + T * e
                A[k1, n2, :] = col[:, k1] * (wbase * W(n2 * T * e))
    # K2: one row k1 per CTA; response row from the transposed copy; inverse = conj(FFT(conj)); W_N^-(k1 m2) factorised
    Ht = H.reshape(Q, P).T.copy()                                        # Ht[k1][k2] = H[k1 + P k2]
    assert all(Ht[k1, k2] == H[k1 + P * k2] for k1 in (0, 1, P - 1) for k2 in (0, 3, Q - 1))
    for k1 in range(P):
        row = np.fft.fft(A[k1], axis=0) * Ht[k1][:, None]                # over n2 -> k2
        inv = np.conj(np.fft.fft(np.conj(row), axis=0))                  # over k2 -> m2 (unnormalised inverse)
        assert np.allclose(inv, np.fft.ifft(row, axis=0) * Q)
        for tid in range(NT):
            wown = np.conj(W(k1 * tid))
            for e in range(16):
                m2 = tid + NT * e
                A[k1, m2, :] = inv[m2] * (wown * np.conj(W(k1 * NT * e)))
    # K3: columns m2; register e of thread j = segment m1 = j + T e: sample Q m1 + m2
    y = np.zeros((2, N), dtype=complex)
    for m2 in range(Q):
        col = np.conj(np.fft.fft(np.conj(A[:, m2, :]), axis=0))          # inverse over k1 -> m1
        for m1 in range(P):
            y[:, Q * m1 + m2] = col[m1]
    ref = np.fft.ifft(np.fft.fft(x, axis=1) * H[None, :], axis=1) * N
    assert np.allclose(y, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    # K3's fold walk: thread t owns the groups t, t + NT3, ... of NC = 4 consecutive samples; group g starts at sample
    # g Q + 4 cb - nfilt_pos of the part's bin plan: a 16-byte boundary of the plan exactly when (plan/4 - nfilt_pos) % 4 == 0
    for npos in (0, 3, 4, 534848):
        for base_words in (0, 1, 4):
            vec = (base_words - npos) % 4 == 0
            starts = [(base_words + g * Q + 4 * cb - npos) % 4 for g in (0, 5) for cb in (0, 7)]
            assert all((s == 0) == vec for s in starts)
