"""Native multi-GPU host parity (include/b200multi.h): ONE process, one worker thread + stream + pipeline per GPU.

  python tests/mgpu_native.py time|channel

Run in its own process by tests/test_gpu_multi.py (the host owns contexts on every device; keeping that out of the
pytest process keeps the other GPU tests' CUDA state untouched).  Prints one JSON line."""
import json
import math
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def W_seconds(mjd, start):
    return (mjd[0] - start[0]) * 86400.0 + (mjd[1] - start[1]) + (mjd[2] - start[2])


def main(mode):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    import synth
    import workloads as W
    from dspsr_b200 import _lib as L, engine as E, hostmath as HM, multi as M, phaseseries as P
    O.build()
    ndev = 2
    host = M.MultiHost(list(range(ndev)))
    assert host.nccl_version() >= 22000
    start = W.utc_to_mjd("2010-04-13-02:05:45")
    pred = HM.Polyco(W.polyco_text())
    opc = O.polyco_parse(W.polyco_text())
    ophase, ofreq = (lambda m: O.polyco_phase(opc, *m)[0]), (lambda m: O.polyco_frequency(opc, *m))
    nbin = 256
    if mode == "time":
        C_, F, npos, nneg, K = 32, 1024, 60, 61, 3
        cfg = dict(W.CFG1, nchan=C_)
        S = W.sizes(cfg, F, npos, nneg)
        ndat = (ndev * K * S["step"] + S["overlap"] + 3) // 4 * 4
        raw = synth.caspsr_bytes(ndat, seed=301)
        rng = np.random.default_rng(302)
        H = np.exp(1j * rng.uniform(-np.pi, np.pi, (C_, F))).astype(np.complex64)
        lut, _ = HM.bittable8()
        robs = P.observation(1, 2, 1, S["rate_in"], start, ndat=ndat, centre_frequency=cfg["freq"], bandwidth=cfg["bw"],
                             dm=cfg["dm"], state=17)
        # constant-period folding (Fold::get_phi, Fold.C:943-950) with a period of the order of one block, so that
        # every bin is hit: phi = fmod(t - start, P) / P
        period = 0.37 * K * S["nkeep"] / S["rate_out"]
        ophase = lambda m: math.fmod(W_seconds(m, start), period) / period
        ofreq = lambda m: 1.0 / period
        for i in range(ndev):
            ctx = host.context(i)
            ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
            fd, keep = E.make_fb_desc(1, 1, 2, C_, F, npos, nneg, H)
            pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
            pipe.set_observation(robs)
            pipe.set_folding_period(period, reference_epoch=start)
            host.set_pipeline(i, pipe)
        nbytes = (K * S["step"] + S["overlap"]) * 2
        inputs = [raw[i * K * S["step"] * 2: i * K * S["step"] * 2 + nbytes] for i in range(ndev)]
        host.execute_host_obs(inputs, [K] * ndev, [i * K * S["step"] for i in range(ndev)])
        out = host.combine(M.SHARD_TIME, C_, 1, 4, nbin)
        ph = [W.block_phase(S, start, i * K * S["step"], ophase, ofreq) for i in range(ndev)]
        luto, _ = O.bittable8()
        f = O.fb_sizes(1, 1, 2, C_, F, npos, nneg)
        op = O.make_pipe(0, 1, 2, 1, luto, 0.0, f, None, H, "Coherence", 4, nbin)
        ref, ref_hits = O.pipe_run(op, raw, ndev, K, [p[0] for p in ph], [p[1] for p in ph], nthread=1)
        assert np.array_equal(out.hits, ref_hits)
        assert synth.relerr(out.data, ref) <= 1e-5
        assert out.ndat_total == ndev * K * S["nkeep"]
        assert out.integration_length == pytest.approx(ndev * K * S["nkeep"] / S["rate_out"], rel=1e-12)
        assert (out.ps.end_time.sec - out.ps.obs.start_time.sec) + (out.ps.end_time.frac - out.ps.obs.start_time.frac) == \
            pytest.approx(((ndev - 1) * K * S["step"]) / S["rate_in"] + K * S["nkeep"] / S["rate_out"], rel=1e-9)
        host.reset()
        assert host.combine(M.SHARD_TIME, C_, 1, 4, nbin).integration_length == 0.0
    else:
        nloc, F, npos, nneg, npart = 3, 4096, 150, 160, 2
        nchan = nloc * ndev
        step, overlap = F - npos - nneg, npos + nneg
        ndat = (npart * step + overlap + 255) // 256 * 256
        raw = synth.meerkat_bytes(ndat, nchan, 2, seed=303)
        rng = np.random.default_rng(304)
        H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
        _, scale = HM.bittable8()
        bw, cf = 856.0 * nchan / 1024, 1284.0
        period = 0.41 * npart * step / (856e6 / 1024)
        ophase = lambda m: math.fmod(W_seconds(m, start), period) / period
        ofreq = lambda m: 1.0 / period
        inputs = []
        for i in range(ndev):
            ctx = host.context(i)
            c0 = i * nloc
            mine = np.ascontiguousarray(raw.reshape(ndat // 256, 2, nchan, 512)[:, :, c0:c0 + nloc, :]).reshape(-1)
            inputs.append(mine)
            ud = E.make_unpack_desc(L.FMT_MEERKAT8, nloc, 2, 2, None, np.float32(scale), 1)
            fd, keep = E.make_fb_desc(False, nloc, 2, 1, F, npos, nneg, H[c0:c0 + nloc])
            pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, nbin)
            sub_bw = bw / ndev
            pipe.set_observation(P.observation(nloc, 2, 2, 856e6 / 1024, start, ndat=ndat,
                                               centre_frequency=cf - 0.5 * bw + (i + 0.5) * sub_bw, bandwidth=sub_bw,
                                               dm=500.0, state=18, machine="MKBF"))
            pipe.set_folding_period(period, reference_epoch=start)
            host.set_pipeline(i, pipe)
        host.execute_host_obs(inputs, [npart] * ndev, [0] * ndev)
        out = host.combine(M.SHARD_CHANNEL, nchan, 1, 4, nbin)
        S = dict(rate_in=856e6 / 1024, rate_out=856e6 / 1024, npos=npos)
        phi, pps = W.block_phase(S, start, 0, ophase, ofreq)
        c = O.conv_sizes(0, nchan, 2, F, npos, nneg)
        op = O.make_pipe(2, nchan, 2, 2, None, np.float32(scale), None, c, H, "Coherence", 4, nbin)
        ref, ref_hits = O.pipe_run(op, raw, 1, npart, [phi], [pps], nthread=1)
        assert np.array_equal(out.hits, ref_hits)
        assert synth.relerr(out.data, ref) <= 1e-5
        assert out.ps.obs.nchan == nchan and out.ps.obs.bandwidth == pytest.approx(bw) and out.ps.obs.centre_frequency == pytest.approx(cf)
        assert out.ndat_total == npart * step
    host.close()
    print(json.dumps({"mode": mode, "ok": True}))


if __name__ == "__main__":
    main(sys.argv[1])
