"""Multi-GPU parity worker: one process per GPU under torchrun, NCCL for the sub-integration combine.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/mgpu_worker.py --case time|channel|subband --out result.json

Every case runs the product pipeline on each rank's shard, combines the per-GPU PhaseSeries on rank 0 the way
`dspsr` combines its threads' (PhaseSeries::combine, PhaseSeries.C:442-480; MultiThread.C:329-342), and rank 0
compares the combined profile / hits / ndat_total with the single-process CPU oracle over the WHOLE problem.

  time     cfg1 at full shape (CASPSR 8-bit, -F 256:D, freq_res 8192, Coherence, 1024 bins): rank g owns parts
           [g*K, (g+1)*K) of one stream, re-reading nsamp_overlap samples at its left edge; NCCL reduce (sum)
  channel  cfg3 shape (MeerKAT heaps, 65536-point convolution, M = 2536 + 2543): the channels of ONE heap-ordered
           stream are split into contiguous ranges, each rank unpacks its own range; NCCL gather (no arithmetic)
  subband  cfg5 shape (UWB 16-bit sub-bands 0, 6, 12, 25 = freq_res 16384 / 2048 / 512 / 64): sub-bands dealt out
           by cost, two sub-integrations, gather at each boundary
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 1e-5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True, choices=["time", "channel", "subband"])
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import synth
    import workloads as W
    from dspsr_b200 import _lib as L
    from dspsr_b200 import engine as E
    from dspsr_b200 import hostmath as HM
    from dspsr_b200 import sharding

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    res = {"case": args.case, "world": world, "backend": dist.get_backend()}

    if args.case == "time":
        cfg = W.CFG1
        d, H = HM.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], True)
        S = W.sizes(cfg, d.ndat, d.impulse_pos, d.impulse_neg)
        K = 2
        ndat = (world * K * S["step"] + S["overlap"] + 3) // 4 * 4
        raw = synth.caspsr_bytes(ndat, seed=101)                      # the same stream on every rank
        pred = HM.Polyco(W.polyco_text())
        start = W.utc_to_mjd(cfg["utc_start"])
        phis = [W.block_phase(S, start, g * K * S["step"], pred.phase, pred.frequency) for g in range(world)]
        lut, _ = HM.bittable8()
        ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, lut)
        fd, keep = E.make_fb_desc(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"], H)
        pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, 1024)
        lo, hi = sharding.part_byte_range(rank * K, K, S["step"], S["overlap"], 2)
        pipe.execute(torch.from_numpy(raw[lo:hi]).cuda(), K, phis[rank][0], phis[rank][1], first_sample=0)
        _, _, ntot_local = pipe.synch()
        prof_dev, hits_dev = pipe.fold.device_profile(), pipe.fold.device_hits()
        _, ntot = sharding.combine_time_sharded(prof_dev, hits_dev, K * S["nkeep"] / S["rate_out"], ntot_local)
        torch.cuda.synchronize()
        if rank == 0:
            import oracle as O
            prof = prof_dev.cpu().numpy().reshape(S["C"], 1, -1)
            hits = hits_dev.cpu().numpy().astype(np.uint32)
            _, Ho = O.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], True)
            luto, _ = O.bittable8()
            f = O.fb_sizes(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"])
            op = O.make_pipe(0, 1, 2, 1, luto, 0.0, f, None, Ho, "Coherence", 4, 1024)
            ref, ref_hits = O.pipe_run(op, raw, world, K, [p[0] for p in phis], [p[1] for p in phis], nthread=min(world, 8))
            res.update(err=synth.relerr(prof, ref), hits_equal=bool(np.array_equal(hits, ref_hits)),
                       ndat_total=int(ntot), ndat_expected=int(world * K * S["nkeep"]))

    elif args.case == "channel":
        nloc, F, npos, nneg, npart = 4, 65536, 2536, 2543, 2
        nchan = nloc * world
        step, overlap = F - npos - nneg, npos + nneg
        ndat = (npart * step + overlap + 255) // 256 * 256
        raw = synth.meerkat_bytes(ndat, nchan, 2, seed=102)          # [heap][pol][chan][256][re,im]
        rng = np.random.default_rng(103)
        H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64)
        _, scale = HM.bittable8()
        c0, n = sharding.shard_channels(nchan, world, rank)
        mine = np.ascontiguousarray(raw.reshape(ndat // 256, 2, nchan, 512)[:, :, c0:c0 + n, :]).reshape(-1)
        ud = E.make_unpack_desc(L.FMT_MEERKAT8, n, 2, 2, None, np.float32(scale), 1)
        fd, keep = E.make_fb_desc(False, n, 2, 1, F, npos, nneg, H[c0:c0 + n])
        pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, 1024)
        phi, pps = 0.3, 1.0 / (0.41 * step * npart)
        pipe.execute(torch.from_numpy(mine).cuda(), npart, phi, pps, first_sample=0)
        prof_local = pipe.fold.device_profile().view(n, 1, -1)
        full = sharding.gather_channel_sharded(prof_local, nchan)
        torch.cuda.synchronize()
        _, hits, ntot = pipe.synch()
        if rank == 0:
            import oracle as O
            c = O.conv_sizes(0, nchan, 2, F, npos, nneg)
            op = O.make_pipe(2, nchan, 2, 2, None, np.float32(scale), None, c, H, "Coherence", 4, 1024)
            ref, ref_hits = O.pipe_run(op, raw, 1, npart, [phi], [pps], nthread=1)
            res.update(err=synth.relerr(full.cpu().numpy(), ref), hits_equal=bool(np.array_equal(hits, ref_hits)),
                       ndat_total=int(ntot), ndat_expected=int(npart * step))

    else:
        subs = [0, 6, 12, 25]
        info = []
        for k in subs:
            cfg = W.cfg5_subband(k)
            d, H = HM.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], False)
            info.append((cfg, d, H, W.sizes(cfg, d.ndat, d.impulse_pos, d.impulse_neg)))
        owner = W.assign_by_cost([W.subband_cost(x[3]) for x in info], world)
        span, nsub = 1 << 22, 2                                       # samples per sub-band and sub-integration
        pred = HM.Polyco(W.polyco_text())
        pipes = {}
        for i in owner[rank]:
            cfg, d, H, S = info[i]
            npart = max(1, span // S["step"])
            ndat = (nsub * npart * S["step"] + S["overlap"] + 2047) // 2048 * 2048
            raw = synth.uwb_bytes(ndat, 2, seed=200 + i)
            ud = E.make_unpack_desc(L.FMT_UWB16, 1, 2, 2)
            fd, keep = E.make_fb_desc(False, 1, 2, S["C"], S["F"], S["npos"], S["nneg"], H)
            pipes[i] = (E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, 1024), torch.from_numpy(raw).cuda(), npart, S)
        start = W.utc_to_mjd(info[0][0]["utc_start"])
        nfl = [info[i][3]["C"] * 4 * 1024 for i in range(len(subs))]
        nmax = max(sum(nfl[i] for i in o) for o in owner)
        hmax = max(len(o) for o in owner) * 1024
        worst, hits_ok, ntot_ok = 0.0, True, True
        for isub in range(nsub):
            pad = torch.zeros(nmax, dtype=torch.float32, device="cuda")
            hpad = torch.zeros(hmax, dtype=torch.int32, device="cuda")
            o = 0
            for j, i in enumerate(owner[rank]):
                pipe, d_raw, npart, S = pipes[i]
                phi, pps = W.block_phase(S, start, isub * npart * S["step"], pred.phase, pred.frequency)
                pipe.zero()
                pipe.execute(d_raw, npart, phi, pps, first_sample=isub * npart * S["step"])
                pad[o:o + nfl[i]].copy_(pipe.fold.device_profile())
                hpad[j * 1024:(j + 1) * 1024].copy_(pipe.fold.device_hits())
                o += nfl[i]
            bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            hbufs = [torch.empty_like(hpad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, bufs, dst=0)                              # the sub-integration boundary
            dist.gather(hpad, hbufs, dst=0)
            torch.cuda.synchronize()
            if rank == 0:
                import oracle as O
                opc = O.polyco_parse(W.polyco_text())

                class OP:
                    phase = staticmethod(lambda m: O.polyco_phase(opc, *m)[0])
                    frequency = staticmethod(lambda m: O.polyco_frequency(opc, *m))
                import ctypes as C
                O.lib().orc_pipe_block.restype = C.c_uint64
                for r in range(world):
                    o = 0
                    for j, i in enumerate(owner[r]):
                        cfg, d, H, S = info[i]
                        npart = max(1, span // S["step"])
                        ndat = (nsub * npart * S["step"] + S["overlap"] + 2047) // 2048 * 2048
                        raw = synth.uwb_bytes(ndat, 2, seed=200 + i)
                        _, Ho = O.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], False)
                        f = O.fb_sizes(0, 1, 2, S["C"], S["F"], S["npos"], S["nneg"])
                        op = O.make_pipe(3, 1, 2, 2, None, 0.0, f, None, Ho, "Coherence", 4, 1024)
                        phi, pps = W.block_phase(S, start, isub * npart * S["step"], OP.phase, OP.frequency)
                        ref = np.zeros((S["C"], 1, 4096), np.float32)
                        rh = np.zeros(1024, np.uint32)
                        O.lib().orc_pipe_block(C.byref(op), raw.ctypes.data_as(C.c_void_p), C.c_uint64(isub * npart),
                                               C.c_uint64(npart), C.c_double(phi), C.c_double(pps),
                                               ref.ctypes.data_as(C.c_void_p), rh.ctypes.data_as(C.c_void_p), None)
                        got = bufs[r][o:o + nfl[i]].cpu().numpy()
                        o += nfl[i]
                        worst = max(worst, synth.relerr(got, ref))
                        gh = hbufs[r][j * 1024:(j + 1) * 1024].cpu().numpy().astype(np.uint32)
                        hits_ok &= bool(np.array_equal(gh, rh))
                        ntot_ok &= int(rh.sum()) == npart * S["nkeep"]
        if rank == 0:
            res.update(err=worst, hits_equal=hits_ok and ntot_ok, owner=owner, ndat_total=0, ndat_expected=0)

    if rank == 0:
        res["ok"] = bool(res["err"] <= TOL and res["hits_equal"] and res["ndat_total"] == res["ndat_expected"])
        with open(args.out, "w") as f:
            json.dump(res, f)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
