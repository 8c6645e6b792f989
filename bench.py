#!/usr/bin/env python
"""bench.py -- throughput of dspsr's baseband hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--parts P] [--batch B]

Workload (config.workload): BASELINE configs[0] / SURVEY cfg1 -- Benchmark/bench.csh's CASPSR
8-bit dual-pol 400 MHz real-sampled band, `dspsr -F 256:D` coherent filterbank at DM 67.99,
Coherence detection, fold into 1024 bins with Benchmark/vela.polyco.  One STEP = one pass of
unpack -> filterbank/dedisperse -> detect -> fold over one block of P overlap-save parts
(P x 3,725,312 new samples per polarisation) of seeded synthetic noise.

Printed JSON line (rank 0):
  value      input MSamples/s (samples per polarisation per second), inputs resident in HBM
  e2e        same metric through b200_pipeline_execute_host: pinned HOST bytes -> device copy ->
             kernels -> device->host read of the folded profile, all inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events
             on the launching stream) vs the measured HBM peak (MEASURED_PEAKS.json)
  roofline_path  the whole 3-kernel path: SURVEY 8(d)'s 10.13 B per sample per pol x samples / step time
  cpu_baseline   the restated reference CPU path (oracle/, NOT FFTW) on a bounded sample
--impl reference times that CPU path alone with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "input MSamples/s (per polarisation); real-time factor = value / 800"
UNIT = "MSamples/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled
    from a thread every few milliseconds (the timed region of a short run lasts only tens of ms, too short for
    `nvidia-smi -lms`); falls back to one `nvidia-smi` query if NVML cannot be loaded."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.nv = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((sm, reasons, power))
            except Exception:
                pass
            time.sleep(0.003)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["no samples"]}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        seen = set()
        for _, r, _ in self.samples:
            for name, bit in bits.items():
                if r & bit:
                    seen.add(name)
        return {"sm_mhz": float(np.median([x[0] for x in self.samples])), "sm_max_mhz": float(mx) if mx else None,
                "reasons": sorted(seen), "power_w_max": max(x[2] for x in self.samples), "samples": len(self.samples),
                "source": "NVML polled every 3 ms from the start of the device-resident timed region to the end of the per-kernel timing pass (same load throughout)"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().split(",")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [], "power_w_max": float(f[2]),
                    "samples": 1, "source": "nvidia-smi once after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


def cfg1_setup(parts):
    """Host-side preparation shared by both arms: response, LUT, raw bytes, per-block fold phase."""
    import synth
    from dspsr_b200 import hostmath as HM
    from dspsr_b200 import workloads as W
    cfg = W.CFG1
    d, H = HM.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], True)
    assert (d.ndat, d.impulse_pos, d.impulse_neg) == (8192, 457, 459)
    lut, _ = HM.bittable8()
    C, F = cfg["nchan"], d.ndat
    nfilt = d.impulse_pos + d.impulse_neg
    nsamp_fft = 2 * C * F
    overlap = 2 * nfilt * C
    step = nsamp_fft - overlap
    nkeep = F - nfilt
    ndat = parts * step + overlap
    rate_in = 1e6 / cfg["tsamp_us"]
    rate_out = rate_in * F / nsamp_fft
    pred = HM.Polyco(W.polyco_text())
    start = HM.utc_to_mjd(cfg["utc_start"])
    return dict(cfg=cfg, H=H, lut=lut, C=C, F=F, npos=d.impulse_pos, nneg=d.impulse_neg, step=step, overlap=overlap,
                nkeep=nkeep, ndat=ndat, rate_in=rate_in, rate_out=rate_out, pred=pred, start=start, HM=HM,
                synth=synth)


def block_phase(S, first_sample):
    """phi, pps of a block whose first input sample is `first_sample` (Filterbank.C:370 + Fold.C:650-657)."""
    HM = S["HM"]
    t_block = HM.mjd_add(S["start"], first_sample / S["rate_in"] + S["npos"] / S["rate_out"])
    return HM.fold_phase(S["pred"], t_block, 0, S["rate_out"])


def make_raw(ndat, seed):
    """Seeded CASPSR bytes; a 16-part random base tiled to length keeps start-up short (throughput is
    data independent; parity is tested elsewhere on fully random data)."""
    import synth
    base_n = min(ndat, 16 * 3725312 + 468992)
    base_n = (base_n + 3) // 4 * 4
    base = synth.caspsr_bytes(base_n, seed=seed)
    nbytes = (ndat + 3) // 4 * 4 * 2
    reps = -(-nbytes // base.size)
    return np.tile(base, reps)[:nbytes].copy()


def run_reference(args):
    """--impl reference: the restated reference CPU path (oracle/) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    S = cfg1_setup(1)
    ncores = os.cpu_count() or 1
    parts_per_block = 1
    nblock = ncores
    f = O.fb_sizes(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"])
    Ho = O.dedispersion(S["cfg"]["freq"], S["cfg"]["bw"], S["cfg"]["dm"], 1, S["C"], True)[1]
    luto, _ = O.bittable8()
    pipe = O.make_pipe(0, 1, 2, 1, luto, 0.0, f, None, Ho, "Coherence", 4, 1024)
    ndat = nblock * parts_per_block * S["step"] + S["overlap"]
    raw = make_raw(ndat, 1234)
    ph = [block_phase(S, b * parts_per_block * S["step"]) for b in range(nblock)]
    phi = [p[0] for p in ph]
    pps = [p[1] for p in ph]
    for _ in range(args.warmup):
        O.pipe_run(pipe, raw, min(nblock, ncores), parts_per_block, phi, pps, nthread=ncores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.pipe_run(pipe, raw, nblock, parts_per_block, phi, pps, nthread=ncores)
    dt = time.perf_counter() - t0
    samples = args.steps * nblock * parts_per_block * S["step"]
    v = samples / dt / 1e6
    sample = "%d blocks x %d part(s) of cfg1 per step on %d threads" % (nblock, parts_per_block, ncores)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": S["cfg"]["name"], "note": "restated reference CPU path (oracle/, not FFTW), dspsr -t P style"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "real_time_factor": v / 800.0,
    }))


def cpu_baseline_leg(S, budget_s=12.0):
    import oracle as O
    ncores = os.cpu_count() or 1
    f = O.fb_sizes(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"])
    Ho = O.dedispersion(S["cfg"]["freq"], S["cfg"]["bw"], S["cfg"]["dm"], 1, S["C"], True)[1]
    luto, _ = O.bittable8()
    pipe = O.make_pipe(0, 1, 2, 1, luto, 0.0, f, None, Ho, "Coherence", 4, 1024)
    nblock = ncores
    raw = make_raw(nblock * S["step"] + S["overlap"], 1234)
    ph = [block_phase(S, b * S["step"]) for b in range(nblock)]
    phi, pps = [p[0] for p in ph], [p[1] for p in ph]
    O.pipe_run(pipe, raw, 1, 1, phi, pps, nthread=1)      # plan/twiddle warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.pipe_run(pipe, raw, nblock, 1, phi, pps, nthread=ncores)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 64:
            break
    dt = time.perf_counter() - t0
    v = reps * nblock * S["step"] / dt / 1e6
    return {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": "%d x (%d blocks x 1 part of cfg1) on %d threads, %.1f s; restated reference path, not FFTW"
                      % (reps, nblock, ncores, dt)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dspsr_b200 import _lib as L
    from dspsr_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    numa_cores = 0
    if world > 1:
        # one process per GPU: stay on the GPU's NUMA node before any pinned buffer is allocated (N = 1 keeps all
        # cores: the cpu_baseline leg runs there)
        from dspsr_b200 import sharding
        numa_cores = sharding.bind_cpu_affinity(local)
    if world > 1:
        # NCCL writes its version / INFO lines to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parts = args.parts
    S = cfg1_setup(parts)
    stream = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        ctx = E.Context(local, stream)
        ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, S["lut"])
        fd, keep = E.make_fb_desc(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"], S["H"], args.batch)
        pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, 1024)
        # time-block sharding (SURVEY 8e ii): rank g owns parts [g*parts, (g+1)*parts) of the stream,
        # re-reading nsamp_overlap samples at its left edge; weak scaling (fixed work per GPU).
        first_sample = rank * parts * S["step"]
        raw = make_raw(S["ndat"], 1234 + rank)
        h_raw = torch.from_numpy(raw).pin_memory()
        d_raw = h_raw.to("cuda", non_blocking=True)
        phi, pps = block_phase(S, first_sample)
        prof_dev = pipe.fold.device_profile()
        hits_dev = pipe.fold.device_hits()
        h_prof = torch.empty(prof_dev.numel(), dtype=torch.float32).pin_memory()

        def reduce_subint():
            # sub-integration boundary: sum the per-GPU PhaseSeries (PhaseSeries::combine, PhaseSeries.C:442-480)
            if world > 1:
                dist.reduce(prof_dev, 0, op=dist.ReduceOp.SUM)
                dist.reduce(hits_dev, 0, op=dist.ReduceOp.SUM)

        def step_resident():
            pipe.execute(d_raw, parts, phi, pps, first_sample=0)
            reduce_subint()

        def step_e2e():
            pipe.execute_host(h_raw, parts, phi, pps, 0)
            reduce_subint()
            h_prof.copy_(prof_dev, non_blocking=True)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        for _ in range(max(args.warmup, 3)):
            step_resident()
        barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        l0 = ctx.launches
        ms = timed(step_resident, args.steps)
        launches = ctx.launches - l0

        # end to end: host bytes in, profile out (the clock sampler keeps running: both timed regions count)
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

        # per-kernel device times (events around every launch, same stream), one extra pass
        pipe.zero()
        ctx.set_timing(True)
        ctx.read_timing()
        nt = max(2, min(args.steps, 5))
        for _ in range(nt):
            pipe.execute(d_raw, parts, phi, pps, first_sample=0)
        kms, kn = ctx.read_timing()
        ctx.set_timing(False)
        clk = clocks.stop() if rank == 0 else None

        # sanity: the folded result is real (hits add up)
        pipe.zero()
        pipe.execute(d_raw, parts, phi, pps, first_sample=0)
        _, hits, ntot = pipe.synch()
        assert int(hits.sum()) == parts * S["nkeep"] == ntot, "fold hit count mismatch"

    samples_step = parts * S["step"]
    value = world * samples_step * args.steps / (ms * 1e-3) / 1e6
    e2e_value = world * samples_step * args.steps / (ms_e2e * 1e-3) / 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    Nc = S["C"] * S["F"]
    npol = 2
    # algorithmic bytes per part (both pols), SURVEY 8(d): raw incl. overlap + spectrum write + spectrum read
    raw_b = 2 * Nc * 1 * npol            # nsamp_fft real samples of 1 byte per pol
    spec_b = 8 * Nc * npol
    alg = {"cols_fwd": raw_b + spec_b, "rows": 2 * spec_b, "inverse": spec_b}
    kinfo = {}
    for k in ("cols_fwd", "rows", "inverse", "bins"):
        if kn.get(k):
            per = kms[k] / kn[k]
            kinfo[k] = {"launches_per_step": kn[k] / nt, "ms_per_launch": per, "ms_per_step": kms[k] / nt}
    tot = sum(v["ms_per_step"] for v in kinfo.values())
    for k, v in kinfo.items():
        v["share"] = v["ms_per_step"] / tot
        if k in alg:
            parts_per_launch = parts / v["launches_per_step"]
            v["alg_bytes_per_launch"] = alg[k] * parts_per_launch
            v["gbs"] = v["alg_bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9
    dom = max((k for k in kinfo if k in alg), key=lambda k: kinfo[k]["ms_per_step"])
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (profiles/)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if dom in tj:
            traffic = tj[dom]["dram_bytes_per_launch"] / tj[dom]["parts_per_launch"] * (parts / kinfo[dom]["launches_per_step"])
    roof = {"bound": "hbm", "kernel": {"cols_fwd": "k_cols_fwd (K1 unpack+column FFT)", "rows": "k_rows (K2 row FFT+split+chirp)",
                                        "inverse": "k_chan_inv (K3 inverse FFT+detect+fold)"}[dom],
            "achieved": kinfo[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kinfo[dom]["gbs"] / peak,
            "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)",
            "alg_bytes_per_launch": kinfo[dom]["alg_bytes_per_launch"],
            "peak_source": peak_src, "share_of_step": kinfo[dom]["share"]}
    b_alg = (raw_b + 2 * spec_b) / npol / S["step"]          # 10.13 B per sample per pol
    path_gbs = b_alg * npol * samples_step / (ms / args.steps * 1e-3) / 1e9 * 1.0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": S["cfg"]["name"], "parts_per_step": parts, "samples_per_pol_per_step": samples_step,
                   "batch_parts": pipe.info.batch_npart, "sharding": "time blocks with overlap re-read (nchan=1)", "numa_bound_cores": numa_cores,
                   "l2": "inputs larger than L2: %d MB raw per step" % (raw.nbytes // 1000000)},
        "real_time_factor": value / 800.0,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(raw.nbytes),
                "d2h_bytes_per_step": int(h_prof.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "roofline_path": {"bound": "hbm", "alg_bytes_per_sample_per_pol": b_alg, "achieved": path_gbs, "peak": peak,
                          "unit": "GB/s", "frac": path_gbs / peak, "peak_source": peak_src,
                          # FP32 side (5 N log2 N convention): 95.7 flop per sample per pol
                          "fp32_alg_flop_per_sample_per_pol": 5.0 * Nc * (np.log2(Nc) + np.log2(S["F"])) / S["step"],
                          "fp32_achieved_tflops": 5.0 * Nc * (np.log2(Nc) + np.log2(S["F"])) / S["step"] * value * 1e6 * npol / 1e12,
                          "fp32_peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12},
        "kernels": kinfo,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_leg(S)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else libraries print (NCCL's
    version banner, torchrun chatter) was redirected to stderr in main()."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--parts", type=int, default=74,
                    help="overlap-save parts per step (per GPU); 74 = two internal batches of 37 parts (full waves on 148 SMs)")
    ap.add_argument("--batch", type=int, default=0, help="parts per internal kernel batch (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
