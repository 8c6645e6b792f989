#!/usr/bin/env python
"""bench.py -- throughput of dspsr's baseband hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1..cfg5]
                  [--parts P] [--blocks B] [--batch b] [--no-cpu]

Workload (config.workload): by default BASELINE configs[0] / SURVEY cfg1 -- Benchmark/bench.csh's CASPSR
8-bit dual-pol 400 MHz real-sampled band, `dspsr -F 256:D` coherent filterbank at DM 67.99, Coherence
detection, fold into 1024 bins with Benchmark/vela.polyco.  --workload cfg2..cfg5 selects the other
BASELINE configurations at their largest single-GPU shard.

One STEP = B blocks; one block = one pass of unpack -> filterbank/convolution (dedispersion) -> detect -> fold
over P overlap-save parts of seeded synthetic noise per pipeline, followed by the sub-integration combine
(NCCL reduce / gather of the PhaseSeries when N > 1).  B is chosen so that a step lasts about 0.1 s: the
timed region of the default run is >= 2 s.

Printed JSON line (rank 0):
  value      input MSamples/s (samples per polarisation per second), inputs resident in HBM
  e2e        same metric through b200_pipeline_execute_host: pinned HOST bytes -> device copy ->
             kernels -> device->host read of the folded profile (cfg2: of the 8-bit filterbank bytes), all inside
             the timed region
  roofline   SURVEY 8(d): algorithmic bytes of the path (ONE spectrum round trip) x samples / step time vs the
             measured HBM peak; the dominant kernel is named with its own algorithmic rate and its measured
             DRAM traffic (ncu, profiles/traffic.json); traffic_over_alg = measured DRAM bytes / algorithmic bytes
  cpu_baseline   the restated reference CPU path (oracle/, NOT FFTW) on a bounded sample
--impl reference times that CPU path alone with all host threads (no import of the product).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

UNIT = "MSamples/s"
# developer build only: cfg4 on the generic three-kernel path (A/B runs); the traffic figures of the long-transform kernels do not apply
W_GENERIC_LONG = os.environ.get("B200_BIG_CONV") == "0"


def metric_name(rate_in):
    return "input MSamples/s (per polarisation); real-time factor = value / %g" % (rate_in / 1e6)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled
    from a thread every few milliseconds; falls back to one `nvidia-smi` query if NVML cannot be loaded."""

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
            time.sleep(0.01)

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
        return {"sm_mhz": float(np.median([x[0] for x in self.samples])), "sm_min_mhz": float(min(x[0] for x in self.samples)),
                "sm_max_mhz": float(mx) if mx else None,
                "reasons": sorted(seen), "power_w_max": max(x[2] for x in self.samples), "samples": len(self.samples),
                "source": "NVML polled every 10 ms over the device-resident and the end-to-end timed regions"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().split(",")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [], "power_w_max": float(f[2]),
                    "samples": 1, "source": "nvidia-smi once after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


# ------------------------------------------------------------------------------------------------------------
# host-maths adapters: the product's (libb200dsp.so) for our arm, the oracle's for the reference arm
# ------------------------------------------------------------------------------------------------------------
class OursApi:
    def __init__(self):
        from dspsr_b200 import hostmath as HM
        self.HM = HM

    def bittable8(self):
        return self.HM.bittable8()

    def dedispersion(self, cf, bw, dm, nin, nchan, real, fres=0, build=True):
        return self.HM.dedispersion(cf, bw, dm, nin, nchan, real, fres, build=build)

    def dedispersion_channels(self, d, chan0, n):
        return self.HM.dedispersion_channels(d, chan0, n)

    def predictor(self, text):
        return self.HM.Polyco(text)


class OracleApi:
    def __init__(self):
        import oracle as O
        self.O = O

    def bittable8(self):
        return self.O.bittable8()

    def dedispersion(self, cf, bw, dm, nin, nchan, real, fres=0, build=True):
        return self.O.dedispersion(cf, bw, dm, nin, nchan, real, fres, build=build)

    def dedispersion_channels(self, d, chan0, n):
        # the oracle builds whole responses (Dedispersion::build); a rank's rows are a slice
        O = self.O
        import ctypes as C
        H = np.zeros((d.nchan, d.ndat), np.complex64)
        O.lib().orc_dedisp_build(C.byref(d), H.ctypes.data_as(C.c_void_p))
        return np.ascontiguousarray(H[chan0:chan0 + n])

    def predictor(self, text):
        O = self.O
        pc = O.polyco_parse(text)

        class P:
            def phase(self, mjd):
                return O.polyco_phase(pc, *mjd)[0]

            def frequency(self, mjd):
                return O.polyco_frequency(pc, *mjd)
        return P()


ORACLE_FMT = {"CASPSR8": 0, "GENERIC8": 1, "MEERKAT8": 2, "UWB16": 3, "TWOBIT": 5}


def oracle_pipe(O, st):
    """The oracle's pipeline object of a stream (make_pipe keeps the arrays alive)."""
    cfg, S = st["cfg"], st["S"]
    fbs = convs = None
    if cfg["filterbank"]:
        fbs = O.fb_sizes(int(cfg["input_real"]), S["nin"], cfg["npol"], S["nin"] * S["C"], S["F"], S["npos"], S["nneg"])
    else:
        convs = O.conv_sizes(int(cfg["input_real"]), S["nin"], cfg["npol"], S["F"], S["npos"], S["nneg"])
    tb = O.TwoBit() if cfg["format"] == "TWOBIT" else None
    return O.make_pipe(ORACLE_FMT[cfg["format"]], S["nin"], cfg["npol"], S["ndim"], st["lut"], float(st["scale"]), fbs, convs,
                       st["H"], st["state"], st["dndim"], st["nbin"], twobit=tb)


def cpu_run(O, plans, nthread, reps=1):
    """One CPU pass: every stream's blocks, `dspsr -t nthread` style.  plans: [(pipe, stream, nblock)]."""
    n = 0
    for _ in range(reps):
        for pipe, st, nblock in plans:
            O.pipe_run(pipe, st["raw"], nblock, st["parts"], [st["phi"]] * nblock, [st["pps"]] * nblock, nthread=nthread)
            n += nblock * st["parts"] * st["S"]["step"] * st["S"]["nin"]
    return n


def cpu_plans(O, W, workload, ncores):
    """Streams sized for the CPU: each of the `ncores` threads takes one block of every stream."""
    api = OracleApi()
    streams, meta = W.plan_streams(workload, 0, 1, api, target="cpu")
    plans = []
    for st in streams:
        nblock = ncores
        ndat = nblock * st["parts"] * st["S"]["step"] + st["S"]["overlap"]
        st["raw"] = W.raw_bytes(st["cfg"], st["S"], ndat, 1234)
        plans.append((oracle_pipe(O, st), st, nblock))
    return plans, meta


def run_reference(args):
    """--impl reference: the restated reference CPU path (oracle/) with all host threads.  Imports nothing of
    the product: set-up arithmetic from tests/workloads.py, host maths and kernels from oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    import workloads as W
    ncores = os.cpu_count() or 1
    plans, meta = cpu_plans(O, W, args.workload, ncores)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_run(O, plans, ncores)
    t0 = time.perf_counter()
    samples = 0
    for _ in range(args.steps):
        samples += cpu_run(O, plans, ncores)
    dt = time.perf_counter() - t0
    v = samples / dt / 1e6
    name = plans[0][1]["cfg"]["name"] if args.workload != "cfg5" else "cfg5: UWL-like 26 x 128 MHz sub-bands, 16-bit dual-pol complex, -F 128:D, fold 1024 bins"
    sample = "%d thread(s) x 1 block of %s part(s) per stream (%d stream(s)) per step" % (
        ncores, "/".join(str(p[1]["parts"]) for p in plans[:3]) + ("/..." if len(plans) > 3 else ""), len(plans))
    emit(json.dumps({
        "impl": "reference", "metric": metric_name(meta["rate_in"]), "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "note": "restated reference CPU path (oracle/, not FFTW), dspsr -t P style"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "real_time_factor": v / (meta["rate_in"] / 1e6 * meta["nchan_samples"]),
    }))


def cpu_baseline_leg(workload, budget_s=12.0):
    import oracle as O
    import workloads as W
    ncores = os.cpu_count() or 1
    plans, meta = cpu_plans(O, W, workload, ncores)
    cpu_run(O, [(p, s, 1) for p, s, _ in plans[:1]], 1)      # plan/twiddle warm-up
    t0 = time.perf_counter()
    reps, samples = 0, 0
    while True:
        samples += cpu_run(O, plans, ncores)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 64:
            break
    dt = time.perf_counter() - t0
    return {"value": samples / dt / 1e6, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": "%d x (%d blocks of %d part(s) x %d stream(s)) on %d threads, %.1f s; restated reference path, not FFTW"
                      % (reps, ncores, plans[0][1]["parts"], len(plans), ncores, dt)}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import workloads as W
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
        # NCCL writes its version / INFO lines to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    streams, meta = W.plan_streams(args.workload, rank, world, OursApi(), parts=args.parts or None, target="gpu")
    FMT = {"CASPSR8": L.FMT_CASPSR8, "GENERIC8": L.FMT_GENERIC8, "MEERKAT8": L.FMT_MEERKAT8, "UWB16": L.FMT_UWB16,
           "TWOBIT": L.FMT_TWOBIT}
    stream = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        ctx = E.Context(local, stream)
        for st in streams:
            cfg, S = st["cfg"], st["S"]
            if cfg["format"] == "TWOBIT":
                st["tb"] = E.make_twobit_desc(npol=cfg["npol"])
                ud = E.make_twobit_unpack_desc(st["tb"])
            else:
                ud = E.make_unpack_desc(FMT[cfg["format"]], S["nin"], cfg["npol"], S["ndim"], st["lut"], float(st["scale"]))
            fd, keep = E.make_fb_desc(cfg["input_real"], S["nin"], cfg["npol"], S["C"], S["F"], S["npos"], S["nneg"],
                                      st["H"], args.batch)
            st["pipe"] = pipe = E.Pipeline(ctx, ud, fd, keep, st["state"], st["dndim"], st["nbin"])
            if args.deterministic and st["nbin"]:
                pipe.set_deterministic()
            st["h_raw"] = torch.from_numpy(st["raw"]).pin_memory()
            st["d_raw"] = st["h_raw"].to("cuda", non_blocking=True)
            st["H"] = None                                    # the plan holds its own copy
            if st["nbin"]:
                st["prof_dev"] = pipe.fold.device_profile()
                st["hits_dev"] = pipe.fold.device_hits()
                st["h_prof"] = torch.empty(st["prof_dev"].numel(), dtype=torch.float32).pin_memory()
            else:
                # digifil tail: detected floats -> Rescale -> SigProcDigitizer 8-bit TFP bytes
                nout = st["parts"] * S["nkeep"]
                st["det"] = torch.empty((pipe.nchan, pipe.dnpol, nout * pipe.dndim), dtype=torch.float32, device="cuda")
                st["rescale"] = E.Rescale(ctx, pipe.nchan, pipe.dnpol, interval_samples=0)
                st["fil"] = torch.empty((nout, pipe.dnpol, pipe.nchan), dtype=torch.uint8, device="cuda")
                st["h_fil"] = torch.empty(st["fil"].numel(), dtype=torch.uint8).pin_memory()
        combine = meta["combine"] if world > 1 else "none"
        gather_bufs = None
        if combine == "gather":
            # disjoint [chan][pol][bin][dim] blocks are concatenated on rank 0 (no arithmetic); ragged shards padded
            nloc = torch.tensor([sum(s["prof_dev"].numel() for s in streams)], device="cuda")
            nmax = nloc.clone()
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
            pad = torch.zeros(int(nmax.item()), dtype=torch.float32, device="cuda")
            gather_bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None

        def combine_subint():
            # sub-integration boundary: PhaseSeries::combine across the GPUs (PhaseSeries.C:442-480)
            if combine == "reduce":
                for s in streams:
                    dist.reduce(s["prof_dev"], 0, op=dist.ReduceOp.SUM)
                    dist.reduce(s["hits_dev"], 0, op=dist.ReduceOp.SUM)
            elif combine == "gather":
                o = 0
                for s in streams:
                    n = s["prof_dev"].numel()
                    pad[o:o + n].copy_(s["prof_dev"])
                    o += n
                dist.gather(pad, gather_bufs, dst=0)

        def step_resident():
            for _ in range(blocks):
                for s in streams:
                    s["pipe"].execute(s["d_raw"], s["parts"], s["phi"], s["pps"], first_sample=0, out=s.get("det"))
            combine_subint()

        def step_e2e():
            for _ in range(blocks):
                for s in streams:
                    if s["nbin"]:
                        s["pipe"].execute_host(s["h_raw"], s["parts"], s["phi"], s["pps"], 0)
                    else:
                        s["pipe"].execute_host(s["h_raw"], s["parts"], 0.0, 0.0, 0, out=s["det"])
                        s["rescale"].transform(s["det"], out=s["det"])
                        E.sigproc_digitize8(ctx, s["det"], out=s["fil"])
                        s["h_fil"].copy_(s["fil"].view(-1), non_blocking=True)
            combine_subint()
            for s in streams:
                if s["nbin"]:
                    s["h_prof"].copy_(s["prof_dev"], non_blocking=True)

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

        # blocks per step: about 0.1 s of device time per step (probe with one block), same on every rank
        blocks = 1
        step_resident()
        step_resident()
        probe = timed(step_resident, 2) / 2
        blocks = args.blocks or max(1, min(256, int(round(args.step_ms / max(probe, 1e-3)))))
        if world > 1:
            t = torch.tensor([blocks], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            blocks = int(t.item())

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
        clk = clocks.stop() if rank == 0 else None

        # per-kernel device times (events around every launch, same stream), one extra pass
        for s in streams:
            if s["nbin"]:
                s["pipe"].zero()
        ctx.set_timing(True)
        ctx.read_timing()
        nt = max(2, min(blocks, 5))
        for _ in range(nt):
            for s in streams:
                s["pipe"].execute(s["d_raw"], s["parts"], s["phi"], s["pps"], first_sample=0, out=s.get("det"))
        kms, kn = ctx.read_timing()
        ctx.set_timing(False)

        # sanity AFTER the combine: the folded result on rank 0 is real (hits add up over the ranks)
        check = None
        if streams[0]["nbin"]:
            for s in streams:
                s["pipe"].zero()
            for s in streams:
                s["pipe"].execute(s["d_raw"], s["parts"], s["phi"], s["pps"], first_sample=0)
            combine_subint()
            torch.cuda.synchronize()
            if rank == 0:
                check = []
                for s in streams:
                    _, hits, _ = s["pipe"].synch()
                    want = s["parts"] * s["S"]["nkeep"] * (world if combine == "reduce" else 1)
                    assert int(hits.sum()) == want, "fold hit count after the combine: %d != %d" % (int(hits.sum()), want)
                    check.append(int(hits.sum()))

    # samples per polarisation: NDAT-samples x channels (SURVEY 8 notation)
    samples_block = sum(s["parts"] * s["S"]["step"] * s["S"]["nin"] for s in streams)
    tot = torch.tensor([float(samples_block)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    samples_step_all = float(tot.item()) * blocks
    value = samples_step_all * args.steps / (ms * 1e-3) / 1e6
    e2e_value = samples_step_all * args.steps / (ms_e2e * 1e-3) / 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    # algorithmic bytes / flops of this rank's streams per block
    alg_block = 0.0
    flop_block = 0.0
    kalg = {"cols_fwd": 0.0, "rows": 0.0, "inverse": 0.0}
    for s in streams:
        cfg, S = s["cfg"], s["S"]
        per_unit, _ = W.algorithmic_bytes(cfg, S)
        units = s["parts"] * cfg["npol"] * S["nin"]
        alg_block += per_unit * units
        flop_block += W.algorithmic_flops(cfg, S) * s["parts"] * S["step"] * cfg["npol"] * S["nin"]
        raw_b = S["nsamp_fft"] * cfg["nbit"] * S["ndim"] / 8.0
        spec_b = 8.0 * S["Nc"]
        # per kernel, its share of the ONE algorithmic round trip: K1 reads the raw samples and writes the spectrum,
        # K3 reads it; K2 (a second pass over the spectrum) has no algorithmic bytes of its own
        kalg["cols_fwd"] += (raw_b + spec_b) * units
        kalg["inverse"] += spec_b * units
    step_s = ms / args.steps * 1e-3
    path_gbs = alg_block * blocks / step_s / 1e9
    kinfo = {}
    for k in ("cols_fwd", "rows", "inverse", "bins", "other"):
        if kn.get(k):
            kinfo[k] = {"launches_per_block": kn[k] / nt, "ms_per_launch": kms[k] / kn[k], "ms_per_block": kms[k] / nt}
    ktot = sum(v["ms_per_block"] for v in kinfo.values())
    for k, v in kinfo.items():
        v["share"] = v["ms_per_block"] / ktot
        if kalg.get(k):
            v["alg_bytes_per_launch"] = kalg[k] / v["launches_per_block"]
            v["alg_gbs"] = kalg[k] / (v["ms_per_block"] * 1e-3) / 1e9
    dom = max(kinfo, key=lambda k: kinfo[k]["ms_per_block"])
    knames = {"cols_fwd": "K1 (unpack + column FFT)", "rows": "K2 (row FFT + real split + response)",
              "inverse": "K3 (inverse FFT + discard + detect + fold)", "bins": "bin plan", "other": "stand-alone unpack/detect/fold"}
    # measured DRAM traffic (ncu --set full, profiles/traffic.json: bytes per part of cfg1 for every kernel)
    traffic = traffic_dom = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if args.workload == "cfg1" and os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        per_part = {k: tj[k]["dram_bytes_per_launch"] / tj[k]["parts_per_launch"] for k in ("cols_fwd", "rows", "inverse") if k in tj}
        if len(per_part) == 3:
            traffic = sum(per_part.values()) * streams[0]["parts"]
        if dom in per_part:
            traffic_dom = per_part[dom] * streams[0]["parts"] / kinfo[dom]["launches_per_block"]
    if args.workload == "cfg3" and os.path.exists(tp) and set(kinfo) <= {"inverse", "bins"}:
        # the one-kernel cluster path (clusterconv.cu): all DRAM traffic of a block is that kernel's
        with open(tp) as f:
            tc = json.load(f).get("cfg3_cluster")
        if tc:
            per_tile = tc["dram_bytes_per_launch"] / (tc["parts_per_launch"] * tc["channels"])
            traffic = per_tile * streams[0]["parts"] * streams[0]["S"]["nin"]
            traffic_dom = traffic / kinfo[dom]["launches_per_block"]
    if args.workload == "cfg4" and os.path.exists(tp) and not W_GENERIC_LONG:
        # the long-transform kernels (longconv.cu k_bc_*): one launch of each per block of 16 parts
        with open(tp) as f:
            tl = json.load(f).get("cfg4_long")
        if tl:
            per_part = {k: v / tl["parts_per_launch"] for k, v in tl["dram_bytes_per_launch"].items()}
            traffic = sum(per_part.values()) * streams[0]["parts"]
            if dom in per_part:
                traffic_dom = per_part[dom] * streams[0]["parts"] / kinfo[dom]["launches_per_block"]
    roof = {"bound": "hbm", "achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
            "what": "whole path, SURVEY 8(d): algorithmic bytes (one spectrum round trip) per block / block time",
            "alg_bytes_per_block": alg_block,
            "traffic": traffic, "traffic_unit": "DRAM bytes per block, all kernels (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)",
            "traffic_over_alg": (traffic / alg_block) if traffic else None,
            "dominant_kernel": {"name": knames[dom], "share_of_block": kinfo[dom]["share"],
                                "ms_per_launch": kinfo[dom]["ms_per_launch"],
                                "alg_bytes_per_launch": kinfo[dom].get("alg_bytes_per_launch"),
                                "alg_gbs": kinfo[dom].get("alg_gbs"), "traffic_per_launch": traffic_dom},
            "peak_source": peak_src + "; burst figure (our kernels draw < 40 % of the board power, clocks stay at maximum: see clocks)",
            "fp32": {"alg_flop_per_block": flop_block, "achieved_tflops": flop_block * blocks / step_s / 1e12,
                     "peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12,
                     "note": "5 N log2 N convention; FFT butterflies are add-dominated: at one flop per lane-cycle the pipe limit is half the nominal figure"}}
    s0 = streams[0]
    name = s0["cfg"]["name"] if args.workload != "cfg5" else "cfg5: UWL-like 26 x 128 MHz sub-bands, 16-bit dual-pol complex, -F 128:D, fold 1024 bins"
    rt = meta["rate_in"] / 1e6 * (meta["nchan_samples"] * world if args.workload == "cfg3" else (26 if args.workload == "cfg5" else 1))
    out = {
        "metric": metric_name(meta["rate_in"]), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if meta.get("strong") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "blocks_per_step": blocks,
                   "parts_per_block": [s["parts"] for s in streams] if len(streams) > 1 else s0["parts"],
                   "samples_per_pol_per_step": samples_step_all, "batch_parts": s0["pipe"].info.batch_npart,
                   "sharding": meta["sharding"], "combine": combine, "numa_bound_cores": numa_cores,
                   "fold_accumulation": "fixed point (reproducible)" if args.deterministic else "float RED (default)",
                   "l2": "inputs larger than L2: %d MB raw per block, %d MB of spectrum scratch per batch"
                         % (sum(s["raw"].nbytes for s in streams) // 1000000, s0["pipe"].info.scratch_bytes // 1000000)},
        "real_time_factor": value / rt,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(s["raw"].nbytes for s in streams)) * blocks,
                "d2h_bytes_per_step": int(sum((s["h_prof"].numel() * 4) if s["nbin"] else s["h_fil"].numel() * blocks for s in streams)),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "kernels": kinfo,
        "hits_after_combine": check,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_leg(args.workload)
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--parts", type=int, default=0,
                    help="overlap-save parts per block and GPU (0 = per-workload default; cfg1: 74 = two internal "
                         "batches of 37 parts, full waves on 148 SMs; cfg5: samples per sub-band and block)")
    ap.add_argument("--blocks", type=int, default=0, help="blocks per step (0 = as many as fill --step-ms)")
    ap.add_argument("--step-ms", type=float, default=100.0, help="target device time of one step")
    ap.add_argument("--batch", type=int, default=0, help="parts per internal kernel batch (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--deterministic", action="store_true",
                    help="fold with the reproducible fixed-point accumulator (b200_pipeline_set_deterministic)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
