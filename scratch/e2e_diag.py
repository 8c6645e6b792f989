import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np
import bench
from dspsr_b200 import _lib as L, engine as E
parts = 32
S = bench.cfg1_setup(parts)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = E.Context(0, stream)
    ud = E.make_unpack_desc(L.FMT_CASPSR8, 1, 2, 1, S["lut"])
    fd, keep = E.make_fb_desc(1, 1, 2, S["C"], S["F"], S["npos"], S["nneg"], S["H"], int(os.environ.get("BATCH", "0")))
    pipe = E.Pipeline(ctx, ud, fd, keep, "Coherence", 4, 1024)
    raw = bench.make_raw(S["ndat"], 1)
    h_raw = torch.from_numpy(raw).pin_memory()
    d_raw = h_raw.cuda()
    phi, pps = bench.block_phase(S, 0)
    prof_dev = pipe.fold.device_profile()
    h_prof = torch.empty(prof_dev.numel(), dtype=torch.float32).pin_memory()
    def run(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    print("resident      %.2f ms" % run(lambda: pipe.execute(d_raw, parts, phi, pps, first_sample=0)))
    print("host          %.2f ms" % run(lambda: pipe.execute_host(h_raw, parts, phi, pps, 0)))
    def both():
        pipe.execute_host(h_raw, parts, phi, pps, 0); h_prof.copy_(prof_dev, non_blocking=True)
    print("host + d2h    %.2f ms" % run(both))
    d2 = torch.empty_like(d_raw)
    print("copy only     %.2f ms" % run(lambda: d2.copy_(h_raw, non_blocking=True)))
    # concurrency check: H2D copy on a second torch stream while the resident pipeline runs
    s2 = torch.cuda.Stream()
    def conc():
        with torch.cuda.stream(s2):
            d2.copy_(h_raw, non_blocking=True)
        pipe.execute(d_raw, parts, phi, pps, first_sample=0)
        pipe.execute(d_raw, parts, phi, pps, first_sample=0)
    print("copy || 2x resident  %.2f ms (serial would be %.2f)" % (run(conc), 4.31 + 2 * 2.11))
