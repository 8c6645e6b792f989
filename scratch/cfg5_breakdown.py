"""Per-sub-band kernel times of cfg5 (one pipeline at a time, events around every launch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import workloads as W
from dspsr_b200 import _lib as L, engine as E
import bench
streams, meta = W.plan_streams(sys.argv[1] if len(sys.argv) > 1 else "cfg5", 0, 1, bench.OursApi(), target="gpu")
ctx = E.Context(0)
FMT = {"CASPSR8": L.FMT_CASPSR8, "GENERIC8": L.FMT_GENERIC8, "MEERKAT8": L.FMT_MEERKAT8, "UWB16": L.FMT_UWB16, "TWOBIT": L.FMT_TWOBIT}
for st in streams:
    cfg, S = st["cfg"], st["S"]
    if cfg["format"] == "TWOBIT":
        tb = E.make_twobit_desc(npol=2); ud = E.make_twobit_unpack_desc(tb)
    else:
        ud = E.make_unpack_desc(FMT[cfg["format"]], S["nin"], cfg["npol"], S["ndim"], st["lut"], float(st["scale"]))
    fd, keep = E.make_fb_desc(cfg["input_real"], S["nin"], cfg["npol"], S["C"], S["F"], S["npos"], S["nneg"], st["H"], int(os.environ.get("BATCH", "0")))
    pipe = E.Pipeline(ctx, ud, fd, keep, st["state"], st["dndim"], st["nbin"])
    d_raw = torch.from_numpy(st["raw"]).cuda()
    out = None
    if not st["nbin"]:
        out = torch.empty((pipe.nchan, pipe.dnpol, st["parts"] * S["nkeep"] * pipe.dndim), dtype=torch.float32, device="cuda")
    for _ in range(2):
        pipe.execute(d_raw, st["parts"], st["phi"], st["pps"], out=out)
    ctx.set_timing(True); ctx.read_timing()
    n = 3
    for _ in range(n):
        pipe.execute(d_raw, st["parts"], st["phi"], st["pps"], out=out)
    ms, cnt = ctx.read_timing(); ctx.set_timing(False)
    tot = sum(ms.values()) / n
    samples = st["parts"] * S["step"] * S["nin"]
    print("F=%6d Nc=%8d P=%5d Q=%5d batch=%4d parts=%5d  %s  total %.3f ms  %.1f GS/s" % (
        S["F"], S["Nc"], pipe.info.fft_rows, pipe.info.fft_cols, pipe.info.batch_npart, st["parts"],
        {k: (round(v / n, 3), cnt[k] // n) for k, v in ms.items() if cnt[k]}, tot, samples / tot / 1e6))
    del pipe
