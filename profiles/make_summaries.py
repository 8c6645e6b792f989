#!/usr/bin/env python
"""Regenerates the tracked summaries of one profile state from the gpurun_out/ captures:
  python profiles/make_summaries.py s11 gpurun_out/launches_r1_s11.csv gpurun_out/prof_r1_s11f.ncu-rep PARTS_PER_LAUNCH
writes profiles/r1_<state>_launches.{csv,summary.txt}, r1_<state>_ncu_full.summary.txt and traffic.json."""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    state, launches, rep, ppl = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    dst_csv = os.path.join(HERE, "r1_%s_launches.csv" % state)
    shutil.copy(launches, dst_csv)
    rows = [r for r in csv.reader(l for l in open(dst_csv) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        t = tot.setdefault(n, [0, 0.0])
        t[0] += 1
        t[1] += v
    s = sum(t[1] for t in tot.values())
    out = ["# per-kernel totals of profiles/r1_%s_launches.csv (ncu --metrics gpu__time_duration.sum --clock-control none, "
           "bench.py --steps 2 --warmup 3 --no-cpu)" % state]
    for n, t in sorted(tot.items(), key=lambda x: -x[1][1]):
        out.append("%-60s launches %4d  total %10.1f us  share %5.1f%%  mean %8.1f us" % (n, t[0], t[1], 100 * t[1] / s, t[1] / t[0]))
    open(os.path.join(HERE, "r1_%s_launches.summary.txt" % state), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    summ = subprocess.run([sys.executable, os.path.join(HERE, "summarize.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(HERE, "r1_%s_ncu_full.summary.txt" % state), "w").write(summ)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, u = rr[0], rr[1]
    tr = {}
    for r in rr[2:]:
        d, un = dict(zip(h, r)), dict(zip(h, u))

        def b(k):
            return float(d[k].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un[k]]
        key = "cols_fwd" if "k1_c2" in d["Kernel Name"] else "rows" if "k2_r32" in d["Kernel Name"] else "inverse"
        tr[key] = {"dram_bytes_per_launch": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"), "parts_per_launch": ppl,
                   "source": "profiles/r1_%s_ncu_full.summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % state}
    json.dump(tr, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print({k: v["dram_bytes_per_launch"] for k, v in tr.items()})


if __name__ == "__main__":
    main()
