#!/usr/bin/env python
"""SASS census of the product library: per kernel, how many of the instructions that carry the design are there.
  python profiles/sass_census.py [lib.so] > profiles/rN_sass_census.txt
Packed FP32 (FFMA2 / FADD2 / FMUL2: Blackwell's two-lane FP32 issue), 128/256-bit global and shared accesses, the
fire-and-forget reductions of the fold (RED), TMA (UTMALDG / UTMASTG tensor, UBLKCP bulk), distributed-shared-memory
stores and cluster barriers of the cluster kernel, named barriers (BAR with an id), and what is
NOT there (no HMMA / tcgen05: FFT butterflies are not a dense contraction)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dspsr_b200/libb200dsp.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern = None
counts = collections.OrderedDict()
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = pat.match(line)
    if m and kern:
        counts[kern][m.group(1)] += 1
keys = [("FFMA2", lambda o: o.startswith("FFMA2")), ("FADD2", lambda o: o.startswith("FADD2")),
        ("FMUL2", lambda o: o.startswith("FMUL2")), ("FFMA", lambda o: o.startswith("FFMA") and not o.startswith("FFMA2")),
        ("FADD", lambda o: o.startswith("FADD") and not o.startswith("FADD2")),
        ("FMUL", lambda o: o.startswith("FMUL") and not o.startswith("FMUL2")),
        ("LDG.256", lambda o: o.startswith("LDG") and ".256" in o), ("LDG.128", lambda o: o.startswith("LDG") and ".128" in o),
        ("LDG.other", lambda o: o.startswith("LDG") and ".128" not in o and ".256" not in o),
        ("STG.128", lambda o: o.startswith("STG") and ".128" in o), ("STG.other", lambda o: o.startswith("STG") and ".128" not in o),
        ("LDS.128", lambda o: o.startswith("LDS") and ".128" in o), ("STS.128", lambda o: o.startswith("STS") and ".128" in o),
        ("LDS.other", lambda o: o.startswith("LDS") and ".128" not in o), ("STS.other", lambda o: o.startswith("STS") and ".128" not in o),
        ("RED", lambda o: o.startswith("RED")), ("ATOM", lambda o: o.startswith("ATOM")),
        ("BAR", lambda o: o.startswith("BAR")), ("SHFL", lambda o: o.startswith("SHFL")),
        ("UTMALDG", lambda o: o.startswith("UTMALDG")), ("UTMASTG", lambda o: o.startswith("UTMASTG")),
        ("UBLKCP", lambda o: o.startswith("UBLKCP")),            # cp.async.bulk: K2's 64 KiB image stores
        ("ST.dsmem", lambda o: o.startswith("ST.E")),            # st.shared::cluster through a mapa address (cluster kernel)
        ("UCGABAR", lambda o: o.startswith("UCGABAR")),          # barrier.cluster arrive / wait
        ("HMMA/tcgen05", lambda o: o.startswith("HMMA") or o.startswith("UTC") or "MMA" in o)]
import subprocess as sp
def demangle(n):
    try:
        return sp.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
    except Exception:
        return n
print("# %s: %d kernels" % (lib, len(counts)))
print("# columns: " + " ".join(k for k, _ in keys) + " | total")
for kname, c in counts.items():
    tot = sum(c.values())
    if tot < 200:
        continue
    row = [sum(v for o, v in c.items() if f(o)) for _, f in keys]
    print("%-112s %s | %d" % (demangle(kname), " ".join("%5d" % x for x in row), tot))
