import csv, sys, collections, re
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix = collections.Counter(); smp = collections.Counter()
tot = 0; tots = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(2).split(".")[0]
    # keep some modifiers
    full = m.group(2)
    if op in ("LDS","STS","LDG","STG","LDL","STL","ATOMS","RED","ATOMG","BAR"): op = ".".join(full.split(".")[:1])
    n = int(float(r[iex] or 0)); s = int(float(r[ismp] or 0))
    mix[op] += n; smp[op] += s; tot += n; tots += s
print("total warp instr", tot, "samples", tots)
for op, n in mix.most_common(28):
    print("%-10s %12d %5.1f%%   samples %5.1f%%" % (op, n, 100.0*n/tot, 100.0*smp[op]/max(tots,1)))
