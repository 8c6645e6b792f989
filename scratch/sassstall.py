"""Top SASS instructions by stall samples, with preceding context, from `ncu --page source --csv --print-source cuda,sass`."""
import csv,sys,re
rows=list(csv.reader(open(sys.argv[1]))); topn=int(sys.argv[2]) if len(sys.argv)>2 else 25; ctx=int(sys.argv[3]) if len(sys.argv)>3 else 0
hdr=rows[2]; ai=hdr.index("Address"); si=hdr.index("# Samples"); srcs=[i for i,h in enumerate(hdr) if h=="Source"]
names=["stall_barrier","stall_lg","stall_long_sb","stall_math","stall_mio","stall_short_sb","stall_wait","stall_not_selected","stall_selected","stall_dispatch","stall_branch_resolving","stall_no_inst"]
idx={n:hdr.index(n) for n in names}
seen=set(); sass=[]
for r in rows[3:]:
    if len(r)>si and re.fullmatch(r'0x[0-9a-f]+', r[ai] or '') and r[ai] not in seen:
        seen.add(r[ai]); sass.append(r)
sass.sort(key=lambda r:int(r[ai],16))
def n(r):
    try: return int(r[si])
    except: return 0
tot=sum(n(r) for r in sass); print("instructions",len(sass),"samples",tot)
order=sorted(range(len(sass)), key=lambda i:-n(sass[i]))[:topn]
for i in order:
    r=sass[i]
    top=sorted(((k[6:],int(r[v] or 0)) for k,v in idx.items()), key=lambda x:-x[1])[:2]
    for c in range(max(0,i-ctx),i): print("      ", sass[c][ai][-5:], sass[c][srcs[1]][:100])
    print("%5.2f%% %s %-100s %s"%(100*n(r)/tot, r[ai][-5:], r[srcs[1]][:100], top))
