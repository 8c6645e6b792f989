"""Aggregates warp-stall samples of one kernel by CUDA source line from `ncu -i X --page source --csv --print-source cuda,sass`."""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
sections=[i for i,r in enumerate(rows) if r and r[0]=="File Path"]
names=["stall_barrier","stall_lg","stall_long_sb","stall_math","stall_mio","stall_short_sb","stall_wait","stall_not_selected","stall_selected","stall_dispatch","stall_branch_resolving","stall_no_inst"]
alld=[]; 
for s in sections:
    fn=rows[s][0+1]; hdr=rows[s+2]
    si=hdr.index("# Samples"); idx={n:hdr.index(n) for n in names}
    j=s+3
    while j<len(rows) and rows[j] and rows[j][0]!="File Path":
        r=rows[j]; j+=1
        try: n=int(r[si]); ln=int(r[0])
        except Exception: continue
        if n>0: alld.append((fn.split('/')[-1],ln,n,r[1][:80],{k:int(r[v] or 0) for k,v in idx.items()}))
tot=sum(d[2] for d in alld); print("total samples",tot)
agg=collections.Counter()
for d in alld:
    for k,v in d[4].items(): agg[k]+=v
print({k:round(100*v/tot,1) for k,v in agg.most_common()})
for d in sorted(alld,key=lambda x:-x[2])[:topn]:
    top=sorted(d[4].items(), key=lambda x:-x[1])[:3]
    print("%-14s %5d %6.2f%%  %-80s %s"%(d[0],d[1],100*d[2]/tot,d[3],[(k[6:],v) for k,v in top]))
