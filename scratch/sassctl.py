"""Decode the scheduling control fields of sm_100 SASS (cuobjdump -sass output): stall, write/read barrier, wait mask."""
import re,sys
lines=open(sys.argv[1]).read().split('\n')
out=[]
i=0
while i<len(lines):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/',lines[i])
    if m and i+1<len(lines):
        m2=re.match(r'\s+/\* 0x([0-9a-f]{16}) \*/',lines[i+1])
        if m2:
            hi=int(m2.group(1),16)
            stall=(hi>>41)&0xf; y=(hi>>45)&1; wb=(hi>>46)&7; rb=(hi>>49)&7; wait=(hi>>52)&0x3f
            out.append((m.group(1),m.group(2).strip(),stall,y,wb,rb,wait))
            i+=2; continue
    i+=1
lo=int(sys.argv[2],16) if len(sys.argv)>2 else 0; hi_=int(sys.argv[3],16) if len(sys.argv)>3 else 1<<30
for a,ins,stall,y,wb,rb,wait in out:
    if lo<=int(a,16)<=hi_:
        print("%5s st%-2d %s W%s R%s wait[%s]  %s"%(a,stall,'Y' if y else ' ', wb if wb!=7 else '-', rb if rb!=7 else '-', ''.join(str(b) for b in range(6) if wait>>b&1) or '-', ins[:110]))
