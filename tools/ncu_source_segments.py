#!/usr/bin/env python3
"""Split an `ncu --page source --csv` export of the stage-1+2 kernel at its BAR.SYNC instructions: executed warp-instructions
per row, stall samples and the opcode mix of every barrier-to-barrier segment, plus the eight hottest SASS lines - "who waits
for whom" (a segment with few instructions and many samples is a wait for the slowest warp of the segment before it).
Usage: ncu_source_segments.py file.src.csv   (B = 65536 rows per launch assumed)"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
for hi,r in enumerate(rows):
    if r and r[0]=="Address": break
hdr=rows[hi]
ia,isrc,ismp,iex=hdr.index("Address"),hdr.index("Source"),hdr.index("# Samples"),hdr.index("Instructions Executed")
data=[r for r in rows[hi+1:] if len(r)>iex and r[ia].startswith("0x")]
B=65536.0
seg=[];cur=[0,0,0,{},None]
for r in data:
    ex=float(r[iex]); sm=float(r[ismp]); src=r[isrc]
    cur[0]+=ex; cur[1]+=sm; cur[2]+=1
    op=src.split()[1] if src.strip().startswith('@') else src.split()[0]
    op=op.split('.')[0]
    cur[3][op]=cur[3].get(op,0)+ex
    if 'BAR.SYNC' in src:
        cur[4]=(r[ia][-5:],ex/B)
        seg.append(cur); cur=[0,0,0,{},None]
seg.append(cur)
tot=sum(s[0] for s in seg); tots=sum(s[1] for s in seg)
print("total samples",tots, "inst/row", tot/B)
for k,s in enumerate(seg):
    if s[0]==0 and s[1]==0: continue
    top=sorted(s[3].items(), key=lambda kv:-kv[1])[:7]
    print(f"seg {k}: {s[0]/B:8.1f} inst/row ({100*s[0]/tot:4.1f}%), samples {s[1]:6.0f} ({100*s[1]/tots:4.1f}%), ends at BAR {s[4]}  top: "+", ".join(f"{o} {v/B:.0f}" for o,v in top))
lines=sorted(data,key=lambda r:-float(r[ismp]))[:8]
for r in lines:
    print(r[ismp], r[ia][-5:], r[isrc][:80])
