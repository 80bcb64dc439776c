#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` export: executed warp-instructions and stall samples by opcode, and the
hottest SASS lines.  Usage: ncu_source_hot.py file.csv [B rows per launch]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
B = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
# find header row
for hi, r in enumerate(rows):
    if r and r[0] == "Address":
        break
hdr = rows[hi]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
seen = False
ops = collections.Counter(); smp = collections.Counter(); lines = []
for r in rows[hi + 1:]:
    if len(r) <= iex or not r[ia].strip().isdigit() and not r[ia].startswith("0x"):
        # second section (source-level) starts: stop at first non-SASS block
        if seen and r and r[0] == "Address":
            break
        continue
    seen = True
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    try:
        ex = float(r[iex]); sm = float(r[ismp])
    except ValueError:
        continue
    ops[op] += ex; smp[op] += sm; lines.append((ex, sm, r[ia], src))
tot = sum(ops.values()); tots = sum(smp.values())
print(f"total warp-inst {tot:.3e} ({tot / B:.1f} per row), samples {tots:.0f}")
for op, ex in ops.most_common(28):
    print(f"  {op:10s} {ex / B:9.1f}/row {100 * ex / tot:5.1f}%   stall-samples {100 * smp[op] / max(tots, 1):5.1f}%")
print("hottest lines by stall samples:")
for ex, sm, a, src in sorted(lines, key=lambda t: -t[1])[:25]:
    print(f"  {sm:7.0f} {ex / B:8.1f}/row  {a} {src[:90]}")
