"""Prints the event trace of CTA 0 of the tcgen05 contraction (cl_set_option dbg = 12 -> oz_trace.txt): per tile, the time of
every event relative to issuer A's tile start.  python tools/oz_trace_view.py oz_trace.txt [first_tile] [n_tiles]"""
import sys
ev = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
names = {10: "A handover seen", 1010: "B touched seen", 20: "A first block issued", 1020: "B first block issued", 70: "** last level complete (MMAs of the tile done)", 1: "A tile", 40: "A last-blk", 41: "A issued", 1001: "B tile", 1040: "B last-blk", 1041: "B issued", }
for g in range(3):
    names[10 + g] = f"A grp{g} free"; names[1010 + g] = f"B grp{g} go"
    for k in range(2):
        names[20 + 4 * k + g] = f"A head b{k} g{g}"; names[1020 + 4 * k + g] = f"B head b{k} g{g}"
for l in range(7):
    names[50 + l] = f"A lvl{l} empty seen"; names[60 + l] = f"A head round {l} issued"
for w in range(8):
    names[2020 + 100 * w] = f"E{w} done"
    for l in range(7):
        names[2000 + 100 * w + l] = f"E{w} lvl{l} full"; names[2010 + 100 * w + l] = f"E{w} lvl{l} read"; names[2030 + 100 * w + l] = f"E{w} lvl{l} folded"
starts = [t for c, t in ev if c == 1]
print("tiles", len(starts), "mean tile", (starts[-1] - starts[0]) / max(1, len(starts) - 1))
ev.sort(key=lambda e: e[1])
for ti in range(first, first + n):
    t0, t1 = starts[ti], starts[ti + 1]
    print(f"--- tile {ti}: {t1 - t0} cycles")
    for c, t in ev:
        if t0 - 3000 <= t < t1:
            print(f"  {t - t0:8d}  {names.get(c, c)}")
