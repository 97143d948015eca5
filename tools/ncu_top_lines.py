"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (every source file of the report).
Usage: python tools/ncu_top_lines.py dump.csv [topn] [sort: samples|inst|smem]"""
import csv
import os
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sort = sys.argv[3] if len(sys.argv) > 3 else "samples"
rows = list(csv.reader(open(path)))
hdr, fname, agg = None, "?", {}
for r in rows:
    if r and r[0] == "File Path":
        fname = os.path.basename(r[1]).replace("mjxb_", "").replace(".cuh", "")
        continue
    if r and r[0] == "Line No":
        hdr = r
        ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
        cb = hdr.index("stall_barrier")
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1], 0, 0, 0, 0])
        a[1] += int(r[ci]); a[2] += int(r[cs]); a[3] += int(r[cw]); a[4] += int(r[cb])
    except ValueError:
        pass
ti, ts, tw = (sum(a[k] for a in agg.values()) for k in (1, 2, 3))
tb = sum(a[4] for a in agg.values())
print(f"total warp-instructions {ti}  samples {ts} (barrier {100 * tb / max(ts, 1):.1f}%)  smem wavefronts {tw}")
k = {"samples": 2, "inst": 1, "smem": 3}[sort]
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][k])[:topn]:
    print(f"{f[:10]:>10}:{ln:<5d} inst {100 * a[1] / ti:5.2f}%  samples {100 * a[2] / ts:5.2f}%  smem {100 * a[3] / max(tw, 1):5.2f}%  {a[0].strip()[:100]}")
