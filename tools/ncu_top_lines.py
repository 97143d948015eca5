"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (first kernel in the report)."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr, data, started = None, [], False
for r in rows:
    if r and r[0] == "Function Name":
        if started:
            break
        started = True
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        data.append((int(r[0]), r[1], int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]),
                     int(r[hdr.index("L1 Wavefronts Shared")])))
    except ValueError:
        pass
agg = {}
for ln, src, ins, smp, wf in data:
    a = agg.setdefault(ln, [src, 0, 0, 0])
    a[1] += ins; a[2] += smp; a[3] += wf
ti, ts, tw = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
print(f"total warp-instructions {ti}  samples {ts}  smem wavefronts {tw}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print(f"L{ln:5d} inst {100 * a[1] / ti:5.1f}%  samples {100 * a[2] / ts:5.1f}%  smem {100 * a[3] / max(tw, 1):5.1f}%  {a[0].strip()[:110]}")
