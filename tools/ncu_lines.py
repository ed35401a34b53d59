#!/usr/bin/env python
"""Per-source-line totals of an `ncu --page source --csv --print-source sass,cuda` dump: warp instructions, stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    if not d["Line No"].strip():
        continue                      # the SASS rows under a source line: counted with the line already
    try:
        ie = int(d["Instructions Executed"] or 0); ns = int(d["# Samples"] or 0); te = int(d["Thread Instructions Executed"] or 0)
    except ValueError:
        continue
    a = agg[(cur_file, d["Line No"])]
    a[0] += ie; a[1] += ns; a[2] += te; a[3] = d["Source"][:110]
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instr {tot_i}, samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{100 * a[1] / max(tot_s, 1):5.1f}% smp {100 * a[0] / max(tot_i, 1):5.1f}% ins lanes {a[2] / max(a[0], 1):4.1f}  {f}:{ln}  {a[3]}")
