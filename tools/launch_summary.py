#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): ms, launches, share."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
skip = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sim_kernel", "dpx_kernel"]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    if any(x in k for x in skip):
        continue
    v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}[d["Metric Unit"]]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1]:9.3f} ms {a[0]:5d} {100 * a[1] / tot:5.1f} %  {k}")
print(f"{tot:9.3f} ms total")
