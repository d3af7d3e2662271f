#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, average, share).
    python scripts/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_launches_*.txt]"""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        c = agg.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {f}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us of kernel time (ncu: serialised, cold cache — compare shares)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n:36s} n={c:5d} total_us={t / 1e3:11.1f} avg_us={t / c / 1e3:10.2f} share={t / tot:.3f}")
