#!/usr/bin/env python
"""Where a kernel's instructions go: groups the SASS of an `ncu --page source --csv` export into runs of similar
execution count (loop bodies) and prints their share of the executed instructions and of the stall samples.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python scripts/sass_segments.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
data = []
for r in rows[rows.index(hdr) + 1:]:            # first kernel of the export only
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r != hdr:
        data.append(r)
iS, iE, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iN]) for r in data)
print("total warp instructions", tot, "samples", ts, "SASS lines", len(data))
seg = []; cur = None
for i, r in enumerate(data):
    e = int(r[iE])
    if cur is None or not (0.7 * cur[2] <= e <= 1.4 * cur[2]):
        cur = [i, i, e, 0, 0, 0]; seg.append(cur)
    cur[1] = i; cur[3] += e; cur[4] += int(r[iN]); cur[5] += int(r[iT])
for s in seg:
    if s[3] > minshare * tot:
        print(f"sass[{s[0]:4d}-{s[1]:4d}] n={s[1]-s[0]+1:4d} exec/inst~{s[2]:>10d} inst={s[3]/tot*100:5.1f}% samples={s[4]/ts*100:5.1f}% thr/inst={s[5]/max(s[3],1):.1f}  first: {data[s[0]][iS].strip()[:60]}")
