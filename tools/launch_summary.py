"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and average duration."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
acc = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v, u = float(r[vi].replace(",", "")), r[ui]
    v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
    acc[r[ki][:48]][0] += 1
    acc[r[ki][:48]][1] += v
tot = sum(t for _, t in acc.values())
for k, (n, t) in sorted(acc.items(), key=lambda x: -x[1][1]):
    print(f"{k:50s} {n:6d} launches {t / 1e3:10.3f} ms total {t / n:9.1f} us avg {100 * t / tot:5.1f} %")
