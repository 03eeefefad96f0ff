#!/usr/bin/env python
"""Per-kernel shares of ONE step from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv):
usage: launch_shares.py launches.csv [steps_in_capture=3] > shares.csv   (takes the last 1/steps of the launches after
the weight load; `bench.py --steps 1 --warmup 1` runs three identical steps: warm-up, timed, profiled)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
recs = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in data if len(r) > vi]
first = [i for i, (k, _) in enumerate(recs) if "gather_blocks" in k][0]
body = recs[first:]
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
step = body[len(body) - len(body) // n_steps:]
agg = collections.OrderedDict()
for k, us in step:
    k = re.sub(r"void tdc::(\(anonymous namespace\)|<unnamed>)::", "", k)
    k = re.sub(r"^void ", "", k).split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_ms,share_pct")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k},{v[0]},{v[1] / 1e3:.3f},{100 * v[1] / tot:.2f}")
print(f"TOTAL,{len(step)},{tot / 1e3:.3f},100.00")
