#!/usr/bin/env python
"""Top stall locations of one kernel of an .ncu-rep (needs --import-source on / -lineinfo).
usage: ncu_stalls.py file.ncu-rep [kernel index in capture] [top N]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r]
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
print(f"{len(blocks)} kernels in capture")
b = blocks[k]
print(b[0][1][:120])
hdr, data = b[1], b[2:]
si, src = hdr.index("# Samples"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in data if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[si]))[:topn]:
    st = {hdr[i][6:]: int(r[i]) for i in stalls if r[i].isdigit() and int(r[i]) > 0}
    main = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  {r[src].strip()[:72]:72s} {main}")
