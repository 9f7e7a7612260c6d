#!/usr/bin/env python
"""Summarises the source page of an .ncu-rep: top SASS lines by stall samples, opcode mix.
    python scripts/ncu_hot.py gpurun_out/prof_x.ncu-rep [top_n] [lo hi]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]     # one table per captured launch
hdr = rows[heads[0]]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
S = lambda r: int(r[idx["# Samples"]])
I = lambda r: int(r[idx["Instructions Executed"]])
print("lines", len(data), "samples", sum(map(S, data)), "inst", sum(map(I, data)))
c = collections.Counter()
for r in data[lo:hi]:
    for k in hdr:
        if k.startswith("stall_") and "(" not in k:
            c[k] += int(r[idx[k]])
print("stalls", [(k, v) for k, v in c.most_common(8)])
top = sorted(range(lo, hi), key=lambda i: -S(data[i]))[:topn]
for i in sorted(top):
    r = data[i]
    st = {k[6:]: int(r[idx[k]]) for k in hdr if k.startswith("stall_") and "(" not in k and int(r[idx[k]]) > 2}
    print(i, r[idx["Source"]].strip()[:64].ljust(64), S(r), I(r), st)
