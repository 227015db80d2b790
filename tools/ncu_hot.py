#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for one kernel of an .ncu-rep (source page)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-skip", str(which), "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
# possibly several kernels: split on "Kernel Name" rows, take the first instance
blocks, cur = [], []
for ln in txt:
    if ln.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [ln]
    else:
        cur.append(ln)
if cur: blocks.append(cur)
b = blocks[0]
print(b[0][:150])
rows = list(csv.reader(b[1:]))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
insts = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "warp instructions", insts, "SASS lines", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    print("%5d %6.2f%% inst=%-10s %s" % (i, 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:110]))
