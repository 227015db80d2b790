#!/usr/bin/env python
"""Share of executed warp instructions and stall samples per 50-line SASS region of one kernel launch."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
which = sys.argv[3] if len(sys.argv) > 3 else "0"
step = int(sys.argv[4]) if len(sys.argv) > 4 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-skip", which, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = [i for i, r in enumerate(rows) if r and "Source" in r][0]
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[ix["Instructions Executed"]].isdigit()]
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
ts = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("warp instructions", tot, "samples", ts)
for a in range(0, len(data), step):
    s = sum(int(r[ix["Instructions Executed"]]) for r in data[a:a + step])
    sm = sum(int(r[ix["# Samples"]] or 0) for r in data[a:a + step])
    print("%5d %5.1f%% inst %5.1f%% samples  %s" % (a, 100 * s / tot, 100 * sm / max(ts, 1), data[a][ix["Source"]].strip()[:70]))
