#!/usr/bin/env python
"""Hot spots of one kernel from `ncu --page source --csv` output saved on the GPU box:
    python tools/src_csv.py gpurun_out/src_<kernel>.csv [top]
Prints totals, then the SASS lines holding the most stall samples (with executed-instruction counts)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = lambda r: int(r[ix["# Samples"]] or 0)
I = lambda r: int(r[ix["Instructions Executed"]] or 0)
tot, insts = sum(map(S, data)), sum(map(I, data))
print(rows[0][1][:100]); print("samples", tot, "warp instructions", insts, "SASS lines", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print("stall mix:", {k[6:]: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -S(data[i]))[:top]
for i in sorted(order):
    r = data[i]
    main = max(stalls, key=lambda h: int(r[ix[h]] or 0))
    print("%5d %6.2f%% inst=%5.2f%% thr=%4s %-12s %s" % (i, 100.0 * S(r) / max(tot, 1), 100.0 * I(r) / max(insts, 1),
          r[ix["Avg. Threads Executed"]][:4], main[6:], r[ix["Source"]].strip()[:100]))
