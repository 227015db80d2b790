#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, time
and share for the LAST step (second half of a --profile-mode --steps 1 --warmup 1 run)."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else (v * 1e6 if u == "s" else v))
        out.append((r["Kernel Name"], v))
    return out


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = load(path)
    half = len(rows) // 2
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows[half:]:
        agg[k[:70]][0] += 1
        agg[k[:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if "pg::" in k)
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k.replace("|", "/"), v[0], v[1], 100 * v[1] / tot))
    print("| total (%d launches; pg:: kernels %.1f us = %.1f%%) | | %.1f | |" % (len(rows) - half, mine, 100 * mine / tot, tot))


if __name__ == "__main__":
    main()
