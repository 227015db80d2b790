#!/usr/bin/env python
"""profiles/<tag>_ncu_summary.md, profiles/<tag>_launches.csv and profiles/traffic.json from one gpurun capture:
    python tools/profile_summary.py TAG gpurun_out/launches_X.csv gpurun_out/prof_X.ncu-rep|gpurun_out/prof_X_raw.csv
The launch list is `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of
`bench.py --profile-mode --steps 1 --warmup 1` (its second half is one step); the report is `ncu --set full` of the
library's main kernels in the same command.  traffic.json = DRAM bytes (read + write) per STEP and kernel."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1:4]


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    if not m:
        return None
    k = m.group(1)
    if k == "k_voxelize_fp_rows":        # the wide-row variant runs under the same timer name as the flat one
        k = "k_voxelize_fp"
    if k == "k_cl_verify":
        k += "<trusted>" if re.search(r"k_cl_verify<[^>]*(true|\(bool\)1|, 1)", name) else "<validating>"
    return k


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v


rows = list(csv.DictReader([l for l in open(launches) if not l.startswith("==")]))
half = len(rows) // 2
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[half:]:
    n = r["Kernel Name"]
    key = ("pg::" + short(n)) if "pg::" in n and short(n) else "torch: " + re.sub(r"^void ", "", n)[:60]
    agg[key][0] += 1
    agg[key][1] += to_us(r["Metric Value"], r["Metric Unit"])
tot = sum(v[1] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if k.startswith("pg::"))
out = ["# %s: ncu launch list and `--set full` summary" % tag, "",
       "Command (1 x B200, under gpurun): `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python "
       "bench.py --profile-mode --steps 1 --warmup 1`; second half of the list = one step of the 8 x 150k-point chain "
       "(%d launches). Times are cold-cache and serialised under the profiler: compare SHARES with bench.py's live "
       "`per_kernel` numbers, not absolutes." % (len(rows) - half), "",
       "| kernel | launches | us | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    out.append("| `%s` | %d | %.1f | %.1f%% |" % (k.replace("|", "/"), v[0], v[1], 100 * v[1] / tot))
out.append("| total (library kernels %.1f us = %.1f%%) | %d | %.1f | |" % (mine, 100 * mine / tot, len(rows) - half, tot))

# `rep` is the .ncu-rep itself or its `ncu -i X.ncu-rep --page raw --csv` dump (a full-step report exceeds what
# gpurun copies back, so the dump is made on the GPU box and the report left there)
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                                                                    capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines()))
hdr, units, data = r[0], r[1], r[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst")]


def val(d, name):
    v, u = d[ix[name]], units[ix[name]]
    f = float(v.replace(",", ""))
    if name == "gpu__time_duration.sum":
        return {"ns": f / 1e6, "nsecond": f / 1e6, "us": f / 1e3, "usecond": f / 1e3, "ms": f, "msecond": f, "s": f * 1e3, "second": f * 1e3}[u]
    if name.startswith("dram__bytes"):
        return f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
    return f


seen = collections.defaultdict(list)
for d in data:
    k = short(d[ix["Kernel Name"]])
    if k:
        seen[k].append(d)
out += ["", "## `ncu --set full --clock-control none` per kernel (all captured launches of one step; `--page raw`)", "",
        "| kernel | launch | " + " | ".join(c[1] for c in cols) + " |", "|---|---|" + "---|" * len(cols)]
traffic = {}
per_step_launches = {k[4:]: v[0] for k, v in agg.items() if k.startswith("pg::")}
for k, ds in sorted(seen.items(), key=lambda kv: -sum(val(d, "gpu__time_duration.sum") for d in kv[1])):
    n = per_step_launches.get(k, len(ds))
    ds = ds[-n:]                      # the launches of the last captured step
    for j, d in enumerate(ds):
        out.append("| `%s` | %d | " % (k, j) + " | ".join("%.4g" % val(d, c[0]) for c in cols) + " |")
    traffic[k] = int(sum(val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum") for d in ds) * 1e6)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.md"), "w").write("\n".join(out) + "\n")
shutil.copy(launches, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1, sort_keys=True)
print("\n".join(out[:60]))
print(traffic)
