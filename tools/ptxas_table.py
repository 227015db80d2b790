#!/usr/bin/env python
"""profiles/<tag>_ptxas.md: registers, shared memory, stack and spills of every kernel (`nvcc -Xptxas -v`, no GPU needed).

    python tools/ptxas_table.py r02
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "ptxas"
rows = []
for src in sorted(glob.glob(os.path.join(ROOT, "d3net_b200", "csrc", "*.cu"))):
    r = subprocess.run(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v", "-diag-suppress", "177",
                        "-I", os.path.join(ROOT, "include"), "-c", src, "-o", "/dev/null"], capture_output=True, text=True)
    cur = None
    for line in r.stderr.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"file": os.path.basename(src), "name": m.group(1), "regs": 0, "smem": 0, "stack": 0, "spill": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["spill"] = int(m.group(1)), int(m.group(2)) + int(m.group(3))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(s.group(1)) if s else 0
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.split("\n")
out = ["# %s: ptxas resource usage per kernel (sm_100a, nvcc -O3 -Xptxas -v)" % tag, "",
       "%d kernels; %d with spills.  Threads per block are in the source (`__launch_bounds__`); 64 K registers and 228 KB of shared"
       % (len(rows), sum(1 for r in rows if r["spill"])),
       "memory per SM bound the resident blocks.", "", "| file | kernel | registers | static smem (B) | stack (B) | spill (B) |", "|---|---|---|---|---|---|"]
for r, n in sorted(zip(rows, names), key=lambda t: (t[0]["file"], -t[0]["regs"])):
    short = re.sub(r"\(.*", "", n).replace("void ", "")
    out.append("| %s | `%s` | %d | %d | %d | %d |" % (r["file"], short[:70], r["regs"], r["smem"], r["stack"], r["spill"]))
path = os.path.join(ROOT, "profiles", "%s_ptxas.md" % tag)
open(path, "w").write("\n".join(out) + "\n")
print(path, len(rows), "kernels")
