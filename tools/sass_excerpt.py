#!/usr/bin/env python
"""profiles/<tag>_sass.md: which sm_100a-specific instructions the built library contains, per kernel, with excerpts.

    python tools/sass_excerpt.py r02

Runs `cuobjdump -sass d3net_b200/libpg_b200.so` (no GPU needed) and counts, per kernel, the mnemonics that only exist
from sm_90 / sm_100 on: packed fp32 math (FADD2 / FMUL2 / FFMA2), bulk asynchronous copies (UBLKCP) with their mbarrier
transaction waits (SYNCS), programmatic-dependent-launch control (ACQBULK / PREEXIT as ptxas emits griddepcontrol),
vector reductions (RED.*.128 / .64), warp match (MATCH) and redux (REDUX)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "sass"
lib = os.path.join(ROOT, "d3net_b200", "libpg_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["FFMA2", "FADD2", "FMUL2", "UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "MATCH", "REDUX", "RED.E", "ATOMG", "LDS.128", "LDG.E.128",
         "STG.E.128", "POPC", "SHFL", "UTMALDG", "UTCHMMA", "LDTM"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows, excerpts = [], {}
for f in funcs:
    name, body = f.split("\n", 1)
    ins = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    cnt = collections.Counter()
    for l in ins:
        for w in WATCH:
            if re.search(r"\b" + re.escape(w), l):
                cnt[w] += 1
    rows.append((demangle(name.strip()), len(ins), cnt))
    for key in ("UBLKCP", "FFMA2", "ACQBULK"):
        if cnt[key] and key not in excerpts:
            k = next(i for i, l in enumerate(ins) if key in l)
            excerpts[key] = (demangle(name.strip()), [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in ins[max(0, k - 6):k + 10]])

arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
out = ["# %s: sm_100a-specific instructions in libpg_b200.so" % tag, "",
       "`cuobjdump -sass d3net_b200/libpg_b200.so` -- %d kernels, code for %s only.  Counts are static SASS instructions." % (len(rows), ", ".join(arch)),
       "", "| kernel | instr | " + " | ".join(WATCH[:10]) + " |", "|---|---|" + "---|" * 10]
tot = collections.Counter()
for name, n, cnt in sorted(rows, key=lambda r: -r[1]):
    tot.update(cnt)
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    out.append("| `%s` | %d | " % (short[:60], n) + " | ".join(str(cnt[w]) if cnt[w] else "" for w in WATCH[:10]) + " |")
out += ["", "Totals over the library: " + ", ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]) + ".",
        "Not present (and not applicable: no contraction on this path, see DESIGN.md section 2): " +
        ", ".join(w for w in ("UTMALDG", "UTCHMMA", "LDTM") if not tot[w]) + ".", ""]
for key, (name, lines) in excerpts.items():
    out += ["## first `%s` -- `%s`" % (key, re.sub(r"\(.*", "", name)[:80]), "", "```"] + lines + ["```", ""]
path = os.path.join(ROOT, "profiles", "%s_sass.md" % tag)
with open(path, "w") as fh:
    fh.write("\n".join(out))
print(path, "kernels", len(rows), dict(tot))
