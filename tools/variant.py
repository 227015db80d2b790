#!/usr/bin/env python
"""Build a kernel-variant copy of libpg_b200.so for A/B runs on the GPU box:
    python tools/variant.py NAME 'file.cu:::old text:::new text' ...
-> d3net_b200/variants/libpg_NAME.so (load it with PG_B200_LIB=...).  The tree itself is untouched."""
import os, shutil, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, edits = sys.argv[1], sys.argv[2:]
tmp = "/tmp/pg_variant_" + name
shutil.rmtree(tmp, ignore_errors=True)
shutil.copytree(os.path.join(ROOT, "d3net_b200", "csrc"), os.path.join(tmp, "d3net_b200", "csrc"))
shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
for e in edits:
    f, old, new = e.split(":::")
    p = os.path.join(tmp, "d3net_b200", "csrc", f)
    s = open(p).read()
    assert old in s, "not found in %s: %s" % (f, old)
    open(p, "w").write(s.replace(old, new))
outdir = os.path.join(ROOT, "d3net_b200", "variants")
os.makedirs(outdir, exist_ok=True)
objs = []
procs = []
for src in sorted(glob.glob(os.path.join(tmp, "d3net_b200", "csrc", "*.cu"))):
    obj = src[:-3] + ".o"
    objs.append(obj)
    procs.append(subprocess.Popen(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                                   "-Xcompiler", "-fPIC", "-diag-suppress", "177", "-I", os.path.join(tmp, "include"), "-c", src, "-o", obj]))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(outdir, "libpg_%s.so" % name)
subprocess.check_call(["nvcc", "-shared", "-o", out] + objs)
print(out)
