#!/usr/bin/env python
"""HARNESS ONLY -- stages an UNMODIFIED copy of the reference's caller-side Python (the files that CALL the hot path)
from /root/reference into baseline/_ref/ so that it can run on the GPU box, where /root/reference does not exist.

baseline/_ref/ is git-ignored (nothing of the reference enters this repo's history) but travels with the gpurun
snapshot -- the same arrangement as oracle/_ref/PG_OP.so, which is the reference's native code compiled in place.
Nothing under d3net_b200/ imports from baseline/_ref; only harness/, tests/ and bench.py's `unchanged_caller` /
`speaker` legs do, and they skip / report "unavailable" when the tree is absent.

What is staged (files are copied byte for byte; MANIFEST.json records their sha256):
  model/*.py                                   PointGroup, SpeakerNet, caption / graph modules (the callers)
  lib/utils/*.py                               bbox / eval / nn_distance helpers those import
  lib/pointgroup_ops/functions/pointgroup_ops.py   the reference's own operator wrapper (INTEGRATION.md option A)
  data/scannet/model_util_scannet.py + two meta_data files ScannetDatasetConfig reads
  conf/*.yaml                                  the hyper-parameters (cluster radius, thresholds, ...)

Usage:  python harness/stage_ref.py [--force]
"""
import glob
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "baseline", "_ref")

PATTERNS = [
    "model/*.py",
    "lib/utils/*.py",
    "lib/pointgroup_ops/functions/pointgroup_ops.py",
    "data/scannet/model_util_scannet.py",
    "data/scannet/meta_data/scannet_reference_means.npz",
    "data/scannet/meta_data/scannetv2-labels.combined.tsv",
    "conf/*.yaml",
]


def ref_available():
    return os.path.isdir(os.path.join(REF, "model"))


def staged():
    return os.path.exists(os.path.join(OUT, "MANIFEST.json"))


def stage(force=False, verbose=True):
    if not ref_available():
        if verbose:
            print("[harness/stage_ref] /root/reference absent: using the staged tree if any")
        return staged()
    if staged() and not force:
        return True
    manifest = {}
    for pat in PATTERNS:
        for src in sorted(glob.glob(os.path.join(REF, pat))):
            rel = os.path.relpath(src, REF)
            dst = os.path.join(OUT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            with open(dst, "rb") as f:
                manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("[harness/stage_ref] staged %d files under %s" % (len(manifest), OUT))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage(force="--force" in sys.argv) else 1)
