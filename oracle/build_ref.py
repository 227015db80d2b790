#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Builds the reference's own native extension ``PG_OP`` from the UNMODIFIED sources where they lie
under ``/root/reference/lib/pointgroup_ops/src`` into ``oracle/_ref/PG_OP.so`` (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  No reference source is copied into this repo; the
only additions are two include-path shims under ``oracle/shim`` (sparsehash's dense_hash_map ->
std::unordered_map, and an empty THC/THC.h).  The reference's own ``setup.py`` is not run.

The recipe mirrors what ``lib/pointgroup_ops/setup.py:4-15`` asks torch's CUDAExtension to do
(g++ on the two host TUs, ``nvcc -O2`` on ``src/cuda.cu``), but targets sm_100a only.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/lib/pointgroup_ops/src"
OUT = os.path.join(HERE, "_ref")


def ref_available():
    return os.path.isdir(REF_SRC)


def built():
    return os.path.exists(os.path.join(OUT, "PG_OP.so"))


def build(force=False, verbose=True):
    if not ref_available():
        if verbose:
            print("[oracle/build_ref] /root/reference absent: using prebuilt oracle/_ref if any")
        return built()
    if built() and not force:
        return True
    import torch  # noqa: F401  (only for include / lib paths)
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(HERE, "shim"), "-I/usr/local/cuda/include",
           "-I" + sysconfig.get_paths()["include"]] + ["-I" + p for p in ce.include_paths()]
    defs = ["-DTORCH_EXTENSION_NAME=PG_OP", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DNDEBUG"]
    objs = []
    jobs = []
    for src in ("pointgroup_ops_api.cpp", "pointgroup_ops.cpp"):
        obj = os.path.join(OUT, src.replace(".cpp", ".o"))
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-w", "-c", os.path.join(REF_SRC, src), "-o", obj] + inc + defs
        jobs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    obj = os.path.join(OUT, "cuda.o")
    cmd = ["nvcc", "-std=c++17", "-O2", "-w", "-c", os.path.join(REF_SRC, "cuda.cu"), "-o", obj,
           "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
           "--expt-relaxed-constexpr", "--compiler-options", "-fPIC",
           "-gencode=arch=compute_100a,code=sm_100a"] + inc + defs
    jobs.append((cmd, subprocess.Popen(cmd)))
    objs.append(obj)
    for cmd, p in jobs:
        if p.wait() != 0:
            raise RuntimeError("reference build step failed: " + " ".join(cmd))
    libdirs = ce.library_paths() + ["/usr/local/cuda/lib64"]
    link = ["g++", "-shared", "-o", os.path.join(OUT, "PG_OP.so")] + objs
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    if verbose:
        print("[oracle/build_ref] built", os.path.join(OUT, "PG_OP.so"))
    return True


def load():
    """Import the reference PG_OP module (None when it was never built)."""
    if not built():
        return None
    import importlib.util
    import torch  # noqa: F401  must be imported first so libtorch symbols resolve
    spec = importlib.util.spec_from_file_location("PG_OP", os.path.join(OUT, "PG_OP.so"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
