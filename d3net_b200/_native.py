"""Loader (and in-tree builder) of libpg_b200.so, the C-ABI library declared in include/pg_b200.h.

The product path has no fallback: if the library is missing or a symbol cannot be bound, importing
the ops raises.  ``build()`` compiles d3net_b200/csrc/*.cu with nvcc for sm_100a only (it
cross-compiles without a GPU) into d3net_b200/libpg_b200.so, which is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
import ctypes
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libpg_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]

ABI_VERSION = 3        # must equal pg_abi_version() of the loaded library (bumped with every signature change)

_lib = None


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a on every .cu, then one shared link."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_vp, _i32, _i64, _sz, _f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
_int = ctypes.c_int

# name -> (restype, argtypes); must list every symbol include/pg_b200.h declares (tests check this)
SIGNATURES = {
    "pg_last_error": (ctypes.c_char_p, []),
    "pg_abi_version": (_int, []),
    "pg_kernel_timing": (None, [_int]),
    "pg_kernel_timing_report": (_sz, [ctypes.c_char_p, _sz]),
    "pg_voxelize_idx_workspace_bytes": (_sz, [_i64]),
    "pg_voxelize_idx_map": (_int, [_vp, _i64, _int, _vp, _vp, _sz, _vp, _vp]),
    "pg_voxelize_idx_fill": (_int, [_vp, _vp, _i64, _i32, _i32, _int, _vp, _sz, _vp, _vp, _vp]),
    "pg_voxelize_fp": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _int, _vp]),
    "pg_voxelize_bp": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _int, _vp]),
    "pg_point_recover_fp": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pg_point_recover_bp": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pg_ballquery_workspace_bytes": (_sz, [_i64]),
    "pg_ballquery_prepare": (_int, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _sz, _vp]),
    "pg_ballquery_count": (_int, [_vp, _i32, _f32, _vp, _vp, _i64, _vp, _sz, _vp, _vp, _vp]),
    "pg_ballquery_fill": (_int, [_vp, _i32, _f32, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "pg_bfs_cluster_workspace_bytes": (_sz, [_i64]),
    "pg_bfs_cluster_count": (_int, [_vp, _vp, _vp, _i32, _i64, _i32, _int, _vp, _sz, _vp, _vp]),
    "pg_bfs_cluster_count_grid": (_int, [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _sz, _vp, _sz, _vp, _vp]),
    "pg_bfs_cluster_count_lazy": (_int, [_vp, _vp, _i32, _i64, _i32, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "pg_bfs_cluster_debug": (None, [_vp]),
    "pg_bfs_cluster_fill": (_int, [_i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "pg_roipool_workspace_bytes": (_sz, [_i32, _i32]),
    "pg_roipool_fp": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _sz, _vp]),
    "pg_roipool_bp": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "pg_sec_mean": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pg_sec_min": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pg_sec_max": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pg_get_iou": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "pg_gather_rows": (_int, [_vp, _vp, _int, _vp, _i64, _i32, _vp]),
    "pg_pack_proposals": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "pg_cross_iou_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "pg_cross_iou": (_int, [_vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "pg_pick_masks": (_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "pg_nms_instances_workspace_bytes": (_sz, [_i32]),
    "pg_nms_instances": (_int, [_vp, _vp, _i32, _f32, _vp, _sz, _vp, _vp, _vp]),
    "pg_collate_points": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pg_cluster_coords_workspace_bytes": (_sz, [_i32]),
    "pg_cluster_coords": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
}


def lib():
    """The loaded library with argtypes bound.  Raises if it is absent -- there is no CPU path."""
    global _lib
    if _lib is None:
        path = os.environ.get("PG_B200_LIB", LIB_PATH)      # kernel-variant experiments load another build
        if not os.path.exists(path):
            raise RuntimeError(
                "d3net_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % path)
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        got = L.pg_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError("d3net_b200: %s reports ABI version %d, this package binds version %d -- a stale build; "
                               "rebuild it (python -c 'import __graft_entry__ as g; g.build()')" % (path, got, ABI_VERSION))
        _lib = L
    return _lib


class PgError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().pg_last_error()
        raise PgError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def kernel_timing(enable):
    """Switch the library's per-kernel CUDA-event timers on (clearing them) or off."""
    lib().pg_kernel_timing(1 if enable else 0)


def kernel_timing_report():
    """{kernel name: (launches, total ms)} of the launches since kernel_timing(True); waits for them."""
    L = lib()
    n = L.pg_kernel_timing_report(None, 0)
    buf = ctypes.create_string_buffer(n + 1)
    L.pg_kernel_timing_report(buf, n + 1)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out
