"""The C-ABI library loads on a machine without a GPU and exports exactly what include/pg_b200.h
declares; host-only entry points (workspace sizing, argument validation, N = 0 early-outs) behave."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from d3net_b200 import _native
    _native.build()
    return _native.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from d3net_b200 import _native
    names = _declared()
    assert len(names) == 38
    for n in names:
        assert hasattr(lib, n), "libpg_b200.so does not export " + n
        assert n in _native.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_native.SIGNATURES) == names      # nothing bound that the header does not declare
    assert lib.pg_abi_version() == _native.ABI_VERSION


def test_header_cites_the_reference_interface():
    src = open(os.path.join(ROOT, "include", "pg_b200.h")).read()
    for ref in ("pointgroup_ops_api.cpp", "voxelize.cpp", "bfs_cluster.h:15", "bfs_cluster.h:18", "roipool.h:15",
                "sec_mean.h:14", "get_iou.h:16"):
        assert ref in src


def test_workspace_sizes_scale(lib):
    for fn in (lib.pg_voxelize_idx_workspace_bytes, lib.pg_ballquery_workspace_bytes, lib.pg_bfs_cluster_workspace_bytes):
        a, b, z = fn(1000), fn(1000000), fn(0)
        assert 0 < z <= a < b
        assert b < 1000000 * 1024          # under 1 KB per point
    assert lib.pg_roipool_workspace_bytes(10, 16) >= 10 * 16 * 8


def test_argument_validation_without_a_gpu(lib):
    """Every path below returns before any CUDA call."""
    sizes = (ctypes.c_int32 * 3)()
    total = ctypes.c_int64(-1)
    assert lib.pg_voxelize_idx_map(None, 0, 4, None, None, 0, sizes, None) == 0 and list(sizes)[:2] == [0, 1]
    assert lib.pg_voxelize_idx_map(None, 0, 9, None, None, 0, sizes, None) == -1           # bad mode
    assert b"mode" in lib.pg_last_error()
    assert lib.pg_voxelize_idx_map(None, 5, 4, None, None, 0, sizes, None) == -1           # null pointers
    assert lib.pg_ballquery_prepare(None, None, None, 0, 1, 0.03, None, 0, None) == 0
    assert lib.pg_ballquery_prepare(None, None, None, -1, 1, 0.03, None, 0, None) == -1
    used = ctypes.c_int(7)
    assert lib.pg_ballquery_count(None, 0, 0.03, None, None, 0, None, 0, ctypes.byref(total), ctypes.byref(used), None) == 0
    assert used.value == 0 and total.value == 0
    assert lib.pg_ballquery_fill(None, 0, 0.03, None, None, None, 0, None, 0, None) == 0
    assert lib.pg_bfs_cluster_count(None, None, None, 0, 0, 50, 0, None, 0, sizes, None) == 0
    assert lib.pg_voxelize_fp(None, None, None, 0, 1, 16, 1, None) == 0
    assert lib.pg_voxelize_fp(None, None, None, 5, 1, 16, 1, None) == -1
    assert lib.pg_sec_mean(None, None, None, 0, 0, 3, None) == 0
    assert lib.pg_get_iou(None, None, None, None, None, 0, 0, None) == 0
    assert lib.pg_roipool_bp(None, None, None, None, 0, 16, None) == 0
    assert lib.pg_gather_rows(None, None, 0, None, 0, 16, None) == 0


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under d3net_b200/ may import or open it."""
    pkg = os.path.join(ROOT, "d3net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("The oracle", "") or f == "chain.py", f
    chain = open(os.path.join(pkg, "chain.py")).read()
    assert "import oracle" not in chain and "from oracle" not in chain


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from d3net_b200 import pointgroup_ops as ops
    with pytest.raises(RuntimeError):
        ops.voxelization_idx(torch.zeros((4, 4), dtype=torch.int64), 1, 4)
    with pytest.raises((AssertionError, RuntimeError, ValueError)):
        ops.sec_mean(torch.zeros((4, 3)), torch.tensor([0, 4], dtype=torch.int32))


def test_no_memory_access_ahead_of_the_dependency_wait():
    """Programmatic dependent launch: every kernel waits for its predecessor with griddepcontrol.wait (SASS ACQBULK)
    before it touches memory.  ptxas moves read-only loads (LDG.E.CONSTANT) above that wait when it can (csrc/common.cuh,
    "HAZARD"); no kernel of the library may have any memory instruction there."""
    import re
    import shutil
    import subprocess
    from d3net_b200 import _native
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    txt = subprocess.run([exe, "-sass", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    assert len(funcs) >= 60
    mem = re.compile(r"\b(LDG|LD|LDL|STG|ST|STL|ATOM|ATOMG|RED|LDS|STS|ATOMS|LDGSTS|UBLKCP)\b")
    offenders = {}
    for f in funcs:
        name, body = f.split("\n", 1)
        ins = [re.sub(r"/\*.*?\*/", "", l).strip() for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        wait = next((i for i, l in enumerate(ins) if "ACQBULK" in l), None)
        assert wait is not None, "kernel without griddepcontrol.wait: " + name
        early = [l for l in ins[:wait] if mem.search(l.replace(".", " "))]
        if early:
            offenders[name.strip()] = early
    assert not offenders, offenders
