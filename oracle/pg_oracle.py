"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the CPU oracle (oracle/pg_oracle.c).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg.
The product package d3net_b200 never imports this module.

Each function mirrors one reference operator (lib/pointgroup_ops/functions/pointgroup_ops.py) on
numpy arrays; see pg_oracle.c for the reference file:line each restates.  Parity status: pinned
against the reference's own compiled PG_OP (oracle/_ref) -- see tests/test_oracle.py and
tests/golden/make_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "pg_oracle.c")
_LIB = os.path.join(HERE, "libpg_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """gcc the C restatement (a few seconds)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
                               _SRC, "-o", _LIB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_ballquery.restype = ctypes.c_int64
        _lib.orc_bfs_cluster.restype = ctypes.c_int64
    return _lib


def set_threads(n):
    """OpenMP threads used by the embarrassingly parallel ops (bench cpu_baseline reports this)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def _p(a, t):
    return a.ctypes.data_as(t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def voxelization_idx(coords, batchsize, mode=4):
    coords = _c(coords, np.int64)
    assert coords.ndim == 2 and coords.shape[1] == 4
    N = coords.shape[0]
    input_map = np.zeros(N, np.int32)
    M = ctypes.c_int32(0)
    mA = ctypes.c_int32(0)
    rc = lib().orc_voxelize_idx_map(_p(coords, _i64p), N, int(mode), _p(input_map, _i32p),
                                    ctypes.byref(M), ctypes.byref(mA))
    assert rc == 0
    M, mA = M.value, mA.value
    output_coords = np.zeros((M, 4), np.int64)
    output_map = np.zeros((M, mA + 1), np.int32)
    lib().orc_voxelize_idx_fill(_p(coords, _i64p), _p(input_map, _i32p), N, M, mA, int(mode),
                                _p(output_coords, _i64p), _p(output_map, _i32p))
    return output_coords, input_map, output_map


def voxelization(feats, map_rule, mode=4):
    feats = _c(feats, np.float32)
    map_rule = _c(map_rule, np.int32)
    M, W = map_rule.shape
    C = feats.shape[1]
    out = np.zeros((M, C), np.float32)
    lib().orc_voxelize_fp(_p(feats, _f32p), _p(out, _f32p), _p(map_rule, _i32p), M, W - 1, C, int(mode == 4))
    return out


def voxelization_bp(d_out, map_rule, N, mode=4):
    d_out = _c(d_out, np.float32)
    map_rule = _c(map_rule, np.int32)
    M, W = map_rule.shape
    C = d_out.shape[1]
    d_feats = np.zeros((N, C), np.float32)
    lib().orc_voxelize_bp(_p(d_out, _f32p), _p(d_feats, _f32p), _p(map_rule, _i32p), M, W - 1, C, int(mode == 4))
    return d_feats


def point_recover(feats, map_rule, nPoint):
    """point_recover_fp = voxelize_bp with average=false (voxelize.cpp:189)."""
    return voxelization_bp(feats, map_rule, nPoint, mode=3)


def point_recover_bp(d_out, map_rule):
    """point_recover_bp = voxelize_fp with average=false (voxelize.cpp:201)."""
    return voxelization(d_out, map_rule, mode=3)


def ballquery_batch_p(coords, batch_idxs, batch_offsets, radius, meanActive=None, use_grid=True):
    coords = _c(coords, np.float32)
    batch_idxs = _c(batch_idxs, np.int32)
    batch_offsets = _c(batch_offsets, np.int32)
    n = coords.shape[0]
    start_len = np.zeros((n, 2), np.int32)
    ptr = _i32p()
    total = lib().orc_ballquery(_p(coords, _f32p), _p(batch_idxs, _i32p), _p(batch_offsets, _i32p), n,
                                len(batch_offsets) - 1, ctypes.c_float(radius), int(bool(use_grid)),
                                _p(start_len, _i32p), ctypes.byref(ptr))
    idx = np.ctypeslib.as_array(ptr, shape=(max(int(total), 1),))[:total].copy()
    lib().orc_free(ptr)
    return idx, start_len


def bfs_cluster(semantic_label, ball_query_idxs, start_len, threshold):
    semantic_label = _c(semantic_label, np.int32)
    ball_query_idxs = _c(ball_query_idxs, np.int32)
    start_len = _c(start_len, np.int32)
    N = start_len.shape[0]
    ci, co = _i32p(), _i32p()
    nC = ctypes.c_int32(0)
    S = lib().orc_bfs_cluster(_p(semantic_label, _i32p), _p(ball_query_idxs, _i32p), _p(start_len, _i32p), N,
                              int(threshold), ctypes.byref(ci), ctypes.byref(co), ctypes.byref(nC))
    cluster_idxs = np.ctypeslib.as_array(ci, shape=(max(int(S), 1) * 2,))[:S * 2].copy().reshape(-1, 2)
    cluster_offsets = np.ctypeslib.as_array(co, shape=(nC.value + 1,)).copy()
    lib().orc_free(ci)
    lib().orc_free(co)
    return cluster_idxs, cluster_offsets


def roipool(feats, proposals_offset):
    feats = _c(feats, np.float32)
    off = _c(proposals_offset, np.int32)
    nP, C = len(off) - 1, feats.shape[1]
    out = np.zeros((nP, C), np.float32)
    maxidx = np.zeros((nP, C), np.int32)
    lib().orc_roipool_fp(_p(feats, _f32p), _p(off, _i32p), _p(out, _f32p), _p(maxidx, _i32p), nP, C)
    return out, maxidx


def roipool_bp(d_out, maxidx, sumNPoint):
    d_out = _c(d_out, np.float32)
    maxidx = _c(maxidx, np.int32)
    nP, C = d_out.shape
    d_feats = np.zeros((sumNPoint, C), np.float32)
    lib().orc_roipool_bp(_p(d_feats, _f32p), _p(maxidx, _i32p), _p(d_out, _f32p), nP, C)
    return d_feats


def _sec(fn, inp, offsets):
    inp = _c(inp, np.float32)
    off = _c(offsets, np.int32)
    nP, C = len(off) - 1, inp.shape[1]
    out = np.zeros((nP, C), np.float32)
    fn(_p(inp, _f32p), _p(off, _i32p), _p(out, _f32p), nP, C)
    return out


def sec_mean(inp, offsets):
    return _sec(lib().orc_sec_mean, inp, offsets)


def sec_min(inp, offsets):
    return _sec(lib().orc_sec_min, inp, offsets)


def sec_max(inp, offsets):
    return _sec(lib().orc_sec_max, inp, offsets)


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum):
    pi = _c(proposals_idx, np.int32)
    po = _c(proposals_offset, np.int32)
    il = _c(instance_labels, np.int64)
    ip = _c(instance_pointnum, np.int32)
    nP, nI = len(po) - 1, len(ip)
    iou = np.zeros((nP, nI), np.float32)
    lib().orc_get_iou(_p(pi, _i32p), _p(po, _i32p), _p(il, _i64p), _p(ip, _i32p), _p(iou, _f32p), nI, nP)
    return iou


# ---------------------------------------------------------------------------------------------
# canonical forms used by the parity tests (neighbour sets sorted, clusters relabelled)
# ---------------------------------------------------------------------------------------------
def relaid_neighbours(idx, start_len, sort=False):
    """Per-point neighbour lists re-laid out in point order (segment placement is the producer's
    choice: the reference places segments by atomicAdd, bfs_cluster.cu:47), each list in its stored
    order -- or sorted ascending with ``sort``.  Also checks that the segments do not overlap."""
    idx = np.asarray(idx)
    start_len = np.asarray(start_len).reshape(-1, 2)
    lens = start_len[:, 1].astype(np.int64)
    starts = start_len[:, 0].astype(np.int64)
    total = int(lens.sum())
    assert total <= idx.shape[0]
    if len(lens):
        o = np.argsort(starts, kind="stable")
        o = o[lens[o] > 0]
        assert (starts[o][1:] >= (starts[o] + lens[o])[:-1]).all(), "neighbour segments overlap"
        assert len(o) == 0 or (starts[o][0] >= 0 and starts[o][-1] + lens[o][-1] <= idx.shape[0])
    new_start = np.concatenate([[0], np.cumsum(lens)])[:-1]
    # gather position p of list i  ->  idx[starts[i] + p]
    owner = np.repeat(np.arange(len(lens)), lens)
    pos = np.arange(total) - np.repeat(new_start, lens)
    flat = idx[np.repeat(starts, lens) + pos]
    if sort:
        flat = flat[np.lexsort((flat, owner))]
    return flat.astype(np.int32), lens.astype(np.int32)


def canonical_neighbours(idx, start_len):
    """Per-point neighbour lists re-laid out in point order with each list sorted ascending."""
    return relaid_neighbours(idx, start_len, sort=True)


def canonical_clusters(cluster_idxs, cluster_offsets):
    """Clusters as (sorted member array) list ordered by smallest member -- invariant to BFS order."""
    ci = np.asarray(cluster_idxs).reshape(-1, 2)
    co = np.asarray(cluster_offsets)
    out = []
    for c in range(len(co) - 1):
        seg = ci[co[c]:co[c + 1]]
        assert (seg[:, 0] == c).all()
        out.append(np.sort(seg[:, 1]))
    return out
