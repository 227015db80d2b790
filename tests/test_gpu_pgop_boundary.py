"""The B2 boundary -- the native module ``PG_OP`` (lib/pointgroup_ops/src/pointgroup_ops_api.cpp:6-24) -- driven the way
the reference's own Python wrapper drives it (lib/pointgroup_ops/functions/pointgroup_ops.py): every one of the 13
functions is called POSITIONALLY with caller-allocated buffers, the ``.new()`` + ``resize_`` protocol for the
variable-size outputs, the zeroed ``n * meanActive`` buffer and the grow-and-retry loop of the ball query, and CPU
tensors where the reference's callers pass CPU tensors.  Results are compared with the oracle.

test_reference_wrapper_over_pg_op goes one step further when the staged reference tree is present (baseline/_ref, see
harness/stage_ref.py): the reference's wrapper module ITSELF, unmodified, runs on top of d3net_b200.PG_OP.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from util import cu, npy, object_subset, small_batch, random_segments, assert_same_floats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PG_OP():
    assert torch.cuda.is_available()
    from d3net_b200 import PG_OP as mod, _native
    _native.lib()
    return mod


def _lists_equal(oracle, idx, sl, ridx, rsl):
    a, al = oracle.canonical_neighbours(npy(idx), npy(sl))
    b, bl = oracle.canonical_neighbours(ridx, rsl)
    np.testing.assert_array_equal(al, bl)
    np.testing.assert_array_equal(a, b)


def test_voxelize_idx_positional_resize_protocol(PG_OP, oracle):
    """functions/pointgroup_ops.py:22-33: CPU coords, ``coords.new()`` / zero-length IntTensor outputs, N-sized
    input_map; the native side resizes output_coords and output_map (voxelize.cpp:22-26)."""
    rng = np.random.default_rng(7)
    N = 20000
    coords_np = np.column_stack([rng.integers(0, 3, N), rng.integers(0, 25, (N, 3))]).astype(np.int64)
    coords = torch.from_numpy(coords_np)
    for mode in (4, 3, 1):
        output_coords = coords.new()
        input_map = torch.IntTensor(N).zero_()
        output_map = input_map.new()
        ret = PG_OP.voxelize_idx(coords, output_coords, input_map, output_map, 3, mode)
        assert ret is None
        roc, rim, rom = oracle.voxelization_idx(coords_np, 3, mode)
        assert not output_coords.is_cuda and output_coords.dtype == torch.int64
        np.testing.assert_array_equal(output_coords.numpy(), roc)
        np.testing.assert_array_equal(input_map.numpy(), rim)
        np.testing.assert_array_equal(output_map.numpy(), rom)
    # CUDA tensors behave the same way (outputs stay where the caller put them)
    c = coords.cuda()
    oc, im, om = c.new(), torch.zeros(N, dtype=torch.int32, device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda")
    PG_OP.voxelize_idx(c, oc, im, om, 3, 4)
    roc, rim, rom = oracle.voxelization_idx(coords_np, 3, 4)
    assert oc.is_cuda and om.is_cuda
    np.testing.assert_array_equal(npy(oc), roc)
    np.testing.assert_array_equal(npy(om), rom)


def _wrapper_ballquery(PG_OP, coords, batch_idxs, batch_offsets, radius, meanActive):
    """functions/pointgroup_ops.py:127-146, statement for statement (torch.zeros instead of the legacy constructors)."""
    n = coords.size(0)
    tries = 0
    while True:
        idx = torch.zeros(n * meanActive, dtype=torch.int32, device="cuda")
        start_len = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
        nActive = PG_OP.ballquery_batch_p(coords, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius)
        tries += 1
        if nActive <= n * meanActive:
            break
        meanActive = int(nActive // n + 1)
    idx = idx[:nActive]
    return idx, start_len, tries


def test_ballquery_positional_retry_protocol(PG_OP, oracle):
    s = object_subset(small_batch(2, 9000))
    x, b, o = cu(s["shifted"]), cu(s["batch_idxs"]), cu(s["batch_offsets"])
    ridx, rsl = oracle.ballquery_batch_p(s["shifted"], s["batch_idxs"], s["batch_offsets"], 0.03)
    # generous buffer: one call
    idx, sl, tries = _wrapper_ballquery(PG_OP, x, b, o, 0.03, 2000)
    assert tries == 1 and idx.numel() == len(ridx)
    _lists_equal(oracle, idx, sl, ridx, rsl)
    # meanActive = 1: the count exceeds n * meanActive, the wrapper regrows and calls again
    idx, sl, tries = _wrapper_ballquery(PG_OP, x, b, o, 0.03, 1)
    assert tries == 2 and idx.numel() == len(ridx)
    _lists_equal(oracle, idx, sl, ridx, rsl)
    # the return value is an int the wrapper can slice with, and start_len is filled even when idx did not fit
    n = x.size(0)
    small = torch.zeros(n, dtype=torch.int32, device="cuda")
    sl0 = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    total = PG_OP.ballquery_batch_p(x, b, o, small, sl0, n, 1, 0.03)
    assert isinstance(total, int) and total == len(ridx) > n
    np.testing.assert_array_equal(npy(sl0)[:, 1], rsl[:, 1])
    assert int(small.abs().sum()) == 0                    # untouched: nothing is clipped into a short buffer


def test_bfs_cluster_positional_new_protocol(PG_OP, oracle):
    """functions/pointgroup_ops.py:166-176 with the caller's CPU tensors (model/pointgroup.py:297): outputs are
    ``semantic_label.new()`` and come back resized on the CPU (bfs_cluster.cpp:103-106)."""
    s = object_subset(small_batch(2, 9000))
    x, b, o = cu(s["coords"]), cu(s["batch_idxs"]), cu(s["batch_offsets"])
    idx, sl, _ = _wrapper_ballquery(PG_OP, x, b, o, 0.03, 50)
    semantic_label = torch.from_numpy(s["sem"])                      # CPU int32
    ball_query_idxs, start_len = idx.cpu(), sl.cpu()
    N = start_len.size(0)
    cluster_idxs = semantic_label.new()
    cluster_offsets = semantic_label.new()
    ret = PG_OP.bfs_cluster(semantic_label, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, 50)
    assert ret is None and not cluster_idxs.is_cuda and cluster_idxs.dtype == torch.int32
    rci, rco = oracle.bfs_cluster(s["sem"], ball_query_idxs.numpy(), start_len.numpy(), 50)
    np.testing.assert_array_equal(cluster_offsets.numpy(), rco)
    assert cluster_idxs.shape == (int(rco[-1]), 2)
    for a, c in zip(oracle.canonical_clusters(cluster_idxs.numpy(), cluster_offsets.numpy()), oracle.canonical_clusters(rci, rco)):
        np.testing.assert_array_equal(a, c)
    # CUDA tensors in, CUDA tensors out
    sem_c = semantic_label.cuda()
    ci, co = sem_c.new(), sem_c.new()
    PG_OP.bfs_cluster(sem_c, idx.contiguous(), sl, ci, co, N, 50)
    assert ci.is_cuda
    np.testing.assert_array_equal(npy(co), rco)
    # a threshold nobody reaches: empty outputs, offsets == [0]
    ci, co = semantic_label.new(), semantic_label.new()
    PG_OP.bfs_cluster(semantic_label, ball_query_idxs, start_len, ci, co, N, 10 ** 7)
    assert ci.numel() == 0 and co.tolist() == [0]


def test_fixed_size_ops_positional(PG_OP, oracle):
    """The ten fixed-size functions with the wrapper's zero-filled, caller-owned outputs."""
    rng = np.random.default_rng(13)
    n, C = 6000, 16
    coords = np.column_stack([rng.integers(0, 2, n), rng.integers(0, 9, (n, 3))]).astype(np.int64)
    _, _, rule = oracle.voxelization_idx(coords, 2, 4)
    M, W = rule.shape
    feats = rng.standard_normal((n, C)).astype(np.float32)
    rt = cu(rule)
    out = torch.zeros((M, C), device="cuda")
    PG_OP.voxelize_fp(cu(feats), out, rt, 4, M, W - 1, C)
    assert_same_floats(npy(out), oracle.voxelization(feats, rule, 4))
    g = rng.standard_normal((M, C)).astype(np.float32)
    d = torch.zeros((n, C), device="cuda")
    PG_OP.voxelize_bp(cu(g), d, rt, 4, M, W - 1, C)
    assert_same_floats(npy(d), oracle.voxelization_bp(g, rule, n, 4))
    rec = torch.zeros((n, C), device="cuda")
    PG_OP.point_recover_fp(cu(g), rec, rt, M, W - 1, C)
    assert_same_floats(npy(rec), oracle.point_recover(g, rule, n))
    dg = torch.zeros((M, C), device="cuda")
    PG_OP.point_recover_bp(cu(feats), dg, rt, M, W - 1, C)
    assert_same_floats(npy(dg), oracle.point_recover_bp(feats, rule))

    off = random_segments(rng, 60, 200, big=3000)
    S, nP = int(off[-1]), len(off) - 1
    x = rng.standard_normal((S, C)).astype(np.float32)
    xt, ot = cu(x), cu(off)
    pooled = torch.zeros((nP, C), device="cuda")
    arg = torch.zeros((nP, C), dtype=torch.int32, device="cuda")
    PG_OP.roipool_fp(xt, ot, pooled, arg, nP, C)
    r_out, r_arg = oracle.roipool(x, off)
    assert_same_floats(npy(pooled), r_out)
    np.testing.assert_array_equal(npy(arg), r_arg)
    gd = rng.standard_normal((nP, C)).astype(np.float32)
    dfe = torch.zeros((S, C), device="cuda")
    PG_OP.roipool_bp(dfe, ot, arg, cu(gd), nP, C)
    assert_same_floats(npy(dfe), oracle.roipool_bp(gd, r_arg, S))
    x3 = rng.standard_normal((S, 3)).astype(np.float32)
    for name in ("sec_mean", "sec_min", "sec_max"):
        o3 = torch.zeros((nP, 3), device="cuda")
        getattr(PG_OP, name)(cu(x3), ot, o3, nP, 3)
        assert_same_floats(npy(o3), getattr(oracle, name)(x3, off))
    N_tot, nI = 20000, 40
    labels = rng.integers(-1, nI, N_tot).astype(np.int64)
    pointnum = np.bincount(labels[labels >= 0], minlength=nI).astype(np.int32)
    pidx = rng.integers(0, N_tot, S).astype(np.int32)
    iou = torch.zeros((nP, nI), device="cuda")
    PG_OP.get_iou(cu(pidx), ot, cu(labels), cu(pointnum), iou, nI, nP)
    assert_same_floats(npy(iou), oracle.get_iou(pidx, off, labels, pointnum))


# ---- the reference's wrapper module itself over d3net_b200.PG_OP (option A) ---------------------------
def _staged_wrapper():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    path = os.path.join(root, "lib", "pointgroup_ops", "functions", "pointgroup_ops.py")
    return path if os.path.exists(path) else None


def test_reference_wrapper_over_pg_op(PG_OP, oracle):
    path = _staged_wrapper()
    if path is None:
        pytest.skip("baseline/_ref not staged (harness/stage_ref.py needs /root/reference)")
    saved = sys.modules.get("PG_OP")
    sys.modules["PG_OP"] = PG_OP                       # `import PG_OP` at functions/pointgroup_ops.py:9
    try:
        spec = importlib.util.spec_from_file_location("_ref_pointgroup_ops", path)
        ref_ops = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_ops)
    finally:
        if saved is None:
            sys.modules.pop("PG_OP", None)
        else:
            sys.modules["PG_OP"] = saved
    s = object_subset(small_batch(2, 9000))
    b = small_batch(2, 9000)
    # voxelization_idx on CPU coords (lib/dataset/pipeline.py:992) + voxelization on the GPU (model/pointgroup.py:472)
    coords = torch.from_numpy(b["locs_scaled"])
    oc, im, om = ref_ops.voxelization_idx(coords, 2, 4)
    roc, rim, rom = oracle.voxelization_idx(b["locs_scaled"], 2, 4)
    np.testing.assert_array_equal(oc.numpy(), roc)
    np.testing.assert_array_equal(im.numpy(), rim)
    np.testing.assert_array_equal(om.numpy(), rom)
    feats = np.random.default_rng(1).standard_normal((coords.size(0), 16)).astype(np.float32)
    ft = cu(feats).requires_grad_(True)
    vf = ref_ops.voxelization(ft, om.cuda(), 4)
    assert_same_floats(npy(vf), oracle.voxelization(feats, rom, 4))
    vf.sum().backward()
    assert_same_floats(npy(ft.grad), oracle.voxelization_bp(np.ones_like(npy(vf)), rom, coords.size(0), 4))
    # ball query with a meanActive that forces the wrapper's own retry loop, then its bfs_cluster on CPU copies
    x, bi, bo = cu(s["shifted"]), cu(s["batch_idxs"]), cu(s["batch_offsets"])
    idx, sl = ref_ops.ballquery_batch_p(x, bi, bo, 0.03, 1)
    ridx, rsl = oracle.ballquery_batch_p(s["shifted"], s["batch_idxs"], s["batch_offsets"], 0.03)
    _lists_equal(oracle, idx, sl, ridx, rsl)
    ci, co = ref_ops.bfs_cluster(torch.from_numpy(s["sem"]), idx.cpu(), sl.cpu(), 50)
    rci, rco = oracle.bfs_cluster(s["sem"], ridx, rsl, 50)
    np.testing.assert_array_equal(co.numpy(), rco)
    for a, c in zip(oracle.canonical_clusters(ci.numpy(), co.numpy()), oracle.canonical_clusters(rci, rco)):
        np.testing.assert_array_equal(a, c)
    # pooled features + IoU + segment statistics through the reference's autograd classes
    off = random_segments(np.random.default_rng(2), 40, 150)
    S = int(off[-1])
    xs = np.random.default_rng(3).standard_normal((S, 16)).astype(np.float32)
    xt = cu(xs).requires_grad_(True)
    pooled = ref_ops.roipool(xt, cu(off))
    r_out, r_arg = oracle.roipool(xs, off)
    assert_same_floats(npy(pooled), r_out)
    pooled.sum().backward()
    assert_same_floats(npy(xt.grad), oracle.roipool_bp(np.ones_like(r_out), r_arg, S))
    x3 = xs[:, :3].copy()
    for name in ("sec_mean", "sec_min", "sec_max"):
        assert_same_floats(npy(getattr(ref_ops, name)(cu(x3), cu(off))), getattr(oracle, name)(x3, off))
