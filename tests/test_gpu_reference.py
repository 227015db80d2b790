"""Pins the oracle (and the product) against the reference's OWN compiled ops on the GPU:
oracle/_ref/PG_OP.so is built by oracle/build_ref.py from the unmodified sources under
/root/reference/lib/pointgroup_ops/src (sm_100a) and travels to the GPU box with the snapshot.
Nothing here reads /root/reference at run time."""
import numpy as np
import pytest
import torch

from util import cu, npy, object_subset, small_batch, random_segments

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/PG_OP.so was not built (needs /root/reference at build time)")
    return mod


def _ref_ballquery(ref, xyz, bi, bo, r, mean_active):
    n = xyz.shape[0]
    x, b, o = cu(xyz), cu(bi), cu(bo)
    while True:
        idx = torch.zeros(n * mean_active, dtype=torch.int32, device="cuda")
        sl = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
        total = ref.ballquery_batch_p(x, b, o, idx, sl, n, mean_active, r)
        torch.cuda.synchronize()
        if total <= n * mean_active:
            return npy(idx[:total]), npy(sl)
        mean_active = total // n + 1


def test_reference_ballquery_vs_oracle_and_product(ref, ops, oracle):
    s = object_subset(small_batch(2, 9000))
    for key, ma in (("coords", 50), ("shifted", 300)):
        ridx, rsl = _ref_ballquery(ref, s[key], s["batch_idxs"], s["batch_offsets"], 0.03, ma)
        oidx, osl = oracle.ballquery_batch_p(s[key], s["batch_idxs"], s["batch_offsets"], 0.03)
        idx, sl = ops.ballquery_batch_p(cu(s[key]), cu(s["batch_idxs"]), cu(s["batch_offsets"]), 0.03, ma)
        ra, rl = oracle.canonical_neighbours(ridx, rsl)          # reference segments sit in atomicAdd order
        oa, ol = oracle.canonical_neighbours(oidx, osl)
        pa, pl = oracle.canonical_neighbours(npy(idx), npy(sl))
        np.testing.assert_array_equal(rl, ol)
        np.testing.assert_array_equal(ra, oa)
        np.testing.assert_array_equal(rl, pl)
        np.testing.assert_array_equal(ra, pa)


def test_reference_ballquery_cap(ref, ops, oracle):
    rng = np.random.default_rng(21)
    xyz = np.concatenate([rng.normal(0, 0.004, (2300, 3)), rng.uniform(-1, 1, (1500, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    bi = np.zeros(len(xyz), np.int32)
    bo = np.array([0, len(xyz)], np.int32)
    ridx, rsl = _ref_ballquery(ref, xyz, bi, bo, 0.03, 700)
    idx, sl = ops.ballquery_batch_p(cu(xyz), cu(bi), cu(bo), 0.03, 700)
    ra, rl = oracle.canonical_neighbours(ridx, rsl)
    pa, pl = oracle.canonical_neighbours(npy(idx), npy(sl))
    assert (rl == 1000).any()
    np.testing.assert_array_equal(rl, pl)
    np.testing.assert_array_equal(ra, pa)
    # the reference's CPU BFS on the reference's own (directed, truncated) lists vs the GPU clustering
    sem = rng.integers(1, 3, len(xyz)).astype(np.int32)
    ci, co = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
    ref.bfs_cluster(torch.from_numpy(sem), torch.from_numpy(ridx), torch.from_numpy(rsl), ci, co, len(xyz), 5)
    pci, pco = ops.bfs_cluster(cu(sem), idx, sl, 5)
    np.testing.assert_array_equal(npy(pco), co.numpy())
    for a, b in zip(oracle.canonical_clusters(npy(pci), npy(pco)), oracle.canonical_clusters(ci.numpy(), co.numpy())):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("C", [3, 16, 134])
def test_reference_voxelize(ref, ops, oracle, C):
    rng = np.random.default_rng(30 + C)
    n = 12000
    coords = np.column_stack([rng.integers(0, 2, n), rng.integers(0, 12, (n, 3))]).astype(np.int64)
    oc, im, om = torch.zeros(0, dtype=torch.int64), torch.zeros(n, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
    ref.voxelize_idx(torch.from_numpy(coords), oc, im, om, 2, 4)
    poc, pim, pom = ops.voxelization_idx(cu(coords), 2, 4)
    np.testing.assert_array_equal(npy(poc), oc.numpy())
    np.testing.assert_array_equal(npy(pim), im.numpy())
    np.testing.assert_array_equal(npy(pom), om.numpy())
    feats = rng.standard_normal((n, C)).astype(np.float32)
    M, W = om.shape
    out = torch.zeros((M, C), device="cuda")
    ref.voxelize_fp(cu(feats), out, om.cuda(), 4, M, W - 1, C)
    torch.cuda.synchronize()
    assert npy(out).tobytes() == oracle.voxelization(feats, om.numpy(), 4).tobytes()
    assert npy(out).tobytes() == npy(ops.voxelization(cu(feats), pom, 4)).tobytes()
    g = rng.standard_normal((M, C)).astype(np.float32)
    d = torch.zeros((n, C), device="cuda")
    ref.voxelize_bp(cu(g), d, om.cuda(), 4, M, W - 1, C)
    torch.cuda.synchronize()
    assert npy(d).tobytes() == oracle.voxelization_bp(g, om.numpy(), n, 4).tobytes()


@pytest.mark.parametrize("C", [3, 16])
def test_reference_segment_ops(ref, ops, oracle, C):
    rng = np.random.default_rng(40 + C)
    off = random_segments(rng, 200, 300, big=20000)
    S = int(off[-1])
    x = rng.standard_normal((S, C)).astype(np.float32)
    x[rng.random(x.shape) < 0.02] = 0.25
    nP = len(off) - 1
    xt, ot = cu(x), cu(off)
    out = torch.zeros((nP, C), device="cuda")
    arg = torch.zeros((nP, C), dtype=torch.int32, device="cuda")
    ref.roipool_fp(xt, ot, out, arg, nP, C)
    torch.cuda.synchronize()
    o_out, o_arg = oracle.roipool(x, off)
    assert npy(out).tobytes() == o_out.tobytes()
    np.testing.assert_array_equal(npy(arg), o_arg)
    assert npy(ops.roipool(xt, ot)).tobytes() == npy(out).tobytes()
    for name in ("sec_mean", "sec_min", "sec_max"):
        r = torch.zeros((nP, C), device="cuda")
        getattr(ref, name)(xt, ot, r, nP, C)
        torch.cuda.synchronize()
        assert npy(r).tobytes() == getattr(oracle, name)(x, off).tobytes(), name
        assert npy(r).tobytes() == npy(getattr(ops, name)(xt, ot)).tobytes(), name


def test_reference_get_iou(ref, ops, oracle):
    rng = np.random.default_rng(50)
    N, nI = 30000, 90
    labels = rng.integers(-1, nI, N).astype(np.int64)
    pointnum = np.bincount(labels[labels >= 0], minlength=nI).astype(np.int32)
    off = random_segments(rng, 80, 500, big=9000)
    pidx = rng.integers(0, N, int(off[-1])).astype(np.int32)
    nP = len(off) - 1
    r = torch.zeros((nP, nI), device="cuda")
    ref.get_iou(cu(pidx), cu(off), cu(labels), cu(pointnum), r, nI, nP)
    torch.cuda.synchronize()
    assert npy(r).tobytes() == oracle.get_iou(pidx, off, labels, pointnum).tobytes()
    assert npy(r).tobytes() == npy(ops.get_iou(cu(pidx), cu(off), cu(labels), cu(pointnum))).tobytes()
