"""GPU parity: every op of d3net_b200.pointgroup_ops (CUDA, through the C ABI) against the CPU
oracle on the same seeded inputs.  Integer / index / max / IoU results must be bit-exact; the
float segment means are bit-exact too here (the kernels keep the reference's operation order), and
additionally checked to the contract's 1e-6 relative."""
import numpy as np
import pytest
import torch

from util import cu, npy, object_subset, small_batch, random_segments, assert_same_floats

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------
# voxelization_idx
# ------------------------------------------------------------------------------------------------
def _check_vox_idx(ops, oracle, coords, batch, mode):
    oc, im, om = ops.voxelization_idx(cu(coords, torch.int64), batch, mode)
    roc, rim, rom = oracle.voxelization_idx(coords, batch, mode)
    assert oc.dtype == torch.int64 and im.dtype == torch.int32 and om.dtype == torch.int32
    np.testing.assert_array_equal(npy(im), rim)
    np.testing.assert_array_equal(npy(om), rom)
    np.testing.assert_array_equal(npy(oc), roc)


def test_voxelize_idx_kat(ops):
    """Known-answer vector recorded from the reference binary."""
    c = torch.tensor([[0, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1], [1, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1]]).cuda()
    oc, im, om = ops.voxelization_idx(c, 2, 4)
    assert oc.tolist() == [[0, 1, 1, 1], [0, 2, 2, 2], [1, 1, 1, 1]]
    assert im.tolist() == [0, 1, 0, 2, 1, 0]
    assert om.tolist() == [[3, 0, 2, 5], [2, 1, 4, 0], [1, 3, 0, 0]]


@pytest.mark.parametrize("mode", [4, 3, 2, 1])
@pytest.mark.parametrize("n,span", [(1, 3), (7, 2), (5000, 12), (200000, 60)])
def test_voxelize_idx_random(ops, oracle, mode, n, span):
    rng = np.random.default_rng(n + mode)
    coords = np.column_stack([rng.integers(0, 3, n), rng.integers(0, span, (n, 3))]).astype(np.int64)
    _check_vox_idx(ops, oracle, coords, 3, mode)


def test_voxelize_idx_mode0_unique(ops, oracle):
    rng = np.random.default_rng(5)
    xyz = rng.permutation(20 ** 3)[:3000]
    coords = np.column_stack([np.zeros(3000), xyz // 400, xyz // 20 % 20, xyz % 20]).astype(np.int64)
    _check_vox_idx(ops, oracle, coords, 1, 0)


def test_voxelize_idx_extreme_coords(ops, oracle):
    """Negative, huge, and int32-wrapping int64 coordinates (the reference narrows to int32)."""
    rng = np.random.default_rng(11)
    base = rng.integers(-5, 5, (4000, 4)).astype(np.int64)
    base[:, 0] = rng.integers(0, 2, 4000)
    big = base.copy()
    big[::3, 1] += 2 ** 32          # same voxel after narrowing
    big[1::7, 2] = 2 ** 31 - 1
    big[2::11, 3] = -2 ** 31
    _check_vox_idx(ops, oracle, big, 2, 4)


def test_voxelize_idx_empty_and_single_voxel(ops, oracle):
    oc, im, om = ops.voxelization_idx(torch.zeros((0, 4), dtype=torch.int64).cuda(), 1, 4)
    assert tuple(oc.shape) == (0, 4) and tuple(im.shape) == (0,) and tuple(om.shape) == (0, 2)
    coords = np.zeros((3000, 4), np.int64)          # everything in one voxel: maxActive = N
    _check_vox_idx(ops, oracle, coords, 1, 4)


def test_voxelize_idx_scene(ops, oracle):
    b = small_batch(3, 30000)
    _check_vox_idx(ops, oracle, b["locs_scaled"], 3, 4)


def test_voxelize_idx_cpu_input_roundtrip(ops, oracle):
    """The reference's callers pass CPU tensors (lib/dataset/pipeline.py:992); results come back on CPU."""
    rng = np.random.default_rng(3)
    coords = np.column_stack([rng.integers(0, 2, 999), rng.integers(0, 6, (999, 3))]).astype(np.int64)
    oc, im, om = ops.voxelization_idx(torch.from_numpy(coords), 2, 4)
    assert not oc.is_cuda and not im.is_cuda and not om.is_cuda
    roc, rim, rom = oracle.voxelization_idx(coords, 2, 4)
    np.testing.assert_array_equal(oc.numpy(), roc)
    np.testing.assert_array_equal(im.numpy(), rim)
    np.testing.assert_array_equal(om.numpy(), rom)


# ------------------------------------------------------------------------------------------------
# voxelization (fp / bp) and point_recover
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [1, 2, 3, 7, 16, 40, 134, 600])
@pytest.mark.parametrize("mode", [4, 3])
def test_voxelize_fp_bp(ops, oracle, C, mode):
    rng = np.random.default_rng(100 + C)
    n = 20000
    coords = np.column_stack([rng.integers(0, 2, n), rng.integers(0, 14, (n, 3))]).astype(np.int64)
    _, _, om = oracle.voxelization_idx(coords, 2, 4)
    feats = rng.standard_normal((n, C)).astype(np.float32)
    f = cu(feats).requires_grad_(True)
    out = ops.voxelization(f, cu(om), mode)
    ref = oracle.voxelization(feats, om, mode)
    np.testing.assert_array_equal(npy(out), ref)                       # bit-exact
    np.testing.assert_allclose(npy(out), ref, rtol=1e-6, atol=0)       # the contract
    g = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(cu(g))
    np.testing.assert_array_equal(npy(f.grad), oracle.voxelization_bp(g, om, n, mode))


@pytest.mark.parametrize("mode", [4, 3])
def test_voxelize_fp_long_lists(ops, oracle, mode):
    """C = 16 rows with a few very long point lists (the cluster grids of the floor-sized proposals): the warp-per-row
    kernel takes lists above its threshold -- here up to 3000 points, many batches of its gather ring --
    the flat kernel the rest; sums must keep the reference's list order bit for bit."""
    rng = np.random.default_rng(77)
    parts = [np.tile(np.array([[0, 1, 1, 1]]), (3000, 1)), np.tile(np.array([[0, 2, 2, 2]]), (1025, 1)),
             np.tile(np.array([[1, 3, 3, 3]]), (1024, 1)), np.tile(np.array([[1, 0, 5, 9]]), (25, 1)),
             np.tile(np.array([[0, 4, 4, 4]]), (24, 1)),
             np.column_stack([rng.integers(0, 2, 6000), rng.integers(0, 3, (6000, 3))]),       # ~110 points per voxel
             np.column_stack([rng.integers(0, 2, 4000), rng.integers(5, 14, (4000, 3))])]      # short lists
    coords = np.concatenate(parts).astype(np.int64)
    coords = coords[rng.permutation(coords.shape[0])]
    n = coords.shape[0]
    _, _, om = oracle.voxelization_idx(coords, 2, 4)
    assert om.shape[1] - 1 >= 3000
    feats = (rng.standard_normal((n, 16)) * np.exp(rng.uniform(-8, 8, (n, 1)))).astype(np.float32)
    out = ops.voxelization(cu(feats), cu(om), mode)
    np.testing.assert_array_equal(npy(out), oracle.voxelization(feats, om, mode))


def test_voxelize_fp_signed_zero_and_specials(ops, oracle):
    om = np.array([[2, 0, 1], [1, 2, 0], [0, 0, 0]], np.int32)
    feats = np.array([[-0.0, np.inf, 1e-45], [-0.0, -np.inf, 1e-45], [-0.0, np.nan, -1e38]], np.float32)
    out = npy(ops.voxelization(cu(feats), cu(om), 4))
    ref = oracle.voxelization(feats, om, 4)
    assert_same_floats(out, ref)


def test_point_recover(ops, oracle):
    rng = np.random.default_rng(8)
    n, C = 5000, 16
    coords = np.column_stack([np.zeros(n), rng.integers(0, 9, (n, 3))]).astype(np.int64)
    _, _, om = oracle.voxelization_idx(coords, 1, 4)
    vf = rng.standard_normal((om.shape[0], C)).astype(np.float32)
    v = cu(vf).requires_grad_(True)
    out = ops.point_recover(v, cu(om), n)
    np.testing.assert_array_equal(npy(out), oracle.point_recover(vf, om, n))
    g = rng.standard_normal((n, C)).astype(np.float32)
    out.backward(cu(g))
    np.testing.assert_array_equal(npy(v.grad), oracle.point_recover_bp(g, om))


# ------------------------------------------------------------------------------------------------
# ballquery_batch_p
# ------------------------------------------------------------------------------------------------
def _check_bq(ops, oracle, xyz, bi, bo, r, mean_active=50):
    idx, sl = ops.ballquery_batch_p(cu(xyz), cu(bi), cu(bo), r, mean_active)
    ridx, rsl = oracle.ballquery_batch_p(xyz, bi, bo, r)
    assert idx.dtype == torch.int32 and sl.dtype == torch.int32
    # every point's list is the oracle's list in the oracle's (ascending) order; where a segment sits
    # in idx is the producer's choice (the reference's atomicAdd placement differs run to run)
    a, la = oracle.relaid_neighbours(npy(idx), npy(sl))
    b, lb = oracle.relaid_neighbours(ridx, rsl)
    np.testing.assert_array_equal(la, lb)
    np.testing.assert_array_equal(a, b)
    assert int(la.astype(np.int64).sum()) == idx.numel()      # the segments tile idx exactly
    # and the canonical form the contract names
    a, la = oracle.canonical_neighbours(npy(idx), npy(sl))
    b, lb = oracle.canonical_neighbours(ridx, rsl)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(la, lb)
    return npy(idx), npy(sl)


def test_ballquery_scene_raw_and_shifted(ops, oracle):
    s = object_subset(small_batch(3, 15000))
    _check_bq(ops, oracle, s["coords"], s["batch_idxs"], s["batch_offsets"], 0.03, 50)
    _, sl = _check_bq(ops, oracle, s["shifted"], s["batch_idxs"], s["batch_offsets"], 0.03, 300)
    assert sl[:, 1].max() > 100          # the shifted coordinates really are dense


def test_ballquery_cap_1000(ops, oracle):
    """More than 1000 points inside one ball: only the first 1000 by index are kept (bfs_cluster.cu:38-43)."""
    rng = np.random.default_rng(2)
    blob = rng.normal(0, 0.004, (2600, 3))
    rest = rng.uniform(-1, 1, (3000, 3))
    xyz = np.concatenate([blob, rest]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    bi = np.zeros(len(xyz), np.int32)
    bo = np.array([0, len(xyz)], np.int32)
    _, sl = _check_bq(ops, oracle, xyz, bi, bo, 0.03)
    assert (sl[:, 1] == 1000).sum() > 500


def test_ballquery_matches_literal_scan(ops, oracle):
    """Against the oracle's literal O(n^2) restatement (no grid) on a small input."""
    rng = np.random.default_rng(4)
    xyz = rng.uniform(0, 0.3, (3000, 3)).astype(np.float32)
    bi = np.repeat(np.arange(3), 1000).astype(np.int32)
    bo = np.array([0, 1000, 2000, 3000], np.int32)
    idx, sl = ops.ballquery_batch_p(cu(xyz), cu(bi), cu(bo), 0.03, 50)
    ridx, rsl = oracle.ballquery_batch_p(xyz, bi, bo, 0.03, use_grid=False)
    a, la = oracle.relaid_neighbours(npy(idx), npy(sl))
    b, lb = oracle.relaid_neighbours(ridx, rsl)
    np.testing.assert_array_equal(la, lb)
    np.testing.assert_array_equal(a, b)


def test_ballquery_specials(ops, oracle):
    """NaN / inf / huge / duplicate coordinates, empty scenes, negative and zero radius."""
    rng = np.random.default_rng(6)
    xyz = rng.uniform(0, 0.2, (2000, 3)).astype(np.float32)
    xyz[5] = np.nan
    xyz[17, 1] = np.inf
    xyz[33] = -np.inf
    xyz[40:60] = xyz[40]                 # exact duplicates
    xyz[100:110] = 3e7                   # far: fp32 spacing above the radius
    xyz[110:120] = [1e30, -1e30, 1e30]
    xyz[120] = [3e7, 3e7, 3e7 + 4]
    bi = np.concatenate([np.zeros(900), np.full(1100, 2)]).astype(np.int32)     # scene 1 is empty
    bo = np.array([0, 900, 900, 2000], np.int32)
    for r in (0.03, -0.03, 0.0, 1e-30, 5.0):
        _check_bq(ops, oracle, xyz, bi, bo, r)


def test_ballquery_empty(ops):
    idx, sl = ops.ballquery_batch_p(torch.zeros((0, 3)).cuda(), torch.zeros(0, dtype=torch.int32).cuda(),
                                    torch.zeros(2, dtype=torch.int32).cuda(), 0.03, 50)
    assert idx.numel() == 0 and tuple(sl.shape) == (0, 2)


# ------------------------------------------------------------------------------------------------
# bfs_cluster
# ------------------------------------------------------------------------------------------------
def _check_bfs(ops, oracle, sem, idx, sl, thr, expect_generic=None):
    from d3net_b200 import PG_OP
    ci, co, generic = PG_OP.bfs_cluster_impl(cu(sem), cu(idx), cu(sl), thr)
    rci, rco = oracle.bfs_cluster(sem, idx, sl, thr)
    np.testing.assert_array_equal(npy(co), rco)                 # same clusters, same order, same sizes
    got = oracle.canonical_clusters(npy(ci), npy(co))
    want = oracle.canonical_clusters(rci, rco)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    # members ascend inside each cluster and the first member is the seed the reference starts from
    cin = npy(ci)
    for c in range(len(rco) - 1):
        seg = cin[rco[c]:rco[c + 1], 1]
        assert (np.diff(seg) > 0).all()
        assert seg[0] == rci[rco[c], 1]
    if expect_generic is not None:
        assert generic == expect_generic
    return len(rco) - 1


def test_bfs_kat(ops):
    """Known-answer vector recorded from the reference binary."""
    sem = torch.tensor([1, 1, 1, 2, 2, 1], dtype=torch.int32)
    lists = [[0, 1], [0, 1, 2], [1, 2, 5], [3, 4], [3, 4], [2, 5]]
    idx = torch.tensor(sum(lists, []), dtype=torch.int32)
    lens = [len(l) for l in lists]
    sl = torch.tensor([[sum(lens[:i]), lens[i]] for i in range(6)], dtype=torch.int32)
    ci, co = ops.bfs_cluster(sem, idx, sl, 2)                    # CPU tensors in, CPU tensors out
    assert not ci.is_cuda
    assert ci.tolist() == [[0, 0], [0, 1], [0, 2], [0, 5], [1, 3], [1, 4]]
    assert co.tolist() == [0, 4, 6]


@pytest.mark.parametrize("thr", [50, 1, 400])
def test_bfs_scene(ops, oracle, thr):
    s = object_subset(small_batch(3, 15000))
    for key, ma in (("coords", 50), ("shifted", 300)):
        idx, sl = oracle.ballquery_batch_p(s[key], s["batch_idxs"], s["batch_offsets"], 0.03)
        n = _check_bfs(ops, oracle, s["sem"], idx, sl, thr, expect_generic=False)
        if thr == 50:
            assert n >= 3


def test_bfs_truncated_lists_are_directed(ops, oracle):
    """Lists cut at 1000 make the graph directed (hard part 2)."""
    rng = np.random.default_rng(12)
    blobs = [rng.normal(c, 0.006, (1800, 3)) for c in ([0, 0, 0], [0.05, 0, 0], [0.3, 0.3, 0])]
    bridge = np.linspace([0, 0, 0], [0.3, 0.3, 0], 40)
    xyz = np.concatenate(blobs + [bridge, rng.uniform(-1, 1, (2000, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    bi = np.zeros(len(xyz), np.int32)
    bo = np.array([0, len(xyz)], np.int32)
    idx, sl = oracle.ballquery_batch_p(xyz, bi, bo, 0.03)
    assert (sl[:, 1] == 1000).any()
    sem = np.ones(len(xyz), np.int32)
    _check_bfs(ops, oracle, sem, idx, sl, 10, expect_generic=False)
    sem2 = rng.integers(1, 3, len(xyz)).astype(np.int32)
    _check_bfs(ops, oracle, sem2, idx, sl, 5, expect_generic=False)


def test_bfs_one_way_edges_between_components(ops, oracle):
    """A hand-built truncated relation where a one-way edge is the ONLY link between two components."""
    # points 0..999 + 1000 form a clique-like star around node 1001 whose list is full (1000 entries:
    # 0..999) so that 1001 -> k has a reverse only for k <= 999; node 1000 lists 1001 but 1001 does not
    # list 1000: a one-way edge 1000 -> 1001.
    N = 1003
    lists = [[k, 1001] for k in range(1000)]          # k <-> 1001 (symmetric: k <= last(1001) = 999)
    lists.append([1000, 1001])                        # 1000 -> 1001 one-way
    lists.append(list(range(1000)))                   # 1001: full list, last = 999
    lists.append([1002])                              # isolated
    idx = np.array(sum(lists, []), np.int32)
    lens = np.array([len(l) for l in lists])
    sl = np.column_stack([np.concatenate([[0], np.cumsum(lens)[:-1]]), lens]).astype(np.int32)
    sem = np.ones(N, np.int32)
    _check_bfs(ops, oracle, sem, idx, sl, 1, expect_generic=False)
    # and the mirror case: the one-way edge leaves the big component (0 -> ... cannot come back)
    lists2 = [[0, 1], [1]] + [[k] for k in range(2, 10)]
    idx2 = np.array(sum(lists2, []), np.int32)
    lens2 = np.array([len(l) for l in lists2])
    sl2 = np.column_stack([np.concatenate([[0], np.cumsum(lens2)[:-1]]), lens2]).astype(np.int32)
    _check_bfs(ops, oracle, np.ones(10, np.int32), idx2, sl2, 1, expect_generic=True)


@pytest.mark.parametrize("seed", range(6))
def test_bfs_random_directed_graphs(ops, oracle, seed):
    """Arbitrary directed graphs (not a ball-query product): the checksum must route them to the
    generic propagation path and the result must still equal the reference's BFS."""
    rng = np.random.default_rng(seed)
    N = int(rng.integers(50, 3000))
    deg = rng.integers(0, 5, N)
    lists = [np.sort(rng.choice(N, d, replace=False)) for d in deg]
    idx = np.concatenate(lists + [np.zeros(0, np.int64)]).astype(np.int32)
    sl = np.column_stack([np.concatenate([[0], np.cumsum(deg)[:-1]]), deg]).astype(np.int32)
    sem = rng.integers(0, 3, N).astype(np.int32)
    _check_bfs(ops, oracle, sem, idx, sl, int(rng.integers(1, 4)), expect_generic=True)


def test_bfs_long_chain(ops, oracle):
    """Component depth: a 60k-point path with shuffled numbering."""
    rng = np.random.default_rng(1)
    N = 60000
    order = rng.permutation(N)
    nb = [[] for _ in range(N)]
    for a, b in zip(order[:-1], order[1:]):
        nb[a].append(b)
        nb[b].append(a)
    lists = [np.sort(np.array(l + [i])) for i, l in enumerate(nb)]
    lens = np.array([len(l) for l in lists])
    idx = np.concatenate(lists).astype(np.int32)
    sl = np.column_stack([np.concatenate([[0], np.cumsum(lens)[:-1]]), lens]).astype(np.int32)
    _check_bfs(ops, oracle, np.ones(N, np.int32), idx, sl, 50, expect_generic=False)
    from d3net_b200 import PG_OP
    ci, co, g = PG_OP.bfs_cluster_impl(cu(np.ones(N, np.int32)), cu(idx), cu(sl), 50, generic=1)
    assert g and co.tolist() == [0, N]


def test_bfs_trusted_lists_by_provenance(ops, oracle):
    """Lists that come straight from this library's ball query carry a provenance stamp and take the
    sweep without validation; any in-place write, slice or copy drops the stamp.  Both sweeps must give
    the reference's clusters (lists with truncation included)."""
    from d3net_b200 import PG_OP, pointgroup_ops as pgo
    rng = np.random.default_rng(3)
    blobs = [rng.normal(c, 0.006, (1700, 3)) for c in ([0, 0, 0], [0.045, 0, 0], [0.4, 0.1, 0])]
    xyz = np.concatenate(blobs + [rng.uniform(-1, 1, (3000, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    n = len(xyz)
    bi, bo = np.zeros(n, np.int32), np.array([0, n], np.int32)
    sem = rng.integers(1, 3, n).astype(np.int32)
    idx, sl = ops.ballquery_batch_p(cu(xyz), cu(bi), cu(bo), 0.03, 50)
    assert pgo._stamped(idx, sl) and int((sl[:, 1] == 1000).sum()) > 0
    ridx, rsl = oracle.ballquery_batch_p(xyz, bi, bo, 0.03)
    rci, rco = oracle.bfs_cluster(sem, ridx, rsl, 5)
    want = oracle.canonical_clusters(rci, rco)
    for trusted in (True, False):
        ci, co, generic = PG_OP.bfs_cluster_impl(cu(sem), idx, sl, 5, trusted=trusted)
        assert not generic
        np.testing.assert_array_equal(npy(co), rco)
        for a, b in zip(oracle.canonical_clusters(npy(ci), npy(co)), want):
            np.testing.assert_array_equal(a, b)
    ci, co = ops.bfs_cluster(cu(sem), idx, sl, 5)                 # the operator picks the trusted sweep itself
    np.testing.assert_array_equal(npy(co), rco)
    assert not pgo._stamped(idx.clone(), sl) and not pgo._stamped(idx, sl.clone()) and not pgo._stamped(idx[:-1], sl)
    idx[0] = idx[0]                                               # an in-place write bumps the version
    assert not pgo._stamped(idx, sl)


def _grid_vs_oracle(ops, oracle, xyz, bi, bo, sem, r, thr):
    """bfs_cluster with the ball query's grid (cells already in one component are not swept), the plain
    trusted sweep and the validating sweep must all give the oracle's clusters."""
    from d3net_b200 import PG_OP, pointgroup_ops as pgo
    idx, sl = ops.ballquery_batch_p(cu(xyz), cu(bi), cu(bo), r, 50)
    assert pgo._stamped(idx, sl) and idx._pg_grid is not None
    ridx, rsl = oracle.ballquery_batch_p(xyz, bi, bo, r)
    rci, rco = oracle.bfs_cluster(sem, ridx, rsl, thr)
    want = oracle.canonical_clusters(rci, rco)
    swept = {}
    for name, kw in (("grid", dict(trusted=True, grid_ws=idx._pg_grid)), ("trusted", dict(trusted=True)), ("auto", {})):
        ci, co, generic = PG_OP.bfs_cluster_impl(cu(sem), idx, sl, thr, **kw)
        assert not generic
        np.testing.assert_array_equal(npy(co), rco)
        got = oracle.canonical_clusters(npy(ci), npy(co))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
        swept[name] = PG_OP.bfs_cluster_debug()[4]
    ci, co = ops.bfs_cluster(cu(sem), idx, sl, thr)               # the operator picks the grid sweep itself
    np.testing.assert_array_equal(npy(co), rco)
    assert swept["trusted"] == len(xyz) and swept["grid"] <= len(xyz)
    return swept["grid"]


def test_bfs_grid_sweep_scene(ops, oracle):
    """Scene-shaped input, noisy labels, raw and shifted coordinates (the shifted blobs are where whole
    cells settle before the sweep)."""
    s = object_subset(small_batch(3, 15000))
    n = len(s["sem"])
    for key in ("coords", "shifted"):
        swept = _grid_vs_oracle(ops, oracle, s[key], s["batch_idxs"], s["batch_offsets"], s["sem"], 0.03, 50)
        if key == "shifted":
            assert swept < n // 2, "the cell pass should settle most of the shifted blobs (%d of %d lists swept)" % (swept, n)


def test_bfs_grid_sweep_truncated_and_bridged(ops, oracle):
    """Blobs dense enough to cut lists at 1000 (one-way edges, cells the ball query left early), joined by a
    thin bridge and surrounded by clutter, two interleaved labels."""
    rng = np.random.default_rng(21)
    blobs = [rng.normal(c, 0.006, (1800, 3)) for c in ([0, 0, 0], [0.05, 0, 0], [0.3, 0.3, 0])]
    bridge = np.linspace([0, 0, 0], [0.3, 0.3, 0], 40)
    xyz = np.concatenate(blobs + [bridge, rng.uniform(-1, 1, (2500, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    n = len(xyz)
    bi, bo = np.zeros(n, np.int32), np.array([0, n], np.int32)
    for labels in (np.ones(n, np.int32), rng.integers(1, 3, n).astype(np.int32),
                   np.where(rng.random(n) < 0.05, 7, 1).astype(np.int32)):
        _grid_vs_oracle(ops, oracle, xyz, bi, bo, labels, 0.03, 5)


@pytest.mark.parametrize("r", [0.0, 5.0, 0.012])
def test_bfs_grid_sweep_radius_extremes(ops, oracle, r):
    """r = 0 (one cell per scene, only duplicates are neighbours), a radius that makes each scene a clique,
    and a radius below the point pitch (isolated points)."""
    rng = np.random.default_rng(5)
    xyz = rng.uniform(0, 1, (1500, 3)).astype(np.float32)
    xyz[100:110] = xyz[100]                                       # duplicates
    bi = np.repeat(np.arange(3, dtype=np.int32), 500)
    bo = np.array([0, 500, 1000, 1500], np.int32)
    sem = rng.integers(1, 4, 1500).astype(np.int32)
    _grid_vs_oracle(ops, oracle, xyz, bi, bo, sem, r, 2)


def test_bfs_grid_sweep_density_gradient(ops, oracle):
    """A blob whose density falls off over several radii: last(j) differs from point to point, so the per-cell
    index thresholds of the grid-assisted sweep are what keeps it exact (tests/test_grid_rule.py shows the rule
    without them is violated on this very input)."""
    from test_grid_rule import _gradient_blob
    g = _gradient_blob()
    n = len(g["xyz"])
    rng = np.random.default_rng(2)
    for labels in (np.ones(n, np.int32), g["sem"], np.where(rng.random(n) < 0.05, 7, 1).astype(np.int32)):
        _grid_vs_oracle(ops, oracle, g["xyz"], g["batch_idxs"], g["batch_offsets"], labels, 0.03, 5)


@pytest.mark.parametrize("bad_row", [(2 ** 30, 700), (-5, 3), (10, 2 ** 30), (0, -1), (2 ** 31 - 8, 16)])
def test_bfs_malformed_start_len_raises_cleanly(ops, oracle, bad_row):
    """Rows of start_len that point outside ball_query_idxs: the reference reads out of bounds there
    (bfs_cluster.cpp:40-42); here the validating sweep never dereferences them and the call fails with a clean error
    -- and the device stays usable (no sticky fault), which the good call afterwards proves."""
    from d3net_b200 import PG_OP, _native
    s = object_subset(small_batch(2, 9000))
    idx, sl = oracle.ballquery_batch_p(s["shifted"], s["batch_idxs"], s["batch_offsets"], 0.03)
    bad = sl.copy()
    rows = np.random.default_rng(1).integers(0, len(bad), 50)
    bad[rows] = np.array(bad_row, np.int32)
    for generic in (0, 1):
        with pytest.raises(_native.PgError, match="start_len"):
            PG_OP.bfs_cluster_impl(cu(s["sem"]), cu(idx), cu(bad), 50, generic=generic)
        torch.cuda.synchronize()
    _check_bfs(ops, oracle, s["sem"], idx, sl, 50, expect_generic=False)


def test_bfs_two_threads_do_not_share_grid_workspaces(ops, oracle):
    """Two threads running ballquery_batch_p + bfs_cluster on different inputs: each bfs_cluster gets the grid of
    ITS ball query (the hand-over is per call, not a module global)."""
    import threading
    sets = [object_subset(small_batch(2, 9000, config_id=11)), object_subset(small_batch(1, 14000, config_id=12))]
    want = []
    for s in sets:
        idx, sl = oracle.ballquery_batch_p(s["shifted"], s["batch_idxs"], s["batch_offsets"], 0.03)
        want.append(oracle.bfs_cluster(s["sem"], idx, sl, 50))
    errors = []

    def work(k):
        try:
            s = sets[k]
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                x, b, o, sem = cu(s["shifted"]), cu(s["batch_idxs"]), cu(s["batch_offsets"]), cu(s["sem"])
                for _ in range(6):
                    idx, sl = ops.ballquery_batch_p(x, b, o, 0.03, 300)
                    assert idx._pg_grid is not None and idx._pg_grid.numel() >= 0
                    ci, co = ops.bfs_cluster(sem, idx, sl, 50)
                    np.testing.assert_array_equal(npy(co), want[k][1])
                    for a, c in zip(oracle.canonical_clusters(npy(ci), npy(co)), oracle.canonical_clusters(*want[k])):
                        np.testing.assert_array_equal(a, c)
        except Exception as e:          # noqa: BLE001  (reported by the main thread)
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_bfs_empty(ops):
    z = torch.zeros(0, dtype=torch.int32).cuda()
    ci, co = ops.bfs_cluster(z, z, torch.zeros((0, 2), dtype=torch.int32).cuda(), 50)
    assert tuple(ci.shape) == (0, 2) and co.tolist() == [0]


# ------------------------------------------------------------------------------------------------
# roipool / sec_mean / sec_min / sec_max
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [1, 3, 5, 16, 32, 33, 40, 134])
def test_roipool_and_sec(ops, oracle, C):
    rng = np.random.default_rng(200 + C)
    off = random_segments(rng, 300, 400, big=60000 if C <= 16 else 3000)
    S = int(off[-1])
    x = rng.standard_normal((S + 17, C)).astype(np.float32)      # rows beyond offsets[-1] must be ignored
    x[rng.random(x.shape) < 0.02] = 0.5                           # ties
    xt = cu(x).requires_grad_(True)
    offt = cu(off)
    out = ops.roipool(xt, offt)
    ref, refidx = oracle.roipool(x, off)
    assert npy(out).tobytes() == ref.tobytes()
    g = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(cu(g))
    np.testing.assert_array_equal(npy(xt.grad), oracle.roipool_bp(g, refidx, x.shape[0]))
    assert npy(ops.sec_max(cu(x), offt)).tobytes() == oracle.sec_max(x, off).tobytes()
    assert npy(ops.sec_min(cu(x), offt)).tobytes() == oracle.sec_min(x, off).tobytes()
    m, rm = npy(ops.sec_mean(cu(x), offt)), oracle.sec_mean(x, off)
    np.testing.assert_allclose(m, rm, rtol=1e-6, atol=0)          # the contract
    assert m.tobytes() == rm.tobytes()                            # and in fact bit-exact


def test_roipool_specials(ops, oracle):
    """NaN is never selected, -inf never beats the -inf start, ties keep the lowest row, empty -> (-inf, -1)."""
    x = np.array([[np.nan, -np.inf, 1.0, 2.0], [3.0, -np.inf, 1.0, np.nan], [np.nan, -np.inf, 0.5, 2.0],
                  [np.nan, np.nan, np.nan, np.nan], [-0.0, 0.0, np.inf, -np.inf]], np.float32)
    off = np.array([0, 3, 3, 4, 5], np.int32)
    from d3net_b200 import PG_OP
    out = torch.empty((4, 4), device="cuda")
    arg = torch.empty((4, 4), dtype=torch.int32, device="cuda")
    PG_OP.roipool_fp(cu(x), cu(off), out, arg, 4, 4)
    ref, refidx = oracle.roipool(x, off)
    assert_same_floats(npy(out), ref)
    np.testing.assert_array_equal(npy(arg), refidx)
    for fn, rf in ((ops.sec_max, oracle.sec_max), (ops.sec_min, oracle.sec_min), (ops.sec_mean, oracle.sec_mean)):
        assert_same_floats(npy(fn(cu(x), cu(off))), rf(x, off))


def test_sec_mean_on_cluster_coords(ops, oracle):
    """The shape the model uses: C = 3 coordinates gathered per cluster (model/pointgroup.py:137-142)."""
    s = object_subset(small_batch(2, 20000))
    idx, sl = oracle.ballquery_batch_p(s["coords"], s["batch_idxs"], s["batch_offsets"], 0.03)
    ci, co = oracle.bfs_cluster(s["sem"], idx, sl, 50)
    pts = s["coords"][ci[:, 1]]
    for fn, rf in ((ops.sec_mean, oracle.sec_mean), (ops.sec_min, oracle.sec_min), (ops.sec_max, oracle.sec_max)):
        assert npy(fn(cu(pts), cu(co))).tobytes() == rf(pts, co).tobytes()


# ------------------------------------------------------------------------------------------------
# get_iou
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nI", [1, 37, 400, 9000])
def test_get_iou(ops, oracle, nI):
    rng = np.random.default_rng(300 + nI)
    N = 50000
    labels = rng.integers(-1, nI, N).astype(np.int64)
    pointnum = np.bincount(labels[labels >= 0], minlength=nI).astype(np.int32)
    off = random_segments(rng, 120, 600, big=20000)
    pidx = rng.integers(0, N, int(off[-1])).astype(np.int32)
    iou = ops.get_iou(cu(pidx), cu(off), cu(labels), cu(pointnum))
    ref = oracle.get_iou(pidx, off, labels, pointnum)
    assert npy(iou).tobytes() == ref.tobytes()


# ------------------------------------------------------------------------------------------------
# the whole chain, traced and replayed op by op
# ------------------------------------------------------------------------------------------------
def test_chain_trace_replay(ops):
    from d3net_b200 import chain
    from oracle import replay
    nb = small_batch(3, 20000, config_id=5)
    batch = chain.batch_to_device(nb, torch.device("cuda"))
    trace = {}
    out = chain.proposal_chain(ops, batch, trace=trace)
    checked = replay.check_trace(trace)
    assert len(checked) == 12
    assert out["proposals_offset"].numel() - 1 >= 10


def test_pack_proposals(ops):
    from d3net_b200 import chain, dist as pgdist
    nb = small_batch(3, 12000, config_id=6)
    batch = chain.batch_to_device(nb, torch.device("cuda"))
    out = chain.proposal_chain(ops, batch)
    packed = pgdist.pack_proposals(out, batch, 64)
    assert tuple(packed.shape) == (3, 64, pgdist.PACK_WIDTH)
    offs = out["proposals_offset"].long()
    scene = batch["locs_scaled"][out["proposals_idx"][offs[:-1], 1].long(), 0]
    for b in range(3):
        k = min(int((scene == b).sum()), 64)
        assert int(packed[b, :, 45].sum()) == k
        first = torch.nonzero(scene == b).view(-1)[0]
        assert torch.equal(packed[b, 0, :16], out["proposals_score_feats"][first])
    assert torch.equal(pgdist.all_gather_proposals(packed), packed)      # world size 1: identity
    for P in (64, 5, 1):                                                 # the pack kernels against the torch restatement
        a, b = pgdist.pack_proposals(out, batch, P), pgdist.pack_proposals_torch(out, batch, P)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_gather_rows(ops):
    rng = np.random.default_rng(77)
    for C in (1, 3, 16, 134):
        src = rng.standard_normal((5000, C)).astype(np.float32)
        for dt in (torch.int32, torch.int64):
            idx = rng.integers(0, 5000, 12345)
            s = cu(src).requires_grad_(True)
            out = ops.gather_rows(s, cu(idx, dt))
            np.testing.assert_array_equal(npy(out), src[idx])
            out.sum().backward()
            np.testing.assert_array_equal(npy(s.grad)[:, 0], np.bincount(idx, minlength=5000).astype(np.float32))


def test_ballquery_mask_and_recompute_paths_agree(ops):
    """The fill phase either decodes hit masks recorded by the count phase or re-evaluates the predicates;
    both must give the same lists (the oracle comparison above runs the mask path)."""
    from d3net_b200 import PG_OP
    s = object_subset(small_batch(2, 15000))
    xyz, bi, bo = cu(s["shifted"]), cu(s["batch_idxs"]), cu(s["batch_offsets"])
    res = []
    # masks on / masks off / a mask buffer too small for the batch (the device falls back by itself)
    for use_masks, words, expect in ((True, None, True), (False, None, False), (True, 4096, False)):
        sl, total, state = PG_OP.ballquery_count_impl(xyz, bi, bo, 0.03, use_masks=use_masks, mask_words=words)
        assert (state[1] is not None) == expect
        idx = torch.empty(total, dtype=torch.int32, device="cuda")
        PG_OP.ballquery_fill_impl(xyz, 0.03, sl, idx, state)
        res.append((sl, idx))
    from util import relaid_on_device
    want = relaid_on_device(res[0][1], res[0][0])
    for r in res[1:]:
        assert torch.equal(res[0][0][:, 1], r[0][:, 1]) and torch.equal(want, relaid_on_device(r[1], r[0]))


def test_fused_cluster_glue_equals_torch_sequence(ops):
    """the fused clusters_voxelization glue gives, bit for bit, what the reference's
    torch op sequence (model/pointgroup.py:125-167) gives on the GPU -- coordinates, centres, sizes and
    everything downstream -- including a non-trivial random offset."""
    from d3net_b200 import chain, scenes
    nb = scenes.make_batch(2, 20000, config_id=5, geometry_points=20000)
    batch = chain.batch_to_device(nb, torch.device("cuda"))
    rand6 = torch.tensor([0.11, 0.52, 0.93, 0.37, 0.08, 0.64], device="cuda")
    a = chain.proposal_chain(ops, batch, rand6)
    b = chain.proposal_chain(ops, batch, rand6, fused_glue=True)
    assert a["proposals_offset"].numel() > 3
    for k in ("proposals_center", "proposals_size", "proposals_voxel_coords", "proposals_voxel_feats",
              "proposals_score_feats", "ious"):
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        if a[k].dtype.is_floating_point:
            assert torch.equal(a[k].view(torch.int32), b[k].view(torch.int32)), k
        else:
            assert torch.equal(a[k], b[k]), k


# ------------------------------------------------------------------------------------------------
# fused ballquery_batch_p + bfs_cluster (lazy neighbour lists)
# ------------------------------------------------------------------------------------------------
def _fused_vs_separate(ops, oracle, xyz, bi, bo, sem, r, thr):
    x, b, o, s_ = cu(xyz), cu(bi), cu(bo), cu(sem)
    idx, sl = ops.ballquery_batch_p(x, b, o, r, 300)
    ci, co = ops.bfs_cluster(s_, idx, sl, thr)
    fi, fo, total = ops.ballquery_bfs_cluster(x, b, o, r, 300, s_, thr)
    assert total == idx.numel()
    assert torch.equal(co, fo) and torch.equal(ci, fi)
    oidx, osl = oracle.ballquery_batch_p(xyz, bi, bo, r)
    rci, rco = oracle.bfs_cluster(sem, oidx, osl, thr)
    np.testing.assert_array_equal(npy(fo), rco)
    for a, c in zip(oracle.canonical_clusters(npy(fi), npy(fo)), oracle.canonical_clusters(rci, rco)):
        np.testing.assert_array_equal(a, c)
    return int(osl[:, 1].max())


def test_fused_cluster_scene(ops, oracle):
    s = object_subset(small_batch(3, 20000))
    for key in ("shifted", "coords"):                      # long lists (lazy path) and short ones (falls back to the two ops)
        _fused_vs_separate(ops, oracle, s[key], s["batch_idxs"], s["batch_offsets"], s["sem"], 0.03, 50)


def test_fused_cluster_truncated_lists(ops, oracle):
    """Blobs whose lists hit the 1000 cap (one-way edges, exact `last` entries from the masks) plus a bridge."""
    rng = np.random.default_rng(12)
    blobs = [rng.normal(c, 0.006, (1800, 3)) for c in ([0, 0, 0], [0.05, 0, 0], [0.3, 0.3, 0])]
    bridge = np.linspace([0, 0, 0], [0.3, 0.3, 0], 40)
    xyz = np.concatenate(blobs + [bridge, rng.uniform(-1, 1, (2000, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    bi = np.zeros(len(xyz), np.int32)
    bo = np.array([0, len(xyz)], np.int32)
    for labels in (np.ones(len(xyz), np.int32), rng.integers(1, 3, len(xyz)).astype(np.int32)):
        mx = _fused_vs_separate(ops, oracle, xyz, bi, bo, labels, 0.03, 5)
        assert mx == 1000


def test_fused_cluster_edge_cases(ops, oracle):
    """Empty input, radius 0 (nobody has a neighbour but itself), one tight blob (every list full), NaN coordinates."""
    z3 = torch.zeros((0, 3), device="cuda")
    zi = torch.zeros(0, dtype=torch.int32, device="cuda")
    ci, co, total = ops.ballquery_bfs_cluster(z3, zi, torch.tensor([0, 0], dtype=torch.int32, device="cuda"), 0.03, 50, zi, 50)
    assert total == 0 and tuple(ci.shape) == (0, 2) and co.tolist() == [0]
    rng = np.random.default_rng(8)
    xyz = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    bi = np.zeros(3000, np.int32)
    bo = np.array([0, 3000], np.int32)
    sem = np.ones(3000, np.int32)
    _fused_vs_separate(ops, oracle, xyz, bi, bo, sem, 0.0, 1)                     # r = 0: 3000 singletons... none (strict <)
    blob = rng.normal(0, 0.003, (2500, 3)).astype(np.float32)                     # every list holds the first 1000 points
    mx = _fused_vs_separate(ops, oracle, blob, np.zeros(2500, np.int32), np.array([0, 2500], np.int32), np.ones(2500, np.int32), 0.03, 10)
    assert mx == 1000
    bad = blob.copy()
    bad[::7] = np.nan
    bad[3::11, 1] = np.inf
    _fused_vs_separate(ops, oracle, bad, np.zeros(2500, np.int32), np.array([0, 2500], np.int32), np.ones(2500, np.int32), 0.03, 10)
