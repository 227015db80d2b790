"""The skip rule of the grid-assisted bfs_cluster sweep (d3net_b200/csrc/cluster.cu), checked
on the CPU against its own claim.

The kernels skip the neighbour list of a point i when i carries its cell's main (label, root) pair and its index
is at most the cell's threshold L.  That is sound iff, for every skipped i and every j in list(i) with the same
label,
        root(j) == root(i)      or      (j is NOT skipped  and  i is in list(j)),
because then every union / parked edge the sweep would have derived from list(i) is derived from list(j) instead.
This test restates the rule in numpy -- cells, main pairs, the three per-cell minima T1/T2/T3 of last(j), the
thresholds L -- on real ball-query output with truncated lists (the oracle's, 1000-entry cap included) and an
arbitrary PARTIAL partition (what the sampling rounds leave behind), and checks the claim edge by edge.  It also
checks that the rule is not vacuous (a good share of the lists is skipped)."""
import os
import sys

import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_inputs  # noqa: E402

CAP = 1000
INF = 0x7fffffff


def _edges(idx, sl):
    n = len(sl)
    src = np.repeat(np.arange(n), sl[:, 1])
    pos = np.concatenate([np.arange(s, s + l) for s, l in sl]) if n else np.zeros(0, np.int64)
    return src, idx[pos].astype(np.int64)


def _partial_partition(rng, n, src, dst, lab, full, last, keep_frac):
    """Connected components of a random subset of the two-way same-label edges; root = smallest member."""
    twoway = (lab[src] == lab[dst]) & (~full[dst] | (src <= last[dst])) & (~full[src] | (dst <= last[src]))
    pick = twoway & (rng.random(len(src)) < keep_frac)
    g = coo_matrix((np.ones(pick.sum(), np.int8), (src[pick], dst[pick])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    root = np.full(comp.max() + 1, n, np.int64)
    np.minimum.at(root, comp, np.arange(n))
    return root[comp]


def _skip_set(xyz, scene, r, lab, root, full, last):
    """numpy restatement of k_cl_cell_main / _thresholds / _settle / _worklist."""
    n = len(xyz)
    s = abs(float(r)) * 1.0001
    cc = np.floor(xyz.astype(np.float64) / s).astype(np.int64)
    key = np.column_stack([scene.astype(np.int64), cc])
    uniq, cell = np.unique(key, axis=0, return_inverse=True)
    cell = cell.reshape(-1)
    nC = len(uniq)
    lookup = {tuple(k): c for c, k in enumerate(uniq)}
    members = [[] for _ in range(nC)]
    for i in range(n):                                   # ascending index inside a cell, like sorted_pt
        members[cell[i]].append(i)
    M = np.zeros(nC, np.int64)
    R = np.zeros(nC, np.int64)
    for c, pts in enumerate(members):
        M[c], R[c] = lab[pts[0]], root[pts[0]]
        if len(pts) >= 3:
            a, b = pts[1], pts[2]
            if lab[a] == lab[b] and root[a] == root[b] and (lab[a] != M[c] or root[a] != R[c]):
                M[c], R[c] = lab[a], root[a]
    T = np.full((nC, 3), INF, np.int64)
    for i in np.nonzero(full)[0]:
        c = cell[i]
        T[c, 2] = min(T[c, 2], last[i])
        if lab[i] == M[c]:
            T[c, 1] = min(T[c, 1], last[i])
            if root[i] != R[c]:
                T[c, 0] = min(T[c, 0], last[i])
    L = np.zeros(nC, np.int64)
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]
    for c in range(nC):
        lim = T[c, 0]
        for dx, dy, dz in offs:
            b = lookup.get((uniq[c, 0], uniq[c, 1] + dx, uniq[c, 2] + dy, uniq[c, 3] + dz))
            if b is None:
                continue
            if M[b] != M[c]:
                lim = min(lim, T[b, 2])
            elif R[b] == R[c]:
                lim = min(lim, T[b, 0])
            elif R[c] < R[b]:
                lim = min(lim, T[b, 1])
            else:
                lim = -1
                break
        L[c] = lim
    idxs = np.arange(n)
    return (idxs <= L[cell]) & (lab == M[cell]) & (root == R[cell])


def _gradient_blob():
    """A blob whose density falls off over several radii: balls in the core hold far more than 1000 points, balls
    at the fringe fewer, so last(j) varies from point to point and the index threshold is what keeps the rule
    sound (without it this input violates the claim)."""
    rng = np.random.default_rng(77)
    xyz = np.concatenate([rng.normal(0, 0.016, (5200, 3)), rng.uniform(-0.2, 0.2, (800, 3))]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    n = len(xyz)
    return {"xyz": xyz, "batch_idxs": np.zeros(n, np.int32), "batch_offsets": np.array([0, n], np.int32), "radius": 0.03,
            "sem": rng.integers(1, 3, n).astype(np.int32)}


@pytest.mark.parametrize("data", ["golden", "gradient"])
@pytest.mark.parametrize("keep_frac,labels", [(1.0, "two"), (0.02, "two"), (0.004, "noisy"), (0.0, "one"), (1.0, "one")])
def test_skipped_lists_are_covered_from_the_other_side(oracle, keep_frac, labels, data):
    # golden: a 2900-point blob (lists cut at 1000) + clutter, two scenes
    g = golden_inputs("graph") if data == "golden" else _gradient_blob()
    xyz, scene, r = g["xyz"], g["batch_idxs"], g["radius"]
    n = len(xyz)
    rng = np.random.default_rng(int(keep_frac * 1000) + len(labels))
    lab = {"two": g["sem"], "one": np.ones(n, np.int32),
           "noisy": np.where(rng.random(n) < 0.05, 7, 1).astype(np.int32)}[labels].astype(np.int64)
    idx, sl = oracle.ballquery_batch_p(xyz, scene, g["batch_offsets"], r)
    full = sl[:, 1] >= CAP
    assert full.sum() > 500
    last = np.where(full, idx[np.maximum(sl[:, 0] + sl[:, 1] - 1, 0)], INF).astype(np.int64)
    src, dst = _edges(idx, sl)
    root = _partial_partition(rng, n, src, dst, lab, full, last, keep_frac)
    skip = _skip_set(xyz, scene, r, lab, root, full, last)
    # the claim, edge by edge
    e = skip[src] & (lab[src] == lab[dst]) & (root[src] != root[dst])
    i, j = src[e], dst[e]
    listed_back = ~full[j] | (i <= last[j])
    assert not skip[j].any(), "a stray that a skipped point leans on is skipped itself"
    assert listed_back.all(), "a skipped point has a same-label neighbour in another set that does not list it back"
    if keep_frac == 1.0 and data == "golden":            # a fully merged forest: most lists need no sweep
        assert skip.mean() > 0.5, skip.mean()
    if keep_frac == 0.0:                                 # no unions yet: every cell's main root is a single point
        assert skip.sum() <= len(np.unique(root))
