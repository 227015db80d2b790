"""Instance NMS of PointGroup.test: cross_iou replaces the dense-mask matmul of
model/pointgroup.py:577-590, nms_instances replaces lib/utils/eval.py:75-97.

CPU: the numpy oracle against golden vectors recorded from the reference's own Python (tests/golden/ref_nms.npz,
written by `make_golden.py nms`).  GPU: the CUDA ops, through the C ABI, against the oracle and the goldens --
IoUs bit-exact (NaN for empty proposals included), picks identical."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_inputs  # noqa: E402
from oracle import nms_oracle  # noqa: E402
from util import assert_same_floats  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_nms.npz"))


def test_oracle_reproduces_reference_python(gold):
    g = golden_inputs("nms")
    cross, npoint = nms_oracle.cross_iou(g["proposals_idx"], g["num_proposals"], g["N"])
    assert_same_floats(cross, gold["cross_ious"])
    np.testing.assert_array_equal(npoint, gold["npoint"])
    assert np.isnan(cross[-1, -1]) and (cross[-1, :-1] == 0).all() and npoint[-1] == 0    # no points: 0 / 0 on the diagonal
    for thr in (0.3, 0.1, 0.6):
        np.testing.assert_array_equal(nms_oracle.nms_instances(gold["cross_ious"], g["scores"], thr), gold["pick_%g" % thr])


def _random_case(rng, N, nP, density, dup=0):
    rows = []
    for p in range(nP):
        k = int(rng.integers(0, max(2, int(density * N))))
        if p % 7 == 3:
            k = 0                                                   # proposals without points
        pts = rng.choice(N, size=min(k, N), replace=False)
        rows += [(p, int(q)) for q in pts]
    rows += [rows[i] for i in rng.integers(0, max(len(rows), 1), dup)] if rows else []
    pidx = np.array(rows, np.int32).reshape(-1, 2)
    return pidx[rng.permutation(len(pidx))]                         # any row order


@pytest.mark.gpu
def test_gpu_golden(gold):
    from d3net_b200 import pointgroup_ops as ops
    g = golden_inputs("nms")
    pidx = torch.from_numpy(g["proposals_idx"]).cuda()
    cross, npoint = ops.cross_iou(pidx, g["num_proposals"], g["N"], want_npoint=True)
    assert cross.dtype == torch.float32 and tuple(cross.shape) == (g["num_proposals"],) * 2
    assert_same_floats(cross.cpu().numpy(), gold["cross_ious"])
    np.testing.assert_array_equal(npoint.cpu().numpy(), gold["npoint"])
    scores = torch.from_numpy(g["scores"]).cuda()
    for thr in (0.3, 0.1, 0.6):
        pick = ops.nms_instances(cross, scores, thr)
        assert pick.dtype == torch.int32
        np.testing.assert_array_equal(pick.cpu().numpy(), gold["pick_%g" % thr])
    # CPU tensors in -> CPU tensors out, like the other ops that the reference runs on the host
    c2 = ops.cross_iou(torch.from_numpy(g["proposals_idx"]), g["num_proposals"], g["N"])
    assert not c2.is_cuda
    assert_same_floats(c2.numpy(), gold["cross_ious"])


@pytest.mark.gpu
@pytest.mark.parametrize("N,nP,density,dup", [(500, 1, 0.5, 0), (2000, 40, 0.05, 100), (300, 97, 0.9, 500), (5000, 300, 0.01, 0)])
def test_gpu_random_vs_oracle(N, nP, density, dup):
    from d3net_b200 import pointgroup_ops as ops
    rng = np.random.default_rng(N + nP)
    pidx = _random_case(rng, N, nP, density, dup)
    want, wnp = nms_oracle.cross_iou(pidx, nP, N)
    cross, npoint = ops.cross_iou(torch.from_numpy(pidx).cuda(), nP, N, want_npoint=True)
    assert_same_floats(cross.cpu().numpy(), want)
    np.testing.assert_array_equal(npoint.cpu().numpy(), wnp)
    scores = rng.permutation(nP).astype(np.float32)
    scores[::5] = scores[0]                                          # ties: lower index first, both sides
    for thr in (0.0, 0.25, 1.0):
        pick = ops.nms_instances(cross, torch.from_numpy(scores).cuda(), thr)
        np.testing.assert_array_equal(pick.cpu().numpy(), nms_oracle.nms_instances(want, scores, thr))


@pytest.mark.gpu
def test_gpu_chain_proposals_and_edge_cases():
    """The proposals the hot path itself produces (two clusterings of the same points), checked against the
    oracle; the empty inputs; rows out of range are an error, not a wild write."""
    from d3net_b200 import chain, pointgroup_ops as ops, _native
    from util import small_batch
    b = chain.batch_to_device(small_batch(2, 12000), torch.device("cuda", 0))
    out = chain.proposal_chain(ops, b)
    pidx, off = out["proposals_idx"], out["proposals_offset"]
    nP, N = off.numel() - 1, b["locs"].shape[0]
    cross, npoint = ops.cross_iou(pidx, nP, N, want_npoint=True)
    want, wnp = nms_oracle.cross_iou(pidx.cpu().numpy(), nP, N)
    assert_same_floats(cross.cpu().numpy(), want)
    np.testing.assert_array_equal(npoint.cpu().numpy(), np.diff(off.cpu().numpy()))      # model/pointgroup.py:342-344
    scores = torch.rand(nP, generator=torch.Generator().manual_seed(1)).cuda()
    pick = ops.nms_instances(cross, scores, 0.3)
    np.testing.assert_array_equal(pick.cpu().numpy(), nms_oracle.nms_instances(want, scores.cpu().numpy(), 0.3))
    assert 0 < pick.numel() < nP
    # empty
    e = ops.cross_iou(torch.zeros((0, 2), dtype=torch.int32).cuda(), 0, 10)
    assert tuple(e.shape) == (0, 0)
    e = ops.cross_iou(torch.zeros((0, 2), dtype=torch.int32).cuda(), 3, 10)
    assert tuple(e.shape) == (3, 3) and torch.isnan(e).all()
    assert ops.nms_instances(torch.zeros((0, 0)).cuda(), torch.zeros(0).cuda(), 0.3).numel() == 0
    with pytest.raises(_native.PgError):
        ops.cross_iou(torch.tensor([[0, 10]], dtype=torch.int32).cuda(), 1, 10)


@pytest.mark.gpu
def test_gpu_pick_masks_equal_dense_mask_rows():
    """model/pointgroup.py:579-580,593: rows of the dense mask for the picked proposals (repeats and an empty
    proposal included); out-of-range picks are an error."""
    from d3net_b200 import pointgroup_ops as ops, _native
    rng = np.random.default_rng(9)
    N, nP = 4000, 37
    lens = rng.integers(0, 300, nP)
    lens[5] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rows = np.concatenate([np.stack([np.full(l, p), rng.choice(N, l, replace=False)], 1) for p, l in enumerate(lens)]).astype(np.int32)
    mask = np.zeros((nP, N), np.int32)
    mask[rows[:, 0], rows[:, 1]] = 1
    pick = np.array([3, 36, 5, 3, 0, 20], np.int32)
    got = ops.pick_masks(torch.from_numpy(rows).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(pick).cuda(), N)
    assert got.dtype == torch.int32 and got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), mask[pick])
    assert tuple(ops.pick_masks(torch.from_numpy(rows).cuda(), torch.from_numpy(off).cuda(), torch.zeros(0, dtype=torch.int32).cuda(), N).shape) == (0, N)
    with pytest.raises(_native.PgError):
        ops.pick_masks(torch.from_numpy(rows).cuda(), torch.from_numpy(off).cuda(), torch.tensor([nP], dtype=torch.int32).cuda(), N)
