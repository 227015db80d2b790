"""Collate-side voxelisation: d3net_b200.collate.sparse_collate_fn against the
dictionary the reference's own sparse_collate_fn (lib/dataset/pipeline.py:917-995) produced for the same scenes
(tests/golden/ref_collate.npz, recorded by `make_golden.py collate` with the reference's compiled voxelize_idx).
Everything is integer / pass-through: bit-exact."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_inputs  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_collate.npz"))


def test_golden_is_consistent(gold):
    """The recorded dictionary itself: offsets, id shifts and the voxel maps agree with the inputs."""
    g = golden_inputs("collate")["batch"]
    counts = [len(b["locs"]) for b in g]
    np.testing.assert_array_equal(gold["batch_offsets"], np.concatenate([[0], np.cumsum(counts)]))
    np.testing.assert_array_equal(gold["locs_scaled"][:, 0], np.repeat(np.arange(3), counts))
    np.testing.assert_array_equal(gold["locs_scaled"][:, 1:], np.concatenate([b["locs_scaled"] for b in g]).astype(np.int64))
    assert gold["instance_ids"].max() == sum(int(b["num_instance"]) for b in g) - 1 and gold["instance_ids"].min() == -1
    assert gold["voxel_locs"].shape[0] < sum(counts)                 # points do share voxels
    np.testing.assert_array_equal(gold["voxel_locs"][gold["p2v_map"]], gold["locs_scaled"])


def test_oracle_reproduces_reference_collate(gold, oracle):
    """oracle/collate_oracle.py (numpy restatement, with the C oracle's voxelization_idx) against the dictionary the
    reference's own sparse_collate_fn produced."""
    from oracle import collate_oracle
    g = golden_inputs("collate")["batch"]
    data = collate_oracle.sparse_collate(g)
    for k in gold.files:
        assert data[k].dtype == gold[k].dtype, (k, data[k].dtype, gold[k].dtype)
        np.testing.assert_array_equal(data[k], gold[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("to_cpu", [False, True])
def test_gpu_collate_equals_reference(gold, to_cpu):
    from d3net_b200 import collate
    g = golden_inputs("collate")["batch"]
    before = copy.deepcopy(g)
    data = collate.sparse_collate_fn(g, to_cpu=to_cpu)
    for k in gold.files:
        got = data[k]
        assert got.is_cuda != to_cpu, k
        got = got.cpu().numpy()
        assert got.dtype == gold[k].dtype, (k, got.dtype, gold[k].dtype)
        np.testing.assert_array_equal(got, gold[k], err_msg=k)
    for k in ("locs", "feats", "instance_info"):
        np.testing.assert_array_equal(data[k].cpu().numpy(), np.concatenate([b[k] for b in before], 0))
    assert data["scene_id"] == [b["scene_id"] for b in before]       # non-array keys: listed, as scannet_collate_fn does
    for a, b in zip(g, before):                                      # unlike the reference, the inputs are left alone
        np.testing.assert_array_equal(a["instance_ids"], b["instance_ids"])


@pytest.mark.gpu
def test_gpu_collate_with_gt_proposals_and_without_instances():
    from d3net_b200 import collate
    rng = np.random.default_rng(4)
    batch = []
    for n, ninst in ((300, 3), (200, 2)):
        locs = rng.uniform(0, 0.3, (n, 3)).astype(np.float32)
        inst = rng.integers(0, ninst, n).astype(np.int32)
        order = np.argsort(inst, kind="stable")
        batch.append({"locs": locs, "locs_scaled": (locs * 50).astype(np.float32), "feats": locs.copy(),
                      "sem_labels": inst.copy(), "instance_ids": inst, "num_instance": np.array(ninst, np.int32),
                      "instance_info": np.zeros((n, 12), np.float32), "instance_num_point": np.bincount(inst, minlength=ninst).astype(np.int32),
                      "gt_proposals_idx": np.stack([inst[order], order], 1).astype(np.int32),
                      "gt_proposals_offset": np.concatenate([[0], np.cumsum(np.bincount(inst, minlength=ninst))]).astype(np.int32)})
    data = collate.sparse_collate_fn(batch)
    gi, go = data["gt_proposals_idx"].cpu().numpy(), data["gt_proposals_offset"].cpu().numpy()
    assert gi.dtype == np.int32 and go.dtype == np.int32
    assert go.tolist() == np.concatenate([[0], np.cumsum(np.concatenate([b["instance_num_point"] for b in batch]))]).tolist()
    inst_all = data["instance_ids"].cpu().numpy()
    np.testing.assert_array_equal(inst_all[gi[:, 1]], gi[:, 0])      # global point -> global instance
    only_points = [{k: b[k] for k in ("locs", "locs_scaled", "feats")} for b in batch]
    d2 = collate.sparse_collate_fn(only_points)
    assert "instance_ids" not in d2 and d2["locs_scaled"].shape == (500, 4) and d2["v2p_map"].shape[0] == d2["voxel_locs"].shape[0]
