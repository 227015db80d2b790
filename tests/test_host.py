"""Host-side logic on CPU: the synthetic scene generator, the chain glue (replayed with the oracle as
the op provider), scene sharding, proposal packing, and the world_size-2 all-gather over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from d3net_b200 import chain, dist as pgdist, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scene_generator_is_deterministic_and_shaped():
    a = scenes.make_scene(20000, seed=5, geometry_points=20000)
    b = scenes.make_scene(20000, seed=5, geometry_points=20000)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert a["locs"].shape == (20000, 3) and a["locs"].dtype == np.float32
    assert a["instance_ids"].min() == -1 and a["instance_ids"].max() + 1 == len(a["instance_pointnum"])
    assert (np.bincount(a["instance_ids"][a["instance_ids"] >= 0]) == a["instance_pointnum"]).all()
    c = scenes.make_scene(20000, seed=6, geometry_points=20000)
    assert not np.array_equal(a["locs"], c["locs"])


def test_batch_collation_mirrors_sparse_collate_fn():
    nb = scenes.make_batch(3, 8000, config_id=3, geometry_points=8000, with_feats=True)
    assert nb["locs_scaled"].shape == (24000, 4) and nb["locs_scaled"].dtype == np.int64
    assert (nb["locs_scaled"][:, 1:] >= 0).all()
    assert nb["batch_offsets"].tolist() == [0, 8000, 16000, 24000]
    assert nb["feats"].shape == (24000, scenes.IN_CHANNELS)
    ids = nb["instance_ids"]
    assert ids.max() + 1 == len(nb["instance_pointnum"])            # batch-global instance ids (pipeline.py:962)
    for b in range(3):
        seg = ids[nb["batch_offsets"][b]:nb["batch_offsets"][b + 1]]
        other = np.concatenate([ids[:nb["batch_offsets"][b]], ids[nb["batch_offsets"][b + 1]:]])
        assert not (set(seg[seg >= 0]) & set(other[other >= 0]))


def test_get_batch_offsets_matches_reference_loop():
    bi = torch.tensor([0, 0, 2, 2, 2, 3], dtype=torch.int32)
    out = chain.get_batch_offsets(bi, 5)
    ref = torch.zeros(6, dtype=torch.int32)
    for i in range(5):                                               # model/pointgroup.py:119-120
        ref[i + 1] = ref[i] + (bi == i).sum()
    assert out.tolist() == ref.tolist()


def test_chain_runs_on_cpu_with_the_oracle_as_op_provider():
    from oracle.ops_adapter import OracleOps
    nb = scenes.make_batch(2, 9000, config_id=4, geometry_points=9000)
    batch = chain.batch_to_device(nb, None)
    out = chain.proposal_chain(OracleOps(use_ref=False), batch)
    nP = out["proposals_offset"].numel() - 1
    assert nP > 5
    assert out["proposals_score_feats"].shape == (nP, scenes.M_CHANNELS)
    assert out["ious"].shape == (nP, len(nb["instance_pointnum"]))
    assert float(out["ious"].max()) > 0.5                            # clusters do line up with the instances
    # every proposal point is an object point of the proposal's own label
    sem = batch["semantic_preds"]
    pidx, off = out["proposals_idx"], out["proposals_offset"]
    for p in range(min(nP, 20)):
        pts = pidx[off[p]:off[p + 1], 1].long()
        assert (sem[pts] == sem[pts[0]]).all() and sem[pts[0]] > 0
        assert off[p + 1] - off[p] >= scenes.CLUSTER_NPOINT_THRE
    packed = pgdist.pack_proposals(out, batch, 32)
    assert packed.shape == (2, 32, pgdist.PACK_WIDTH)
    assert packed[:, :, 45].sum() == min(nP, 64) or packed[:, :, 45].sum() <= nP


def test_scene_shard_partitions():
    for n, w in ((64, 8), (8, 8), (10, 4), (3, 8), (0, 2)):
        blocks = [pgdist.scene_shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
            assert b == c and a <= b
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from d3net_b200 import dist as pgdist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = pgdist.scene_shard(6, rank, world)
packed = torch.full((hi - lo, 4, pgdist.PACK_WIDTH), float(rank + 1))
packed[:, :, 0] = torch.arange(lo, hi, dtype=torch.float32)[:, None]
out = pgdist.all_gather_proposals(packed)
assert out.shape == (6, 4, pgdist.PACK_WIDTH), out.shape
assert out[:, 0, 0].tolist() == [0., 1., 2., 3., 4., 5.]
assert out[:3, :, 1].eq(1).all() and out[3:, :, 1].eq(2).all()
# uneven split (7 scenes over 2 ranks: 4 + 3): padded to ceil(7 / 2) per rank, padding cut out of the result
lo, hi = pgdist.scene_shard(7, rank, world)
packed = torch.full((hi - lo, 4, pgdist.PACK_WIDTH), float(rank + 1))
packed[:, :, 0] = torch.arange(lo, hi, dtype=torch.float32)[:, None]
out = pgdist.all_gather_proposals(packed, n_scenes_total=7)
assert out.shape == (7, 4, pgdist.PACK_WIDTH), out.shape
assert out[:, 0, 0].tolist() == [0., 1., 2., 3., 4., 5., 6.]
assert out[:4, :, 1].eq(1).all() and out[4:, :, 1].eq(2).all()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to the GPU arm): exactly one line on stdout, a
    JSON object with the contract's keys; nothing of the product's CUDA library is needed for it."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-budget-s", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scenes/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["e2e"] == {"value": d["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0,
                                                   "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and "workload" in d["config"]
