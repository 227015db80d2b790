"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the PointGroup part of the reference's collate function.

Allowed importers: tests/ (the product package d3net_b200 never imports this module).

  sparse_collate   lib/dataset/pipeline.py:917-995 (``sparse_collate_fn``) for the keys of the detector path:
                   per-scene arrays are concatenated (:937-968); the voxel-scale coordinates get the scene index as
                   column 0 and are truncated to int64 (:939-943); instance ids other than -1 move up by the number
                   of instances in the scenes before (:963-964); the two offset tables count points and instances
                   (:944,966); voxelization_idx(mode 4) runs on the result (:992).

Parity status: pinned by tests/golden/ref_collate.npz -- the reference's own ``sparse_collate_fn`` source executed in
place with its compiled ``voxelize_idx`` (tests/golden/make_golden.py collate)."""
import numpy as np

from . import pg_oracle


def sparse_collate(batch, mode=4):
    out = {}
    counts = [len(b["locs_scaled"]) for b in batch]
    out["locs"] = np.concatenate([b["locs"] for b in batch], 0).astype(np.float32)
    out["feats"] = np.concatenate([b["feats"] for b in batch], 0)
    scene = np.repeat(np.arange(len(batch), dtype=np.int64), counts)
    scaled = np.concatenate([b["locs_scaled"] for b in batch], 0)
    out["locs_scaled"] = np.column_stack([scene, np.trunc(scaled).astype(np.int64)])        # .long(): toward zero
    out["batch_offsets"] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    if "instance_ids" in batch[0]:
        ninst = [int(np.asarray(b["num_instance"]).item()) for b in batch]
        base = np.concatenate([[0], np.cumsum(ninst)])
        ids = []
        for b, off in zip(batch, base[:-1]):
            i = np.asarray(b["instance_ids"]).astype(np.int64)
            ids.append(np.where(i != -1, i + off, i))
        out["instance_ids"] = np.concatenate(ids)
        out["sem_labels"] = np.concatenate([b["sem_labels"] for b in batch]).astype(np.int64)
        out["instance_info"] = np.concatenate([b["instance_info"] for b in batch], 0).astype(np.float32)
        out["instance_num_point"] = np.concatenate([np.asarray(b["instance_num_point"]).reshape(-1) for b in batch]).astype(np.int32)
        out["instance_offsets"] = base.astype(np.int32)
    out["voxel_locs"], out["p2v_map"], out["v2p_map"] = pg_oracle.voxelization_idx(out["locs_scaled"], len(batch), mode)
    return out
