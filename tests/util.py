"""Shared helpers for the parity tests (numpy <-> torch, small synthetic inputs)."""
import numpy as np
import torch

from d3net_b200 import scenes


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def npy(t):
    return t.detach().cpu().numpy()


def object_subset(batch):
    """What model/pointgroup.py:288-295 feeds the clustering: points with semantic_preds > 0."""
    mask = batch["semantic_preds"] > 0
    object_idxs = np.nonzero(mask)[0]
    batch_idxs = batch["locs_scaled"][:, 0].astype(np.int32)
    b_ = batch_idxs[object_idxs]
    B = batch["n_scenes"]
    offs = np.zeros(B + 1, np.int32)
    offs[1:] = np.cumsum(np.bincount(b_, minlength=B))
    return {
        "object_idxs": object_idxs,
        "coords": batch["locs"][object_idxs].astype(np.float32),
        "shifted": (batch["locs"][object_idxs] + batch["pt_offsets"][object_idxs]).astype(np.float32),
        "batch_idxs": b_.astype(np.int32),
        "batch_offsets": offs,
        "sem": batch["semantic_preds"][object_idxs].astype(np.int32),
    }


def small_batch(n_scenes=2, n_points=12000, config_id=9):
    return scenes.make_batch(n_scenes, n_points, config_id=config_id)


def random_segments(rng, n_seg, max_len, empty_frac=0.1, big=None):
    lens = rng.integers(1, max_len + 1, n_seg)
    lens[rng.random(n_seg) < empty_frac] = 0
    if big is not None and n_seg > 2:
        lens[n_seg // 2] = big
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
