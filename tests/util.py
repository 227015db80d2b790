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
    # geometry scaled with the point count: the 2 cm pitch (hence connectivity at r = 0.03) is kept
    return scenes.make_batch(n_scenes, n_points, config_id=config_id, geometry_points=n_points)


def random_segments(rng, n_seg, max_len, empty_frac=0.1, big=None):
    lens = rng.integers(1, max_len + 1, n_seg)
    lens[rng.random(n_seg) < empty_frac] = 0
    if big is not None and n_seg > 2:
        lens[n_seg // 2] = big
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


def assert_same_floats(a, b):
    """Bitwise equality, except that any NaN equals any NaN: x86 and the GPU generate different
    quiet-NaN bit patterns for inf - inf (0xFFC00000 vs 0x7FFFFFFF); signed zeros still have to match."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    assert a.shape == b.shape
    na, nb = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(na, nb)
    np.testing.assert_array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def relaid_on_device(idx, sl):
    """Every point's neighbour list re-laid out in point order (torch, on the device): where a producer puts the
    segments inside ``idx`` is its own choice (the reference places them by atomicAdd, bfs_cluster.cu:47)."""
    lens = sl[:, 1].long()
    owner = torch.repeat_interleave(torch.arange(lens.numel(), device=idx.device), lens)
    pos = torch.arange(owner.numel(), device=idx.device) - (torch.cumsum(lens, 0) - lens)[owner]
    return idx[sl[:, 0].long()[owner] + pos]
