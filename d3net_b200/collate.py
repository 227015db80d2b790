"""Collate-side voxelisation on the GPU.

The reference collates a batch in forked CPU worker processes (lib/dataset/pipeline.py:917-995,
``sparse_collate_fn``): a Python loop concatenates the per-scene arrays, prefixes the scaled coordinates
with the scene id, shifts the instance ids, and runs the single-threaded ``voxelization_idx`` (0.19 s per
150k-point scene).  ``sparse_collate_fn`` here produces the same dictionary -- same keys, dtypes and
values -- with the per-point work and the voxelisation on the device:

  host    one concatenation per key into pinned memory, the two offset tables (B + 1 integers each);
  device  pg_collate_points (scene column + truncation, label widening, instance-id shift: one kernel),
          then voxelization_idx.

The tensors come back on ``device`` (the next thing the caller does with them is ``.cuda()``,
model/pointgroup.py:463-472); pass ``to_cpu=True`` for the reference's placement.  Only the keys of the
PointGroup path are handled (``locs`` ... ``instance_num_point``, optional ``gt_proposals_*``); the
captioning / grounding keys of ``scannet_collate_fn`` (:892-915) are out of scope and passed through
stacked exactly as the reference stacks numpy arrays.
"""
import ctypes

import numpy as np
import torch

from . import PG_OP, pointgroup_ops
from ._native import check

_POINT_KEYS = ("locs", "locs_scaled", "feats", "sem_labels", "instance_ids", "num_instance", "instance_info",
               "instance_num_point", "gt_proposals_idx", "gt_proposals_offset")


def _cat(arrays, dtype, pin):
    """np.concatenate into one (pinned) host tensor of `dtype`."""
    n = sum(a.shape[0] for a in arrays)
    out = torch.empty((n,) + tuple(arrays[0].shape[1:]), dtype=dtype, pin_memory=pin)
    o = out.numpy()
    at = 0
    for a in arrays:
        o[at:at + a.shape[0]] = a
        at += a.shape[0]
    return out


def collate_points(locs_scaled, batch_offsets, sem_labels=None, instance_ids=None, instance_offsets=None):
    """Device tensors: locs_scaled fp32 [N,3], batch_offsets int32 [B+1], sem_labels / instance_ids int32 [N]
    (optional), instance_offsets int32 [B+1] -> (locs_scaled int64 [N,4], sem_labels int64, instance_ids int64)."""
    PG_OP._need(locs_scaled, "locs_scaled", torch.float32)
    PG_OP._need(batch_offsets, "batch_offsets", torch.int32)
    N, B = locs_scaled.size(0), batch_offsets.numel() - 1
    dev = locs_scaled.device
    out_locs = torch.empty((N, 4), dtype=torch.int64, device=dev)
    out_sem = out_inst = None
    if sem_labels is not None:
        PG_OP._need(sem_labels, "sem_labels", torch.int32)
        out_sem = torch.empty(N, dtype=torch.int64, device=dev)
    if instance_ids is not None:
        PG_OP._need(instance_ids, "instance_ids", torch.int32)
        PG_OP._need(instance_offsets, "instance_offsets", torch.int32)
        out_inst = torch.empty(N, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(PG_OP._L().pg_collate_points(PG_OP._p(locs_scaled), PG_OP._p(sem_labels), PG_OP._p(instance_ids),
                                           PG_OP._p(batch_offsets), PG_OP._p(instance_offsets), N, B, PG_OP._p(out_locs),
                                           PG_OP._p(out_sem), PG_OP._p(out_inst), PG_OP._stream()), "collate_points")
    return out_locs, out_sem, out_inst


def sparse_collate_fn(batch, device=None, to_cpu=False, mode=4):
    """lib/dataset/pipeline.py:917-995 for the PointGroup keys.  ``batch``: list of per-scene dicts of numpy
    arrays as PipelineDataset.__getitem__ builds them (:180-187)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    data = {}
    for key in batch[0].keys():                                   # scannet_collate_fn (:892-915), numpy / tensor / list cases
        if key in _POINT_KEYS:
            continue
        v0 = batch[0][key]
        if isinstance(v0, np.ndarray):
            data[key] = torch.stack([torch.from_numpy(s[key]) for s in batch], 0)
        elif isinstance(v0, torch.Tensor):
            data[key] = torch.stack([s[key] for s in batch], 0)
        else:
            data[key] = [s[key] for s in batch]
    if "locs" not in batch[0]:
        return data
    B = len(batch)
    counts = [int(b["locs_scaled"].shape[0]) for b in batch]
    batch_offsets = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)            # :944
    has_inst = "instance_ids" in batch[0]
    up = lambda t: t.to(device, non_blocking=True)
    locs = up(_cat([np.asarray(b["locs"], np.float32) for b in batch], torch.float32, True))               # :970
    feats = up(_cat([b["feats"] for b in batch], torch.from_numpy(batch[0]["feats"]).dtype, True))         # :972 (dtype kept)
    scaled = up(_cat([np.asarray(b["locs_scaled"], np.float32) for b in batch], torch.float32, True))
    d_off = up(batch_offsets)
    sem = inst = inst_off_d = None
    if has_inst:
        ninst = [int(np.asarray(b["num_instance"]).item()) for b in batch]
        instance_offsets = torch.tensor(np.concatenate([[0], np.cumsum(ninst)]), dtype=torch.int32)      # :966
        sem = up(_cat([np.asarray(b["sem_labels"], np.int32) for b in batch], torch.int32, True))
        inst = up(_cat([np.asarray(b["instance_ids"], np.int32) for b in batch], torch.int32, True))
        inst_off_d = up(instance_offsets)
    locs_scaled, sem64, inst64 = collate_points(scaled, d_off, sem, inst, inst_off_d)
    data["locs"], data["locs_scaled"], data["feats"], data["batch_offsets"] = locs, locs_scaled, feats, d_off
    if has_inst:
        data["sem_labels"], data["instance_ids"] = sem64, inst64
        data["instance_info"] = up(_cat([np.asarray(b["instance_info"], np.float32) for b in batch], torch.float32, True))
        data["instance_num_point"] = up(_cat([np.asarray(b["instance_num_point"], np.int32).reshape(-1) for b in batch],
                                             torch.int32, True))
        data["instance_offsets"] = inst_off_d
    if "gt_proposals_idx" in batch[0]:                            # :946-955
        gi, go = [], []
        inst_at = pts_at = off_at = 0
        for i, b in enumerate(batch):
            g = np.array(b["gt_proposals_idx"], np.int64, copy=True)
            g[:, 0] += inst_at
            g[:, 1] += pts_at
            gi.append(g)
            o = np.asarray(b["gt_proposals_offset"], np.int64)
            go.append(o if i == 0 else o[1:] + off_at)
            off_at = int(go[-1][-1])
            inst_at += int(np.asarray(b["num_instance"]).item())
            pts_at += counts[i]
        data["gt_proposals_idx"] = up(torch.from_numpy(np.concatenate(gi, 0)).to(torch.int32))
        data["gt_proposals_offset"] = up(torch.from_numpy(np.concatenate(go, 0)).to(torch.int32))
    # :992 -- the voxelisation, on the device
    data["voxel_locs"], data["p2v_map"], data["v2p_map"] = pointgroup_ops.voxelization_idx(locs_scaled, B, mode)
    if to_cpu:
        torch.cuda.current_stream(device).synchronize()
        data = {k: (v.cpu() if torch.is_tensor(v) and v.is_cuda else v) for k, v in data.items()}
    return data
