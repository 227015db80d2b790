"""The proposal hot path as the detector runs it, restated around the ops of this package.

This is the caller side of the path BASELINE.json names -- input voxelisation, dual-set clustering,
cluster re-voxelisation, proposal pooling, proposal/instance IoU -- following
model/pointgroup.py:266-370 (forward), :125-178 (clusters_voxelization), :445 (get_iou in the loss)
and lib/dataset/pipeline.py:992 (collate-side voxelization_idx) line by line, with two differences:

  * the MinkowskiEngine backbone and score U-Net are out of scope, so the three tensors they would
    produce (per-point features [N, m], semantic_preds, pt_offsets; per-voxel score features) are
    inputs / pass-throughs here;
  * every tensor stays on the device: the reference moves the neighbour lists to the CPU for its BFS
    (model/pointgroup.py:297,305) and the cluster coordinates to the CPU for voxelization_idx (:167);
    with GPU clustering and GPU voxelisation neither round trip is needed.  Python-side loops of the
    reference (get_batch_offsets :112-122) are replaced by their vectorised equivalents.

``ops`` is any module exposing the reference's operator API (d3net_b200.pointgroup_ops in the
product; the tests also trace this function to replay each op's inputs through the oracle).
"""
import torch

from . import scenes


class SectionTimer:
    """CUDA-event timing of named sections on the current stream (bench.py's per-op breakdown)."""

    def __init__(self, enabled=True):
        self.enabled = enabled
        self.events = {}
        self.current = ""

    def start(self, name):
        if not self.enabled:
            return None
        self.current = name
        return self._begin(name)

    def start_sub(self, name):
        """A phase inside the current section (used by pointgroup_ops for its two-phase ops)."""
        if not self.enabled:
            return None
        return self._begin(self.current + "." + name)

    def _begin(self, name):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        self.events.setdefault(name, []).append((a, b))
        return b

    def stop(self, tok):
        if tok is not None:
            tok.record()

    def totals_ms(self):
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.events.items()}

    def counts(self):
        return {k: len(v) for k, v in self.events.items()}


class _NoTimer(SectionTimer):
    def __init__(self):
        super().__init__(False)


def _gather_rows(ops, feats, idx):
    """feats[idx]; through the package's coalesced gather kernel when `ops` offers one."""
    fn = getattr(ops, "gather_rows", None)
    return fn(feats, idx) if fn is not None and feats.is_cuda else feats[idx]


def get_batch_offsets(batch_idxs, batch_size):
    """model/pointgroup.py:112-122 without the Python loop.  The points of a collated batch are laid out
    scene after scene (lib/dataset/pipeline.py:955-963), so batch_idxs ascends and the offsets are the
    positions where each scene id would be inserted."""
    bounds = torch.arange(batch_size + 1, dtype=batch_idxs.dtype, device=batch_idxs.device)
    return torch.searchsorted(batch_idxs, bounds).int()


def _clusters_voxelization_tail(ops, clusters_coords, clusters_feats, n_clusters, mode, timer, trace, box):
    t = timer.start("voxelization_idx(clusters)")
    voxel_coords, p2v_map, v2p_map = ops.voxelization_idx(clusters_coords, n_clusters, mode)
    timer.stop(t)
    if trace is not None:
        trace["voxelization_idx(clusters)"] = (clusters_coords, n_clusters, voxel_coords, p2v_map, v2p_map)
    t = timer.start("voxelization(clusters)")
    voxel_feats = ops.voxelization(clusters_feats, v2p_map, mode)                         # (M, C)
    timer.stop(t)
    if trace is not None:
        trace["voxelization(clusters)"] = (clusters_feats, v2p_map, voxel_feats)
    return voxel_feats, voxel_coords, p2v_map, v2p_map, box


def clusters_voxelization(ops, clusters_idx, clusters_offset, feats, coords, fullscale, scale, mode, rand6,
                          timer=None, trace=None, fused_glue=False):
    """model/pointgroup.py:125-178.  ``rand6`` replaces the two torch.rand(3) draws (:161) so runs are
    reproducible.  Returns (voxel_feats [M,C], voxel_coords int64 [M,4], p2v_map, v2p_map, (center, size)).
    ``fused_glue``: use the package's fused kernel for everything between the gathers and
    voxelization_idx (bit-identical; an edit a caller makes to this function, not to the operator API)."""
    timer = timer or _NoTimer()
    c_idxs = clusters_idx[:, 1].long()
    clusters_feats = _gather_rows(ops, feats, c_idxs)
    if fused_glue:
        t = timer.start("cluster_voxel_coords(fused glue)")
        clusters_coords, center, size = ops.cluster_voxel_coords(coords, clusters_idx, clusters_offset, fullscale, scale,
                                                                 rand6)
        timer.stop(t)
        return _clusters_voxelization_tail(ops, clusters_coords, clusters_feats, clusters_offset.numel() - 1, mode, timer,
                                           trace, (center, size))
    clusters_coords = coords[c_idxs]
    cid = clusters_idx[:, 0].long()

    t = timer.start("sec_mean")
    clusters_coords_mean = ops.sec_mean(clusters_coords, clusters_offset)                # (nCluster, 3)
    timer.stop(t)
    if trace is not None:
        trace["sec_mean"] = (clusters_coords.clone(), clusters_offset, clusters_coords_mean)
    clusters_coords = clusters_coords - torch.index_select(clusters_coords_mean, 0, cid)

    t = timer.start("sec_minmax")
    clusters_coords_min = ops.sec_min(clusters_coords, clusters_offset)
    clusters_coords_max = ops.sec_max(clusters_coords, clusters_offset)
    timer.stop(t)
    if trace is not None:
        trace["sec_minmax"] = (clusters_coords.clone(), clusters_offset, clusters_coords_min, clusters_coords_max)

    clusters_size = clusters_coords_max - clusters_coords_min
    clusters_center = (clusters_coords_max + clusters_coords_min) / 2 + clusters_coords_mean

    clusters_scale = 1 / ((clusters_coords_max - clusters_coords_min) / fullscale).max(1)[0] - 0.01
    clusters_scale = torch.clamp(clusters_scale, min=None, max=scale)
    min_xyz = clusters_coords_min * clusters_scale.unsqueeze(-1)
    max_xyz = clusters_coords_max * clusters_scale.unsqueeze(-1)
    clusters_scale = torch.index_select(clusters_scale, 0, cid)
    clusters_coords = clusters_coords * clusters_scale.unsqueeze(-1)
    rng_ = max_xyz - min_xyz
    offset = -min_xyz + torch.clamp(fullscale - rng_ - 0.001, min=0) * rand6[:3] \
        + torch.clamp(fullscale - rng_ + 0.001, max=0) * rand6[3:]
    clusters_coords = clusters_coords + torch.index_select(offset, 0, cid)
    clusters_coords = clusters_coords.long()
    clusters_coords = torch.cat([cid.view(-1, 1), clusters_coords], 1).contiguous()       # (sumNPoint, 1 + 3)

    return _clusters_voxelization_tail(ops, clusters_coords, clusters_feats, clusters_offset.numel() - 1, mode, timer,
                                       trace, (clusters_center, clusters_size))


_pool = None
_streams = {}


class _Fork:
    """Independent parts of one step on side streams: ``spawn(fn)`` runs fn in a worker thread on a side stream of its own,
    ``here(fn)`` in this thread on another; both start after the work the current stream holds at that moment, and
    ``join()`` makes the current stream wait for all of them.  Each fn returns a tuple of tensors."""

    def __init__(self, dev):
        global _pool
        if _pool is None:
            import os
            import sys
            from concurrent.futures import ThreadPoolExecutor
            _pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="pg-chain")
            # three threads issue short launches: hand the interpreter lock over more often than every 5 ms (the default
            # switch interval is as long as the whole step)
            sys.setswitchinterval(float(os.environ.get("PG_CHAIN_SWITCH_S", "0.0002")))
        if dev not in _streams:
            _streams[dev] = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
        self.dev, self.main, self.side, self.used, self.jobs = dev, torch.cuda.current_stream(), _streams[dev], 0, []

    def _in_stream(self, fn, st):
        torch.cuda.set_device(self.dev)
        with torch.cuda.stream(st):
            return fn()

    def _next(self):
        st = self.side[self.used]
        self.used += 1
        st.wait_stream(self.main)
        return st

    def spawn(self, fn):
        st = self._next()
        self.jobs.append((st, _pool.submit(self._in_stream, fn, st)))
        return len(self.jobs) - 1

    def here(self, fn):
        st = self._next()
        self.jobs.append((st, self._in_stream(fn, st)))
        return len(self.jobs) - 1

    def join(self):
        res = []
        for st, r in self.jobs:
            r = r if isinstance(r, tuple) else r.result()
            self.main.wait_stream(st)
            for t in r:
                if torch.is_tensor(t):
                    t.record_stream(self.main)
            res.append(r)
        return res


def proposal_chain(ops, batch, rand6=None, timer=None, trace=None, fused_glue=False, overlap=True, fused_cluster=False):
    """One pass of the hot path over one collated batch (all tensors on one CUDA device).

    batch: locs fp32 [N,3], locs_scaled int64 [N,4], feats fp32 [N,134], pt_feats fp32 [N,16],
    semantic_preds int64 [N], pt_offsets fp32 [N,3], instance_ids int64 [N], instance_pointnum int32
    [nInst], n_scenes.  Returns a dict of the tensors the rest of the detector consumes.
    ``fused_cluster``: each ballquery_batch_p + bfs_cluster pair as ONE op (pointgroup_ops.ballquery_bfs_cluster: the
    neighbour lists are never materialised where the clustering does not read them) -- a caller edit, like ``fused_glue``.
    ``overlap`` (CUDA only; ignored while a timer or a trace is attached): the input cloud's voxelisation and the two
    clusterings, which share inputs and nothing else, are issued from three host threads on three streams -- scheduling
    only, the op calls and their results are the same."""
    timer = timer or _NoTimer()
    dev = batch["locs"].device
    B = int(batch["n_scenes"])
    if rand6 is None:
        rand6 = torch.full((6,), 0.5, device=dev)
    out = {}

    # The three independent parts of the step -- the input cloud's voxelisation and the two clusterings -- share their
    # inputs and nothing else, and each needs the host for its exact output sizes.  Issued from three host threads on three
    # streams, one's kernels fill the others' synchronisation bubbles and launch-bound stretches.  Same calls, same
    # results -- scheduling only (bench.py reports the one-stream schedule beside it).
    fork = _Fork(dev) if (overlap and dev.type == "cuda" and timer.enabled is False and trace is None) else None

    # ---- collate-side voxelisation of the input cloud (pipeline.py:992, pointgroup.py:472)
    def voxelize_scene():
        t = timer.start("voxelization_idx(scene)")
        voxel_locs, p2v_map, v2p_map = ops.voxelization_idx(batch["locs_scaled"], B, scenes.SCORE_MODE)
        timer.stop(t)
        if trace is not None:
            trace["voxelization_idx(scene)"] = (batch["locs_scaled"], B, voxel_locs, p2v_map, v2p_map)
        t = timer.start("voxelization(scene)")
        voxel_feats = ops.voxelization(batch["feats"], v2p_map, scenes.SCORE_MODE)
        timer.stop(t)
        if trace is not None:
            trace["voxelization(scene)"] = (batch["feats"], v2p_map, voxel_feats)
        return voxel_locs, voxel_feats, p2v_map, v2p_map

    def keep_scene(r):
        out["voxel_locs"], out["voxel_feats"], out["p2v_map"] = r[0], r[1], r[2]
        out["v2p_map_numel"] = r[3].numel()

    if fork is not None:
        fork.spawn(voxelize_scene)
    else:
        keep_scene(voxelize_scene())

    # ---- clustering on the predicted-object points (pointgroup.py:284-316)
    semantic_preds = batch["semantic_preds"]
    batch_idxs = batch["locs_scaled"][:, 0].int()
    object_idxs = torch.nonzero(semantic_preds > 0, as_tuple=False).view(-1)
    batch_idxs_ = batch_idxs[object_idxs].contiguous()
    batch_offsets_ = get_batch_offsets(batch_idxs_, B)
    coords_ = batch["locs"][object_idxs].contiguous()
    pt_offsets_ = batch["pt_offsets"][object_idxs]
    sem_ = semantic_preds[object_idxs].int().contiguous()

    shifted = (coords_ + pt_offsets_).contiguous()

    fused = fused_cluster and trace is None and getattr(ops, "ballquery_bfs_cluster", None) is not None

    def cluster_fused(pts, mean_active, key):
        t = timer.start("ballquery_bfs_cluster(%s)" % key)
        pidx, poff, n_active = ops.ballquery_bfs_cluster(pts, batch_idxs_, batch_offsets_, scenes.CLUSTER_RADIUS, mean_active,
                                                         sem_, scenes.CLUSTER_NPOINT_THRE)
        timer.stop(t)
        out["nActive_" + key] = n_active
        pidx[:, 1] = object_idxs[pidx[:, 1].long()].int()
        return pidx, poff

    def cluster_shift():
        if fused:
            return cluster_fused(shifted, scenes.CLUSTER_SHIFT_MEANACTIVE, "shift")
        t = timer.start("ballquery(shift)")
        idx_shift, start_len_shift = ops.ballquery_batch_p(shifted, batch_idxs_, batch_offsets_, scenes.CLUSTER_RADIUS,
                                                           scenes.CLUSTER_SHIFT_MEANACTIVE)
        timer.stop(t)
        t = timer.start("bfs_cluster(shift)")
        pidx, poff = ops.bfs_cluster(sem_, idx_shift, start_len_shift, scenes.CLUSTER_NPOINT_THRE)
        timer.stop(t)
        if trace is not None:
            trace["ballquery(shift)"] = (shifted, batch_idxs_, batch_offsets_, idx_shift, start_len_shift)
            trace["bfs_cluster(shift)"] = (sem_, idx_shift, start_len_shift, pidx.clone(), poff.clone())
        out["nActive_shift"] = idx_shift.numel()
        pidx[:, 1] = object_idxs[pidx[:, 1].long()].int()
        return pidx, poff

    def cluster_raw():
        if fused:
            return cluster_fused(coords_, scenes.CLUSTER_MEANACTIVE, "raw")
        t = timer.start("ballquery(raw)")
        idx, start_len = ops.ballquery_batch_p(coords_, batch_idxs_, batch_offsets_, scenes.CLUSTER_RADIUS,
                                               scenes.CLUSTER_MEANACTIVE)
        timer.stop(t)
        t = timer.start("bfs_cluster(raw)")
        pidx, poff = ops.bfs_cluster(sem_, idx, start_len, scenes.CLUSTER_NPOINT_THRE)
        timer.stop(t)
        if trace is not None:
            trace["ballquery(raw)"] = (coords_, batch_idxs_, batch_offsets_, idx, start_len)
            trace["bfs_cluster(raw)"] = (sem_, idx, start_len, pidx.clone(), poff.clone())
        out["nActive_raw"] = idx.numel()
        pidx[:, 1] = object_idxs[pidx[:, 1].long()].int()
        return pidx, poff

    if fork is not None:                  # model/pointgroup.py:296-298 and :304-306
        fork.spawn(cluster_raw)
        fork.here(cluster_shift)
        scene_r, (proposals_idx, proposals_offset), (proposals_idx_shift, proposals_offset_shift) = fork.join()
        keep_scene(scene_r)
    else:
        proposals_idx_shift, proposals_offset_shift = cluster_shift()
        proposals_idx, proposals_offset = cluster_raw()

    proposals_idx_shift[:, 0] += (proposals_offset.size(0) - 1)
    proposals_offset_shift = proposals_offset_shift + proposals_offset[-1]
    proposals_idx = torch.cat((proposals_idx, proposals_idx_shift), dim=0).contiguous()
    proposals_offset = torch.cat((proposals_offset, proposals_offset_shift[1:])).contiguous()
    out["proposals_idx"], out["proposals_offset"] = proposals_idx, proposals_offset
    out["n_object_points"] = object_idxs.numel()

    # ---- proposal re-voxelisation (pointgroup.py:326 -> :125-178)
    (prop_voxel_feats, prop_voxel_coords, prop_p2v_map, _v2p,
     (proposals_center, proposals_size)) = clusters_voxelization(
        ops, proposals_idx, proposals_offset, batch["pt_feats"], batch["locs"], scenes.SCORE_FULLSCALE,
        scenes.SCORE_SCALE, scenes.SCORE_MODE, rand6, timer, trace, fused_glue)
    out["proposals_center"], out["proposals_size"] = proposals_center, proposals_size
    out["proposals_voxel_feats"], out["proposals_voxel_coords"] = prop_voxel_feats, prop_voxel_coords
    out["proposals_v2p_map_numel"] = _v2p.numel()

    # ---- score features per point and proposal pooling (pointgroup.py:332-334; the score U-Net is
    #      out of scope, its per-voxel output is stood in for by its input)
    pt_score_feats = _gather_rows(ops, prop_voxel_feats, prop_p2v_map)
    t = timer.start("roipool")
    proposals_score_feats = ops.roipool(pt_score_feats, proposals_offset)
    timer.stop(t)
    if trace is not None:
        trace["roipool"] = (pt_score_feats, proposals_offset, proposals_score_feats)
    out["proposals_score_feats"] = proposals_score_feats

    # ---- proposal / ground-truth IoU (pointgroup.py:445)
    t = timer.start("get_iou")
    ious = ops.get_iou(proposals_idx[:, 1].contiguous(), proposals_offset, batch["instance_ids"],
                       batch["instance_pointnum"])
    timer.stop(t)
    if trace is not None:
        trace["get_iou"] = (proposals_idx[:, 1].contiguous(), proposals_offset, batch["instance_ids"],
                            batch["instance_pointnum"], ious)
    out["ious"] = ious
    return out


def batch_to_device(np_batch, device, pt_feat_seed=0, pin=False):
    """numpy batch from scenes.make_batch -> the tensor dict proposal_chain expects."""
    import numpy as np
    g = torch.Generator().manual_seed(1234 + pt_feat_seed)
    n = np_batch["locs"].shape[0]
    host = {
        "locs": torch.from_numpy(np_batch["locs"]),
        "locs_scaled": torch.from_numpy(np_batch["locs_scaled"]),
        "feats": torch.from_numpy(np_batch["feats"]) if "feats" in np_batch
        else torch.randn((n, scenes.IN_CHANNELS), generator=g),
        "pt_feats": torch.randn((n, scenes.M_CHANNELS), generator=g),
        "semantic_preds": torch.from_numpy(np_batch["semantic_preds"]),
        "pt_offsets": torch.from_numpy(np_batch["pt_offsets"]),
        "instance_ids": torch.from_numpy(np_batch["instance_ids"]),
        "instance_pointnum": torch.from_numpy(np.ascontiguousarray(np_batch["instance_pointnum"], dtype=np.int32)),
    }
    if pin:
        host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    if device is None:
        host["n_scenes"] = np_batch["n_scenes"]
        return host
    dev = {k: v.to(device, non_blocking=pin).contiguous() for k, v in host.items()}
    dev["n_scenes"] = np_batch["n_scenes"]
    return dev


# ---- convert_stack_to_batch (model/pointgroup.py:223-263) without the per-scene Python loop ---------------------------------
_BOX_SIGNS = ((1, 1, 1), (1, -1, 1), (-1, -1, 1), (-1, 1, 1), (1, 1, -1), (1, -1, -1), (-1, -1, -1), (-1, 1, -1))   # lib/utils/bbox.py:67-69


def box_corners(center, size, heading=None):
    """get_3d_box_batch (lib/utils/bbox.py:54-74) on the device, bit for bit: half extents in fp32, corners rotated and
    translated in fp64 (numpy's promotion), the caller rounds to fp32.  The detector only ever passes heading 0
    (proposal_crop_bbox[:, 6] is never written, model/pointgroup.py:355-361), where the rotation is the identity."""
    signs = torch.tensor(_BOX_SIGNS, dtype=torch.float64, device=center.device)            # [8, 3]
    half = (size.float() / 2).double()                                                      # l/2, w/2, h/2 as fp32 values
    c = signs[None] * half[:, None, :]                                                      # [n, 8, 3]
    if heading is not None and bool((heading != 0).any()):
        t = heading.double()
        cs, sn = torch.cos(t)[:, None], torch.sin(t)[:, None]
        x, y, z = c[..., 0], c[..., 1], c[..., 2]
        c = torch.stack([x * cs + z * sn, y, -x * sn + z * cs], -1)                         # corners @ roty(t)^T
    return c + center.double()[:, None, :]


def convert_stack_to_batch(proposals_batchId, proposal_feats, proposal_crop_bbox, proposal_objectness_scores, batch_size,
                           max_num_proposal, perms=None):
    """The six batched tensors of PointGroup.convert_stack_to_batch with IDENTICAL values, in a handful of device ops:
    per scene the first ``max_num_proposal`` proposals (in stack order), zero padding, then the per-scene shuffle.
    ``perms``: int64 [batch_size, max_num_proposal]; ``None`` draws ``torch.randperm(max_num_proposal)`` once per scene
    from the CPU generator, exactly the draws (and their order) of model/pointgroup.py:250."""
    dev, B, P = proposal_feats.device, int(batch_size), int(max_num_proposal)
    n = proposal_feats.size(0)
    if perms is None:
        perms = torch.stack([torch.randperm(P) for _ in range(B)]) if B > 0 else torch.zeros((0, P), dtype=torch.long)
    perms = perms.to(dev)
    out = {
        "proposal_feats_batched": torch.zeros((B, P, proposal_feats.size(1)), dtype=proposal_feats.dtype, device=dev),
        "proposal_bbox_batched": torch.zeros((B, P, 8, 3), dtype=proposal_feats.dtype, device=dev),
        "proposal_center_batched": torch.zeros((B, P, 3), dtype=proposal_feats.dtype, device=dev),
        "proposal_sem_cls_batched": torch.zeros((B, P), dtype=proposal_feats.dtype, device=dev),
        "proposal_scores_batched": torch.zeros((B, P), dtype=proposal_feats.dtype, device=dev),
        "proposal_batch_mask": torch.zeros((B, P), dtype=proposal_feats.dtype, device=dev),
    }
    if n > 0 and B > 0:
        bid = proposals_batchId.to(dev).long()
        order = torch.argsort(bid, stable=True)                                      # scene by scene, stack order inside
        sb = bid[order]
        start = torch.searchsorted(sb, torch.arange(B, device=dev))
        slot = torch.arange(n, device=dev) - start[sb]
        keep = (slot < P) & (sb >= 0) & (sb < B)
        src, b, s = order[keep], sb[keep], slot[keep]
        crop = proposal_crop_bbox
        corners = box_corners(crop[:, :3], crop[:, 3:6], crop[:, 6]).to(proposal_feats.dtype)
        out["proposal_feats_batched"][b, s] = proposal_feats[src]
        out["proposal_bbox_batched"][b, s] = corners[src]
        out["proposal_center_batched"][b, s] = crop[src, :3]
        out["proposal_sem_cls_batched"][b, s] = crop[src, 7]
        out["proposal_scores_batched"][b, s] = proposal_objectness_scores[src]
        out["proposal_batch_mask"][b, s] = 1
    rows = torch.arange(B, device=dev)[:, None]
    return {k: v[rows, perms] for k, v in out.items()}


def proposals_npoint(proposals_offset):
    """model/pointgroup.py:341-344 -- a Python loop of `(proposals_idx[:, 0] == i).sum()` over every proposal, O(nProposal
    x sumNPoint) on the CPU (0.8 s of the reference's forward at 8 x 150k points) -- is the difference of the offsets."""
    off = proposals_offset
    return (off[1:] - off[:-1]).float()
