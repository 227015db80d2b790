"""Drop-in stand-in for the reference's native module ``PG_OP``
(lib/pointgroup_ops/src/pointgroup_ops_api.cpp:6-24): the same 13 functions, the same positional
signatures and output conventions, implemented by the sm_100a kernels behind the C ABI of
include/pg_b200.h.  ``import d3net_b200.PG_OP as PG_OP`` (or ``sys.modules["PG_OP"] = ...``) lets the
reference's own functions/pointgroup_ops.py run unchanged; d3net_b200.pointgroup_ops is the
leaner wrapper that skips the zero fills and retry loop this library does not need.

Ownership follows the reference: fixed-size outputs are caller-allocated tensors written in place;
variable-size outputs (voxelize_idx, bfs_cluster) arrive as empty tensors and are ``resize_``d here
(voxelize.cpp:22-26, bfs_cluster.cpp:103-106).  Differences, all on the safe side: arguments are
validated (the reference reads raw ``data_ptr``s unchecked) and failures raise instead of calling
``exit`` (bfs_cluster.cu:82-86).  There is no CPU path: the two ops the reference runs on the CPU
(voxelize_idx, bfs_cluster) accept CPU tensors by staging them through the current CUDA device.
"""
import ctypes

import torch

from . import _native
from ._native import check

__all__ = ["voxelize_idx", "voxelize_fp", "voxelize_bp", "point_recover_fp", "point_recover_bp",
           "ballquery_batch_p", "bfs_cluster", "roipool_fp", "roipool_bp", "get_iou",
           "sec_mean", "sec_min", "sec_max"]


def _L():
    return _native.lib()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _need(t, name, dtype, cuda=True):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if cuda and not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU path)" % name)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _compute_device(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("d3net_b200: no CUDA device -- these ops have no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


# -------------------------------------------------------------------------------------------------
# implementations on device tensors (shared with d3net_b200.pointgroup_ops)
# -------------------------------------------------------------------------------------------------
def voxelize_idx_impl(coords, mode):
    """coords int64 [N,4] CUDA -> (output_coords int64 [M,4], input_map int32 [N], output_map int32 [M,W])."""
    _need(coords, "coords", torch.int64)
    if coords.dim() != 2 or coords.size(1) != 4:
        # the reference's 3-column branch reads coords with stride 4 afterwards (voxelize.cpp:43)
        raise ValueError("coords must be [N, 4] (batch, x, y, z); got %s" % (tuple(coords.shape),))
    N = coords.size(0)
    dev = coords.device
    with torch.cuda.device(dev):
        L = _L()
        nws = L.pg_voxelize_idx_workspace_bytes(N)
        ws = _ws(nws, dev)
        input_map = torch.empty(N, dtype=torch.int32, device=dev)
        sizes = (ctypes.c_int32 * 2)()
        check(L.pg_voxelize_idx_map(_p(coords), N, int(mode), _p(input_map), _p(ws), nws, sizes, _stream()),
              "voxelize_idx(map)")
        M, maxActive = int(sizes[0]), int(sizes[1])
        output_coords = torch.empty((M, 4), dtype=torch.int64, device=dev)
        output_map = torch.empty((M, maxActive + 1), dtype=torch.int32, device=dev)
        check(L.pg_voxelize_idx_fill(_p(coords), _p(input_map), N, M, maxActive, int(mode), _p(ws), nws,
                                     _p(output_coords), _p(output_map), _stream()), "voxelize_idx(fill)")
    return output_coords, input_map, output_map


# Hit masks cost 1 bit per tested (query, candidate) pair -- about 25 words per point at 330 neighbours per
# point.  The buffer is sized from n alone (no round trip to learn the exact need); a batch that needs more
# runs without masks (the fill phase then evaluates the predicates again).
BALLQUERY_MASK_WORDS_PER_POINT = 512
BALLQUERY_MASK_BUDGET_BYTES = 4 << 30


def ballquery_count_impl(xyz, batch_idxs, batch_offsets, radius, use_masks=True, mask_words=None):
    """Phases 1+2: returns (start_len int32 [n,2], total, state) where state carries the workspace and the
    hit-mask buffer (None when it was not used) to ballquery_fill_impl."""
    _need(xyz, "coords", torch.float32)
    _need(batch_idxs, "batch_idxs", torch.int32)
    _need(batch_offsets, "batch_offsets", torch.int32)
    n = xyz.size(0)
    if xyz.dim() != 2 or xyz.size(1) != 3 or batch_idxs.numel() != n:
        raise ValueError("coords must be [n,3] and batch_idxs [n]")
    dev = xyz.device
    with torch.cuda.device(dev):
        L = _L()
        nws = L.pg_ballquery_workspace_bytes(n)
        ws = _ws(nws, dev)
        check(L.pg_ballquery_prepare(_p(xyz), _p(batch_idxs), _p(batch_offsets), n, batch_offsets.numel() - 1,
                                     float(radius), _p(ws), nws, _stream()), "ballquery_batch_p(prepare)")
        masks = None
        if use_masks and n > 0:
            if mask_words is None:
                mask_words = min(BALLQUERY_MASK_WORDS_PER_POINT * n + 1024, BALLQUERY_MASK_BUDGET_BYTES // 4)
            try:
                masks = torch.empty(int(mask_words), dtype=torch.int32, device=dev)
            except torch.cuda.OutOfMemoryError:
                masks = None          # the kernels run without: the fill phase then evaluates the predicates again
        start_len = torch.empty((n, 2), dtype=torch.int32, device=dev)
        total = ctypes.c_int64(0)
        used = ctypes.c_int(0)
        check(L.pg_ballquery_count(_p(xyz), n, float(radius), _p(start_len), _p(masks),
                                   masks.numel() if masks is not None else 0, _p(ws), nws, ctypes.byref(total),
                                   ctypes.byref(used), _stream()), "ballquery_batch_p(count)")
        if not used.value:
            masks = None
    return start_len, int(total.value), (ws, masks)


def ballquery_fill_impl(xyz, radius, start_len, idx, state):
    ws, masks = state
    with torch.cuda.device(xyz.device):
        check(_L().pg_ballquery_fill(_p(xyz), xyz.size(0), float(radius), _p(start_len), _p(masks), _p(idx),
                                     idx.numel(), _p(ws), ws.numel(), _stream()), "ballquery_batch_p(fill)")


BFS_AUTO, BFS_GENERIC, BFS_TRUSTED = 0, 1, 2      # include/pg_b200.h, `mode` of pg_bfs_cluster_count


def bfs_cluster_impl(semantic_label, ball_query_idxs, start_len, threshold, generic=0, trusted=False, grid_ws=None):
    """All CUDA int32 -> (cluster_idxs [S,2], cluster_offsets [nC+1], used_generic_path).
    ``trusted``: the lists are ballquery_batch_p output of this library, untouched -- the sweep skips the
    validation it needs for foreign neighbour lists (d3net_b200.pointgroup_ops decides this by provenance).
    ``grid_ws`` (with ``trusted``): the workspace tensor of the ball query that produced the lists, untouched --
    cells whose neighbourhood is already one component are not swept at all (pg_bfs_cluster_count_grid)."""
    _need(semantic_label, "semantic_label", torch.int32)
    _need(ball_query_idxs, "ball_query_idxs", torch.int32)
    _need(start_len, "start_len", torch.int32)
    N = start_len.size(0)
    if semantic_label.numel() < N:
        raise ValueError("semantic_label has fewer entries than start_len has rows")
    dev = semantic_label.device
    with torch.cuda.device(dev):
        L = _L()
        # whatever the buffer holds beyond the minimum becomes parking room for one-way edges (lists cut at 1000
        # entries): one 8-byte slot per 8 neighbours, at most 512 MB
        nws = L.pg_bfs_cluster_workspace_bytes(N) + 8 * min(ball_query_idxs.numel() // 8, 64 << 20)
        ws = _ws(nws, dev)
        sizes = (ctypes.c_int32 * 3)()
        if trusted and not generic and grid_ws is not None and N > 0:
            check(L.pg_bfs_cluster_count_grid(_p(semantic_label), _p(ball_query_idxs), _p(start_len), N,
                                              ball_query_idxs.numel(), int(threshold), _p(ws), nws, _p(grid_ws),
                                              grid_ws.numel(), sizes, _stream()), "bfs_cluster(count, grid)")
        else:
            check(L.pg_bfs_cluster_count(_p(semantic_label), _p(ball_query_idxs), _p(start_len), N,
                                         ball_query_idxs.numel(), int(threshold),
                                         BFS_GENERIC if generic else (BFS_TRUSTED if trusted else BFS_AUTO), _p(ws), nws,
                                         sizes, _stream()), "bfs_cluster(count)")
        nC, S = int(sizes[0]), int(sizes[1])
        cluster_idxs = torch.empty((S, 2), dtype=torch.int32, device=dev)
        cluster_offsets = torch.empty(nC + 1, dtype=torch.int32, device=dev)
        check(L.pg_bfs_cluster_fill(N, nC, S, _p(ws), nws, _p(cluster_idxs), _p(cluster_offsets), _stream()),
              "bfs_cluster(fill)")
    return cluster_idxs, cluster_offsets, bool(sizes[2])


def ballquery_bfs_cluster_impl(xyz, batch_idxs, batch_offsets, radius, semantic_label, threshold):
    """Fused ballquery_batch_p + bfs_cluster for callers that need the clusters, not the lists (model/pointgroup.py:296-297
    passes idx straight on and never looks at it again): the neighbour lists stay in the form the count phase leaves them
    in (hit masks + merged candidates per cell) and only the lists the clustering sweep actually reads are decoded.
    Returns (cluster_idxs, cluster_offsets, nActive); same clusters as the two ops called one after the other."""
    _need(semantic_label, "semantic_label", torch.int32)
    start_len, total, state = ballquery_count_impl(xyz, batch_idxs, batch_offsets, radius)
    ws_bq, masks = state
    n = xyz.size(0)
    dev = xyz.device
    lazy_ok = masks is not None and n > 0 and total >= 12 * n        # long lists only (the sweep's cell pass needs them)
    with torch.cuda.device(dev):
        L = _L()
        if lazy_ok:
            nws = L.pg_bfs_cluster_workspace_bytes(n) + 8 * min(total // 8, 64 << 20)
            ws = _ws(nws, dev)
            idx = torch.empty(total, dtype=torch.int32, device=dev)      # only the swept lists get written
            sizes = (ctypes.c_int32 * 3)()
            need = ctypes.c_int(0)
            check(L.pg_bfs_cluster_count_lazy(_p(semantic_label), _p(start_len), n, total, int(threshold), _p(ws), nws,
                                              _p(ws_bq), ws_bq.numel(), _p(masks), _p(idx), sizes, ctypes.byref(need),
                                              _stream()), "ballquery_bfs_cluster(count)")
            if not need.value:
                nC, S = int(sizes[0]), int(sizes[1])
                cluster_idxs = torch.empty((S, 2), dtype=torch.int32, device=dev)
                cluster_offsets = torch.empty(nC + 1, dtype=torch.int32, device=dev)
                check(L.pg_bfs_cluster_fill(n, nC, S, _p(ws), nws, _p(cluster_idxs), _p(cluster_offsets), _stream()),
                      "ballquery_bfs_cluster(fill)")
                return cluster_idxs, cluster_offsets, total
        else:
            idx = torch.empty(total, dtype=torch.int32, device=dev)
    # short lists, no mask buffer, or a parking-lot overflow: materialise everything, cluster as usual
    ballquery_fill_impl(xyz, radius, start_len, idx, state)
    ci, co, _ = bfs_cluster_impl(semantic_label, idx, start_len, threshold, trusted=True, grid_ws=ws_bq)
    return ci, co, total


def bfs_cluster_debug():
    """Diagnostics of this thread's last bfs_cluster count phase: [checksum != 0, bad lists, parked one-way
    edges, propagation sweeps, neighbour lists the edge sweep read]."""
    d = (ctypes.c_longlong * 5)()
    _L().pg_bfs_cluster_debug(d)
    return list(d)


def _vox(fn, src, dst, rules, M, maxActive, C, *extra):
    _need(src, "feats", torch.float32)
    _need(dst, "output", torch.float32)
    _need(rules, "map_rule", torch.int32)
    if rules.numel() < M * (maxActive + 1):
        raise ValueError("map_rule is smaller than nActive x (maxActive + 1)")
    with torch.cuda.device(src.device):
        check(fn(_p(src), _p(dst), _p(rules), int(M), int(maxActive), int(C), *extra, _stream()), fn.__name__)


def _seg(fn, inp, offsets, out, nProposal, C):
    _need(inp, "inp", torch.float32)
    _need(offsets, "offsets", torch.int32)
    _need(out, "out", torch.float32)
    if offsets.numel() < nProposal + 1 or out.numel() < nProposal * C:
        raise ValueError("offsets / out too small for nProposal, C")
    nRows = inp.numel() // C if C > 0 else 0
    with torch.cuda.device(inp.device):
        check(fn(_p(inp), _p(offsets), _p(out), nRows, int(nProposal), int(C), _stream()), fn.__name__)


# -------------------------------------------------------------------------------------------------
# the 13 PG_OP functions (src/pointgroup_ops_api.cpp:6-24)
# -------------------------------------------------------------------------------------------------
def voxelize_idx(coords, output_coords, input_map, output_map, batchSize, mode):
    """src/pointgroup_ops.cpp:13 / voxelize.cpp:11-31.  ``batchSize`` only pre-sizes the reference's
    per-batch hash maps (voxelize.cpp:64,92-94) and does not affect the result."""
    dev = _compute_device(coords)
    oc, im, om = voxelize_idx_impl(coords.to(dev), mode)
    output_coords.resize_(oc.shape).copy_(oc)
    input_map.resize_(im.shape).copy_(im)
    output_map.resize_(om.shape).copy_(om)


def voxelize_fp(feats, output_feats, output_map, mode, nActive, maxActive, nPlane):
    """src/pointgroup_ops.cpp:18 / voxelize.cpp:157-166.  Overwrites output_feats (the reference
    accumulates onto the zeros its wrapper put there -- same result)."""
    _vox(_L().pg_voxelize_fp, feats, output_feats, output_map, nActive, maxActive, nPlane, int(mode == 4))


def voxelize_bp(d_output_feats, d_feats, output_map, mode, nActive, maxActive, nPlane):
    """src/pointgroup_ops.cpp:25 / voxelize.cpp:169-178.  Accumulates into d_feats."""
    _vox(_L().pg_voxelize_bp, d_output_feats, d_feats, output_map, nActive, maxActive, nPlane, int(mode == 4))


def point_recover_fp(feats, output_feats, idx_map, nActive, maxActive, nPlane):
    """src/pointgroup_ops.cpp:30 / voxelize.cpp:182-190.  Accumulates into output_feats."""
    _vox(_L().pg_point_recover_fp, feats, output_feats, idx_map, nActive, maxActive, nPlane)


def point_recover_bp(d_output_feats, d_feats, idx_map, nActive, maxActive, nPlane):
    """src/pointgroup_ops.cpp:35 / voxelize.cpp:193-202.  Overwrites d_feats."""
    _vox(_L().pg_point_recover_bp, d_output_feats, d_feats, idx_map, nActive, maxActive, nPlane)


def ballquery_batch_p(xyz, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius):
    """bfs_cluster.h:15 / bfs_cluster.cpp:15-25.  Returns the total neighbour count; ``idx`` is filled
    only when it fits ``n * meanActive`` (the reference clips silently, bfs_cluster.cu:51-55, and its
    wrapper retries with a larger buffer, functions/pointgroup_ops.py:135-142)."""
    _need(idx, "idx", torch.int32)
    _need(start_len, "start_len", torch.int32)
    if xyz.size(0) != n or start_len.numel() < 2 * n:
        raise ValueError("n does not match coords / start_len")
    sl, total, ws = ballquery_count_impl(xyz, batch_idxs, batch_offsets, radius)
    start_len.view(-1)[:2 * n].copy_(sl.view(-1))
    if total <= n * meanActive and total <= idx.numel():
        ballquery_fill_impl(xyz, radius, sl, idx, ws)
    return total


def bfs_cluster(semantic_label, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, threshold):
    """bfs_cluster.h:18 / bfs_cluster.cpp:93-112.  Inputs may live on the CPU (as the reference's caller
    passes them, model/pointgroup.py:297); outputs are resized on the device of the tensors passed in."""
    dev = _compute_device(semantic_label)
    if start_len.size(0) != N:
        raise ValueError("N does not match start_len")
    ci, co, _ = bfs_cluster_impl(semantic_label.to(dev), ball_query_idxs.to(dev), start_len.to(dev), threshold)
    cluster_idxs.resize_(ci.shape).copy_(ci)
    cluster_offsets.resize_(co.shape).copy_(co)


def roipool_fp(feats, proposals_offset, output_feats, output_maxidx, nProposal, C):
    """roipool.h:15 / roipool.cu:12-39."""
    _need(feats, "feats", torch.float32)
    _need(proposals_offset, "proposals_offset", torch.int32)
    _need(output_feats, "output_feats", torch.float32)
    _need(output_maxidx, "output_maxidx", torch.int32)
    if proposals_offset.numel() < nProposal + 1 or output_feats.numel() < nProposal * C:
        raise ValueError("proposals_offset / output_feats too small")
    with torch.cuda.device(feats.device):
        L = _L()
        nws = L.pg_roipool_workspace_bytes(nProposal, C)
        ws = _ws(nws, feats.device)
        nRows = feats.numel() // C if C > 0 else 0
        check(L.pg_roipool_fp(_p(feats), _p(proposals_offset), _p(output_feats), _p(output_maxidx), nRows,
                              int(nProposal), int(C), _p(ws), nws, _stream()), "roipool_fp")


def roipool_bp(d_feats, proposals_offset, output_maxidx, d_output_feats, nProposal, C):
    """roipool.h:21 / roipool.cu:42-57.  Accumulates into d_feats."""
    _need(d_feats, "d_feats", torch.float32)
    _need(output_maxidx, "output_maxidx", torch.int32)
    _need(d_output_feats, "d_output_feats", torch.float32)
    with torch.cuda.device(d_feats.device):
        check(_L().pg_roipool_bp(_p(d_feats), _p(proposals_offset), _p(output_maxidx), _p(d_output_feats),
                                 int(nProposal), int(C), _stream()), "roipool_bp")


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance,
            nProposal):
    """get_iou.h:16 / get_iou.cu:12-38."""
    _need(proposals_idx, "proposals_idx", torch.int32)
    _need(proposals_offset, "proposals_offset", torch.int32)
    _need(instance_labels, "instance_labels", torch.int64)
    _need(instance_pointnum, "instance_pointnum", torch.int32)
    _need(proposals_iou, "proposals_iou", torch.float32)
    if proposals_iou.numel() < nInstance * nProposal:
        raise ValueError("proposals_iou too small")
    with torch.cuda.device(proposals_idx.device):
        check(_L().pg_get_iou(_p(proposals_idx), _p(proposals_offset), _p(instance_labels), _p(instance_pointnum),
                              _p(proposals_iou), int(nInstance), int(nProposal), _stream()), "get_iou")


def sec_mean(inp, offsets, out, nProposal, C):
    """sec_mean.h:14 / sec_mean.cu:12-34."""
    _seg(_L().pg_sec_mean, inp, offsets, out, nProposal, C)


def sec_min(inp, offsets, out, nProposal, C):
    """sec_mean.h:17 / sec_mean.cu:38-60."""
    _seg(_L().pg_sec_min, inp, offsets, out, nProposal, C)


def sec_max(inp, offsets, out, nProposal, C):
    """sec_mean.h:20 / sec_mean.cu:64-86."""
    _seg(_L().pg_sec_max, inp, offsets, out, nProposal, C)


# -------------------------------------------------------------------------------------------------
# additions (not in the reference's PG_OP)
# -------------------------------------------------------------------------------------------------
def gather_rows(src, idx, out=None):
    """out[i] = src[idx[i]] for fp32 [N, C] rows and an int32 / int64 index vector, all CUDA."""
    _need(src, "src", torch.float32)
    if idx.dtype not in (torch.int32, torch.int64) or not idx.is_cuda or not idx.is_contiguous():
        raise TypeError("idx must be a contiguous CUDA int32 / int64 tensor")
    if src.dim() != 2:
        raise ValueError("src must be [N, C]")
    n, C = idx.numel(), src.size(1)
    if out is None:
        out = torch.empty((n, C), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(_L().pg_gather_rows(_p(src), _p(idx), int(idx.dtype == torch.int64), _p(out), n, C, _stream()),
              "gather_rows")
    return out


def cluster_coords(coords, cluster_idxs, cluster_offsets, fullscale, scale, rand6):
    """Fused glue of clusters_voxelization (model/pointgroup.py:125-167): -> (clusters_coords int64 [S,4],
    center fp32 [nC,3], size fp32 [nC,3]); bit-identical to the torch op sequence it replaces."""
    _need(coords, "coords", torch.float32)
    _need(cluster_idxs, "cluster_idxs", torch.int32)
    _need(cluster_offsets, "cluster_offsets", torch.int32)
    _need(rand6, "rand6", torch.float32)
    if coords.dim() != 2 or coords.size(1) != 3 or rand6.numel() != 6:
        raise ValueError("coords must be [N,3] and rand6 hold 6 values")
    S, nC = cluster_idxs.size(0), cluster_offsets.numel() - 1
    dev = coords.device
    out = torch.empty((S, 4), dtype=torch.int64, device=dev)
    center = torch.empty((nC, 3), dtype=torch.float32, device=dev)
    size = torch.empty((nC, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L = _L()
        nws = L.pg_cluster_coords_workspace_bytes(nC)
        ws = _ws(nws, dev)
        check(L.pg_cluster_coords(_p(coords), _p(cluster_idxs), _p(cluster_offsets), S, nC, int(fullscale), float(scale),
                                  _p(rand6), _p(ws), nws, _p(out), _p(center), _p(size), _stream()), "cluster_coords")
    return out, center, size


def cross_iou(proposals_idx, nProposal, N, want_npoint=False):
    """Sparse replacement of the dense-mask matmul of PointGroup.test (model/pointgroup.py:577-590):
    proposals_idx int32 [S, 2] CUDA rows (proposal, point) -> cross_ious fp32 [nProposal, nProposal]
    (and, on request, the distinct point count per proposal, proposals_mask.sum(1) of :582)."""
    _need(proposals_idx, "proposals_idx", torch.int32)
    if proposals_idx.dim() != 2 or proposals_idx.size(1) != 2:
        raise ValueError("proposals_idx must be [sumNPoint, 2]")
    S, nP, N = proposals_idx.size(0), int(nProposal), int(N)
    dev = proposals_idx.device
    out = torch.empty((nP, nP), dtype=torch.float32, device=dev)
    npoint = torch.empty(nP, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L = _L()
        nws = L.pg_cross_iou_workspace_bytes(S, nP, N)
        ws = _ws(nws, dev)
        check(L.pg_cross_iou(_p(proposals_idx), S, nP, N, _p(ws), nws, _p(out), _p(npoint), _stream()), "cross_iou")
    return (out, npoint) if want_npoint else out


def nms_instances(cross_ious, scores, threshold):
    """get_nms_instances (lib/utils/eval.py:75-97) on the device: -> int32 [nPick] kept proposals, best first."""
    _need(cross_ious, "cross_ious", torch.float32)
    _need(scores, "scores", torch.float32)
    n = scores.numel()
    if cross_ious.dim() != 2 or cross_ious.size(0) != n or cross_ious.size(1) != n:
        raise ValueError("cross_ious must be [n, n] for n scores")
    dev = scores.device
    pick = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = ctypes.c_int32(0)
    with torch.cuda.device(dev):
        L = _L()
        nws = L.pg_nms_instances_workspace_bytes(n)
        ws = _ws(nws, dev)
        check(L.pg_nms_instances(_p(cross_ious), _p(scores), n, float(threshold), _p(ws), nws, _p(pick),
                                 ctypes.byref(cnt), _stream()), "nms_instances")
    return pick[:cnt.value]


def pick_masks(proposals_idx, proposals_offset, pick, N):
    """clusters_mask = proposals_mask[pick_idxs] (model/pointgroup.py:593) without the dense [nProposal, N] mask:
    int32 [nPick, N], row k = membership of proposal pick[k] (ids in the numbering of proposals_offset)."""
    _need(proposals_idx, "proposals_idx", torch.int32)
    _need(proposals_offset, "proposals_offset", torch.int32)
    _need(pick, "pick", torch.int32)
    dev = proposals_idx.device
    out = torch.empty((pick.numel(), int(N)), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _ws(8, dev)
        check(_L().pg_pick_masks(_p(proposals_idx), _p(proposals_offset), proposals_offset.numel() - 1, _p(pick),
                                 pick.numel(), int(N), _p(ws), _p(out), _stream()), "pick_masks")
    return out
