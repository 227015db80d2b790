"""The reference's Python operator API (lib/pointgroup_ops/functions/pointgroup_ops.py) on the
B200 kernels: the same ten callables -- voxelization_idx, voxelization, point_recover,
ballquery_batch_p, bfs_cluster, roipool, get_iou, sec_mean, sec_min, sec_max -- with the same
argument meaning, return conventions, autograd behaviour and assertion behaviour, so
model/pointgroup.py, lib/dataset/pipeline.py and lib/solver/pointgroup.py can import this module in
place of the original.

What changed underneath:
  * outputs that the kernels fully overwrite are allocated uninitialised instead of zero-filled;
  * ball query is count -> exact allocation -> fill, so the reference's grow-and-retry loop
    (functions/pointgroup_ops.py:135-142) is gone and ``meanActive`` is only a hint;
  * voxelization_idx and bfs_cluster -- CPU-only in the reference -- run on the GPU.  CPU inputs (what
    the reference's callers pass, model/pointgroup.py:169,297) are staged through the current CUDA
    device and the results come back on the input's device; CUDA inputs stay on the device.
Two differences ARE visible, both inside what the reference leaves undefined or what its caller never reads:
  * ``ballquery_batch_p``: every point's list is the reference's (ascending, first 1000), but the SEGMENTS sit in
    ``idx`` in grid-cell order; the reference places them by ``atomicAdd`` (bfs_cluster.cu:47), differently on
    every run.  Scene membership is taken from ``batch_idxs``; ``batch_offsets`` (the reference's scan range,
    bfs_cluster.cu:27-32) is only consulted for the scene count -- the two agree for every input the caller builds
    (model/pointgroup.py:291-292).
  * ``bfs_cluster``: clusters, their order and their members are the reference's; inside a cluster the members are
    listed in ASCENDING point order, the reference lists them in BFS visit order (bfs_cluster.cpp:44-51).  Everything
    downstream that is integer-valued is unaffected (voxel sets, argmax rows, IoUs); the row order of
    ``clusters_coords`` changes, hence the first-appearance numbering of the cluster voxels (a consistent relabelling
    of voxel_coords / p2v_map / v2p_map rows) and the fp32 summation order inside ``sec_mean`` / the voxel means --
    within the 1e-6 relative contract, not bit-identical to a reference run.
There is no CPU fallback: without the CUDA library or a GPU these functions raise.
"""
import torch
from torch.autograd import Function

from . import PG_OP


# bench.py sets this to a chain.SectionTimer to get CUDA-event timings of the phases of the two-phase ops
_section_timer = None


def _sub(name):
    return _section_timer.start_sub(name) if _section_timer is not None else None


def _end(tok):
    if tok is not None:
        tok.record()


def _cuda_of(t):
    return t if t.is_cuda else t.to(PG_OP._compute_device(t))


class Voxelization_Idx(Function):
    @staticmethod
    def forward(ctx, coords, batchsize, mode=4):
        '''
        :param coords:  long (N, dimension + 1), dimension = 3, column 0 = batch index
        :param batchsize
        :param mode: int 4=mean
        :return: output_coords:  long (M, dimension + 1) (M <= N)
        :return: input_map: int (N,)
        :return: output_map: int (M, (maxActive + 1))
        (functions/pointgroup_ops.py:11-39; all three on coords.device)
        '''
        assert coords.is_contiguous()
        oc, im, om = PG_OP.voxelize_idx_impl(_cuda_of(coords), mode)
        if not coords.is_cuda:
            oc, im, om = oc.cpu(), im.cpu(), om.cpu()
        ctx.mark_non_differentiable(oc, im, om)
        return oc, im, om

    @staticmethod
    def backward(ctx, a=None, b=None, c=None):
        return None, None, None


voxelization_idx = Voxelization_Idx.apply


class Voxelization(Function):
    @staticmethod
    def forward(ctx, feats, map_rule, mode=4):
        '''
        :param map_rule: cuda int (M, (maxActive + 1))
        :param feats: cuda float (N, C)
        :return: output_feats: cuda float (M, C)
        (functions/pointgroup_ops.py:42-75)
        '''
        assert map_rule.is_contiguous()
        assert feats.is_contiguous()
        N, C = feats.size()
        M = map_rule.size(0)
        maxActive = map_rule.size(1) - 1
        output_feats = torch.empty((M, C), dtype=torch.float32, device=feats.device)
        ctx.for_backwards = (map_rule, mode, maxActive, N)
        PG_OP.voxelize_fp(feats, output_feats, map_rule, mode, M, maxActive, C)
        return output_feats

    @staticmethod
    def backward(ctx, d_output_feats):
        map_rule, mode, maxActive, N = ctx.for_backwards
        M, C = d_output_feats.size()
        d_feats = torch.zeros((N, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.voxelize_bp(d_output_feats.contiguous(), d_feats, map_rule, mode, M, maxActive, C)
        return d_feats, None, None


voxelization = Voxelization.apply


class PointRecover(Function):
    @staticmethod
    def forward(ctx, feats, map_rule, nPoint):
        '''
        :param feats: cuda float M * C
        :param map_rule: cuda int M * (maxActive + 1)
        :param nPoint: int
        :return: output_feats: cuda float N * C
        (functions/pointgroup_ops.py:78-112)
        '''
        assert map_rule.is_contiguous()
        assert feats.is_contiguous()
        M, C = feats.size()
        maxActive = map_rule.size(1) - 1
        output_feats = torch.zeros((nPoint, C), dtype=torch.float32, device=feats.device)
        ctx.for_backwards = (map_rule, maxActive, M)
        PG_OP.point_recover_fp(feats, output_feats, map_rule, M, maxActive, C)
        return output_feats

    @staticmethod
    def backward(ctx, d_output_feats):
        map_rule, maxActive, M = ctx.for_backwards
        N, C = d_output_feats.size()
        d_feats = torch.empty((M, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.point_recover_bp(d_output_feats.contiguous(), d_feats, map_rule, M, maxActive, C)
        return d_feats, None, None


point_recover = PointRecover.apply


class BallQueryBatchP(Function):
    @staticmethod
    def forward(ctx, coords, batch_idxs, batch_offsets, radius, meanActive, ws_out=None):
        '''
        :param coords: (n, 3) float
        :param batch_idxs: (n) int
        :param batch_offsets: (B+1) int
        :param radius: float
        :param meanActive: int (the reference's initial buffer guess; unused here)
        :param ws_out: optional list; receives the call's grid workspace tensor (this module's own hand-over to
                       bfs_cluster -- a per-call object, so concurrent callers cannot see each other's workspace)
        :return: idx (nActive), int -- per point ascending, segments laid out in query (cell) order
        :return: start_len (n, 2), int
        (functions/pointgroup_ops.py:115-150)
        '''
        assert coords.is_contiguous() and coords.is_cuda
        assert batch_idxs.is_contiguous() and batch_idxs.is_cuda
        assert batch_offsets.is_contiguous() and batch_offsets.is_cuda
        t = _sub("count")
        start_len, nActive, ws = PG_OP.ballquery_count_impl(coords, batch_idxs, batch_offsets, radius)
        _end(t)
        idx = torch.empty(nActive, dtype=torch.int32, device=coords.device)
        t = _sub("fill")
        PG_OP.ballquery_fill_impl(coords, radius, start_len, idx, ws)
        _end(t)
        if ws_out is not None:
            ws_out.append(ws[0])
        ctx.mark_non_differentiable(idx, start_len)
        return idx, start_len

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None, None


# bfs_cluster on stamped lists also gets the ball query's uniform grid (its workspace tensor, kept alive by the
# stamp): cells that are already one component are not swept (pg_bfs_cluster_count_grid).  Results are identical.
USE_GRID_SWEEP = True


def _tag(idx, start_len):
    return (start_len.data_ptr(), start_len._version, idx.data_ptr(), idx._version, tuple(start_len.shape), idx.numel(),
            str(idx.device), str(start_len.device))


def _stamp(idx, start_len, grid_ws=None):
    """Provenance of a neighbour-list pair: the two tensors (storage, shape, device), as produced here and never
    written since (torch bumps ``_version`` on every in-place write).  bfs_cluster may then skip the validation it
    runs on foreign lists."""
    idx._pg_lists = _tag(idx, start_len)
    idx._pg_grid = grid_ws


def _stamped(idx, start_len):
    tag = getattr(idx, "_pg_lists", None)
    return (tag is not None and idx.is_cuda and start_len.is_cuda and idx.device == start_len.device
            and tag == _tag(idx, start_len))


def ballquery_batch_p(coords, batch_idxs, batch_offsets, radius, meanActive):
    ws = []
    idx, start_len = BallQueryBatchP.apply(coords, batch_idxs, batch_offsets, radius, meanActive, ws)
    _stamp(idx, start_len, ws[0] if ws else None)
    return idx, start_len


class BFSCluster(Function):
    @staticmethod
    def forward(ctx, semantic_label, ball_query_idxs, start_len, threshold):
        '''
        :param semantic_label: (N), int
        :param ball_query_idxs: (nActive), int
        :param start_len: (N, 2), int
        :return: cluster_idxs:  int (sumNPoint, 2), dim 0 for cluster_id, dim 1 for corresponding point idxs in N
        :return: cluster_offsets: int (nCluster + 1)
        (functions/pointgroup_ops.py:153-182; outputs on semantic_label.device like semantic_label.new())
        '''
        assert semantic_label.is_contiguous()
        assert ball_query_idxs.is_contiguous()
        assert start_len.is_contiguous()
        # compute where the lists already are (they are the big operand); otherwise where semantic_label is / the
        # current device.  Lists that have to be moved are copies: no provenance, no grid.
        if ball_query_idxs.is_cuda:
            dev = ball_query_idxs.device
        else:
            dev = PG_OP._compute_device(semantic_label)
        trusted = _stamped(ball_query_idxs, start_len) and ball_query_idxs.device == dev
        grid_ws = getattr(ball_query_idxs, "_pg_grid", None) if (trusted and USE_GRID_SWEEP) else None
        if grid_ws is not None and grid_ws.device != dev:
            grid_ws = None
        ci, co, _ = PG_OP.bfs_cluster_impl(semantic_label.to(dev), ball_query_idxs.to(dev), start_len.to(dev),
                                           threshold, trusted=trusted, grid_ws=grid_ws)
        if not semantic_label.is_cuda:
            ci, co = ci.cpu(), co.cpu()
        ctx.mark_non_differentiable(ci, co)
        return ci, co

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None


bfs_cluster = BFSCluster.apply


def ballquery_bfs_cluster(coords, batch_idxs, batch_offsets, radius, meanActive, semantic_label, threshold):
    """Not part of the reference's operator API: ``bfs_cluster(semantic_label, *ballquery_batch_p(coords, ...), threshold)``
    as ONE op for callers that hand the neighbour lists straight on (model/pointgroup.py:296-297 and :304-305).  The
    lists are clustered from the ball query's hit masks; only the ones the sweep has to read are ever decoded.  Returns
    (cluster_idxs, cluster_offsets, nActive) -- the same clusters, bit for bit."""
    assert coords.is_contiguous() and coords.is_cuda
    assert batch_idxs.is_contiguous() and batch_idxs.is_cuda
    assert batch_offsets.is_contiguous() and batch_offsets.is_cuda
    dev = coords.device
    ci, co, total = PG_OP.ballquery_bfs_cluster_impl(coords, batch_idxs, batch_offsets, radius,
                                                     semantic_label.to(dev).contiguous(), threshold)
    if not semantic_label.is_cuda:
        ci, co = ci.cpu(), co.cpu()
    return ci, co, total


class RoiPool(Function):
    @staticmethod
    def forward(ctx, feats, proposals_offset):
        '''
        :param feats: (sumNPoint, C) float
        :param proposals_offset: (nProposal + 1) int
        :return: output_feats (nProposal, C) float
        (functions/pointgroup_ops.py:185-221)
        '''
        nProposal = proposals_offset.size(0) - 1
        sumNPoint, C = feats.size()
        assert feats.is_contiguous()
        assert proposals_offset.is_contiguous()
        output_feats = torch.empty((nProposal, C), dtype=torch.float32, device=feats.device)
        output_maxidx = torch.empty((nProposal, C), dtype=torch.int32, device=feats.device)
        PG_OP.roipool_fp(feats, proposals_offset, output_feats, output_maxidx, nProposal, C)
        ctx.for_backwards = (output_maxidx, proposals_offset, sumNPoint)
        return output_feats

    @staticmethod
    def backward(ctx, d_output_feats):
        nProposal, C = d_output_feats.size()
        output_maxidx, proposals_offset, sumNPoint = ctx.for_backwards
        d_feats = torch.zeros((sumNPoint, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.roipool_bp(d_feats, proposals_offset, output_maxidx, d_output_feats.contiguous(), nProposal, C)
        return d_feats, None


roipool = RoiPool.apply


class GetIoU(Function):
    @staticmethod
    def forward(ctx, proposals_idx, proposals_offset, instance_labels, instance_pointnum):
        '''
        :param proposals_idx: (sumNPoint), int
        :param proposals_offset: (nProposal + 1), int
        :param instance_labels: (N), long, 0~total_nInst-1, -1
        :param instance_pointnum: (total_nInst), int
        :return: proposals_iou: (nProposal, total_nInst), float
        (functions/pointgroup_ops.py:224-253)
        '''
        nInstance = instance_pointnum.size(0)
        nProposal = proposals_offset.size(0) - 1
        assert proposals_idx.is_contiguous() and proposals_idx.is_cuda
        assert proposals_offset.is_contiguous() and proposals_offset.is_cuda
        assert instance_labels.is_contiguous() and instance_labels.is_cuda
        assert instance_pointnum.is_contiguous() and instance_pointnum.is_cuda
        proposals_iou = torch.empty((nProposal, nInstance), dtype=torch.float32, device=proposals_idx.device)
        PG_OP.get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou,
                      nInstance, nProposal)
        return proposals_iou

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


get_iou = GetIoU.apply


def _sec_forward(fn, inp, offsets):
    nProposal = offsets.size(0) - 1
    C = inp.size(1)
    assert inp.is_contiguous()
    assert offsets.is_contiguous()
    out = torch.empty((nProposal, C), dtype=torch.float32, device=inp.device)
    fn(inp, offsets, out, nProposal, C)
    return out


class SecMean(Function):
    @staticmethod
    def forward(ctx, inp, offsets):
        '''
        :param inp: (N, C) float
        :param offsets: (nProposal + 1) int
        :return: out (nProposal, C) float
        (functions/pointgroup_ops.py:256-281)
        '''
        return _sec_forward(PG_OP.sec_mean, inp, offsets)

    @staticmethod
    def backward(ctx, a=None):
        return None, None


sec_mean = SecMean.apply


class SecMin(Function):
    @staticmethod
    def forward(ctx, inp, offsets):
        '''(functions/pointgroup_ops.py:284-309)'''
        return _sec_forward(PG_OP.sec_min, inp, offsets)

    @staticmethod
    def backward(ctx, a=None):
        return None, None


sec_min = SecMin.apply


class SecMax(Function):
    @staticmethod
    def forward(ctx, inp, offsets):
        '''(functions/pointgroup_ops.py:312-337)'''
        return _sec_forward(PG_OP.sec_max, inp, offsets)

    @staticmethod
    def backward(ctx, a=None):
        return None, None


sec_max = SecMax.apply


class GatherRows(Function):
    """ADDITION (not in the reference API): feats[idx] for fp32 [N, C] rows as one coalesced kernel.
    torch's own row gather is launch-rate bound on this shape (one tiny block per row); the
    reference's caller does this gather between the ops (model/pointgroup.py:133-134,333)."""

    @staticmethod
    def forward(ctx, feats, idx):
        ctx.save_for_backward(idx)
        ctx.n_rows = feats.size(0)
        return PG_OP.gather_rows(feats.contiguous(), idx.contiguous())

    @staticmethod
    def backward(ctx, d_out):
        (idx,) = ctx.saved_tensors
        d_feats = torch.zeros((ctx.n_rows, d_out.size(1)), dtype=d_out.dtype, device=d_out.device)
        d_feats.index_add_(0, idx.long(), d_out)
        return d_feats, None


gather_rows = GatherRows.apply


def cluster_voxel_coords(coords, cluster_idxs, cluster_offsets, fullscale, scale, rand6):
    """Not part of the reference's operator API: the glue of clusters_voxelization (model/pointgroup.py:125-167)
    as one fused op, for callers willing to edit that function.  Returns (clusters_coords int64 [S,4] for
    voxelization_idx, center [nC,3], size [nC,3]); not differentiable (the reference's `.long()` is not either)."""
    return PG_OP.cluster_coords(coords.contiguous(), cluster_idxs.contiguous(), cluster_offsets.contiguous(), fullscale,
                                scale, rand6.contiguous())


def cross_iou(proposals_idx, num_proposals, N, want_npoint=False):
    """Not part of the reference's operator API: the proposal-vs-proposal IoU matrix of the instance NMS in
    PointGroup.test (model/pointgroup.py:577-590) without the dense [nProposal, N] mask and its matmul.
    ``proposals_idx`` int32 [sumNPoint, 2] rows (proposal id, point id) as bfs_cluster returns them (CPU
    tensors are staged through the current CUDA device); returns fp32 [num_proposals, num_proposals],
    bit-identical to ``inter / (n_h + n_v - inter)`` of the reference."""
    dev = PG_OP._compute_device(proposals_idx)
    res = PG_OP.cross_iou(proposals_idx.to(dev).contiguous(), num_proposals, N, want_npoint)
    if proposals_idx.is_cuda:
        return res
    return tuple(r.cpu() for r in res) if want_npoint else res.cpu()


def nms_instances(cross_ious, scores, threshold):
    """get_nms_instances (lib/utils/eval.py:75-97) on the device; torch tensors in, int32 pick indices out
    (the reference takes and returns numpy arrays after copying the matrix to the host)."""
    dev = PG_OP._compute_device(scores)
    pick = PG_OP.nms_instances(cross_ious.to(dev).float().contiguous(), scores.to(dev).float().contiguous(), threshold)
    return pick if scores.is_cuda else pick.cpu()


def pick_masks(proposals_idx, proposals_offset, pick_idxs, N):
    """Not part of the reference's operator API: ``proposals_mask[pick_idxs]`` of PointGroup.test
    (model/pointgroup.py:579-580,593) for the picked proposals only -- int32 [nPick, N] on the inputs' device.
    ``pick_idxs`` are proposal ids in the numbering of ``proposals_offset`` (after a score / size filter:
    ``keep[pick]``)."""
    dev = PG_OP._compute_device(proposals_idx)
    out = PG_OP.pick_masks(proposals_idx.to(dev).contiguous(), proposals_offset.to(dev).int().contiguous(),
                           torch.as_tensor(pick_idxs).to(dev).int().contiguous(), N)
    return out if proposals_idx.is_cuda else out.cpu()
