"""TEST INFRASTRUCTURE ONLY -- the oracle dressed as the reference's operator API on CPU torch
tensors, so d3net_b200.chain.proposal_chain can be replayed on the host (bench.py's cpu_baseline and
--impl reference legs, and the chain parity tests).  Never imported by the product package.

With ``use_ref=True`` and oracle/_ref/PG_OP.so present, the two ops the reference itself implements on
the CPU -- voxelize_idx and bfs_cluster -- run through the reference's own compiled code; the ops that
exist only as CUDA kernels in the reference run through the C restatement (oracle/pg_oracle.c)."""
import numpy as np
import torch

from . import build_ref, pg_oracle as o


def _n(t):
    return t.detach().cpu().numpy()


class OracleOps:
    def __init__(self, use_ref=True):
        o.build()
        self.ref = build_ref.load() if use_ref else None

    def voxelization_idx(self, coords, batchsize, mode=4):
        if self.ref is not None:
            oc, im, om = coords.new(), torch.zeros(coords.size(0), dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
            self.ref.voxelize_idx(coords.contiguous(), oc, im, om, batchsize, mode)
            return oc, im, om
        oc, im, om = o.voxelization_idx(_n(coords), batchsize, mode)
        return torch.from_numpy(oc), torch.from_numpy(im), torch.from_numpy(om)

    def voxelization(self, feats, map_rule, mode=4):
        return torch.from_numpy(o.voxelization(_n(feats), _n(map_rule), mode))

    def ballquery_batch_p(self, coords, batch_idxs, batch_offsets, radius, meanActive):
        idx, sl = o.ballquery_batch_p(_n(coords), _n(batch_idxs), _n(batch_offsets), radius)
        return torch.from_numpy(idx), torch.from_numpy(sl)

    def bfs_cluster(self, semantic_label, ball_query_idxs, start_len, threshold):
        if self.ref is not None:
            ci, co = semantic_label.new(), semantic_label.new()
            self.ref.bfs_cluster(semantic_label.contiguous(), ball_query_idxs.contiguous(), start_len.contiguous(),
                                 ci, co, start_len.size(0), threshold)
            return ci, co
        ci, co = o.bfs_cluster(_n(semantic_label), _n(ball_query_idxs), _n(start_len), threshold)
        return torch.from_numpy(ci), torch.from_numpy(co)

    def roipool(self, feats, proposals_offset):
        return torch.from_numpy(o.roipool(_n(feats), _n(proposals_offset))[0])

    def get_iou(self, proposals_idx, proposals_offset, instance_labels, instance_pointnum):
        return torch.from_numpy(o.get_iou(_n(proposals_idx), _n(proposals_offset), _n(instance_labels),
                                          _n(instance_pointnum)))

    def sec_mean(self, inp, offsets):
        return torch.from_numpy(o.sec_mean(_n(inp), _n(offsets)))

    def sec_min(self, inp, offsets):
        return torch.from_numpy(o.sec_min(_n(inp), _n(offsets)))

    def sec_max(self, inp, offsets):
        return torch.from_numpy(o.sec_max(_n(inp), _n(offsets)))


class RefGpuOps:
    """The reference's lib/pointgroup_ops as a D3Net user runs it today, on the GPU box: its nine CUDA kernels
    (compiled unmodified for sm_100a into oracle/_ref) on device tensors and its two CPU ops after the D2H copies its
    callers make (model/pointgroup.py:167-169,297,305), with the wrapper's protocol (zero-filled outputs, the ball
    query's grow-and-retry loop, functions/pointgroup_ops.py:135-142).  bench.py's `reference_mixed` leg and the GPU
    reference tests only; needs a CUDA device."""

    def __init__(self):
        self.ref = build_ref.load()

    def voxelization_idx(self, coords, batchsize, mode=4):
        c = coords.cpu().contiguous()                                     # the op is CPU-only in the reference
        oc, im, om = c.new(), torch.zeros(c.size(0), dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
        self.ref.voxelize_idx(c, oc, im, om, batchsize, mode)
        return oc.to(coords.device), im.to(coords.device), om.to(coords.device)

    def voxelization(self, feats, map_rule, mode=4):
        M, W = map_rule.shape
        out = torch.zeros((M, feats.size(1)), dtype=torch.float32, device=feats.device)
        self.ref.voxelize_fp(feats.contiguous(), out, map_rule.contiguous(), mode, M, W - 1, feats.size(1))
        return out

    def ballquery_batch_p(self, coords, batch_idxs, batch_offsets, radius, meanActive):
        n = coords.size(0)
        while True:
            idx = torch.zeros(n * meanActive, dtype=torch.int32, device=coords.device)
            start_len = torch.zeros((n, 2), dtype=torch.int32, device=coords.device)
            nActive = self.ref.ballquery_batch_p(coords, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius)
            if nActive <= n * meanActive:
                break
            meanActive = int(nActive // n + 1)
        return idx[:nActive], start_len

    def bfs_cluster(self, semantic_label, ball_query_idxs, start_len, threshold):
        sem = semantic_label.cpu().contiguous()
        ci, co = sem.new(), sem.new()
        self.ref.bfs_cluster(sem, ball_query_idxs.cpu().contiguous(), start_len.cpu().contiguous(), ci, co,
                             start_len.size(0), threshold)
        return ci.to(semantic_label.device), co.to(semantic_label.device)

    def roipool(self, feats, proposals_offset):
        nP, C = proposals_offset.numel() - 1, feats.size(1)
        out = torch.zeros((nP, C), dtype=torch.float32, device=feats.device)
        arg = torch.zeros((nP, C), dtype=torch.int32, device=feats.device)
        self.ref.roipool_fp(feats.contiguous(), proposals_offset.contiguous(), out, arg, nP, C)
        return out

    def get_iou(self, proposals_idx, proposals_offset, instance_labels, instance_pointnum):
        nP, nI = proposals_offset.numel() - 1, instance_pointnum.numel()
        out = torch.zeros((nP, nI), dtype=torch.float32, device=proposals_idx.device)
        self.ref.get_iou(proposals_idx.contiguous(), proposals_offset.contiguous(), instance_labels.contiguous(),
                         instance_pointnum.contiguous(), out, nI, nP)
        return out

    def _sec(self, fn, inp, offsets):
        nP, C = offsets.numel() - 1, inp.size(1)
        out = torch.zeros((nP, C), dtype=torch.float32, device=inp.device)
        fn(inp.contiguous(), offsets.contiguous(), out, nP, C)
        return out

    def sec_mean(self, inp, offsets):
        return self._sec(self.ref.sec_mean, inp, offsets)

    def sec_min(self, inp, offsets):
        return self._sec(self.ref.sec_min, inp, offsets)

    def sec_max(self, inp, offsets):
        return self._sec(self.ref.sec_max, inp, offsets)
