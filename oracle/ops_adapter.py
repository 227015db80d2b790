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
