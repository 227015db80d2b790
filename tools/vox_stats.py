"""Sizes of the cluster re-voxelisation of the bench workload: M, maxActive and the distribution of points per voxel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes, pointgroup_ops as ops
dev = torch.device("cuda", 0)
nb = scenes.make_batch(8, 150_000, config_id=2, with_feats=False)
batch = chain.batch_to_device(nb, dev)
trace = {}
out = chain.proposal_chain(ops, batch, torch.full((6,), 0.5, device=dev), trace=trace)
cf, v2p, vf = trace["voxelization(clusters)"]
n = v2p[:, 0].long()
print("M", v2p.shape[0], "W", v2p.shape[1], "points", int(n.sum()), "C", cf.shape[1])
for t in (4, 8, 16, 24, 48, 96, 200, 400, 800, 1600):
    m = n > t
    print("  rows with n > %4d: %7d  holding %8d points" % (t, int(m.sum()), int(n[m].sum())))
print("  largest rows:", torch.sort(n, descending=True).values[:12].tolist())
