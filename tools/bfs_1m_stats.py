#!/usr/bin/env python
"""bfs_cluster on the 1M-point scene (BASELINE configs[3]): parked one-way edges, propagation rounds, kernel times."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes, pointgroup_ops as ops, PG_OP, _native

dev = torch.device("cuda", 0)
nb = scenes.make_batch(1, 1000000, config_id=2, with_feats=False)
b = chain.batch_to_device(nb, dev)
sem = b["semantic_preds"]
obj = torch.nonzero(sem > 0).view(-1)
bi = b["locs_scaled"][:, 0].int()[obj].contiguous()
bo = chain.get_batch_offsets(bi, 1)
xyz = b["locs"][obj].contiguous()
sh = (xyz + b["pt_offsets"][obj]).contiguous()
s32 = sem[obj].int().contiguous()
idx, sl = ops.ballquery_batch_p(sh, bi, bo, 0.03, 300)
print("n", sh.shape[0], "nA", idx.numel(), "full lists", int((sl[:, 1] >= 1000).sum()))
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if it == 2:
        _native.kernel_timing(True)
    ci, co = ops.bfs_cluster(s32, idx, sl, 50)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
rep = _native.kernel_timing_report(); _native.kernel_timing(False)
d = PG_OP.bfs_cluster_debug()
print("wall ms", round(dt * 1e3, 2), "pending", d[2], "host rounds", d[3], "swept lists", d[4], "clusters", co.numel() - 1,
      {k: round(v[1], 3) for k, v in rep.items()})
