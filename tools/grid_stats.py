#!/usr/bin/env python
"""How much of the edge sweep the cell pass (k_cl_cells) saves on the bench batch (8 x 150k points), and what
the kernels of bfs_cluster cost with it:  python tools/grid_stats.py
(set d3net_b200.pointgroup_ops.USE_GRID_SWEEP = False for the plain trusted sweep)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes, pointgroup_ops as ops, PG_OP, _native

dev = torch.device("cuda", 0)
nb = scenes.make_batch(8, 150000, config_id=2, with_feats=False)
b = chain.batch_to_device(nb, dev)
sem = b["semantic_preds"]
obj = torch.nonzero(sem > 0).view(-1)
bi = b["locs_scaled"][:, 0].int()[obj].contiguous()
bo = chain.get_batch_offsets(bi, 8)
xyz = b["locs"][obj].contiguous()
sh = (xyz + b["pt_offsets"][obj]).contiguous()
s32 = sem[obj].int().contiguous()
for name, pts, ma in (("shift", sh, 300), ("raw", xyz, 50)):
    idx, sl = ops.ballquery_batch_p(pts, bi, bo, 0.03, ma)
    for it in range(3):
        if it == 2:
            _native.kernel_timing(True)
        ci, co = ops.bfs_cluster(s32, idx, sl, 50)
    torch.cuda.synchronize()
    rep = _native.kernel_timing_report()
    _native.kernel_timing(False)
    d = PG_OP.bfs_cluster_debug()
    print(name, "n", pts.shape[0], "nA", idx.numel(), "swept lists", d[4], "pending", d[2], "clusters", co.numel() - 1,
          {k: round(v[1], 3) for k, v in rep.items()})
