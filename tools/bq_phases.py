#!/usr/bin/env python
"""CUDA-event timing of the ball query's phases (prepare+count, fill) on the bench batch, with the CUPTI kernel table:
    python tools/bq_phases.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes, PG_OP
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
nb = scenes.make_batch(8, 150000, config_id=2, with_feats=False)
b = chain.batch_to_device(nb, dev)
obj = torch.nonzero(b["semantic_preds"] > 0).view(-1)
bi = b["locs_scaled"][:, 0].int()[obj].contiguous()
bo = chain.get_batch_offsets(bi, 8)
xyz = b["locs"][obj].contiguous()
sh = (xyz + b["pt_offsets"][obj]).contiguous()
for name, pts in (("shift", sh), ("raw", xyz)):
    for it in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        sl, total, st = PG_OP.ballquery_count_impl(pts, bi, bo, 0.03)
        e[1].record()
        idx = torch.empty(total, dtype=torch.int32, device=dev)
        PG_OP.ballquery_fill_impl(pts, 0.03, sl, idx, st)
        e[2].record()
        torch.cuda.synchronize()
        print(name, it, "count %.3f fill %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])), total)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sl, total, st = PG_OP.ballquery_count_impl(pts, bi, bo, 0.03)
        torch.cuda.synchronize()
    rows = sorted(((e.key[:60], e.count, e.device_time_total) for e in prof.key_averages()), key=lambda r: -r[2])
    for r in rows[:14]:
        print("   %-60s x%-3d %9.1f us" % r)
