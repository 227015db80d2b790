#!/usr/bin/env python
"""Ball-query grid statistics of the bench batch (8 x 150k points), computed with torch: cells, candidates per cell (K),
queries per cell (nq), the dense / medium / small classes of csrc/ballquery.cu and the distance tests each performs.
    python tools/cell_stats.py [scenes] [points]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
dev = torch.device("cuda", 0)
nb = scenes.make_batch(ns, npts, config_id=2, with_feats=False)
b = chain.batch_to_device(nb, dev)
obj = torch.nonzero(b["semantic_preds"] > 0).view(-1)
bi = b["locs_scaled"][:, 0][obj]
xyz = b["locs"][obj].double()
sh = xyz + b["pt_offsets"][obj].double()
s = 0.03 * 1.0001
for name, pts in (("shift", sh), ("raw", xyz)):
    for mult in (1, 2):
        c = torch.floor(pts / (s * mult)).long() + 4096
        key = ((bi * 8192 + c[:, 0]) * 8192 + c[:, 1]) * 8192 + c[:, 2]
        uk, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
        K = torch.zeros_like(cnt)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    q = uk + (dx * 8192 + dy) * 8192 + dz
                    pos = torch.searchsorted(uk, q).clamp(max=uk.numel() - 1)
                    K += torch.where(uk[pos] == q, cnt[pos], torch.zeros_like(cnt))
        nq = cnt
        dense = (K > 256) | ((K > 128) & (nq > 16))
        medium = ~dense & (K > 128)
        small = ~dense & ~medium
        pad = lambda x: (x + 31) // 32 * 32
        def rep(tag, m):
            if int(m.sum()) == 0:
                print("   %-7s none" % tag); return
            print("   %-7s cells %8d  points %8d  sumK %10d  tests %12d  padded tests %12d  nq mean %.1f max %d  K mean %.0f max %d"
                  % (tag, int(m.sum()), int(nq[m].sum()), int(K[m].sum()), int((nq[m] * K[m]).sum()), int((pad(nq[m]) * pad(K[m])).sum()),
                     float(nq[m].float().mean()), int(nq[m].max()), float(K[m].float().mean()), int(K[m].max())))
        print("%s  cell edge %d r: n %d cells %d" % (name, mult, pts.shape[0], uk.numel()))
        rep("small", small); rep("medium", medium); rep("dense", dense)
        if mult == 1 and name == "shift":
            d = dense
            for lo, hi in ((0, 16), (16, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 100000)):
                m = d & (nq > lo) & (nq <= hi)
                print("      dense nq in (%d,%d]: cells %d  tests %d  K mean %.0f" % (lo, hi, int(m.sum()), int((nq[m] * K[m]).sum()),
                      float(K[m].float().mean()) if int(m.sum()) else 0))
