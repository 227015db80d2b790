"""GPU timeline of one chain step (CUPTI through torch.profiler): busy / idle time and the largest gaps.

    python tools/timeline.py [--overlap] [--fused-cluster] [--fused-glue] [--out gpurun_out/timeline.json]

Prints the union-busy time of the device over one step, the idle gaps above 8 us with the kernels either side of them,
and per-kernel totals.  Not a benchmark: the profiler adds host overhead per launch (the gaps are upper bounds).
Programmatic dependent launch is switched off unless --pdl is given: with it a kernel's CUPTI interval starts when its
first block becomes resident, i.e. it includes the wait for its predecessor, and per-kernel totals are over-counted.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--fused-cluster", action="store_true")
    ap.add_argument("--fused-glue", action="store_true")
    ap.add_argument("--pdl", action="store_true")
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--points", type=int, default=150_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    a = ap.parse_args()
    if not a.pdl:
        os.environ["PG_B200_NO_PDL"] = "1"
    from torch.profiler import profile, ProfilerActivity
    from d3net_b200 import chain, scenes, pointgroup_ops as ops
    dev = torch.device("cuda", 0)
    nb = scenes.make_batch(a.scenes, a.points, config_id=2, with_feats=False)
    batch = chain.batch_to_device(nb, dev)
    rand6 = torch.full((6,), 0.5, device=dev)

    def step():
        return chain.proposal_chain(ops, batch, rand6, overlap=a.overlap, fused_cluster=a.fused_cluster, fused_glue=a.fused_glue)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, e.name))
    ev.sort()
    if not ev:
        print("no device events")
        return
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy, cur_s, cur_e = 0.0, ev[0][0], ev[0][1]
    gaps = []
    last_name = ev[0][2]
    for s, e, name in ev[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, cur_e - t0, last_name, name))
            cur_s, cur_e = s, e
            last_name = name
        elif e > cur_e:
            cur_e = e
            last_name = name
    busy += cur_e - cur_s
    span = t1 - t0
    print("span %.1f us, busy %.1f us (%.1f %%), idle %.1f us in %d gaps; %d device events"
          % (span, busy, 100 * busy / span, span - busy, len(gaps), len(ev)))
    big = sorted(gaps, reverse=True)[:40]
    print("largest gaps (us, at us, after -> before):")
    for g, at, a_, b_ in big:
        print("  %7.1f  @%8.1f  %s -> %s" % (g, at, a_[:48], b_[:48]))
    hist = {}
    for g, *_ in gaps:
        k = "<4" if g < 4 else "<8" if g < 8 else "<16" if g < 16 else "<32" if g < 32 else "<64" if g < 64 else ">=64"
        hist.setdefault(k, [0, 0.0])
        hist[k][0] += 1
        hist[k][1] += g
    print("gap histogram:", {k: (v[0], round(v[1], 1)) for k, v in hist.items()})
    agg = {}
    for s_, e_, n_ in ev:
        k = n_.split("(")[0][:70]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += e_ - s_
    print("per kernel (calls, us):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-70s %3d %8.1f" % (k, v[0], v[1]))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"span_us": span, "busy_us": busy, "events": [(s - t0, e - t0, n[:80]) for s, e, n in ev]}, f)


if __name__ == "__main__":
    main()
