#!/usr/bin/env python
"""bench.py -- scenes/sec of the PointGroup proposal hot path (voxelize + cluster + roipool + IoU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of d3net_b200.chain.proposal_chain over one collated batch of 8 synthetic
150k-point ScanNet-shaped scenes per GPU (BASELINE.json configs[1]; configs[2] under torchrun: each
rank owns its own 8 scenes -- weak scaling, scene-per-GPU -- and the packed per-scene proposal
tensors are all-gathered over NCCL every step).  One JSON line is printed by rank 0.

  value      scenes/sec over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the public API with HOST (pinned) inputs: H2D of the step's inputs and
             D2H of its results inside the timed region
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the same chain (oracle/, with the reference's own compiled
             voxelize_idx / bfs_cluster when oracle/_ref exists) on a bounded sample, rank 0, N=1
  --impl reference  times only that CPU path (the reference has no GPU implementation of
             voxelize_idx / bfs_cluster and no multi-GPU path of its own)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from d3net_b200 import chain, scenes  # noqa: E402

METRIC = "scenes/sec (voxelize+cluster+roipool)"
UNIT = "scenes/s"
SCENES_PER_GPU = 8
POINTS = 150_000


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


KERNEL_NOTES = {
    "k_bq_cells_dense": "ball-query count over dense cells: 0.8 G exact fp32 distance tests per step (9 SASS instructions each) "
                        "on coordinates that stay in shared memory -- FP32-issue / barrier bound, so its share of the HBM roofline "
                        "says little; the streaming kernels are listed under hbm_kernels (DESIGN.md section 5)",
    "k_bq_fill_mask": "decodes 1 bit per (query, candidate) into the neighbour index lists: writes 4 B per neighbour, instruction-bound "
                      "on the bit -> position arithmetic",
    "k_cl_verify<trusted>": "edge sweep of the union-find; algorithmic bytes count every list once although settled cells are never read",
}

# ----------------------------------------------------------------------------------------------------
# algorithmic bytes per op (restated in DESIGN.md)
# ----------------------------------------------------------------------------------------------------
def kernel_algorithmic_bytes(batch, out):
    """Algorithmic bytes per STEP of the library's main kernels: the part of the op's
    compulsory traffic that the kernel is there to move, summed over its launches in a step."""
    N = batch["locs"].shape[0]
    C = batch["feats"].shape[1]
    M = out["voxel_feats"].shape[0]
    n = out["n_object_points"]
    S = out["proposals_idx"].shape[0]
    nP = out["proposals_offset"].numel() - 1
    Mc = out["proposals_voxel_feats"].shape[0]
    nA = out["nActive_shift"] + out["nActive_raw"]
    maps = out["v2p_map_numel"] + out["proposals_v2p_map_numel"]      # int32 [M, maxActive + 1] rows
    return {
        # ball query, count phase: coords + scene ids in, (start, len) rows out -- per set
        "k_bq_cells_dense": 2 * (12 * n + 4 * n + 8 * n),
        # ball query, fill phase: (start, len) rows in, neighbour indices out
        "k_bq_fill_mask": 2 * (8 * n + 4 * n) + 4 * nA,
        # bfs_cluster, edge sweep: neighbour indices + (start, len) + labels in
        "k_cl_verify<trusted>": 4 * nA + 2 * (8 * n + 4 * n),
        "k_cl_verify<validating>": 4 * nA + 2 * (8 * n + 4 * n),
        # voxelization: features in, voxel means out, one map row per voxel (scene C=134, clusters C=16)
        "k_voxelize_fp": 4 * N * C + 4 * M * C + 4 * S * 16 + 4 * Mc * 16 + 4 * maps,
        # voxelization_idx, fill phase: coords in (first point of each voxel), coords + zero-padded map rows out
        "k_vox_fill": 32 * M + 32 * M + 32 * Mc + 32 * Mc + 4 * maps,
        "k_sec_mean": 4 * S * 3 + 4 * (nP + 1) + 4 * nP * 3,
    }


def algorithmic_bytes(batch, out):
    N = batch["locs"].shape[0]
    C = batch["feats"].shape[1]
    M, W = out["voxel_feats"].shape[0], None
    n = out["n_object_points"]
    S = out["proposals_idx"].shape[0]
    nP = out["proposals_offset"].numel() - 1
    nI = batch["instance_pointnum"].numel()
    b = {}
    b["voxelization(scene)"] = 4 * N * C + 4 * M * C + 4 * M * 2          # + the map (>= 2 ints per row)
    b["ballquery(shift).fill"] = 16 * n + 8 * n + 4 * out["nActive_shift"]
    b["ballquery(raw).fill"] = 16 * n + 8 * n + 4 * out["nActive_raw"]
    b["bfs_cluster(shift)"] = 4 * n + 4 * out["nActive_shift"] + 8 * n + 8 * S // 2
    b["bfs_cluster(raw)"] = 4 * n + 4 * out["nActive_raw"] + 8 * n + 8 * S // 2
    b["roipool"] = 4 * S * 16 + 4 * (nP + 1) + 8 * nP * 16
    b["get_iou"] = 4 * S + 8 * S + 4 * nI + 4 * (nP + 1) + 4 * nP * nI
    return b


# ----------------------------------------------------------------------------------------------------
# CPU leg (oracle / reference-compiled ops) -- the only place bench.py touches oracle/
# ----------------------------------------------------------------------------------------------------
def cpu_chain_rate(budget_s, steps, warmup):
    """Times the CPU restatement of the chain on a bounded sample.  Returns (scene-equivalents/s,
    ms per step, description, cores, kind)."""
    from oracle.ops_adapter import OracleOps
    from oracle import pg_oracle
    cores = os.cpu_count() or 1
    pg_oracle.set_threads(cores)
    torch.set_num_threads(cores)
    ops = OracleOps(use_ref=True)
    kind = "port"
    # calibrate on a 15k-point scene, then pick the sample size that fits the budget
    cal = chain.batch_to_device(scenes.make_batch(1, 15000, config_id=2, geometry_points=15000), None)
    t0 = time.perf_counter()
    chain.proposal_chain(ops, cal)
    per_point = (time.perf_counter() - t0) / 15000
    n_steps = steps + warmup
    pts = int(min(POINTS, max(5000, budget_s / max(n_steps, 1) / per_point)))
    nb = scenes.make_batch(1, pts, config_id=2, geometry_points=pts)
    b = chain.batch_to_device(nb, None)
    for _ in range(warmup):
        chain.proposal_chain(ops, b)
    t0 = time.perf_counter()
    for _ in range(steps):
        chain.proposal_chain(ops, b)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    rate = (pts / POINTS) / dt
    desc = ("1 synthetic scene of %d points (=%.3f of a 150k-point scene) per step through the CPU chain: "
            "voxelize_idx + bfs_cluster via %s, CUDA-only ops via oracle/pg_oracle.c (OpenMP)"
            % (pts, pts / POINTS, "oracle/_ref (reference sources compiled)" if ops.ref is not None
               else "oracle/pg_oracle.c"))
    return rate, dt * 1e3, desc, cores, kind


def run_reference(args, rank, world, emit=print):
    if rank != 0:
        return
    rate, ms, desc, cores, kind = cpu_chain_rate(args.cpu_budget_s, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: full pointgroup_ops chain on synthetic 150k-point scenes (CPU sample)",
                   "points_per_scene": POINTS},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU leg
# ----------------------------------------------------------------------------------------------------
def count_pg_kernels(fn):
    """Kernels of this library (namespace pg) launched by one call of fn, counted with CUPTI."""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if "pg::" in e.key]
        return sum(r[1] for r in rows), sorted(rows, key=lambda r: -r[2])
    except Exception as e:  # CUPTI can be unavailable under another profiler
        return None, [("profiler unavailable: %r" % (e,), 0, 0)]


def run_b200(args, rank, world, local, emit=print):
    import torch.distributed as dist
    from d3net_b200 import pointgroup_ops as ops, dist as pgdist, _native
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    _native.lib()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / warnings belong on stderr
        dist.init_process_group("nccl", device_id=dev)

    n_scenes = args.scenes
    first_scene = rank * n_scenes                      # scene-per-GPU sharding: each rank its own scenes
    nb = scenes.make_batch(n_scenes, args.points, config_id=2, first_scene=first_scene, with_feats=False)
    host = chain.batch_to_device(nb, None, pt_feat_seed=rank, pin=True)
    batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
    rand6 = torch.full((6,), 0.5, device=dev)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))

    def step_device(timer=None, fused_glue=False):
        out = chain.proposal_chain(ops, batch, rand6, timer, fused_glue=fused_glue)
        packed = pgdist.pack_proposals(out, batch, args.max_proposals)
        gathered = pgdist.all_gather_proposals(packed)
        return out, gathered

    d2h_keep = {}

    # End to end: every step's inputs start in pinned HOST memory.  The upload of step k+1 runs on a copy
    # stream while step k computes (what a serving loop does); every byte of every upload is inside the
    # timed region, and each step ends with the device->host read of its results.
    copy_stream = torch.cuda.Stream(device=dev)

    def upload():
        with torch.cuda.stream(copy_stream):
            b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    e2e_state = {"next": None, "left": 0}

    def step_e2e():
        if e2e_state["next"] is None:
            e2e_state["next"] = upload()
        b, ev = e2e_state["next"]
        e2e_state["left"] -= 1
        e2e_state["next"] = upload() if e2e_state["left"] > 0 else None
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for v in b.values():
            if torch.is_tensor(v):
                v.record_stream(cur)
        out = chain.proposal_chain(ops, b, rand6)
        packed = pgdist.pack_proposals(out, b, args.max_proposals)
        gathered = pgdist.all_gather_proposals(packed)
        res = [gathered, out["ious"], out["proposals_offset"]]
        hostres = [r.to("cpu", non_blocking=False) for r in res]     # the device->host read of the step's result
        d2h_keep["bytes"] = sum(r.numel() * r.element_size() for r in res)
        return hostres

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.profile_mode:
        for _ in range(args.warmup + args.steps):
            step_device()
        torch.cuda.synchronize()
        return
    # warm-up (also sizes the caching allocator)
    for _ in range(max(args.warmup, 3)):
        out, _ = step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region 1: device-resident inputs, with the per-op event timers live
    timer = chain.SectionTimer(True)
    ops._section_timer = timer
    ms_total = timed(lambda: step_device(timer), args.steps)
    ops._section_timer = None
    # timed region 2: end to end from pinned host memory
    e2e_state["left"] = 2
    for _ in range(2):
        step_e2e()
    e2e_state["left"] = args.steps
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # timed region 2b: the caller's clusters_voxelization glue replaced by the fused op (an
    # edit to the caller, so it is reported next to `value`, not as `value`)
    for _ in range(2):
        step_device(fused_glue=True)
    ms_fused = timed(lambda: step_device(fused_glue=True), args.steps)

    # timed region 3: the same steps with the library's per-kernel CUDA-event timers on (events on the
    # launching stream around every main kernel; include/pg_b200.h, pg_kernel_timing)
    _native.kernel_timing(True)
    barrier()
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize()
    ktimes = _native.kernel_timing_report()
    _native.kernel_timing(False)

    # kernels of this library launched per step (CUPTI count of one untimed step; every rank takes part
    # because the step contains the collective)
    n_launch, top = count_pg_kernels(lambda: step_device())

    total_scenes = n_scenes * world
    value = total_scenes * args.steps / (ms_total / 1e3)
    e2e_value = total_scenes * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        sec_ms = {k: v / args.steps for k, v in timer.totals_ms().items()}
        algo = algorithmic_bytes(batch, out)
        peak, peak_kind = peaks()
        # the roofline is quoted for the kernel that takes the most time per step, from the library's own
        # event timers; its DRAM traffic comes from the committed ncu capture of the same workload
        kalgo = kernel_algorithmic_bytes(batch, out)
        traffic_db = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic_db = json.load(f)
        per_kernel = {}
        for name, (cnt, ms) in sorted(ktimes.items(), key=lambda kv: -kv[1][1]):
            ms_step = ms / args.steps
            row = {"launches_per_step": cnt / args.steps, "ms_per_step": round(ms_step, 4)}
            if name in kalgo:
                row["algorithmic_bytes_per_step"] = int(kalgo[name])
                row["GBps"] = round(kalgo[name] / (ms_step / 1e3) / 1e9, 1)
                row["frac"] = round(row["GBps"] / peak, 4)
            if name in traffic_db:
                row["dram_bytes_per_step"] = traffic_db[name]
            per_kernel[name] = row
        dom = next(iter(per_kernel))
        dom_ms = per_kernel[dom]["ms_per_step"]
        dom_launches = max(per_kernel[dom]["launches_per_step"], 1)
        achieved = per_kernel[dom].get("GBps", 0.0)
        per_op = {k: {"ms": round(v, 4), "GBps": (round(algo[k] / (v / 1e3) / 1e9, 1) if k in algo else None)}
                  for k, v in sorted(sec_ms.items(), key=lambda kv: -kv[1])}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, ms, desc, cores, kind = cpu_chain_rate(20.0, 1, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "points_per_sec": value * args.points,
            "config": {"workload": "configs[1]: full pointgroup_ops chain (voxelization_idx, voxelization C=134, "
                                   "ballquery+bfs_cluster on shifted and raw coords, sec_mean/min/max, cluster "
                                   "re-voxelisation C=16, roipool, get_iou) on %d synthetic %dk-point scenes per GPU"
                                   % (n_scenes, args.points // 1000),
                       "scenes_per_gpu": n_scenes, "points_per_scene": args.points, "parallelism": "scene-per-GPU x%d" % world,
                       "cluster_radius": scenes.CLUSTER_RADIUS, "npoint_thre": scenes.CLUSTER_NPOINT_THRE,
                       "l2": "inputs larger than L2 (%.0f MB of point features per step)" % (batch["feats"].numel() * 4 / 1e6),
                       "collective": "all_gather of [scenes, %d, 46] fp32 proposal block" % args.max_proposals if world > 1 else "none (N=1)",
                       "n_points": int(batch["locs"].shape[0]), "n_object_points": int(out["n_object_points"]),
                       "nActive_shift": int(out["nActive_shift"]), "nActive_raw": int(out["nActive_raw"]),
                       "n_proposals": int(out["proposals_offset"].numel() - 1), "sumNPoint": int(out["proposals_idx"].shape[0])},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "H2D of step k+1 on a copy stream overlaps step k; all uploads inside the timed region",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_keep.get("bytes", 0))},
            "fused_glue": {"value": total_scenes * args.steps / (ms_fused / 1e3), "unit": UNIT, "ms_per_step": ms_fused / args.steps,
                           "what": "same chain with the caller-side clusters_voxelization glue (model/pointgroup.py:125-167) "
                                   "done by pointgroup_ops.cluster_voxel_coords; bit-identical outputs"},
            "gpu_launches": (n_launch * args.steps) if n_launch is not None else None,
            "gpu_launches_per_step": n_launch,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (traffic_db[dom] / dom_launches) if dom in traffic_db else None,
                         "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else peak_kind,
                         "algorithmic_bytes_per_launch": int(kalgo.get(dom, 0) / dom_launches),
                         "ms_per_launch": dom_ms / dom_launches, "launches_per_step": dom_launches,
                         "timing": "CUDA events on the launching stream around every launch of the kernel, %d steps" % args.steps,
                         "note": KERNEL_NOTES.get(dom, "")},
            # the kernels that stream their operands once (the ones an HBM roofline is the right yardstick for)
            "hbm_kernels": {k: {"GBps": v["GBps"], "frac": v["frac"]} for k, v in per_kernel.items()
                            if k in ("k_voxelize_fp", "k_vox_fill", "k_bq_fill_mask") and "GBps" in v},
            "per_kernel": per_kernel,
            "per_op": per_op,
            "top_kernels": [{"name": k[:80], "calls": c, "us": round(t, 1)} for k, c, t in top[:8]],
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=SCENES_PER_GPU, help="scenes per GPU per step")
    ap.add_argument("--points", type=int, default=POINTS)
    ap.add_argument("--max-proposals", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="--impl reference: CPU seconds the whole run may take (sizes the sample of the workload)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="only W warm-up + K device-resident steps, no JSON line (for runs under ncu)")
    args = ap.parse_args()
    rank, world, local = dist_env()
    # stdout carries exactly one JSON line: while the run is going, file descriptor 1 points at stderr, so
    # nothing a library prints (NCCL's version banner, for one) can land in front of it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lines.append
    try:
        if args.impl == "reference":
            run_reference(args, rank, world, emit)
        else:
            run_b200(args, rank, world, local, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
