#!/usr/bin/env python
"""bench.py -- scenes/sec of the PointGroup proposal hot path (voxelize + cluster + roipool + IoU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config 1|2|3|4] [--scaling weak|strong]

A "step" is one pass of d3net_b200.chain.proposal_chain over one collated batch.  --config picks the BASELINE.json
workload (0-based index into its `configs`):
  1 (default)  8 synthetic 150k-point ScanNet-shaped scenes per GPU.  Under torchrun every rank owns its own 8 scenes
               (weak scaling, scene-per-GPU) and the packed per-scene proposal tensors are all-gathered over NCCL.
  2            the 64-scene batch of configs[2], STRONG scaling: 64 scenes per step in total, 64 / N per rank, run as
               sub-batches of 8 scenes (a 64-scene batch holds 2.7 G neighbours -- more than int32 offsets address, in
               the reference's format as much as here).  Same as --scaling strong.
  3            one 1M-point dense room (configs[3]).
  4            the reference's own caller (`PointGroup.feed`, staged unmodified under baseline/_ref) with the ops in the
               loop, 8 scenes, 256 proposals per scene (the detector half of configs[4]; see harness/d3net_stub.py).
One JSON line is printed by rank 0.

  value      scenes/sec over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the public API with HOST (pinned) inputs: H2D of the step's host inputs and D2H of its
             results inside the timed region
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the same chain (oracle/, with the reference's own compiled voxelize_idx /
             bfs_cluster when oracle/_ref exists) on a bounded sample, rank 0, N=1
  unchanged_caller / reference_mixed (N=1, config 1): the reference's caller over these ops with its CPU-tensor call
             pattern, and the reference's own kernels + CPU ops (oracle/_ref) on this GPU -- what a D3Net user has today
  --impl reference  times only the CPU path (the reference has no GPU implementation of voxelize_idx / bfs_cluster and
             no multi-GPU path of its own)
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
STRONG_TOTAL = 64

WORKLOADS = {
    1: "configs[1]: full pointgroup_ops chain (voxelization_idx, voxelization C=134, ballquery+bfs_cluster on shifted and "
       "raw coords, sec_mean/min/max, cluster re-voxelisation C=16, roipool, get_iou) on %d synthetic %dk-point scenes per GPU",
    2: "configs[2]: the same chain on a 64-scene batch sharded scene-per-GPU (%d scenes per rank per step, in sub-batches "
       "of %d scenes of %dk points) with an NCCL all-gather of the packed proposal block",
    3: "configs[3]: the same chain on %d dense %dk-point scene(s) (one room, ~7.7 mm pitch: neighbour lists at the "
       "1000 cap on shifted coordinates, one giant floor component)",
    4: "configs[4]: D3Net speaker forward -- the reference's own PointGroup.feed + SpeakerNet (model/pointgroup.py, "
       "model/speaker.py, model/caption_module.py unmodified; stand-in backbone / graph module) with the new ops in the loop "
       "on %d synthetic %dk-point scenes, 256 proposals per scene",
}


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


def bind_near_gpu(local):
    """Run this process (and first-touch its pinned buffers) on the CPUs NVML names as closest to GPU `local`."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want:
            os.sched_setaffinity(0, want)
            return "%d cpus near gpu %d" % (len(want), local)
        return "nvml affinity outside the allowed cpu set: not bound"
    except Exception as e:          # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


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


# Kernels whose time is not set by DRAM traffic: bench reports them against the limit that does bind them.
KERNEL_NOTES = {
    "k_bq_test_dense": "ball-query distance tests on dense cells: exact fp32 predicate on candidate tiles staged in shared "
                       "memory (packed f32x2 math); bound by FP32 issue, not by DRAM (DESIGN.md section 2a)",
    "k_bq_cells_dense": "ball-query count over dense cells: exact fp32 distance tests on coordinates staged in shared memory "
                        "-- FP32-issue / barrier bound, not DRAM bound (DESIGN.md section 2a)",
    "k_bq_fill_mask": "decodes 1 bit per (query, candidate) into the neighbour index lists: writes 4 B per neighbour",
    "k_cl_verify<trusted>": "edge sweep of the union-find; `GBps` / `frac` credit every list once (the op's compulsory bytes) "
                            "although the grid-assisted sweep never reads the lists of settled cells -- an algorithmic saving, "
                            "not bandwidth: what it moves is `dram_GBps` (latency-bound on one snapshot load per edge)",
}
COMPUTE_BOUND = ("k_bq_cells_dense", "k_bq_test_dense", "k_bq_cells_small", "k_bq_cells_medium")


# ----------------------------------------------------------------------------------------------------
# algorithmic bytes (restated in DESIGN.md section 2)
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
    count = 2 * (12 * n + 4 * n + 8 * n)        # ball query, count phase: coords + scene ids in, (start, len) rows out
    return {
        "k_bq_cells_dense": count, "k_bq_test_dense": count,
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
        "k_gather_rows": 2 * (4 * S * 16 * 2 + 8 * S),
    }


def algorithmic_bytes(batch, out):
    N = batch["locs"].shape[0]
    C = batch["feats"].shape[1]
    M = out["voxel_feats"].shape[0]
    n = out["n_object_points"]
    S = out["proposals_idx"].shape[0]
    nP = out["proposals_offset"].numel() - 1
    nI = batch["instance_pointnum"].numel()
    b = {}
    b["voxelization(scene)"] = 4 * N * C + 4 * M * C + 4 * out["v2p_map_numel"]
    b["ballquery(shift)"] = 16 * n + 4 * 9 + 8 * n + 4 * out["nActive_shift"]
    b["ballquery(raw)"] = 16 * n + 4 * 9 + 8 * n + 4 * out["nActive_raw"]
    b["ballquery(shift).fill"] = 16 * n + 8 * n + 4 * out["nActive_shift"]
    b["ballquery(raw).fill"] = 16 * n + 8 * n + 4 * out["nActive_raw"]
    b["bfs_cluster(shift)"] = 4 * n + 4 * out["nActive_shift"] + 8 * n + 8 * S // 2
    b["bfs_cluster(raw)"] = 4 * n + 4 * out["nActive_raw"] + 8 * n + 8 * S // 2
    b["roipool"] = 4 * S * 16 + 4 * (nP + 1) + 8 * nP * 16
    b["get_iou"] = 4 * S + 8 * S + 4 * nI + 4 * (nP + 1) + 4 * nP * nI
    b["voxelize_bp(scene, C=134)"] = 4 * M * C + 4 * out["v2p_map_numel"] + 4 * N * C
    b["roipool_bp"] = 8 * nP * 16 + 4 * S * 16
    return b


# ----------------------------------------------------------------------------------------------------
# CPU leg (oracle / reference-compiled ops) -- the only place bench.py touches oracle/
# ----------------------------------------------------------------------------------------------------
def cpu_chain_rate(budget_s, steps, warmup, points=POINTS, dense_room=False):
    """Times the CPU restatement of the chain on a bounded sample.  Returns (scene-equivalents/s,
    ms per step, description, cores, kind, detail)."""
    from oracle.ops_adapter import OracleOps
    from oracle import pg_oracle
    cores = os.cpu_count() or 1
    pg_oracle.set_threads(cores)
    torch.set_num_threads(cores)
    ops = OracleOps(use_ref=True)
    # `kind`: "port" -- the chain is this repo's CPU restatement; two of its ops (the only two the reference
    # implements on the CPU) run through the reference's own compiled code when oracle/_ref exists
    kind = "port"
    detail = ("voxelize_idx + bfs_cluster: %s (1 thread, as the reference); CUDA-only ops: oracle/pg_oracle.c, OpenMP x%d "
              "(BASELINE.md asks for torch-CPU there; the C port is faster, i.e. a conservative baseline)"
              % ("oracle/_ref = reference sources compiled" if ops.ref is not None else "oracle/pg_oracle.c", cores))
    # a sample keeps the DENSITY of the workload: an n-point sample of the 150k-point room is an n-point room at the same
    # 2 cm pitch; a sample of the 1M-point dense room is a smaller room at its 7.7 mm pitch
    geom = (lambda n: n) if not dense_room else (lambda n: int(n * scenes._GEOM_N / float(points)))
    cal_n = 100_000 if dense_room else 15000          # the generator needs a room of >= ~10k geometry points
    cal = chain.batch_to_device(scenes.make_batch(1, cal_n, config_id=2, geometry_points=geom(cal_n)), None)
    t0 = time.perf_counter()
    chain.proposal_chain(ops, cal)
    per_point = (time.perf_counter() - t0) / cal_n
    n_steps = steps + warmup
    pts = int(min(points, max(cal_n if dense_room else 5000, budget_s / max(n_steps, 1) / per_point)))
    nb = scenes.make_batch(1, pts, config_id=2, geometry_points=geom(pts))
    b = chain.batch_to_device(nb, None)
    for _ in range(warmup):
        chain.proposal_chain(ops, b)
    t0 = time.perf_counter()
    for _ in range(steps):
        chain.proposal_chain(ops, b)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    rate = (pts / points) / dt
    desc = ("1 synthetic scene of %d points (=%.3f of a %dk-point scene, same point density) per step through the CPU chain"
            % (pts, pts / points, points // 1000))
    return rate, dt * 1e3, desc, cores, kind, detail


def run_reference(args, rank, world, emit=print):
    if rank != 0:
        return
    points = 1_000_000 if args.config == 3 else args.points
    rate, ms, desc, cores, kind, detail = cpu_chain_rate(args.cpu_budget_s, args.steps, args.warmup, points,
                                                         dense_room=args.config == 3)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[%d] on the CPU: the same chain on a bounded sample (see cpu_baseline.sample)" % args.config,
                   "points_per_scene": points, "n_procs": 1,
                   "note": "one CPU process on rank 0 whatever --gpus says: the reference has no multi-process CPU path"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "kind_detail": detail, "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU leg
# ----------------------------------------------------------------------------------------------------
def count_pg_kernels(fn):
    """Kernels of this library (namespace pg) launched by one call of fn, counted with CUPTI; also every kernel."""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        ev = [e for e in prof.key_averages() if e.device_time_total > 0 or "pg::" in e.key]
        rows = [(e.key, e.count, e.device_time_total) for e in ev if "pg::" in e.key]
        n_all = sum(e.count for e in ev if not e.key.startswith("Memcpy") and not e.key.startswith("Memset"))
        return sum(r[1] for r in rows), sorted(rows, key=lambda r: -r[2]), n_all
    except Exception as e:  # CUPTI can be unavailable under another profiler
        return None, [("profiler unavailable: %r" % (e,), 0, 0)], None


def wall_ms(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


def event_ms(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def backward_rows(ops, batch, out, v2p_scene, peak):
    """The two backward kernels of the training step (functions/pointgroup_ops.py:66-73,210-219), timed as the autograd
    wrappers run them: zero-filled gradient buffer + kernel."""
    from d3net_b200 import PG_OP
    dev = batch["locs"].device
    N, C = batch["feats"].shape
    M = out["voxel_feats"].shape[0]
    maxA = v2p_scene.size(1) - 1
    g = torch.randn((M, C), device=dev)
    S = out["proposals_idx"].shape[0]
    nP = out["proposals_offset"].numel() - 1
    feats16 = torch.randn((S, 16), device=dev)
    pooled = torch.empty((nP, 16), device=dev)
    arg = torch.empty((nP, 16), dtype=torch.int32, device=dev)
    PG_OP.roipool_fp(feats16, out["proposals_offset"], pooled, arg, nP, 16)
    gp = torch.randn((nP, 16), device=dev)

    def vbp():
        d = torch.zeros((N, C), dtype=torch.float32, device=dev)
        PG_OP.voxelize_bp(g, d, v2p_scene, 4, M, maxA, C)

    def rbp():
        d = torch.zeros((S, 16), dtype=torch.float32, device=dev)
        PG_OP.roipool_bp(d, out["proposals_offset"], arg, gp, nP, 16)

    algo = algorithmic_bytes(batch, out)
    rows = {}
    for name, fn in (("voxelize_bp(scene, C=134)", vbp), ("roipool_bp", rbp)):
        fn()
        ms = event_ms(fn, 5)
        gbs = algo[name] / (ms / 1e3) / 1e9
        rows[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4),
                      "what": "zero fill of the gradient buffer + kernel, as the autograd wrapper runs it"}
    return rows


def h2d_probe(dev, world):
    """Bare pinned-host -> device copy rate of this rank while every rank copies at once (the ceiling of e2e)."""
    import torch.distributed as dist
    nbytes = 512 << 20
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    gbs = 4 * nbytes / (a.elapsed_time(b) / 1e3) / 1e9
    if world > 1:
        t = torch.tensor([gbs], device=dev)
        lo = t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(lo.item()), float(t.item())
    return gbs, gbs


def unchanged_caller_leg(n_scenes, points, reps=2):
    """The reference's own PointGroup.feed (staged unmodified, harness/d3net_stub.py) over d3net_b200.pointgroup_ops,
    with the caller's CPU-tensor call pattern (model/pointgroup.py:297,305,167-169) and no caller edit."""
    from harness import d3net_stub as H
    if not H.available():
        return {"unavailable": "baseline/_ref not staged (harness/stage_ref.py needs /root/reference)"}
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = H.load_cfg(max_num_proposal=256)
    model = H.build_detector(cfg, dev)
    nb = scenes.make_batch(n_scenes, points, config_id=2)
    dd = H.collate(nb, dev)
    out = H.run_feed(model, dd)                                         # warm-up
    ms = wall_ms(lambda: H.run_feed(model, dd), reps)
    with H.OpTimer() as ot:
        t0 = time.perf_counter()
        out = H.run_feed(model, dd)
        torch.cuda.synchronize()
        total_inst = (time.perf_counter() - t0) * 1e3
    ops_ms = sum(ot.ms.values())
    scores, pidx, poff = out["proposal_scores"]
    return {"value": n_scenes / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "what": "model/pointgroup.py PointGroup.feed, unmodified (forward + clusters_voxelization + "
                    "convert_stack_to_batch), over d3net_b200.pointgroup_ops; neighbour lists cross PCIe to the CPU and "
                    "back as the caller asks (idx.cpu(), model/pointgroup.py:297); stand-in pass-through backbone; wall clock",
            "ops_ms": round(ops_ms, 2), "ops_share": round(ops_ms / total_inst, 3),
            "per_op_ms": {k: round(v, 2) for k, v in sorted(ot.ms.items(), key=lambda kv: -kv[1])},
            "caller_ms_outside_ops": round(total_inst - ops_ms, 2),
            "n_proposals": int(poff.numel() - 1), "sumNPoint": int(pidx.shape[0]),
            "proposals_kept_per_batch": int(out["proposal_batch_mask"].sum())}


def speaker_leg(n_scenes, points, reps=2, chunk=8):
    """BASELINE configs[4]: the reference's detector caller + its caption module (model/speaker.py,
    model/caption_module.py, unmodified; graph module stood in for, harness/d3net_stub.py) with the ops in the loop."""
    from harness import d3net_stub as H
    if not H.available():
        return {"unavailable": "baseline/_ref not staged (harness/stage_ref.py needs /root/reference)"}
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = H.load_cfg(max_num_proposal=256)
    det = H.build_detector(cfg, dev)
    spk = H.build_speaker(cfg, dev)
    nb = scenes.make_batch(n_scenes, points, config_id=2)
    dd = H.collate(nb, dev)
    lang = H.speaker_inputs(nb, cfg, dev, chunk=chunk)
    out = H.run_speaker(det, spk, dd, lang)                               # warm-up
    ms = wall_ms(lambda: H.run_speaker(det, spk, dd, lang), reps)
    with H.OpTimer() as ot:
        t0 = time.perf_counter()
        out = H.run_speaker(det, spk, dd, lang)
        torch.cuda.synchronize()
        total_inst = (time.perf_counter() - t0) * 1e3
    ops_ms = sum(ot.ms.values())
    return {"value": n_scenes / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "what": "PointGroup.feed + graph stand-in + SpeakerNet forward (teacher forcing, %d descriptions per scene, 256 "
                    "proposals per scene), reference model files unmodified over d3net_b200.pointgroup_ops; wall clock" % chunk,
            "ops_ms": round(ops_ms, 2), "ops_share": round(ops_ms / total_inst, 3),
            "per_op_ms": {k: round(v, 2) for k, v in sorted(ot.ms.items(), key=lambda kv: -kv[1])},
            "captions": list(out["lang_cap"].shape), "proposals_kept_per_batch": int(out["proposal_batch_mask"].sum())}


def reference_mixed_leg(batch, rand6, reps=1):
    """What a D3Net user runs today on this GPU: the reference's own nine CUDA kernels + its two CPU ops, compiled from
    its sources (oracle/_ref), driven through the same chain."""
    from oracle.ops_adapter import RefGpuOps
    ops = RefGpuOps()
    if ops.ref is None:
        return {"unavailable": "oracle/_ref/PG_OP.so was not built"}
    # one thread, one stream: the reference's bfs_cluster keeps a 4-byte-per-point array on the stack (bfs_cluster.cpp:61),
    # more than a worker thread's stack holds at a million points
    chain.proposal_chain(ops, batch, rand6, overlap=False)
    ms = wall_ms(lambda: chain.proposal_chain(ops, batch, rand6, overlap=False), reps)
    return {"value": int(batch["n_scenes"]) / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "what": "oracle/_ref = the reference's lib/pointgroup_ops compiled unmodified for sm_100a: brute-force "
                    "ballquery_batch_p and the other eight kernels on this GPU, voxelize_idx / bfs_cluster on one CPU "
                    "thread after D2H copies, as model/pointgroup.py calls them; wall clock"}


def run_b200(args, rank, world, local, emit=print):
    import torch.distributed as dist
    from d3net_b200 import pointgroup_ops as ops, dist as pgdist, _native, collate
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    _native.lib()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / warnings belong on stderr
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    n_scenes = args.scenes                                  # scenes per chain call
    if strong:
        per_rank = STRONG_TOTAL // world
        n_scenes = min(SCENES_PER_GPU, per_rank)
        n_sub = per_rank // n_scenes
    else:
        n_sub = 1
    first_scene = rank * n_scenes * n_sub                   # scene-per-GPU sharding: each rank its own scenes
    subs = []
    for s in range(n_sub):
        nb = scenes.make_batch(n_scenes, args.points, config_id=2, first_scene=first_scene + s * n_scenes, with_feats=False)
        host = chain.batch_to_device(nb, None, pt_feat_seed=rank * 16 + s, pin=True)
        # the scaled coordinates as the dataset hands them over (fp32 [N,3], lib/dataset/pipeline.py:155); the collate
        # kernel prefixes the scene id and truncates (pipeline.py:941-942)
        host["locs_scaled_f"] = host["locs_scaled"][:, 1:].float().contiguous().pin_memory()
        host["instance_ids32"] = host["instance_ids"].int().contiguous().pin_memory()
        host["batch_offsets"] = torch.from_numpy(nb["batch_offsets"].astype(np.int32)).pin_memory()
        # the point features as the dataset hands them over: [N, 131]; PointGroup.feed appends the coordinates on the
        # device (model/pointgroup.py:469-470)
        host["feats131"] = host["feats"][:, :scenes.IN_CHANNELS - 3].contiguous().pin_memory()
        batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
        subs.append((nb, host, batch))
    nb, host, batch = subs[0]
    rand6 = torch.full((6,), 0.5, device=dev)
    # host inputs of a step: what the DataLoader produces.  pt_feats / semantic_preds / pt_offsets are outputs of the
    # backbone (model/pointgroup.py:271-281) -- born on the device in the real pipeline, resident here.
    HOST_KEYS = ("locs", "locs_scaled_f", "feats131", "instance_ids32", "instance_pointnum", "batch_offsets")
    h2d_bytes = n_sub * sum(host[k].numel() * host[k].element_size() for k in HOST_KEYS)

    def one_chain(b, timer=None, fused_glue=False, overlap=True, fused_cluster=False):
        out = chain.proposal_chain(ops, b, rand6, timer, fused_glue=fused_glue, overlap=overlap, fused_cluster=fused_cluster)
        packed = pgdist.pack_proposals(out, b, args.max_proposals)
        return out, packed

    def step_device(timer=None, fused_glue=False, overlap=True, fused_cluster=False):
        outs = [one_chain(b, timer, fused_glue, overlap, fused_cluster) for _, _, b in subs]
        packed = outs[0][1] if n_sub == 1 else torch.cat([p for _, p in outs], 0)
        gathered = pgdist.all_gather_proposals(packed)
        return outs[0][0], gathered

    d2h_keep = {}

    # End to end: every step's inputs start in pinned HOST memory.  The upload of the next batch runs on a copy
    # stream while the current one computes (what a serving loop does); every byte of every upload is inside the
    # timed region, and each step ends with the device->host read of its results.
    copy_stream = torch.cuda.Stream(device=dev)

    def upload(h, resident):
        with torch.cuda.stream(copy_stream):
            b = {k: h[k].to(dev, non_blocking=True) for k in HOST_KEYS}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        b.update({k: resident[k] for k in ("pt_feats", "semantic_preds", "pt_offsets")})
        b["n_scenes"] = resident["n_scenes"]
        return b, ev

    e2e_state = {"queue": [], "left": 0}

    def e2e_refill():
        while len(e2e_state["queue"]) < 2 and e2e_state["left"] > 0:
            k = e2e_state["issued"] % n_sub
            e2e_state["queue"].append(upload(subs[k][1], subs[k][2]))
            e2e_state["issued"] += 1
            e2e_state["left"] -= 1

    def step_e2e():
        packs, first = [], None
        for _ in range(n_sub):
            e2e_refill()
            b, ev = e2e_state["queue"].pop(0)
            e2e_refill()
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for v in b.values():
                if torch.is_tensor(v):
                    v.record_stream(cur)
            # collate on the device (pg_collate_points): scene column + truncation, instance ids widened
            locs_scaled, _, inst64 = collate.collate_points(b["locs_scaled_f"], b["batch_offsets"], None, b["instance_ids32"],
                                                            torch.zeros(int(b["n_scenes"]) + 1, dtype=torch.int32, device=dev))
            b["locs_scaled"], b["instance_ids"] = locs_scaled, inst64
            b["feats"] = torch.cat((b["feats131"], b["locs"]), 1)          # model/pointgroup.py:469-470
            out, packed = one_chain(b)
            packs.append(packed)
            first = first or out
        packed = packs[0] if n_sub == 1 else torch.cat(packs, 0)
        gathered = pgdist.all_gather_proposals(packed)
        res = [gathered, first["ious"], first["proposals_offset"]]
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
    # timed region 1: device-resident inputs.  The per-op CUDA-event timers run in a separate pass below so that their
    # ~60 event records per step stay out of `value`.
    ms_total = timed(step_device, args.steps)
    # timed region 2: end to end from pinned host memory
    # warm-up of its own: the uploads allocate on the copy stream and the chain on its side streams, and a caching-allocator
    # miss inside the timed region (cudaMalloc, or worse a cudaFree round) stalls the whole pipeline once
    n_e2e_warm = max(args.warmup, 3) + 2
    e2e_state.update(left=n_e2e_warm * n_sub, issued=0)
    for _ in range(n_e2e_warm):
        step_e2e()
    e2e_state.update(left=args.steps * n_sub, issued=0)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d_rank_gbs, h2d_sum_gbs = h2d_probe(dev, world)

    extras = world == 1 and not args.no_extras
    timer = chain.SectionTimer(True)
    ms_fused = None
    ktimes = {}
    if rank == 0 or world > 1:
        # per-op breakdown (CUDA events around every op call), on the single-stream schedule; two untimed steps first:
        # the caching allocator keeps one pool per stream, and `value` ran the clusterings on side streams
        for _ in range(2):
            step_device(overlap=False)
        ops._section_timer = timer
        timed(lambda: step_device(timer), args.steps)
        ops._section_timer = None
    if extras or world > 1:
        # timed region 2b: the caller's clusters_voxelization glue replaced by the fused op (an
        # edit to the caller, so it is reported next to `value`, not as `value`)
        for _ in range(2):
            step_device(fused_glue=True)
        ms_fused = timed(lambda: step_device(fused_glue=True), args.steps)
    ms_fc = ms_all = None
    if extras or world > 1 or args.overlap_variant:
        for _ in range(3):
            step_device(fused_cluster=True)
        ms_fc = timed(lambda: step_device(fused_cluster=True), args.steps)
        for _ in range(2):
            step_device(fused_cluster=True, fused_glue=True)
        ms_all = timed(lambda: step_device(fused_cluster=True, fused_glue=True), args.steps)
    ms_overlap = None
    if extras or world > 1 or args.overlap_variant:
        for _ in range(3):
            step_device(overlap=False)
        ms_overlap = timed(lambda: step_device(overlap=False), args.steps)
    # timed region 3: the same steps with the library's per-kernel CUDA-event timers on (events on the
    # launching stream around every main kernel; include/pg_b200.h, pg_kernel_timing)
    _native.kernel_timing(True)
    barrier()
    for _ in range(args.steps):
        step_device(overlap=False)               # one stream: a kernel's events then bracket that kernel alone
    torch.cuda.synchronize()
    ktimes = _native.kernel_timing_report()
    _native.kernel_timing(False)

    # kernels of this library launched per step (CUPTI count of one untimed step; every rank takes part
    # because the step contains the collective)
    n_launch, top, n_all = count_pg_kernels(lambda: step_device())

    scenes_per_step = n_scenes * n_sub * world
    value = scenes_per_step * args.steps / (ms_total / 1e3)
    e2e_value = scenes_per_step * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        sec_ms = {k: v / args.steps for k, v in timer.totals_ms().items()}
        algo = algorithmic_bytes(batch, out)
        peak, peak_kind = peaks()
        kalgo = {k: v * n_sub for k, v in kernel_algorithmic_bytes(batch, out).items()}
        # measured DRAM traffic per kernel comes from the committed ncu capture of THIS workload only (config 1)
        traffic_db = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if args.config == 1 and not strong and args.points == POINTS and n_scenes == SCENES_PER_GPU and os.path.exists(tpath):
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
            if name in COMPUTE_BOUND:
                row["bound"] = "fp32-issue (not DRAM): see roofline.note"
            if name in traffic_db:
                # what the kernel really moved (ncu dram__bytes, committed capture) over the time measured now
                row["dram_bytes_per_step"] = traffic_db[name]
                row["dram_GBps"] = round(traffic_db[name] / (ms_step / 1e3) / 1e9, 1) if ms_step > 0 else None
                row["dram_frac"] = round(row["dram_GBps"] / peak, 4) if row["dram_GBps"] else None
            if name in KERNEL_NOTES:
                row["note"] = KERNEL_NOTES[name]
            per_kernel[name] = row
        dom = next(iter(per_kernel))
        dom_ms = per_kernel[dom]["ms_per_step"]
        dom_launches = max(per_kernel[dom]["launches_per_step"], 1)
        achieved = per_kernel[dom].get("GBps", 0.0)
        per_op = {}
        for k, v in sorted(sec_ms.items(), key=lambda kv: -kv[1]):
            v1 = v / n_sub
            gbs = algo[k] / (v1 / 1e3) / 1e9 if k in algo and v1 > 0 else None
            per_op[k] = {"ms": round(v, 4), "GBps": round(gbs, 1) if gbs else None, "frac": round(gbs / peak, 4) if gbs else None}
        if extras:
            v2p_scene = ops.voxelization_idx(batch["locs_scaled"], n_scenes, 4)[2]
            per_op.update(backward_rows(ops, batch, out, v2p_scene, peak))
        chain_bytes = sum(algo[k] for k in ("voxelization(scene)", "ballquery(shift)", "ballquery(raw)", "bfs_cluster(shift)",
                                            "bfs_cluster(raw)", "roipool", "get_iou"))
        cpu = unchanged = mixed = None
        if extras:
            if not args.no_cpu_baseline:
                rate, ms, desc, cores, kind, detail = cpu_chain_rate(20.0, 1, 1, args.points, dense_room=args.config == 3)
                cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "kind_detail": detail, "sample": desc}
            if args.config == 1:
                try:
                    mixed = reference_mixed_leg(batch, rand6)
                except Exception as e:      # noqa: BLE001  (a baseline leg must not take the line down)
                    mixed = {"unavailable": repr(e)[:300]}
                try:
                    unchanged = unchanged_caller_leg(n_scenes, args.points)
                except Exception as e:      # noqa: BLE001
                    unchanged = {"unavailable": repr(e)[:300]}
        if args.config == 2 or strong:
            workload = WORKLOADS[2] % (n_scenes * n_sub, n_scenes, args.points // 1000)
        elif args.config == 3:
            workload = WORKLOADS[3] % (n_scenes, args.points // 1000)
        else:
            workload = WORKLOADS[1] % (n_scenes, args.points // 1000)
            if world > 1:
                workload = ("configs[2], weak-scaling form (8 scenes per GPU at every N; `--scaling strong` runs the fixed "
                            "64-scene batch): ") + workload.split(": ", 1)[1] + ", NCCL all-gather of the packed proposal block"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "points_per_sec": value * args.points,
            "config": {"workload": workload,
                       "scenes_per_gpu_per_step": n_scenes * n_sub, "scenes_per_chain_call": n_scenes,
                       "points_per_scene": args.points, "parallelism": "scene-per-GPU x%d" % world,
                       "cluster_radius": scenes.CLUSTER_RADIUS, "npoint_thre": scenes.CLUSTER_NPOINT_THRE,
                       "l2": "inputs larger than L2 (%.0f MB of point features per chain call)" % (batch["feats"].numel() * 4 / 1e6),
                       "collective": "all_gather of [scenes, %d, 46] fp32 proposal block" % args.max_proposals if world > 1 else "none (N=1)",
                       "n_points": int(batch["locs"].shape[0]), "n_object_points": int(out["n_object_points"]),
                       "nActive_shift": int(out["nActive_shift"]), "nActive_raw": int(out["nActive_raw"]),
                       "n_proposals": int(out["proposals_offset"].numel() - 1), "sumNPoint": int(out["proposals_idx"].shape[0]),
                       "cpu_binding": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "H2D of the next batch on a copy stream overlaps the current one; all uploads inside the timed region",
                    "host_inputs": "locs, scaled coords (fp32), feats [N,131] (+ locs appended on the device as PointGroup.feed does), instance ids, instance_pointnum, batch offsets; "
                                   "collated on the device (pg_collate_points).  pt_feats / semantic_preds / pt_offsets are "
                                   "backbone OUTPUTS (model/pointgroup.py:271-281): device-resident, not uploaded",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_keep.get("bytes", 0)),
                    "h2d_GBps_per_rank": round(h2d_bytes / (ms_e2e / args.steps / 1e3) / 1e9, 1),
                    "h2d_ceiling_gbs": {"per_rank_min": round(h2d_rank_gbs, 1), "all_ranks": round(h2d_sum_gbs, 1),
                                        "how": "bare 512 MB pinned->device copies issued by every rank at once"},
                    "frac_of_h2d_ceiling": round(h2d_bytes / (ms_e2e / args.steps / 1e3) / 1e9 / max(h2d_rank_gbs, 1e-9), 3)},
            "gpu_launches": (n_launch * args.steps) if n_launch is not None else None,
            "gpu_launches_per_step": n_launch,
            "all_kernel_launches_per_step": n_all,
            "chain": {"algorithmic_bytes_per_step": int(chain_bytes * n_sub),
                      "GBps": round(chain_bytes * n_sub / (ms_total / args.steps / 1e3) / 1e9, 1),
                      "frac_of_hbm_peak": round(chain_bytes * n_sub / (ms_total / args.steps / 1e3) / 1e9 / peak, 4),
                      "what": "SURVEY 8(d) compulsory bytes of the ops of one step / ms_per_step"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (traffic_db[dom] / dom_launches) if dom in traffic_db else None,
                         "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else peak_kind,
                         "algorithmic_bytes_per_launch": int(kalgo.get(dom, 0) / dom_launches),
                         "ms_per_launch": dom_ms / dom_launches, "launches_per_step": dom_launches,
                         "timing": "CUDA events on the launching stream around every launch of the kernel, %d steps" % args.steps,
                         "binding_limit": "fp32-issue" if dom in COMPUTE_BOUND else "hbm",
                         "note": KERNEL_NOTES.get(dom, "")},
            # the kernels that stream their operands once (the ones an HBM roofline is the right yardstick for)
            "hbm_kernels": {k: {"GBps": v["GBps"], "frac": v["frac"]} for k, v in per_kernel.items()
                            if k in ("k_voxelize_fp", "k_vox_fill", "k_bq_fill_mask", "k_gather_rows") and "GBps" in v},
            "per_kernel": per_kernel,
            "per_op": per_op,
            "top_kernels": [{"name": k[:80], "calls": c, "us": round(t, 1)} for k, c, t in top[:8]],
            "clocks": clocks,
        }
        if ms_fused is not None:
            line["fused_glue"] = {"value": scenes_per_step * args.steps / (ms_fused / 1e3), "unit": UNIT,
                                  "ms_per_step": ms_fused / args.steps,
                                  "what": "same chain with the caller-side clusters_voxelization glue (model/pointgroup.py:125-167) "
                                          "done by pointgroup_ops.cluster_voxel_coords; bit-identical outputs"}
        if ms_fc is not None:
            line["fused_cluster"] = {"value": scenes_per_step * args.steps / (ms_fc / 1e3), "unit": UNIT, "ms_per_step": ms_fc / args.steps,
                                     "what": "same chain with each ballquery_batch_p + bfs_cluster pair (model/pointgroup.py:296-297, "
                                             ":304-305) as one op, pointgroup_ops.ballquery_bfs_cluster: neighbour lists are only "
                                             "decoded where the clustering sweep reads them; identical clusters"}
            line["fused_cluster_and_glue"] = {"value": scenes_per_step * args.steps / (ms_all / 1e3), "unit": UNIT,
                                              "ms_per_step": ms_all / args.steps, "what": "both caller edits together"}
        if ms_overlap is not None:
            line["single_stream"] = {"value": scenes_per_step * args.steps / (ms_overlap / 1e3), "unit": UNIT,
                                     "ms_per_step": ms_overlap / args.steps,
                                     "what": "same chain with its three independent parts -- the input cloud's voxelisation "
                                             "(pipeline.py:992) and the two clusterings (model/pointgroup.py:296-298, :304-306) "
                                             "-- issued one after the other on one stream; `value` issues them from three "
                                             "host threads on three CUDA streams (identical calls and outputs; the per_op / "
                                             "per_kernel tables are timed on the single-stream schedule)"}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if unchanged is not None:
            line["unchanged_caller"] = unchanged
        if mixed is not None:
            line["reference_mixed"] = mixed
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config4(args, rank, world, local, emit=print):
    """configs[4], detector half: the reference's caller with the ops in the loop."""
    if rank != 0:
        return
    from d3net_b200 import _native
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    _native.lib()
    torch.cuda.set_device(local)
    sampler = ClockSampler(local)
    sampler.start()
    det = unchanged_caller_leg(args.scenes, args.points, reps=max(args.steps, 1))
    r = speaker_leg(args.scenes, args.points, reps=max(args.steps, 1))
    clocks = sampler.stop()
    if "unavailable" in r:
        emit(json.dumps({"metric": METRIC, "unavailable": r["unavailable"], "config": {"workload": WORKLOADS[4] % (args.scenes, args.points // 1000)}}))
        return
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": max(args.steps, 1), "warmup": 1,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[4] % (args.scenes, args.points // 1000), "max_num_proposal": 256},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "wall clock of the whole caller; the batch is device-resident as Lightning leaves it"},
            "ops_share": r["ops_share"], "speaker_forward": r, "detector_only": det, "clocks": clocks, "gpu_launches": None}
    emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="index into BASELINE.json configs")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--scenes", type=int, default=None, help="scenes per GPU per chain call")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--max-proposals", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the N=1 side legs (fused glue, backward rows, unchanged caller, ...)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="--impl reference: CPU seconds the whole run may take (sizes the sample of the workload)")
    ap.add_argument("--overlap-variant", action="store_true", help="also time the single-stream variant with --no-extras")
    ap.add_argument("--profile-mode", action="store_true",
                    help="only W warm-up + K device-resident steps, no JSON line (for runs under ncu)")
    args = ap.parse_args()
    if args.scaling is None:
        args.scaling = "strong" if args.config == 2 else "weak"
    if args.scaling == "strong":
        args.config = 2
    if args.points is None:
        args.points = 1_000_000 if args.config == 3 else POINTS
    if args.scenes is None:
        args.scenes = 1 if args.config == 3 else SCENES_PER_GPU
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
        elif args.config == 4:
            run_config4(args, rank, world, local, emit)
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
