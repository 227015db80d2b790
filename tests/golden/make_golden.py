#!/usr/bin/env python
"""Generates the committed golden vectors from the REFERENCE's own compiled ops (oracle/_ref/PG_OP.so,
built by oracle/build_ref.py from the unmodified sources under /root/reference).

    python tests/golden/make_golden.py cpu     # this container: voxelize_idx, bfs_cluster (the reference's CPU ops)
    python tests/golden/make_golden.py gpu OUT # on the B200 box (gpurun): the nine CUDA kernels -> OUT/ref_gpu.npz
    python tests/golden/make_golden.py collate # this container: the reference's sparse_collate_fn -> ref_collate.npz
    python tests/golden/make_golden.py nms     # this container: instance NMS (reference Python, torch CPU) -> ref_nms.npz

Inputs are NOT stored: golden_inputs(case) regenerates them from fixed seeds, so the fixtures stay small.
The reference ships no tests or vectors of its own; these files are what pins the
oracle (tests/test_oracle.py) and, through it, the CUDA kernels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def golden_inputs(case):
    """Deterministic inputs shared by the generator and the tests."""
    rng = np.random.default_rng(20261017 + sum(map(ord, case)))
    if case == "voxelize_idx":
        n = 4000
        coords = np.column_stack([rng.integers(0, 3, n), rng.integers(0, 11, (n, 3))]).astype(np.int64)
        coords[::97, 1] += 2 ** 32                       # narrowing to int32 (datatype.h:9)
        return {"coords": coords, "batchsize": 3}
    if case == "graph":                                   # a small ball-query product incl. truncated lists
        blob = rng.normal(0, 0.004, (2900, 3))
        rest = rng.uniform(-0.4, 0.4, (1500, 3))
        xyz = np.concatenate([blob, rest]).astype(np.float32)
        xyz = xyz[rng.permutation(len(xyz))]
        n = len(xyz)
        bi = (np.arange(n) >= 2600).astype(np.int32)
        bo = np.array([0, 2600, n], np.int32)
        sem = rng.integers(1, 3, n).astype(np.int32)
        return {"xyz": xyz, "batch_idxs": bi, "batch_offsets": bo, "radius": 0.03, "sem": sem}
    if case == "feats":
        n, C = 3000, 7
        coords = np.column_stack([rng.integers(0, 2, n), rng.integers(0, 8, (n, 3))]).astype(np.int64)
        return {"coords": coords, "feats": rng.standard_normal((n, C)).astype(np.float32),
                "grad": rng.standard_normal((n, C)).astype(np.float32)}
    if case == "segments":
        lens = rng.integers(0, 120, 60)
        lens[7] = 5000
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        x = rng.standard_normal((int(off[-1]), 16)).astype(np.float32)
        x[rng.random(x.shape) < 0.03] = 0.5
        N, nI = 20000, 23
        labels = rng.integers(-1, nI, N).astype(np.int64)
        return {"offsets": off, "x": x, "x3": np.ascontiguousarray(x[:, :3]),
                "grad": rng.standard_normal((len(lens), 16)).astype(np.float32),
                "pidx": rng.integers(0, N, int(off[-1])).astype(np.int32), "labels": labels,
                "pointnum": np.bincount(labels[labels >= 0], minlength=nI).astype(np.int32)}
    if case == "nms":
        # two overlapping partitions of a point set (what the raw and the shifted clustering produce), a few
        # repeated rows, one proposal without points, distinct scores
        N, nA, nB = 6000, 23, 19
        la = rng.integers(-1, nA, N)
        stick = (np.maximum(la, 0) + 1) / nA              # per-cluster overlap: IoUs from ~0.03 to ~0.9
        lb = np.where(rng.random(N) < stick, la % nB, rng.integers(-1, nB, N))
        rows = [(a, p) for p, a in enumerate(la) if a >= 0] + [(nA + b, p) for p, b in enumerate(lb) if b >= 0]
        rows += rows[:50]
        pidx = np.array(rows, np.int32)
        pidx = pidx[np.argsort(pidx[:, 0], kind="stable")]
        nP = nA + nB + 1                                  # the last proposal has no rows
        scores = rng.permutation(nP).astype(np.float32) / nP
        return {"proposals_idx": pidx, "num_proposals": nP, "N": N, "scores": scores, "threshold": 0.3}
    if case == "collate":
        # three scenes as PipelineDataset.__getitem__ hands them to the collate function (pipeline.py:180-187)
        batch = []
        for n, ninst in ((1500, 7), (2200, 11), (900, 3)):
            locs = rng.uniform(0, 0.45, (n, 3)).astype(np.float32)     # ~22^3 voxels: points share voxels
            scaled = ((locs - locs.min(0)) * 50).astype(np.float32)
            inst = rng.integers(-1, ninst, n).astype(np.int32)
            batch.append({"locs": locs, "locs_scaled": scaled, "feats": rng.standard_normal((n, 3)).astype(np.float32),
                          "sem_labels": rng.integers(-1, 20, n).astype(np.int32), "instance_ids": inst,
                          "num_instance": np.array(ninst).astype(np.int32),
                          "instance_info": rng.standard_normal((n, 12)).astype(np.float32),
                          "instance_num_point": np.bincount(inst[inst >= 0], minlength=ninst).astype(np.int32),
                          "scene_id": "scene%04d_00" % n})
        return {"batch": batch}
    raise KeyError(case)


def _load_ref():
    from oracle import build_ref
    build_ref.build(verbose=False)
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/PG_OP.so is not built (needs /root/reference)"
    return ref


def make_cpu():
    import torch
    ref = _load_ref()
    from oracle import pg_oracle as o
    out = {}
    g = golden_inputs("voxelize_idx")
    for mode in (4, 1, 2):
        oc, im, om = torch.zeros(0, dtype=torch.int64), torch.zeros(len(g["coords"]), dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
        ref.voxelize_idx(torch.from_numpy(g["coords"]), oc, im, om, g["batchsize"], mode)
        out["vox_oc_m%d" % mode], out["vox_im_m%d" % mode], out["vox_om_m%d" % mode] = oc.numpy(), im.numpy(), om.numpy()
    g = golden_inputs("graph")
    # neighbour lists: the oracle's restatement (the reference ball query is CUDA-only); the golden is the
    # reference's CPU BFS on them, for two thresholds
    idx, sl = o.ballquery_batch_p(g["xyz"], g["batch_idxs"], g["batch_offsets"], g["radius"])
    assert (sl[:, 1] == 1000).any()
    for thr in (5, 50):
        ci, co = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
        ref.bfs_cluster(torch.from_numpy(g["sem"]), torch.from_numpy(idx), torch.from_numpy(sl), ci, co, len(sl), thr)
        out["bfs_ci_t%d" % thr], out["bfs_co_t%d" % thr] = ci.numpy(), co.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_cpu.npz"), **out)
    print("wrote ref_cpu.npz:", {k: v.shape for k, v in out.items()})


def _reference_function(path, name):
    """The reference's own source of one top-level function, compiled in place (the module around it imports
    packages that are not installed here)."""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def make_nms():
    """cross_ious: the statements of model/pointgroup.py:577-590 executed by torch on the CPU; picks: the
    reference's get_nms_instances (lib/utils/eval.py:75-97) run on that matrix."""
    import torch
    g = golden_inputs("nms")
    pidx = torch.from_numpy(g["proposals_idx"])
    nP, N = g["num_proposals"], g["N"]
    mask = torch.zeros((nP, N), dtype=torch.int)
    mask[pidx[:, 0].long(), pidx[:, 1].long()] = 1
    npoint_int = mask.sum(1)
    mf = mask.float()
    inter = torch.mm(mf, mf.t())
    npoint = mf.sum(1)
    h = npoint.unsqueeze(-1).repeat(1, nP)
    v = npoint.unsqueeze(0).repeat(nP, 1)
    cross = inter / (h + v - inter)
    nms = _reference_function("/root/reference/lib/utils/eval.py", "get_nms_instances")
    out = {"cross_ious": cross.numpy(), "npoint": npoint_int.numpy().astype(np.int32)}
    for thr in (0.3, 0.1, 0.6):
        out["pick_%g" % thr] = nms(cross.numpy(), g["scores"], thr)
    np.savez_compressed(os.path.join(HERE, "ref_nms.npz"), **out)
    print("wrote ref_nms.npz:", {k: v.shape for k, v in out.items()})


def make_collate():
    """The reference's own sparse_collate_fn / scannet_collate_fn (lib/dataset/pipeline.py:892-995) executed in
    place on the golden scenes, with its voxelization_idx bound to the reference's compiled op (oracle/_ref)."""
    import copy
    import torch
    from oracle.ops_adapter import OracleOps
    ops = OracleOps(use_ref=True)
    assert ops.ref is not None, "oracle/_ref/PG_OP.so is not built (needs /root/reference)"
    path = "/root/reference/lib/dataset/pipeline.py"
    ns = {"np": np, "torch": torch, "pointgroup_ops": ops, "batched_coordinates": None}
    import ast
    tree = ast.parse(open(path).read())
    for name in ("scannet_collate_fn", "sparse_collate_fn"):
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
        exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    g = golden_inputs("collate")
    data = ns["sparse_collate_fn"](copy.deepcopy(g["batch"]))           # the reference shifts instance ids in place
    out = {k: v.numpy() for k, v in data.items() if torch.is_tensor(v)}
    cat = lambda k: np.concatenate([b[k] for b in g["batch"]], 0)
    for k in ("locs", "feats", "instance_info"):       # pure concatenations: checked here, not stored
        assert out[k].dtype == np.float32 and np.array_equal(out.pop(k), cat(k))
    np.savez_compressed(os.path.join(HERE, "ref_collate.npz"), **out)
    print("wrote ref_collate.npz:", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


def make_gpu(outdir):
    import torch
    ref = _load_ref()
    from oracle import pg_oracle as o
    assert torch.cuda.is_available()
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = {}
    g = golden_inputs("graph")
    n = len(g["xyz"])
    ma = 400
    while True:
        idx = torch.zeros(n * ma, dtype=torch.int32, device="cuda")
        sl = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
        tot = ref.ballquery_batch_p(cu(g["xyz"]), cu(g["batch_idxs"]), cu(g["batch_offsets"]), idx, sl, n, ma, g["radius"])
        torch.cuda.synchronize()
        if tot <= n * ma:
            break
        ma = tot // n + 1
    flat, lens = o.canonical_neighbours(idx[:tot].cpu().numpy(), sl.cpu().numpy())     # atomicAdd placement is not defined
    out["bq_flat"], out["bq_lens"] = flat, lens
    g = golden_inputs("feats")
    oc, im, om = torch.zeros(0, dtype=torch.int64), torch.zeros(len(g["coords"]), dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
    ref.voxelize_idx(torch.from_numpy(g["coords"]), oc, im, om, 2, 4)
    M, W = om.shape
    C = g["feats"].shape[1]
    for mode in (4, 3):
        v = torch.zeros((M, C), device="cuda")
        ref.voxelize_fp(cu(g["feats"]), v, om.cuda(), mode, M, W - 1, C)
        d = torch.zeros((len(g["coords"]), C), device="cuda")
        ref.voxelize_bp(cu(g["grad"][:M]), d, om.cuda(), mode, M, W - 1, C)
        torch.cuda.synchronize()
        out["vox_fp_m%d" % mode], out["vox_bp_m%d" % mode] = v.cpu().numpy(), d.cpu().numpy()
    g = golden_inputs("segments")
    nP = len(g["offsets"]) - 1
    off = cu(g["offsets"])
    r = torch.zeros((nP, 16), device="cuda")
    a = torch.zeros((nP, 16), dtype=torch.int32, device="cuda")
    ref.roipool_fp(cu(g["x"]), off, r, a, nP, 16)
    d = torch.zeros(g["x"].shape, device="cuda")
    ref.roipool_bp(d, off, a, cu(g["grad"]), nP, 16)
    torch.cuda.synchronize()
    out["roi_out"], out["roi_arg"], out["roi_bp"] = r.cpu().numpy(), a.cpu().numpy(), d.cpu().numpy()
    for name in ("sec_mean", "sec_min", "sec_max"):
        for key, C in (("x", 16), ("x3", 3)):
            r = torch.zeros((nP, C), device="cuda")
            getattr(ref, name)(cu(g[key]), off, r, nP, C)
            torch.cuda.synchronize()
            out["%s_%s" % (name, key)] = r.cpu().numpy()
    nI = len(g["pointnum"])
    r = torch.zeros((nP, nI), device="cuda")
    ref.get_iou(cu(g["pidx"]), off, cu(g["labels"]), cu(g["pointnum"]), r, nI, nP)
    torch.cuda.synchronize()
    out["iou"] = r.cpu().numpy()
    os.makedirs(outdir, exist_ok=True)
    np.savez_compressed(os.path.join(outdir, "ref_gpu.npz"), **out)
    print("wrote ref_gpu.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gpu":
        make_gpu(sys.argv[2] if len(sys.argv) > 2 else HERE)
    elif len(sys.argv) > 1 and sys.argv[1] == "nms":
        make_nms()
    elif len(sys.argv) > 1 and sys.argv[1] == "collate":
        make_collate()
    else:
        make_cpu()
