"""Synthetic ScanNet-shaped scenes for the parity tests and bench.py.

Host-side numpy only.  There is no dataset in this environment, so every workload is generated:
a partial indoor scan made of surface samples -- a floor patch (semantic label 1), short wall strips
(label 0) and 30-60 box / cylinder "objects" (labels 2..19, one instance id each).  Points sit on a
jittered grid with ~2 cm pitch, like ScanNet mesh vertices, so that a 3 cm ball query sees ~8
neighbours and surfaces stay connected (uniform random sampling at the same density does not
percolate at r = 0.03).  The fields mirror what the reference's collate function hands the detector
(lib/dataset/pipeline.py:937-992): ``locs`` fp32 xyz, ``locs_scaled`` = floor((xyz - min) * 50) with
the batch index in column 0, ``instance_ids`` (-1 = ignore, batch-global ids, pipeline.py:962,980),
``instance_pointnum``, ``sem_labels`` and ``batch_offsets``; plus stand-ins for the two network
outputs the proposal path consumes: ``semantic_preds`` (labels with 5 % noise) and ``pt_offsets``
(0.9 * (instance centroid - xyz) + N(0, 0.01), model/pointgroup.py:280-296).

Seeds: ``seed = 1000 * config_id + scene_index`` (numpy default_rng).
"""
import numpy as np

SCALE = 50                 # conf/pointgroup.yaml:28  -> 2 cm voxels
CLUSTER_RADIUS = 0.03      # conf/pointgroup.yaml:157
CLUSTER_MEANACTIVE = 50
CLUSTER_SHIFT_MEANACTIVE = 300
CLUSTER_NPOINT_THRE = 50
SCORE_SCALE = 50           # conf/pointgroup.yaml:130-132
SCORE_FULLSCALE = 14
SCORE_MODE = 4
IN_CHANNELS = 134          # 128 multiview + 3 normal + 3 coords (model/pointgroup.py:38)
M_CHANNELS = 16            # conf/pointgroup.yaml:58

_PITCH0 = 0.02             # grid pitch at the 150k-point geometry
_GEOM_N = 150_000


def _grid_patch(rng, w, h, pitch):
    """Jittered grid on a w x h rectangle -> (k, 2) local coordinates."""
    nx = max(int(round(w / pitch)), 1)
    ny = max(int(round(h / pitch)), 1)
    u, v = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    p = np.stack([u.ravel(), v.ravel()], 1).astype(np.float64) + 0.5
    p += rng.uniform(-0.2, 0.2, p.shape)
    return p * np.array([w / nx, h / ny])


def make_scene(n_points=150_000, seed=0, geometry_points=_GEOM_N):
    """One scene with exactly ``n_points`` points.  The room geometry is fixed by
    ``geometry_points`` (150k -> ~60 m^2 of surface); a larger ``n_points`` densifies the same room
    (1M points -> ~7.7 mm pitch, ~50 raw neighbours, config 4)."""
    rng = np.random.default_rng(seed)
    area = geometry_points * _PITCH0 ** 2
    pitch = _PITCH0 * np.sqrt(geometry_points / float(n_points)) * 0.985   # ~3 % surplus, trimmed below
    fx = np.sqrt(0.30 * area * 4.0 / 3.0)
    fy = 0.30 * area / fx
    wall_h = 0.15 * area / (2 * (fx + fy))

    xyz, sem, inst = [], [], []

    def add(p, s, i):
        xyz.append(p)
        sem.append(np.full(len(p), s, np.int32))
        inst.append(np.full(len(p), i, np.int64))

    # floor (label 1, no instance)
    uv = _grid_patch(rng, fx, fy, pitch)
    add(np.column_stack([uv, np.zeros(len(uv))]), 1, -1)
    # walls (label 0, no instance)
    for axis, length, fixed in ((0, fx, 0.0), (0, fx, fy), (1, fy, 0.0), (1, fy, fx)):
        uv = _grid_patch(rng, length, wall_h, pitch)
        p = np.zeros((len(uv), 3))
        p[:, axis] = uv[:, 0]
        p[:, 1 - axis] = fixed
        p[:, 2] = uv[:, 1]
        add(p, 0, -1)
    # objects until the point budget is met
    n_obj_target = int(rng.integers(30, 61))
    obj_area = 0.55 * area / n_obj_target
    budget = int(n_points * 1.03)
    k = 0
    while sum(len(p) for p in xyz) < budget:
        label = int(rng.integers(2, 20))
        cx, cy = rng.uniform(0.3, fx - 0.3), rng.uniform(0.3, fy - 0.3)
        a = obj_area * rng.uniform(0.6, 1.4)
        if rng.random() < 0.6:      # box: top + 4 sides, footprint w x d, height h
            w = np.sqrt(a / 4.5) * rng.uniform(0.8, 1.25)
            d = a / 4.5 / w
            h = (a - w * d) / (2 * (w + d))
            parts = []
            uv = _grid_patch(rng, w, d, pitch)
            parts.append(np.column_stack([uv[:, 0] - w / 2, uv[:, 1] - d / 2, np.full(len(uv), h)]))
            for sx, sy, ln in ((0, -d / 2, w), (0, d / 2, w), (-w / 2, 0, d), (w / 2, 0, d)):
                uv = _grid_patch(rng, ln, h, pitch)
                if sx == 0:
                    parts.append(np.column_stack([uv[:, 0] - ln / 2, np.full(len(uv), sy), uv[:, 1]]))
                else:
                    parts.append(np.column_stack([np.full(len(uv), sx), uv[:, 0] - ln / 2, uv[:, 1]]))
            p = np.concatenate(parts)
            th = rng.uniform(0, np.pi)
            rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            p[:, :2] = p[:, :2] @ rot.T
        else:                        # cylinder: side + top disc
            r = np.sqrt(a / (3 * np.pi)) * rng.uniform(0.8, 1.2)
            h = max((a - np.pi * r * r) / (2 * np.pi * r), 2 * pitch)
            uv = _grid_patch(rng, 2 * np.pi * r, h, pitch)
            side = np.column_stack([r * np.cos(uv[:, 0] / r), r * np.sin(uv[:, 0] / r), uv[:, 1]])
            uv = _grid_patch(rng, 2 * r, 2 * r, pitch) - r
            uv = uv[(uv ** 2).sum(1) <= r * r]
            p = np.concatenate([side, np.column_stack([uv, np.full(len(uv), h)])])
        p[:, 0] += cx
        p[:, 1] += cy
        add(p, label, k)
        k += 1

    xyz = np.concatenate(xyz)
    sem = np.concatenate(sem)
    inst = np.concatenate(inst)
    perm = rng.permutation(len(xyz))[:n_points]      # shuffle + trim to exactly n_points
    if len(perm) < n_points:                          # (cannot happen with the 3 % surplus; be safe)
        perm = np.concatenate([perm, rng.integers(0, len(xyz), n_points - len(perm))])
    xyz, sem, inst = xyz[perm], sem[perm], inst[perm]
    xyz = xyz + rng.normal(0, 0.0005, xyz.shape)      # sensor noise, breaks exact ties
    # compact instance ids (an object may have lost all its points to the trim -- practically never)
    ids = np.unique(inst[inst >= 0])
    remap = -np.ones(k + 1, np.int64)
    remap[ids] = np.arange(len(ids))
    inst = np.where(inst >= 0, remap[np.maximum(inst, 0)], -1)
    n_inst = len(ids)

    xyz32 = xyz.astype(np.float32)
    pointnum = np.bincount(inst[inst >= 0], minlength=n_inst).astype(np.int32)
    cent = np.zeros((n_inst, 3))
    for d in range(3):
        cent[:, d] = np.bincount(inst[inst >= 0], weights=xyz[inst >= 0, d], minlength=n_inst) / np.maximum(pointnum, 1)
    pt_offsets = rng.normal(0, 0.01, xyz.shape)
    has = inst >= 0
    pt_offsets[has] += 0.9 * (cent[inst[has]] - xyz[has])
    noisy = rng.random(n_points) < 0.05
    semantic_preds = np.where(noisy, rng.integers(0, 20, n_points), sem).astype(np.int64)
    return {
        "locs": xyz32,
        "sem_labels": sem,
        "instance_ids": inst.astype(np.int64),
        "instance_pointnum": pointnum,
        "pt_offsets": pt_offsets.astype(np.float32),
        "semantic_preds": semantic_preds,
    }


def make_batch(n_scenes, n_points=150_000, config_id=2, first_scene=0, with_feats=False, feat_seed=None,
               geometry_points=_GEOM_N):
    """Collate ``n_scenes`` scenes the way sparse_collate_fn does (lib/dataset/pipeline.py:937-992):
    stacked points, batch index column, batch-global instance ids."""
    scenes = [make_scene(n_points, 1000 * config_id + first_scene + i, geometry_points) for i in range(n_scenes)]
    locs = np.concatenate([s["locs"] for s in scenes])
    batch_idx = np.concatenate([np.full(len(s["locs"]), i, np.int64) for i, s in enumerate(scenes)])
    inst, off = [], 0
    for s in scenes:
        ii = s["instance_ids"].copy()
        ii[ii >= 0] += off
        off += len(s["instance_pointnum"])
        inst.append(ii)
    scaled = np.concatenate([np.floor((s["locs"] - s["locs"].min(0)) * SCALE).astype(np.int64) for s in scenes])
    batch = {
        "locs": locs,
        "locs_scaled": np.column_stack([batch_idx, scaled]).astype(np.int64),
        "batch_offsets": np.concatenate([[0], np.cumsum([len(s["locs"]) for s in scenes])]).astype(np.int32),
        "sem_labels": np.concatenate([s["sem_labels"] for s in scenes]),
        "instance_ids": np.concatenate(inst),
        "instance_pointnum": np.concatenate([s["instance_pointnum"] for s in scenes]),
        "pt_offsets": np.concatenate([s["pt_offsets"] for s in scenes]),
        "semantic_preds": np.concatenate([s["semantic_preds"] for s in scenes]),
        "n_scenes": n_scenes,
    }
    if with_feats:
        rng = np.random.default_rng(7 + (feat_seed if feat_seed is not None else 1000 * config_id + first_scene))
        batch["feats"] = rng.standard_normal((len(locs), IN_CHANNELS), dtype=np.float32)
    return batch
