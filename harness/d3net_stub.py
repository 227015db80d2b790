"""HARNESS ONLY (not product, not oracle): runs the reference's OWN caller of the hot path -- `PointGroup.feed` /
`forward` / `clusters_voxelization` / `convert_stack_to_batch` of model/pointgroup.py, unmodified, from the staged
tree baseline/_ref (harness/stage_ref.py) -- on top of d3net_b200.pointgroup_ops.

What is stood in for, and why (hard part 6 -- none of these packages exists in this image and
none is on the hot path):
  MinkowskiEngine     the sparse-conv backbone / score U-Net.  Stand-in: a PASS-THROUGH network -- every
                      MinkowskiConvolution is a parameter-free channel adapter (truncate / zero-pad the channel axis, no
                      spatial mixing), MinkowskiBatchNorm the identity, MinkowskiReLU a ReLU.  On non-negative inputs the
                      reference's U-Net wiring (residual adds, ME.cat) then returns gain * x[:, :m] with a fixed gain (a
                      power of two, measured at build time).  A random-init backbone would predict noise labels, nothing
                      would cluster, and the ops would run on empty inputs; with the pass-through the synthetic scene's
                      labels and offsets are planted in the first input channels and DECODED by hand-set weights in the
                      reference's own `sem_seg` / `offset_net` heads, so `forward` sees realistic `semantic_preds` and
                      `pt_offsets` and drives the ops with BASELINE-shaped tensors.
  pytorch_lightning   LightningModule -> nn.Module with no-op save_hyperparameters / log.
  omegaconf           the yaml files are read with PyYAML into attribute dicts.
  trimesh, plyfile, matplotlib, torch_geometric, imageio, h5py
                      imported at module level by lib/utils/{bbox,pc}.py and model/graph_module.py, never called here.
`lib.pointgroup_ops.functions.pointgroup_ops` resolves to d3net_b200.pointgroup_ops (INTEGRATION.md option B) or, with
``wrapper="reference"``, to the reference's own wrapper file running over d3net_b200.PG_OP (option A).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")

N_SEM_CHANNEL = 6          # feats[:, 6] carries the label; [:, 0:3] / [:, 3:6] the positive / negative part of the offset


def available():
    return os.path.exists(os.path.join(STAGED, "model", "pointgroup.py"))


# ---- stand-in packages ----------------------------------------------------------------------------------------------
class _Anything(types.ModuleType):
    """A module whose every attribute is a dummy class (for packages that are imported but never used)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (object,), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class SparseTensor:
    """MinkowskiEngine.SparseTensor as far as model/pointgroup.py uses it: features + coordinates."""

    def __init__(self, features, coordinates=None, **_):
        self.features = features
        self.coordinates = coordinates

    F = property(lambda self: self.features)
    C = property(lambda self: self.coordinates)

    def __add__(self, other):
        return SparseTensor(self.features + other.features, self.coordinates)

    __iadd__ = __add__


class _ChannelAdapter(nn.Module):
    def __init__(self, in_channels, out_channels, *a, **k):
        super().__init__()
        self.cin, self.cout = in_channels, out_channels

    def forward(self, x):
        f = x.features
        if self.cout <= f.size(1):
            f = f[:, :self.cout]
        else:
            f = torch.cat([f, f.new_zeros(f.size(0), self.cout - f.size(1))], 1)
        return SparseTensor(f, x.coordinates)


class _Identity(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


class _ReLU(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return SparseTensor(torch.relu(x.features), x.coordinates)


def _me_module():
    me = _Anything("MinkowskiEngine")
    me.SparseTensor = SparseTensor
    me.MinkowskiConvolution = _ChannelAdapter
    me.MinkowskiConvolutionTranspose = _ChannelAdapter
    me.MinkowskiBatchNorm = _Identity
    me.MinkowskiReLU = _ReLU
    me.cat = lambda *ts: SparseTensor(torch.cat([t.features for t in ts], 1), ts[0].coordinates)
    return me


def _pl_module():
    pl = _Anything("pytorch_lightning")

    class LightningModule(nn.Module):
        current_epoch = 0

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    return pl


_installed = {}


def install_stubs(wrapper="d3net_b200"):
    """Puts the stand-ins and the staged tree on the import path.  Idempotent per `wrapper`."""
    assert available(), "baseline/_ref is not staged: run harness/stage_ref.py where /root/reference exists"
    if _installed.get("wrapper") == wrapper:
        return
    for name, mod in (("MinkowskiEngine", _me_module()), ("pytorch_lightning", _pl_module())):
        sys.modules[name] = mod
    for name in ("trimesh", "plyfile", "matplotlib", "matplotlib.pyplot", "imageio", "h5py", "torch_geometric",
                 "torch_geometric.utils", "torch_geometric.data", "torch_geometric.nn", "torch_geometric.typing"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Anything(name)
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if wrapper == "reference":
        # option A: the reference's wrapper, unmodified, over this repo's native-module mirror
        from d3net_b200 import PG_OP
        sys.modules["PG_OP"] = PG_OP
        path = os.path.join(STAGED, "lib", "pointgroup_ops", "functions", "pointgroup_ops.py")
        spec = importlib.util.spec_from_file_location("lib.pointgroup_ops.functions.pointgroup_ops", path)
        ops = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ops)
    else:
        from d3net_b200 import pointgroup_ops as ops
    sys.modules["lib.pointgroup_ops.functions.pointgroup_ops"] = ops
    for m in [k for k in sys.modules if k == "model.pointgroup" or k.startswith("model.pointgroup.")]:
        del sys.modules[m]                         # re-import against the newly selected wrapper
    pkg = sys.modules.get("lib.pointgroup_ops.functions")
    if pkg is not None:
        setattr(pkg, "pointgroup_ops", ops)
    _installed["wrapper"] = wrapper
    _installed["ops"] = ops


def ops_module():
    return _installed["ops"]


def use_ops(ops):
    """Point `lib.pointgroup_ops.functions.pointgroup_ops` at another provider of the operator API (the tests run the
    caller over the CPU oracle this way) and make the model files re-import against it."""
    sys.modules["lib.pointgroup_ops.functions.pointgroup_ops"] = ops
    pkg = sys.modules.get("lib.pointgroup_ops.functions")
    if pkg is not None:
        setattr(pkg, "pointgroup_ops", ops)
    for m in [k for k in sys.modules if k in ("model.pointgroup", "model.speaker", "model.caption_module")]:
        del sys.modules[m]
    _installed["ops"] = ops


# ---- configuration -----------------------------------------------------------------------------------------------------
class Cfg(dict):
    """Attribute access over the yaml dicts (what the reference gets from omegaconf)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    return x


def load_cfg(max_num_proposal=256, task="train"):
    import yaml
    with open(os.path.join(STAGED, "conf", "pointgroup.yaml")) as f:
        cfg = _wrap(yaml.safe_load(f))
    cfg.general.task = task
    cfg.SCANNETV2_PATH = Cfg(meta_data=os.path.join(STAGED, "data", "scannet", "meta_data"))
    cfg.model.max_num_proposal = max_num_proposal          # BASELINE configs[4]: 256 proposals per scene
    cfg.model.num_graph_steps = 0                          # SURVEY section 7, hard part 6: no torch_geometric
    cfg.model.use_relation = False
    return cfg


# ---- the detector with decodable heads -----------------------------------------------------------------------------
def build_detector(cfg, device):
    install_stubs(_installed.get("wrapper", "d3net_b200"))
    PointGroup = importlib.import_module("model.pointgroup").PointGroup
    torch.manual_seed(0)
    model = PointGroup(cfg).to(device).eval()
    m = cfg.model.m
    with torch.no_grad():
        probe = SparseTensor(torch.ones(4, 134, device=device), None)
        gain = float(model.backbone(probe).features[0, 0])       # pass-through gain of the stand-in U-Net wiring
        assert gain > 0 and torch.equal(model.backbone(probe).features, torch.full((4, m), gain, device=device))
        # sem_seg: argmax_c (2 c x - c^2) = the integer nearest to x, x = label channel / gain
        model.sem_seg.weight.zero_()
        model.sem_seg.bias.zero_()
        c = torch.arange(cfg.data.classes, dtype=torch.float32, device=device)
        model.sem_seg.weight[:, N_SEM_CHANNEL] = 2.0 * c / gain
        model.sem_seg.bias.copy_(-c * c)
        # offset_net: Linear (undo the gain) -> BatchNorm1d (eval: x / sqrt(1 + eps)) -> ReLU -> Linear (off+ - off-)
        lin0, bn, _, lin1 = model.offset_net
        lin0.weight.copy_(torch.eye(m, device=device) / gain)
        lin0.bias.zero_()
        s = float(torch.sqrt(bn.running_var[0] + bn.eps))
        lin1.weight.zero_()
        lin1.bias.zero_()
        for j in range(3):
            lin1.weight[j, j] = s
            lin1.weight[j, 3 + j] = -s
    model.pass_through_gain = gain
    return model


def planted_labels(np_batch, p2v_map, v2p_map):
    """The label the detector will predict for every point: the label of the first point of its input voxel.  The
    network sees voxel MEANS of the point features (mode 4), so a planted scalar survives only if it is constant on
    each voxel -- which is also what the real detector does: one prediction per voxel, broadcast to its points."""
    first = v2p_map[:, 1].long()[p2v_map.long()].numpy()
    return np_batch["semantic_preds"][first]


def planted_feats(np_batch, p2v_map, v2p_map, seed=0):
    """fp32 [N, 131] point features (the reference appends `locs` as the last three of its 134 input channels,
    model/pointgroup.py:469-470) whose first channels carry what the decodable heads read back."""
    n = np_batch["locs"].shape[0]
    rng = np.random.default_rng(seed)
    f = np.abs(rng.standard_normal((n, 131))).astype(np.float32)
    off = np_batch["pt_offsets"].astype(np.float32)
    f[:, 0:3] = np.maximum(off, 0)
    f[:, 3:6] = np.maximum(-off, 0)
    f[:, N_SEM_CHANNEL] = planted_labels(np_batch, p2v_map, v2p_map).astype(np.float32)
    return f


def collate(np_batch, device, seed=0, max_instances=128):
    """The collated batch as lib/dataset/pipeline.py:937-992 hands it to the model: CPU voxelization_idx through the
    operator API in use (as the DataLoader workers do), then everything moved to the device (as Lightning does)."""
    ops = ops_module()
    B = int(np_batch["n_scenes"])
    locs_scaled = torch.from_numpy(np_batch["locs_scaled"])
    voxel_locs, p2v_map, v2p_map = ops.voxelization_idx(locs_scaled, B, 4)            # CPU in, CPU out
    bidx = np_batch["locs_scaled"][:, 0]
    batch_offsets = np.zeros(B + 1, np.int32)
    batch_offsets[1:] = np.cumsum(np.bincount(bidx, minlength=B))
    # GT object centres per scene, padded (data_dict["center_label"], used by get_object_assignments)
    inst = np_batch["instance_ids"]
    center = np.zeros((B, max_instances, 3), np.float32)
    for b in range(B):
        sel = (bidx == b) & (inst >= 0)
        ids = np.unique(inst[sel])[:max_instances]
        for k, i in enumerate(ids):
            center[b, k] = np_batch["locs"][sel & (inst == i)].mean(0)
    d = {
        "locs": torch.from_numpy(np_batch["locs"]),
        "locs_scaled": locs_scaled,
        "feats": torch.from_numpy(planted_feats(np_batch, p2v_map.cpu(), v2p_map.cpu(), seed)),
        "voxel_locs": voxel_locs, "p2v_map": p2v_map, "v2p_map": v2p_map,
        "batch_offsets": torch.from_numpy(batch_offsets),
        "instance_ids": torch.from_numpy(np_batch["instance_ids"]),
        "instance_num_point": torch.from_numpy(np.ascontiguousarray(np_batch["instance_pointnum"], dtype=np.int32)),
        "center_label": torch.from_numpy(center),
    }
    return {k: v.to(device) for k, v in d.items()}


class OpTimer:
    """Wall-clock share of the pointgroup_ops calls inside the caller (synchronising; for the `ops share` figure only)."""

    NAMES = ("voxelization_idx", "voxelization", "ballquery_batch_p", "bfs_cluster", "roipool", "get_iou", "sec_mean",
             "sec_min", "sec_max")

    def __init__(self):
        self.ms = {}
        self.calls = []            # (name, args, result) when recording

    def __enter__(self):
        import time
        ops = ops_module()
        self._saved = {n: getattr(ops, n) for n in self.NAMES}

        def wrap(name, fn):
            def timed(*a, **k):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                self.ms[name] = self.ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
                self.calls.append((name, a, r))
                return r
            return timed
        for n, fn in self._saved.items():
            setattr(ops, n, wrap(n, fn))
        return self

    def __exit__(self, *exc):
        ops = ops_module()
        for n, fn in self._saved.items():
            setattr(ops, n, fn)
        return False


def run_feed(model, data_dict, epoch=1, seed=1234):
    """`PointGroup.feed` exactly as training_step / test call it (model/pointgroup.py:466-479, :509)."""
    torch.manual_seed(seed)                 # clusters_voxelization draws torch.rand(3) twice (:161)
    with torch.no_grad():
        return model.feed(dict(data_dict), epoch)


def rand6_for(seed=1234):
    torch.manual_seed(seed)
    return torch.cat([torch.rand(3), torch.rand(3)])


# ---- the speaker (BASELINE configs[4]) -------------------------------------------------------------------------------
# model/speaker.py's SpeakerNet = GraphModule (torch_geometric, absent here) -> TopDownSceneCaptionModule.  The caption
# module reads three tensors only the graph module writes (`bbox_feature`, `edge_feature`, `adjacent_mat`,
# model/graph_module.py:315-318), so the reference itself cannot run with `num_graph_steps: 0`.  The graph module is one
# more out-of-scope network: its stand-in below is a fixed linear lift of the detector's batched proposal features to the
# caption module's feature size, no message passing.  The caption module runs unmodified.
SPK_FEAT = 128


def build_speaker(cfg, device, vocab_size=3433, seed=0):
    install_stubs(_installed.get("wrapper", "d3net_b200"))
    SpeakerNet = importlib.import_module("model.speaker").SpeakerNet
    words = ["sos", "eos", "unk", "pad_"] + ["w%d" % i for i in range(vocab_size - 4)]
    vocabulary = {"word2idx": {w: i for i, w in enumerate(words)}, "idx2word": {str(i): w for i, w in enumerate(words)}}
    rng = np.random.default_rng(seed)
    embeddings = rng.standard_normal((vocab_size, 300)).astype(np.float32)
    cfg.model.no_captioning = False
    cfg.model.no_detection = False
    torch.manual_seed(seed)
    spk = SpeakerNet(cfg, vocabulary, embeddings).to(device).eval()
    g = torch.Generator().manual_seed(seed)
    spk.graph_stand_in = (torch.randn((cfg.model.m, SPK_FEAT), generator=g) / cfg.model.m ** 0.5).to(device)
    return spk


def speaker_inputs(np_batch, cfg, device, chunk=8, seed=0, max_instances=128, vocab_size=3433):
    """The language / ground-truth-box keys of a captioning batch (lib/dataset/pipeline.py), synthetic."""
    B = int(np_batch["n_scenes"])
    rng = np.random.default_rng(seed)
    L = cfg.data.max_spk_len + 2
    bidx = np_batch["locs_scaled"][:, 0]
    inst = np_batch["instance_ids"]
    gt = np.zeros((B, max_instances, 8, 3), np.float32)
    ref_label = np.zeros((B, chunk, max_instances), np.float32)
    ref_corner = np.zeros((B, chunk, 8, 3), np.float32)
    signs = np.array([[sx, sy, sz] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)], np.float32)
    for b in range(B):
        sel = (bidx == b) & (inst >= 0)
        ids = np.unique(inst[sel])[:max_instances]
        for k, i in enumerate(ids):
            p = np_batch["locs"][sel & (inst == i)]
            lo, hi = p.min(0), p.max(0)
            gt[b, k] = (lo + hi) / 2 + 0.5 * (hi - lo) * signs
        pick = rng.integers(0, max(len(ids), 1), chunk)
        ref_label[b, np.arange(chunk), pick] = 1
        ref_corner[b] = gt[b, pick]
    lens = rng.integers(6, cfg.data.max_spk_len, (B, chunk))
    ids_ = rng.integers(4, vocab_size, (B, chunk, L))
    ids_[:, :, 0] = 0                                                   # sos
    d = {"annotated": np.ones((B, chunk), np.int64), "lang_ids": ids_.astype(np.int64), "lang_len": lens.astype(np.int64),
         "ref_box_label": ref_label, "ref_box_corner_label": ref_corner, "gt_bbox": gt}
    return {k: torch.from_numpy(v).to(device) for k, v in d.items()}


def run_speaker(det, spk, data_dict, lang, epoch=1, seed=1234):
    """Detector feed + graph stand-in + SpeakerNet forward (teacher forcing, as the captioning training step runs it,
    model/pipeline.py:154-157)."""
    torch.manual_seed(seed)
    with torch.no_grad():
        d = det.feed(dict(data_dict), epoch)
        d.update(lang)
        feats = d["proposal_feats_batched"]                              # [B, P, m]
        B, P, _ = feats.shape
        d["bbox_feature"] = feats @ spk.graph_stand_in
        d["edge_feature"] = feats.new_zeros((B, P, spk.cfg.model.num_locals, SPK_FEAT))
        d["adjacent_mat"] = feats.new_zeros((B, P, P))
        return spk(d, use_tf=True, use_rl=False, is_eval=False)
