"""The reference's own caller (model/pointgroup.py, staged unmodified under baseline/_ref) against this repo's
restatement of it (d3net_b200/chain.py):

  CPU  (-m "not gpu")  both run over the ORACLE dressed as the operator API (oracle/ops_adapter.py), with `.cuda()`
                       patched to a no-op -- checks the harness, the stand-in backbone and that chain.py issues the same
                       op calls with the same tensors as the real `PointGroup.forward` / `clusters_voxelization`;
  GPU  (-m gpu)        the real caller over d3net_b200.pointgroup_ops with its own CPU-tensor call pattern
                       (model/pointgroup.py:297,305,167-169), compared with chain.py's device-resident pass.
Skipped when baseline/_ref is absent (it is staged by harness/stage_ref.py where /root/reference exists).
"""
import numpy as np
import pytest
import torch

from d3net_b200 import chain, scenes
from harness import d3net_stub as H

needs_staged = pytest.mark.skipif(not H.available(), reason="baseline/_ref not staged (harness/stage_ref.py)")


def _chain_batch(np_batch, data_dict, captured, device):
    """The tensors chain.proposal_chain takes as inputs, taken from what the real caller computed."""
    return {
        "locs": data_dict["locs"], "locs_scaled": data_dict["locs_scaled"],
        "feats": torch.cat((data_dict["feats"], data_dict["locs"]), 1).contiguous(),
        "pt_feats": captured["pt_feats"].contiguous(),
        "semantic_preds": captured["semantic_preds"], "pt_offsets": captured["pt_offsets"].contiguous(),
        "instance_ids": data_dict["instance_ids"], "instance_pointnum": data_dict["instance_num_point"],
        "n_scenes": int(np_batch["n_scenes"]),
    }


def _run_both(ops, device, n_scenes, n_points, seed=1234):
    cfg = H.load_cfg()
    model = H.build_detector(cfg, device)
    nb = scenes.make_batch(n_scenes, n_points, config_id=5, geometry_points=n_points)
    dd = H.collate(nb, device)
    captured = {}
    hook = model.backbone.register_forward_hook(
        lambda mod, inp, out: captured.__setitem__("pt_feats", out.features[dd["p2v_map"].long()]))
    out = H.run_feed(model, dd, epoch=1, seed=seed)
    hook.remove()
    captured["semantic_preds"] = out["semantic_scores"].max(1)[1]
    captured["pt_offsets"] = out["pt_offsets"]
    batch = _chain_batch(nb, dd, captured, device)
    mine = chain.proposal_chain(ops, batch, rand6=H.rand6_for(seed).to(device))
    return cfg, model, nb, out, mine, captured, dd


def _compare(cfg, out, mine, captured, nb, dd):
    scores, proposals_idx, proposals_offset = out["proposal_scores"]
    nP = proposals_offset.numel() - 1
    assert nP > 10
    # the planted labels / offsets came back through the reference's own heads
    planted = H.planted_labels(nb, dd["p2v_map"].cpu(), dd["v2p_map"].cpu())
    assert (captured["semantic_preds"].cpu().numpy() == planted).all()
    assert torch.equal(proposals_idx.cpu().int(), mine["proposals_idx"].cpu().int())
    assert torch.equal(proposals_offset.cpu().int(), mine["proposals_offset"].cpu().int())
    mask = out["proposal_thres_mask"].cpu()
    crop = out["proposal_crop_bbox"].cpu()
    assert torch.equal(crop[:, :3], mine["proposals_center"].cpu()[mask])
    assert torch.equal(crop[:, 3:6], mine["proposals_size"].cpu()[mask])
    # score_net is a pass-through with a power-of-two gain: pooled features differ by exactly that factor
    gain = float(out["proposal_feats"].sum() / mine["proposals_score_feats"].cpu()[mask].sum()) if mask.any() else 1.0
    if mask.any():
        assert gain == 2.0 ** round(np.log2(gain))
        assert torch.equal(out["proposal_feats"].cpu(), mine["proposals_score_feats"].cpu()[mask] * gain)
    B, P = len(out["batch_offsets"]) - 1, cfg.model.max_num_proposal
    # convert_stack_to_batch (model/pointgroup.py:223-263) vs the loop-free helper: identical tensors, the shuffle included
    torch.manual_seed(1234)
    torch.rand(3), torch.rand(3)                                  # the two draws of clusters_voxelization come first (:161)
    mine_b = chain.convert_stack_to_batch(out["proposals_batchId"], out["proposal_feats"], out["proposal_crop_bbox"],
                                          out["proposal_objectness_scores"], B, P)
    for k, v in mine_b.items():
        assert torch.equal(v.cpu(), out[k].cpu()), k
    assert torch.equal(out["proposals_npoint"].cpu(), chain.proposals_npoint(proposals_offset).cpu())
    assert tuple(out["proposal_feats_batched"].shape) == (B, P, cfg.model.m)
    assert int(out["proposal_batch_mask"].sum()) == min(int(mask.sum()), int(out["proposal_batch_mask"].numel()))


@needs_staged
def test_real_caller_equals_chain_on_cpu_oracle(monkeypatch):
    from oracle.ops_adapter import OracleOps
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ops = OracleOps(use_ref=True)
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    H.use_ops(ops)
    try:
        cfg, model, nb, out, mine, captured, dd = _run_both(ops, torch.device("cpu"), 2, 9000)
        _compare(cfg, out, mine, captured, nb, dd)
    finally:
        H._installed.clear()


@needs_staged
@pytest.mark.gpu
@pytest.mark.parametrize("wrapper", ["d3net_b200", "reference"])
def test_real_caller_on_gpu(wrapper):
    """`wrapper="reference"`: the reference's own functions/pointgroup_ops.py over d3net_b200.PG_OP (option A);
    `"d3net_b200"`: this repo's operator API mirror (option B).  Same caller, same results as chain.py."""
    from d3net_b200 import pointgroup_ops as ops
    H._installed.clear()
    H.install_stubs(wrapper=wrapper)
    try:
        cfg, model, nb, out, mine, captured, dd = _run_both(ops, torch.device("cuda"), 2, 20000)
        _compare(cfg, out, mine, captured, nb, dd)
        assert not out["proposal_scores"][1].is_cuda            # bfs_cluster got CPU tensors and answered on the CPU
    finally:
        H._installed.clear()


@needs_staged
def test_speaker_forward_on_cpu_oracle(monkeypatch):
    """BASELINE configs[4] end to end on the CPU: the reference's detector caller AND its caption module
    (model/speaker.py, model/caption_module.py, unmodified) over the oracle ops -- the harness wiring check."""
    import sys
    from oracle.ops_adapter import OracleOps
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ops = OracleOps(use_ref=True)
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    H.use_ops(ops)
    try:
        dev = torch.device("cpu")
        cfg = H.load_cfg(max_num_proposal=64)
        det = H.build_detector(cfg, dev)
        spk = H.build_speaker(cfg, dev, vocab_size=300)
        nb = scenes.make_batch(2, 9000, config_id=5, geometry_points=9000)
        dd = H.collate(nb, dev)
        lang = H.speaker_inputs(nb, cfg, dev, chunk=2, vocab_size=300)
        out = H.run_speaker(det, spk, dd, lang)
        caps = out["lang_cap"]
        assert caps.shape[0] == 4 and caps.shape[2] == 300 and torch.isfinite(caps).all()
        assert out["good_bbox_masks"].shape == (4,)
    finally:
        H._installed.clear()


@needs_staged
@pytest.mark.gpu
def test_speaker_forward_on_gpu():
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    try:
        dev = torch.device("cuda")
        cfg = H.load_cfg(max_num_proposal=256)
        det = H.build_detector(cfg, dev)
        spk = H.build_speaker(cfg, dev)
        nb = scenes.make_batch(2, 20000, config_id=5, geometry_points=20000)
        dd = H.collate(nb, dev)
        lang = H.speaker_inputs(nb, cfg, dev, chunk=4)
        out = H.run_speaker(det, spk, dd, lang)
        caps = out["lang_cap"]
        assert caps.shape[0] == 8 and caps.is_cuda and torch.isfinite(caps).all()
        assert int(out["proposal_batch_mask"].sum()) > 10
    finally:
        H._installed.clear()


@needs_staged
@pytest.mark.parametrize("P", [4, 64])
def test_convert_stack_to_batch_helper_matches_the_reference_method(P):
    """The reference's own method on random stacks (truncation at max_num_proposal, empty scenes, unsorted scene ids)."""
    H._installed.clear()
    H.install_stubs(wrapper="d3net_b200")
    try:
        cfg = H.load_cfg(max_num_proposal=P, task="test")             # "test": skips the GT assignment (needs center_label)
        det = H.build_detector(cfg, torch.device("cpu"))
        g = torch.Generator().manual_seed(3)
        n, B = 90, 5
        bid = torch.randint(0, B, (n,), generator=g)
        bid[bid == 3] = 2                                             # scene 3 has no proposals
        dd = {"batch_offsets": torch.arange(B + 1), "proposals_batchId": bid.int(),
              "proposal_feats": torch.randn((n, cfg.model.m), generator=g),
              "proposal_crop_bbox": torch.cat([torch.randn((n, 3), generator=g), torch.rand((n, 3), generator=g) * 2,
                                               torch.zeros(n, 1), torch.randint(0, 20, (n, 1), generator=g).float(),
                                               torch.rand((n, 1), generator=g)], 1),
              "proposal_objectness_scores": torch.rand(n, generator=g)}
        torch.manual_seed(77)
        ref = det.convert_stack_to_batch(dict(dd))
        torch.manual_seed(77)
        mine = chain.convert_stack_to_batch(dd["proposals_batchId"], dd["proposal_feats"], dd["proposal_crop_bbox"],
                                            dd["proposal_objectness_scores"], B, P)
        for k, v in mine.items():
            assert torch.equal(v, ref[k]), k
        assert int(mine["proposal_batch_mask"].sum()) == sum(min(int((bid == b).sum()), P) for b in range(B))
    finally:
        H._installed.clear()
