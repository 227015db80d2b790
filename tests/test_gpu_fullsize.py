"""BASELINE.json's full sizes on the GPU:
  configs[1]  the whole chain on 8 x 150k-point scenes
  configs[3]  one 1M-point dense scene through both clusterings (truncated lists, one giant component)
Each is checked twice: bit-exact against the CPU oracle (the C restatement, plus the reference's own compiled
voxelize_idx / bfs_cluster from oracle/_ref when it was built) by replaying the trace of every op call, and
through size-independent properties and independent re-computations (torch / scipy).
"""
import numpy as np
import pytest
import torch

from d3net_b200 import chain, scenes

pytestmark = pytest.mark.gpu


def _seg_ids(offsets):
    lens = (offsets[1:] - offsets[:-1]).long()
    return torch.repeat_interleave(torch.arange(lens.numel(), device=offsets.device), lens)


def _check_voxel_maps(coords, voxel_coords, p2v, v2p):
    N, M = coords.size(0), voxel_coords.size(0)
    assert torch.equal(voxel_coords[p2v.long()], coords)                      # every point sits in its voxel
    cnt = v2p[:, 0].long()
    assert int(cnt.sum()) == N and int(cnt.min()) >= 1
    assert int(cnt.max()) + 1 == v2p.size(1)                                   # width = maxActive + 1
    cols = torch.arange(1, v2p.size(1), device=v2p.device)[None, :]
    live = cols <= cnt[:, None]
    body = v2p[:, 1:].long()
    assert int((body * ~live).abs().sum()) == 0                                # zero padding
    assert torch.equal(torch.sort(body[live])[0], torch.arange(N, device=v2p.device))   # a partition of the points
    inc = (body[:, 1:] > body[:, :-1]) | ~live[:, 1:]
    assert bool(inc.all())                                                     # ascending inside a row
    assert bool((v2p[1:, 1] > v2p[:-1, 1]).all())                              # voxels numbered by first occurrence
    assert torch.equal(p2v.long()[body[live]], torch.arange(M, device=v2p.device).repeat_interleave(cnt))
    assert voxel_coords.unique(dim=0).size(0) == M                             # no voxel split in two


def _owners(sl, n_idx):
    """Owner point of every idx entry.  Segments may sit anywhere (the reference places them by
    atomicAdd, bfs_cluster.cu:47) but must tile idx exactly, without gaps or overlap."""
    starts, lens = sl[:, 0].long(), sl[:, 1].long()
    order = torch.argsort(starts, stable=True)
    order = order[lens[order] > 0]
    ls = lens[order]
    assert torch.equal(starts[order], torch.cumsum(ls, 0) - ls) and int(ls.sum()) == n_idx
    return torch.repeat_interleave(order, ls)


def _check_neighbours(xyz, batch_idxs, idx, sl, r, rng, n_sample=300):
    n = xyz.size(0)
    starts, lens = sl[:, 0].long(), sl[:, 1].long()
    assert int(lens.max()) <= 1000
    owner = _owners(sl, idx.numel())
    j = idx.long()
    assert bool((batch_idxs[owner] == batch_idxs[j]).all())                    # never across scenes
    same_list = owner[1:] == owner[:-1]
    assert bool((j[1:] > j[:-1])[same_list].all())                             # ascending, no duplicates
    d = (xyz[owner].double() - xyz[j].double()).pow(2).sum(1).sqrt()
    assert float(d.max()) < r * (1 + 1e-5)
    finite = torch.isfinite(xyz).all(1)
    has_self = torch.zeros(n, dtype=torch.bool, device=xyz.device)
    has_self[owner[j == owner]] = True
    assert bool((has_self | (lens == 1000) | ~finite).all())                   # self included unless truncated away
    # completeness on a sample: brute force in float64 with a guard band around r
    for q in rng.integers(0, n, n_sample):
        q = int(q)
        same = torch.nonzero(batch_idxs == batch_idxs[q]).view(-1)
        dd = (xyz[same].double() - xyz[q].double()).pow(2).sum(1).sqrt()
        sure_in, sure_out = same[dd < r * (1 - 1e-5)], same[dd > r * (1 + 1e-5)]
        got = j[starts[q]:starts[q] + lens[q]]
        if lens[q] < 1000:
            assert bool(torch.isin(sure_in, got).all()) and not bool(torch.isin(sure_out, got).any())
        else:                                                                  # the 1000 lowest indices in range
            lim = got[-1]
            assert bool(torch.isin(sure_in[sure_in <= lim], got).all())
            assert int((sure_in <= lim).sum()) <= 1000


def _cluster_of(ci, n):
    c = torch.full((n,), -1, dtype=torch.long, device=ci.device)
    c[ci[:, 1].long()] = ci[:, 0].long()
    return c


def _check_clusters(sem, idx, sl, ci, co, thr):
    n = sl.size(0)
    sizes = (co[1:] - co[:-1]).long()
    assert int(co[0]) == 0 and int(co[-1]) == ci.size(0) and (sizes.numel() == 0 or int(sizes.min()) >= thr)
    assert torch.equal(ci[:, 0].long(), _seg_ids(co))
    assert ci[:, 1].unique().numel() == ci.size(0)                             # clusters are disjoint
    first = ci[co[:-1].long(), 1]
    assert bool((first[1:] > first[:-1]).all())                                # numbered by ascending seed
    same_c = ci[1:, 0] == ci[:-1, 0]
    assert bool((ci[1:, 1] > ci[:-1, 1])[same_c].all())                        # members ascend
    cl = _cluster_of(ci, n)
    lab = sem[ci[:, 1].long()]
    assert torch.equal(lab, lab[co[:-1].long()].repeat_interleave(sizes))      # one label per cluster
    # every edge between equal labels stays inside one component (sizes of both ends agree)
    lens = sl[:, 1].long()
    owner = _owners(sl, idx.numel())
    j = idx.long()
    e = sem[owner] == sem[j]
    two_way = (lens[j] < 1000)                                                 # reverse edge certainly present
    assert bool((cl[owner] == cl[j])[e & two_way].all())
    return cl


@pytest.fixture(scope="module")
def full_batch():
    nb = scenes.make_batch(8, 150_000, config_id=2)
    return nb, chain.batch_to_device(nb, torch.device("cuda"))


def test_config2_chain_properties(ops, full_batch):
    nb, batch = full_batch
    rng = np.random.default_rng(0)
    trace = {}
    out = chain.proposal_chain(ops, batch, trace=trace)
    # --- voxelisation of the input cloud
    coords, B, vc, p2v, v2p = trace["voxelization_idx(scene)"]
    _check_voxel_maps(coords, vc, p2v, v2p)
    feats, rule, vf = trace["voxelization(scene)"]
    ref = torch.zeros((vc.size(0), feats.size(1)), dtype=torch.float64, device="cuda").index_add_(0, p2v.long(), feats.double())
    ref = ref / rule[:, :1].double()
    torch.testing.assert_close(vf.double(), ref, rtol=1e-5, atol=1e-6)
    # --- ball query + clustering, both coordinate sets
    for tag in ("shift", "raw"):
        xyz, bi, bo, idx, sl = trace["ballquery(%s)" % tag]
        _check_neighbours(xyz, bi, idx, sl, 0.03, rng, 40)
        sem, _, _, ci, co = trace["bfs_cluster(%s)" % tag]
        _check_clusters(sem, idx, sl, ci, co, 50)
        assert co.numel() - 1 > 100
    # --- cluster statistics against torch's own segment reductions
    x, off, mean = trace["sec_mean"]
    lens = (off[1:] - off[:-1]).long()
    m64 = torch.segment_reduce(x.double(), "mean", lengths=lens, axis=0)
    torch.testing.assert_close(mean.double(), m64, rtol=1e-4, atol=1e-5)
    x, off, mn, mx = trace["sec_minmax"]
    assert torch.equal(mn, torch.segment_reduce(x, "min", lengths=lens, axis=0))
    assert torch.equal(mx, torch.segment_reduce(x, "max", lengths=lens, axis=0))
    ccoords, nC, cvc, cp2v, cv2p = trace["voxelization_idx(clusters)"]
    _check_voxel_maps(ccoords, cvc, cp2v, cv2p)
    x, off, pooled = trace["roipool"]
    assert torch.equal(pooled, torch.segment_reduce(x, "max", lengths=lens, axis=0))
    # --- IoU recomputed with torch in the reference's mixed precision
    pidx, off, labels, pointnum, iou = trace["get_iou"]
    nP, nI = off.numel() - 1, pointnum.numel()
    lab = labels[pidx.long()]
    keep = lab >= 0
    inter = torch.bincount(_seg_ids(off)[keep] * nI + lab[keep], minlength=nP * nI).view(nP, nI)
    den = (lens[:, None] + pointnum[None, :].long() - inter).float().double() + 1e-5
    assert torch.equal(iou, (inter.float().double() / den).float())
    assert float(iou.max()) > 0.9
    # --- determinism: a second pass gives the same bits
    out2 = chain.proposal_chain(ops, batch)
    for k in ("proposals_idx", "proposals_offset", "proposals_score_feats", "ious", "voxel_feats"):
        assert torch.equal(out[k], out2[k]), k


def test_config4_one_million_point_scene(ops):
    """Dense 1M-point room: long neighbour lists (many at the 1000 cap on the shifted coordinates) and a
    giant floor component.  The union-find path and the generic propagation path are two independent
    algorithms: they must agree; on the raw coordinates scipy's connected components is a third opinion."""
    from d3net_b200 import PG_OP
    s = scenes.make_scene(1_000_000, seed=4000)
    dev = torch.device("cuda")
    sem_all = torch.from_numpy(s["semantic_preds"]).to(dev)
    obj = torch.nonzero(sem_all > 0).view(-1)
    xyz = torch.from_numpy(s["locs"]).to(dev)[obj].contiguous()
    shifted = (xyz + torch.from_numpy(s["pt_offsets"]).to(dev)[obj]).contiguous()
    sem = sem_all[obj].int().contiguous()
    n = xyz.size(0)
    bi = torch.zeros(n, dtype=torch.int32, device=dev)
    bo = torch.tensor([0, n], dtype=torch.int32, device=dev)
    rng = np.random.default_rng(1)
    for tag, pts in (("raw", xyz), ("shift", shifted)):
        idx, sl = ops.ballquery_batch_p(pts, bi, bo, 0.03, 300)
        _check_neighbours(pts, bi, idx, sl, 0.03, rng, 25)
        ci, co, generic = PG_OP.bfs_cluster_impl(sem, idx, sl, 50)
        assert not generic
        cl = _check_clusters(sem, idx, sl, ci, co, 50)
        if tag == "shift":
            assert int((sl[:, 1] == 1000).sum()) > 1000                        # the cap really is exercised
            ci2, co2, g2 = PG_OP.bfs_cluster_impl(sem, idx, sl, 50, generic=1)
            assert g2 and torch.equal(co, co2) and torch.equal(ci, ci2)
        else:
            assert int(sl[:, 1].max()) < 1000
            assert int((co[1:] - co[:-1]).max()) > 100_000                     # the floor is one component
            import scipy.sparse as sp
            from scipy.sparse.csgraph import connected_components
            owner = _owners(sl, idx.numel())
            j = idx.long()
            e = sem[owner] == sem[j]
            a, b = owner[e].cpu().numpy(), j[e].cpu().numpy()
            g = sp.coo_matrix((np.ones(len(a), np.int8), (a, b)), shape=(n, n)).tocsr()
            ncomp, comp = connected_components(g, directed=False)
            comp = torch.from_numpy(comp).to(dev)
            size = torch.bincount(comp, minlength=ncomp)
            kept = size[comp] >= 50
            assert torch.equal(kept, cl >= 0)
            # same partition: component id <-> cluster id is a bijection on the kept points
            pairs = torch.stack([comp[kept], cl[kept]], 1).unique(dim=0)
            assert pairs.size(0) == co.numel() - 1 == pairs[:, 0].unique().numel() == pairs[:, 1].unique().numel()


# ---- bit-exact oracle replay at BASELINE sizes --------------------------------------------------------------------
def _ref_module():
    from oracle import build_ref
    return build_ref.load()          # None when oracle/_ref was never built (needs /root/reference at build time)


def test_config2_chain_oracle_replay(ops, full_batch):
    """configs[1]: every op call of the 8 x 150k chain replayed through the oracle on the same inputs --
    voxel maps, every neighbour list (338 M + 9 M entries), both clusterings, segment reductions, roipool, IoU."""
    from oracle import replay
    nb, batch = full_batch
    trace = {}
    chain.proposal_chain(ops, batch, trace=trace)
    torch.cuda.synchronize()
    checked = replay.check_trace(trace, ref=_ref_module(), big=True)
    assert len(checked) == 12


def _one_million_scene(dev):
    s = scenes.make_scene(1_000_000, seed=4000)
    sem_all = torch.from_numpy(s["semantic_preds"]).to(dev)
    obj = torch.nonzero(sem_all > 0).view(-1)
    xyz = torch.from_numpy(s["locs"]).to(dev)[obj].contiguous()
    shifted = (xyz + torch.from_numpy(s["pt_offsets"]).to(dev)[obj]).contiguous()
    sem = sem_all[obj].int().contiguous()
    n = xyz.size(0)
    bi = torch.zeros(n, dtype=torch.int32, device=dev)
    bo = torch.tensor([0, n], dtype=torch.int32, device=dev)
    return xyz, shifted, sem, bi, bo


def test_config4_one_million_point_scene_oracle_replay(ops):
    """configs[3]: both clusterings of the 1M-point dense room against the oracle, bit for bit, through the
    operator API (trusted, grid-assisted sweep) -- with hundreds of thousands of lists cut at the 1000 cap."""
    from oracle import replay
    dev = torch.device("cuda")
    xyz, shifted, sem, bi, bo = _one_million_scene(dev)
    ref = _ref_module()
    for tag, pts in (("raw", xyz), ("shift", shifted)):
        idx, sl = ops.ballquery_batch_p(pts, bi, bo, 0.03, 300)
        ci, co = ops.bfs_cluster(sem, idx, sl, 50)
        if tag == "shift":
            assert int((sl[:, 1] == 1000).sum()) > 100_000                     # the cap really is exercised
        trace = {"ballquery(%s)" % tag: (pts, bi, bo, idx, sl), "bfs_cluster(%s)" % tag: (sem, idx, sl, ci, co)}
        replay.check_trace(trace, ref=ref, big=True)
        # the fused op (lists decoded only where the sweep reads them; here most lists are full): same clusters
        fi, fo, total = ops.ballquery_bfs_cluster(pts, bi, bo, 0.03, 300, sem, 50)
        assert total == idx.numel() and torch.equal(fo, co) and torch.equal(fi, ci)
        del fi, fo
        del idx, sl, trace
        torch.cuda.empty_cache()


def test_config2_maskless_fill_matches(ops, full_batch):
    """The fill path taken when the hit-mask buffer does not fit (it re-evaluates the predicates): same lists,
    same layout as the masked path on the full 8 x 150k shifted set."""
    from d3net_b200 import PG_OP
    nb, batch = full_batch
    obj = torch.nonzero(batch["semantic_preds"] > 0).view(-1)
    bi = batch["locs_scaled"][:, 0].int()[obj].contiguous()
    bo = chain.get_batch_offsets(bi, 8)
    shifted = (batch["locs"][obj] + batch["pt_offsets"][obj]).contiguous()
    idx, sl = ops.ballquery_batch_p(shifted, bi, bo, 0.03, 300)

    def relaid(idx_, sl_):                       # every point's list, re-laid out in point order (placement is free)
        lens = sl_[:, 1].long()
        owner = torch.repeat_interleave(torch.arange(lens.numel(), device=idx_.device), lens)
        pos = torch.arange(owner.numel(), device=idx_.device) - (torch.cumsum(lens, 0) - lens)[owner]
        return idx_[sl_[:, 0].long()[owner] + pos]

    want = relaid(idx, sl)
    for kw in ({"use_masks": False}, {"mask_words": 1 << 20}):                 # no buffer / a buffer that is too small
        sl2, total, state = PG_OP.ballquery_count_impl(shifted, bi, bo, 0.03, **kw)
        assert state[1] is None and total == idx.numel()
        idx2 = torch.empty(total, dtype=torch.int32, device=idx.device)
        PG_OP.ballquery_fill_impl(shifted, 0.03, sl2, idx2, state)
        assert torch.equal(sl[:, 1], sl2[:, 1]) and torch.equal(want, relaid(idx2, sl2))


def test_config2_fused_cluster_chain_is_identical(ops, full_batch):
    """chain.proposal_chain(fused_cluster=True) on the full 8 x 150k batch: every output tensor equals the plain chain's."""
    nb, batch = full_batch
    a = chain.proposal_chain(ops, batch, overlap=False)
    b = chain.proposal_chain(ops, batch, fused_cluster=True, fused_glue=True)
    torch.cuda.synchronize()
    for k in ("proposals_idx", "proposals_offset", "proposals_score_feats", "ious", "proposals_center", "proposals_size",
              "proposals_voxel_feats", "proposals_voxel_coords"):
        assert torch.equal(a[k], b[k]), k
    assert a["nActive_shift"] == b["nActive_shift"] and a["nActive_raw"] == b["nActive_raw"]


def test_chain_two_stream_variant_is_identical(ops):
    """chain.proposal_chain(overlap=True): the scene voxelisation and the two clusterings issued from three host threads
    on three streams give the same tensors as the sequential pass."""
    nb = scenes.make_batch(3, 40_000, config_id=6, geometry_points=40_000)
    batch = chain.batch_to_device(nb, torch.device("cuda"))
    a = chain.proposal_chain(ops, batch, overlap=False)
    for _ in range(3):
        b = chain.proposal_chain(ops, batch, overlap=True)
        torch.cuda.synchronize()
        for k in ("proposals_idx", "proposals_offset", "proposals_score_feats", "ious", "proposals_center", "proposals_size",
                  "voxel_locs", "voxel_feats", "p2v_map"):
            assert torch.equal(a[k], b[k]), k
        assert a["nActive_shift"] == b["nActive_shift"] and a["nActive_raw"] == b["nActive_raw"]
        assert a["v2p_map_numel"] == b["v2p_map_numel"]
