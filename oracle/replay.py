"""TEST INFRASTRUCTURE ONLY -- replays a trace of d3net_b200.chain.proposal_chain through the oracle.

proposal_chain(..., trace={}) records, for every op call, the exact tensors it was given and what it
returned.  check_trace() feeds the same inputs to the CPU oracle and compares: bit-exact for every op
(cluster memberships up to the canonical member order; NaNs compare equal to NaNs)."""
import numpy as np

from . import pg_oracle as o


def _n(t):
    return t.detach().cpu().numpy()


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        assert (na == nb).all(), what + ": NaN pattern differs"
        assert (a.view(np.uint32)[~na] == b.view(np.uint32)[~nb]).all(), what + ": values differ"
    else:
        assert (a == b).all(), what + ": values differ"


def _same_lists_on_device(idx, sl, ridx, rsl, what):
    """Full-size neighbour lists (hundreds of millions of entries): the per-point comparison of
    relaid_neighbours() done with torch on the device that already holds the product's lists.  The
    oracle lays its segments out in point order; the product's may sit anywhere."""
    import torch
    dev = idx.device
    rsl_t = torch.from_numpy(np.ascontiguousarray(rsl)).to(dev)
    lens = rsl_t[:, 1].long()
    assert torch.equal(sl[:, 1].long(), lens), what + " start_len lengths"
    total = int(lens.sum())
    assert total == idx.numel() == len(ridx), what + " total"
    # the product's segments must tile idx without overlap
    o_ = torch.argsort(sl[:, 0].long(), stable=True)
    o_ = o_[lens[o_] > 0]
    st = sl[:, 0].long()[o_]
    assert torch.equal(st, torch.cumsum(lens[o_], 0) - lens[o_]), what + ": segments do not tile idx"
    del o_, st
    step = 1 << 26                                             # oracle entries per slice (memory bound)
    rstart = torch.cumsum(lens, 0) - lens
    owner_all = torch.repeat_interleave(torch.arange(lens.numel(), device=dev, dtype=torch.int32), lens)
    for a in range(0, total, step):
        b = min(total, a + step)
        owner = owner_all[a:b].long()
        pos = torch.arange(a, b, device=dev) - rstart[owner]
        got = idx[sl[:, 0].long()[owner] + pos]
        want = torch.from_numpy(ridx[a:b]).to(dev)
        assert torch.equal(got, want), what + " idx"


def check_trace(trace, ref=None, big=False):
    """``ref``: the reference's own compiled PG_OP (oracle/_ref) -- its two CPU ops, voxelize_idx and
    bfs_cluster, are then run next to the C restatement.  ``big``: BASELINE-size traces; neighbour lists
    are compared on the device."""
    import torch
    checked = []
    for name, rec in trace.items():
        if name.startswith("voxelization_idx"):
            coords, B, oc, im, om = rec
            roc, rim, rom = o.voxelization_idx(_n(coords), B, 4)
            _same(_n(oc), roc, name + " output_coords")
            _same(_n(im), rim, name + " input_map")
            _same(_n(om), rom, name + " output_map")
            if ref is not None:
                c = coords.detach().cpu().contiguous()
                xoc, xim, xom = c.new(), torch.zeros(c.size(0), dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
                ref.voxelize_idx(c, xoc, xim, xom, int(B), 4)
                _same(roc, xoc.numpy(), name + " output_coords (reference binary)")
                _same(rim, xim.numpy(), name + " input_map (reference binary)")
                _same(rom, xom.numpy(), name + " output_map (reference binary)")
        elif name.startswith("voxelization("):
            feats, rule, out = rec
            _same(_n(out), o.voxelization(_n(feats), _n(rule), 4), name)
        elif name.startswith("ballquery"):
            xyz, bi, bo, idx, sl = rec
            ridx, rsl = o.ballquery_batch_p(_n(xyz), _n(bi), _n(bo), 0.03)
            # segment placement is free (bfs_cluster.cu:47); lengths and each point's ascending list are not
            if big:
                _same_lists_on_device(idx, sl, ridx, rsl, name)
            else:
                got, glen = o.relaid_neighbours(_n(idx), _n(sl))
                want, wlen = o.relaid_neighbours(ridx, rsl)
                _same(glen, wlen, name + " start_len lengths")
                _same(got, want, name + " idx")
            del ridx, rsl
        elif name.startswith("bfs_cluster"):
            sem, idx, sl, ci, co = rec
            h_sem, h_idx, h_sl = _n(sem), _n(idx), _n(sl)
            rci, rco = o.bfs_cluster(h_sem, h_idx, h_sl, 50)
            _same(_n(co), rco, name + " cluster_offsets")
            got, want = o.canonical_clusters(_n(ci), _n(co)), o.canonical_clusters(rci, rco)
            for a, b in zip(got, want):
                _same(a, b, name + " members")
            if ref is not None:           # the reference's own BFS (bfs_cluster.cpp:60-75) on the same lists
                xci, xco = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
                ref.bfs_cluster(torch.from_numpy(h_sem), torch.from_numpy(h_idx), torch.from_numpy(h_sl), xci, xco,
                                h_sl.shape[0], 50)
                _same(rco, xco.numpy(), name + " cluster_offsets (reference binary)")
                _same(rci, xci.numpy(), name + " cluster_idxs incl. BFS order (reference binary)")
            del h_idx
        elif name == "sec_mean":
            x, off, out = rec
            _same(_n(out), o.sec_mean(_n(x), _n(off)), name)
        elif name == "sec_minmax":
            x, off, mn, mx = rec
            _same(_n(mn), o.sec_min(_n(x), _n(off)), "sec_min")
            _same(_n(mx), o.sec_max(_n(x), _n(off)), "sec_max")
        elif name == "roipool":
            x, off, out = rec
            _same(_n(out), o.roipool(_n(x), _n(off))[0], name)
        elif name == "get_iou":
            pidx, off, lab, pn, iou = rec
            _same(_n(iou), o.get_iou(_n(pidx), _n(off), _n(lab), _n(pn)), name)
        else:
            raise AssertionError("unknown trace entry " + name)
        checked.append(name)
    return checked
