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


def check_trace(trace):
    checked = []
    for name, rec in trace.items():
        if name.startswith("voxelization_idx"):
            coords, B, oc, im, om = rec
            roc, rim, rom = o.voxelization_idx(_n(coords), B, 4)
            _same(_n(oc), roc, name + " output_coords")
            _same(_n(im), rim, name + " input_map")
            _same(_n(om), rom, name + " output_map")
        elif name.startswith("voxelization("):
            feats, rule, out = rec
            _same(_n(out), o.voxelization(_n(feats), _n(rule), 4), name)
        elif name.startswith("ballquery"):
            xyz, bi, bo, idx, sl = rec
            ridx, rsl = o.ballquery_batch_p(_n(xyz), _n(bi), _n(bo), 0.03)
            # segment placement is free (bfs_cluster.cu:47); lengths and each point's ascending list are not
            got, glen = o.relaid_neighbours(_n(idx), _n(sl))
            want, wlen = o.relaid_neighbours(ridx, rsl)
            _same(glen, wlen, name + " start_len lengths")
            _same(got, want, name + " idx")
        elif name.startswith("bfs_cluster"):
            sem, idx, sl, ci, co = rec
            rci, rco = o.bfs_cluster(_n(sem), _n(idx), _n(sl), 50)
            _same(_n(co), rco, name + " cluster_offsets")
            got, want = o.canonical_clusters(_n(ci), _n(co)), o.canonical_clusters(rci, rco)
            for a, b in zip(got, want):
                _same(a, b, name + " members")
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
