// Workspace layout of the ball query (ballquery.cu), shared with bfs_cluster's grid-assisted sweep
// (cluster.cu), which reads the uniform grid the producing ball query left behind.
#pragma once
#include "common.cuh"

namespace pg {

struct BqWs {
    int4 *keys;
    GroupTable tab;
    int32_t *pslot, *cell, *ccnt, *cstart, *kc, *kb, *cand_start, *counts, *nbr, *mbase, *dense;
    int32_t *dbase;     // dense cells, by slot of `dense`: where the cell's candidate coordinates start in cand_xy / cand_z
    uint2 *crange;      // dense cells: (smallest, largest) candidate index
    int32_t *qpos;      // point -> its position in the query order (inverse of sorted_pt)
    uint32_t *cand_idx;
    // dense cells: candidate coordinates in merged (ascending index) order, laid out in PAIRS for the packed f32x2
    // predicate and padded per cell to a multiple of 32 with +inf:  cand_xy[p] = (x0, x1, y0, y1), cand_z[p] = (z0, z1)
    float4 *cand_xy;
    float2 *cand_z;
    size_t cand_cap;    // capacity of cand_xy / cand_z in candidates
    uint32_t *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    // [0] nCells [1] total candidates [2] total neighbours [3] nDense [4] test work counter [5] merge work counter
    // [6] mask words [7] cursor into cand_xy / cand_z
    int64_t *scalars;
    bool ok;
    size_t used;
};

inline BqWs bq_layout(void *ws, size_t ws_bytes, int64_t n_) {
    Arena a(ws, ws_bytes);
    BqWs w;
    const size_t n = (size_t)(n_ > 0 ? n_ : 1);
    w.tab.cap = group_table_cap(n_);
    w.keys = a.take<int4>(n);
    w.tab.slot_rep = a.take<int32_t>(w.tab.cap);
    w.tab.slot_gid = a.take<int32_t>(w.tab.cap);
    w.tab.slot_key = a.take<int4>(w.tab.cap);
    w.pslot = a.take<int32_t>(n);
    w.cell = a.take<int32_t>(n);
    w.ccnt = a.take<int32_t>(n + 1);
    w.cstart = a.take<int32_t>(n + 1);
    w.kc = a.take<int32_t>(n + 1);
    w.kb = a.take<int32_t>(n + 1);
    w.cand_start = a.take<int32_t>(n + 1);
    w.counts = a.take<int32_t>(n + 1);
    w.nbr = a.take<int32_t>(n * 27);
    w.dense = a.take<int32_t>(n);
    w.dbase = a.take<int32_t>(n);
    w.crange = a.take<uint2>(n);
    w.qpos = a.take<int32_t>(n);
    w.kA = a.take<uint32_t>(n);
    w.vA = a.take<uint32_t>(n);
    w.kB = a.take<uint32_t>(n);
    w.vB = a.take<uint32_t>(n);
    w.hist = a.take<int32_t>(radix_tmp_count(n_));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(n + radix_tmp_count(n_))));
    w.scalars = a.take<int64_t>(8);
    w.cand_idx = a.take<uint32_t>(n * 27);
    w.mbase = a.take<int32_t>(n + 1);
    // a point is a candidate of at most 27 cells; a dense cell has more than 128 candidates, so at most 27 n / 129 cells
    // are dense and padding each to a multiple of 32 adds at most 31 entries per dense cell: 27 n + 31 * 27 n / 129 < 34 n
    w.cand_cap = n * 34 + 64;
    w.cand_xy = a.take<float4>(w.cand_cap / 2);
    w.cand_z = a.take<float2>(w.cand_cap / 2);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

// the points grouped by cell (cell c: sorted_pt[cstart[c] .. cstart[c] + ccnt[c]), in no particular order inside a cell);
// this is also the QUERY order.
inline const uint32_t *bq_sorted(const BqWs &w, int32_t) { return w.kA; }

// Lazy lists (cluster.cu's fused path): the neighbour lists stay in the form the count phase left them in -- one bit per
// (query, candidate) plus every cell's merged candidate indices -- and are only turned into index lists where the
// clustering sweep reads them.
//   bq_list_samples  per point (first, second, a middle, last) entry of its list, straight from the masks; `last` is the
//                    1000th hit when the list is full.  What k_cl_prep / k_cl_sample need instead of idx.
//   bq_fill_lists    the lists of the points in `worklist[0 .. *count)` (device count), written at their start_len
//                    positions of idx; everything else in idx stays untouched.
int bq_list_samples(const BqWs &w, const uint32_t *masks, int32_t n, int4 *samples, cudaStream_t st);
int bq_fill_lists(const BqWs &w, const uint32_t *masks, const int2 *start_len, const uint32_t *worklist,
                  const unsigned long long *count, int32_t n, int32_t *idx, cudaStream_t st);

}  // namespace pg
