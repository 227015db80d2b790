// ballquery_batch_p: uniform-grid radius search with the reference's exact fp32 predicate and its
// "first 1000 neighbours by ascending index" rule.  Reference behaviour:
// lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.cu:15-90 (brute-force scan of the whole scene).
//
// Pipeline (all on `stream`, no host round trip until the total is read):
//   1. cell key per point: (scene, floor(x/s), floor(y/s), floor(z/s)), s = 1.0001 * |r| in fp64, so
//      every pair that can pass the predicate sits in adjacent cells;
//   2. hash-group the keys -> dense cell ids, stable radix sort of (cell, point) -> every cell's
//      point list in ASCENDING original index;
//   3. per cell: the ids of its 27 neighbour cells (hash lookups) and the size of its candidate set;
//   4. merge: every cell gets ONE candidate array -- the union of its 27 neighbour lists, in
//      ascending original index, as float4 (x, y, z, index) -- built by rank-merging (a point's slot
//      is the sum of its lower-bound ranks in the 27 sorted lists);
//   5. count: a warp takes up to 32 query points of one cell (lane = query), streams the cell's
//      candidate array (warp-uniform 16-byte loads) and counts hits, capped at 1000;
//   6. exclusive scan of the counts -> start_len, total;
//   7. fill: a warp per query streams the same candidate array lane-per-candidate (coalesced 512-byte
//      loads), ballots the hits and writes them compacted -- ascending by construction -- until 1000.
#include <math.h>

#include "common.cuh"

namespace pg {

constexpr int kCap = PG_BALLQUERY_CAP;

struct BqWs {
    int4 *keys;
    GroupTable tab;
    int32_t *pslot, *cell, *ccnt, *cstart, *kc, *cand_start, *counts, *nbr;
    uint32_t *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    int64_t *scalars;   // [0] nCells, [1] total candidates, [2] total neighbours, [3] sorted-buffer id
    float4 *cand;
    bool ok;
    size_t used;
};

static BqWs bq_layout(void *ws, size_t ws_bytes, int64_t n_) {
    Arena a(ws, ws_bytes);
    BqWs w;
    const size_t n = (size_t)(n_ > 0 ? n_ : 1);
    w.tab.cap = group_table_cap(n_);
    w.keys = a.take<int4>(n);
    w.tab.slot_rep = a.take<int32_t>(w.tab.cap);
    w.tab.slot_gid = a.take<int32_t>(w.tab.cap);
    w.pslot = a.take<int32_t>(n);
    w.cell = a.take<int32_t>(n);
    w.ccnt = a.take<int32_t>(n + 1);
    w.cstart = a.take<int32_t>(n + 1);
    w.kc = a.take<int32_t>(n + 1);
    w.cand_start = a.take<int32_t>(n + 1);
    w.counts = a.take<int32_t>(n + 1);
    w.nbr = a.take<int32_t>(n * 27);
    w.kA = a.take<uint32_t>(n);
    w.vA = a.take<uint32_t>(n);
    w.kB = a.take<uint32_t>(n);
    w.vB = a.take<uint32_t>(n);
    w.hist = a.take<int32_t>(radix_tmp_count(n_));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(n + radix_tmp_count(n_))));
    w.scalars = a.take<int64_t>(8);
    w.cand = a.take<float4>(n * 27);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

// Far coordinates (|x/s| >= 2^30) have an fp32 spacing above the radius, so two of them can only be
// neighbours along that axis when they are the SAME float: any function of the bit pattern is a
// valid cell coordinate there.  Non-finite values land in the same branch and never pass the
// predicate anyway.
__device__ __forceinline__ int cell_coord(float x, double inv_s) {
    const double q = (double)x * inv_s;
    if (fabs(q) < 1073741824.0) return (int)floor(q);
    return (int)__float_as_uint(x);
}

__global__ void k_bq_keys(const float *__restrict__ xyz, const int32_t *__restrict__ batch_idxs, int32_t n,
                          double inv_s, int4 *__restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(xyz + 3 * (int64_t)i), y = __ldg(xyz + 3 * (int64_t)i + 1), z = __ldg(xyz + 3 * (int64_t)i + 2);
    keys[i] = make_int4(__ldg(batch_idxs + i), cell_coord(x, inv_s), cell_coord(y, inv_s), cell_coord(z, inv_s));
}

// one thread per cell: 27 neighbour ids (-1 when absent; slot 13 is the cell itself) + candidate count
__global__ void k_bq_neighbours(const int4 *__restrict__ keys, GroupTable tab, const uint32_t *__restrict__ sorted_pt,
                                const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                const int64_t *__restrict__ nCells, int32_t *__restrict__ nbr, int32_t *__restrict__ kc) {
    const int64_t nc = *nCells;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        const int4 k = keys[sorted_pt[cstart[c]]];
        int total = 0;
        int j = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++, j++) {
                    int id;
                    if (j == 13) id = (int)c;
                    else {
                        // wrapping adds: far-coordinate cells may sit at the int32 limits
                        const int4 q = make_int4(k.x, (int)((unsigned)k.y + (unsigned)dx), (int)((unsigned)k.z + (unsigned)dy),
                                                 (int)((unsigned)k.w + (unsigned)dz));
                        id = group_lookup(keys, tab, q);
                    }
                    nbr[c * 27 + j] = id;
                    if (id >= 0) total += ccnt[id];
                }
        kc[c] = total;
    }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t *__restrict__ a, int n, uint32_t v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per (sorted position q, neighbour slot j): place point k = sorted_pt[q] into the
// candidate array of the j-th neighbour cell of its own cell
__global__ void __launch_bounds__(256) k_bq_merge(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                  const int32_t *__restrict__ cell, const int32_t *__restrict__ cstart,
                                                  const int32_t *__restrict__ ccnt, const int32_t *__restrict__ nbr,
                                                  const int32_t *__restrict__ cand_start, int32_t n,
                                                  float4 *__restrict__ cand) {
    const int64_t total = (int64_t)n * 27;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(t / 27), j = (int)(t - (int64_t)q * 27);
        const uint32_t k = sorted_pt[q];
        const int s = cell[k];
        const int target = nbr[(int64_t)s * 27 + j];
        if (target < 0) continue;
        const int32_t *tn = nbr + (int64_t)target * 27;
        int pos = 0;
#pragma unroll 1
        for (int jj = 0; jj < 27; jj++) {
            const int src = __ldg(tn + jj);
            if (src < 0) continue;
            if (src == s) pos += q - cstart[s];
            else pos += lower_bound_u32(sorted_pt + cstart[src], ccnt[src], k);
        }
        const float *p = xyz + 3 * (int64_t)k;
        cand[(int64_t)cand_start[target] + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float((int)k));
    }
}

// bfs_cluster.cu:36 as nvcc compiles it (-fmad=true): d2 = fma(dz, dz, fma(dx, dx, dy * dy)), with
// dx = o_x - x etc.; the compare is strict and false for NaN.
__device__ __forceinline__ bool bq_hit(float ox, float oy, float oz, float4 c, float r2) {
    const float dx = __fsub_rn(ox, c.x), dy = __fsub_rn(oy, c.y), dz = __fsub_rn(oz, c.z);
    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    return d2 < r2;
}

__global__ void k_bq_clear_tail(int32_t *kc, const int64_t *__restrict__ nCells, int32_t n1) {
    const int64_t nc = *nCells;
    for (int64_t c = nc + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n1; c += (int64_t)gridDim.x * blockDim.x) kc[c] = 0;
}

// chunks of <= 32 queries per cell
__global__ void k_bq_chunks(const int32_t *__restrict__ ccnt, const int64_t *__restrict__ nCells, int32_t n,
                            int32_t *__restrict__ chunks) {
    const int64_t nc = *nCells;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x)
        chunks[c] = c < nc ? (ccnt[c] + 31) >> 5 : 0;
}

__global__ void __launch_bounds__(256) k_bq_count(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                  const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                  const int32_t *__restrict__ chunk_start, const int64_t *__restrict__ scalars,
                                                  const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kc,
                                                  const float4 *__restrict__ cand, float r2, int32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t nCells = scalars[0], nItems = scalars[4];
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nItems; w += nWarps) {
        // cell owning work item w: largest c with chunk_start[c] <= w
        int64_t lo = 0, hi = nCells;
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (__ldg(chunk_start + mid) <= w) lo = mid; else hi = mid;
        }
        const int c = (int)lo;
        const int first = (int)(w - __ldg(chunk_start + c)) * 32;
        const int nq = min(32, __ldg(ccnt + c) - first);
        const bool on = lane < nq;
        uint32_t k = 0;
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (on) {
            k = sorted_pt[__ldg(cstart + c) + first + lane];
            ox = __ldg(xyz + 3 * (int64_t)k); oy = __ldg(xyz + 3 * (int64_t)k + 1); oz = __ldg(xyz + 3 * (int64_t)k + 2);
        }
        const float4 *cl = cand + __ldg(cand_start + c);
        const int K = __ldg(kc + c);
        int cnt = 0;
        int e = 0;
        for (; e + 4 <= K; e += 4) {
            const float4 c0 = __ldg(cl + e), c1 = __ldg(cl + e + 1), c2 = __ldg(cl + e + 2), c3 = __ldg(cl + e + 3);
            cnt += bq_hit(ox, oy, oz, c0, r2) + bq_hit(ox, oy, oz, c1, r2) + bq_hit(ox, oy, oz, c2, r2) + bq_hit(ox, oy, oz, c3, r2);
        }
        for (; e < K; e++) cnt += bq_hit(ox, oy, oz, __ldg(cl + e), r2);
        if (on) counts[k] = min(cnt, kCap);
    }
}

__global__ void k_bq_start_len(const int32_t *__restrict__ counts, const int32_t *__restrict__ starts, int32_t n,
                               int2 *__restrict__ start_len) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) start_len[i] = make_int2(starts[i], counts[i]);
}

__global__ void __launch_bounds__(256) k_bq_fill(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                 const int32_t *__restrict__ cell, const int32_t *__restrict__ cand_start,
                                                 const int32_t *__restrict__ kc, const float4 *__restrict__ cand,
                                                 const int2 *__restrict__ start_len, float r2, int32_t n,
                                                 int32_t *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n; q += nWarps) {
        const uint32_t k = sorted_pt[q];
        const int len = __ldg(&start_len[k].y);
        if (len == 0) continue;
        const int c = __ldg(cell + k);
        const float ox = __ldg(xyz + 3 * (int64_t)k), oy = __ldg(xyz + 3 * (int64_t)k + 1), oz = __ldg(xyz + 3 * (int64_t)k + 2);
        const float4 *cl = cand + __ldg(cand_start + c);
        const int K = __ldg(kc + c);
        int32_t *out = idx + __ldg(&start_len[k].x);
        int written = 0;
        for (int base = 0; base < K && written < len; base += 32) {
            const int e = base + lane;
            bool hit = false;
            float4 cd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < K) {
                cd = __ldg(cl + e);
                hit = bq_hit(ox, oy, oz, cd, r2);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const int pos = written + __popc(m & lt);
            if (hit && pos < len) out[pos] = __float_as_int(cd.w);
            written += __popc(m);
        }
    }
}

}  // namespace pg

using namespace pg;

extern "C" size_t pg_ballquery_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return bq_layout(nullptr, 0, n).used + 256;
}

extern "C" int pg_ballquery_count(const float *xyz, const int32_t *batch_idxs, const int32_t *batch_offsets, int32_t n,
                                  int32_t B, float radius, int32_t *start_len, void *ws, size_t ws_bytes,
                                  int64_t *host_total, void *stream) {
    (void)batch_offsets; (void)B;   // scene membership comes from batch_idxs
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_total, "null host_total");
    *host_total = 0;
    PG_CHECK_ARG(n >= 0 && n <= (1 << 26), "n out of range (0 .. 2^26)");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(xyz && batch_idxs && start_len && ws, "null pointer");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_count: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }

    const float r2 = radius * radius;
    const double s = fabs((double)radius) * 1.0001;
    const double inv_s = (s > 0.0 && isfinite(s)) ? 1.0 / s : 0.0;   // r = 0 / inf / NaN: one cell per scene
    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 8 * sizeof(int64_t), st));
    k_bq_keys<<<(unsigned)div_up(n, 256), 256, 0, st>>>(xyz, batch_idxs, n, inv_s, w.keys);
    PG_TRY(group_int4(w.keys, n, w.tab, w.pslot, w.cell, w.ccnt, w.scalars, w.scan_tmp, st));
    int bits = 0;
    while ((1ll << bits) < (long long)n) bits++;
    int res = 0;
    PG_TRY(radix_sort_pairs(reinterpret_cast<const uint32_t *>(w.cell), nullptr, w.kA, w.vA, w.kB, w.vB, n, bits,
                            w.hist, w.scan_tmp, st, &res));
    const uint32_t *sorted_pt = res == 0 ? w.vA : w.vB;
    uint32_t *spare = res == 0 ? w.kB : w.kA;   // a free n-sized int array (chunk starts)
    const int64_t flag = res;
    PG_CUDA(cudaMemcpyAsync(w.scalars + 3, &flag, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(w.ccnt + n, 0, sizeof(int32_t), st));
    PG_TRY(scan_exclusive_i32(w.ccnt, w.cstart, (int64_t)n + 1, nullptr, w.scan_tmp, st));
    const unsigned gsm = kNumSM * 8;
    k_bq_neighbours<<<gsm, 256, 0, st>>>(w.keys, w.tab, sorted_pt, w.cstart, w.ccnt, w.scalars, w.nbr, w.kc);
    k_bq_clear_tail<<<gsm, 256, 0, st>>>(w.kc, w.scalars, n + 1);   // kc beyond nCells must scan as 0
    PG_TRY(scan_exclusive_i32(w.kc, w.cand_start, (int64_t)n + 1, w.scalars + 1, w.scan_tmp, st));
    k_bq_merge<<<kNumSM * 16, 256, 0, st>>>(xyz, sorted_pt, w.cell, w.cstart, w.ccnt, w.nbr, w.cand_start, n, w.cand);
    k_bq_chunks<<<gsm, 256, 0, st>>>(w.ccnt, w.scalars, n, (int32_t *)spare);
    PG_TRY(scan_exclusive_i32((int32_t *)spare, (int32_t *)spare, n, w.scalars + 4, w.scan_tmp, st));
    k_bq_count<<<kNumSM * 8, 256, 0, st>>>(xyz, sorted_pt, w.cstart, w.ccnt, (int32_t *)spare, w.scalars, w.cand_start,
                                           w.kc, w.cand, r2, w.counts);
    // starts (reuse pslot) and the interleaved (start, len) rows
    PG_TRY(scan_exclusive_i32(w.counts, w.pslot, n, w.scalars + 2, w.scan_tmp, st));
    k_bq_start_len<<<(unsigned)div_up(n, 256), 256, 0, st>>>(w.counts, w.pslot, n, (int2 *)start_len);
    PG_LAUNCH_CHECK();
    int64_t total = 0;
    PG_CUDA(cudaMemcpyAsync(&total, w.scalars + 2, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *host_total = total;
    if (total > 0x7fffffffLL) {
        set_error("pg_ballquery_count: %lld neighbours do not fit the int32 start offsets of start_len", (long long)total);
        return PG_EOVERFLOW;
    }
    return PG_OK;
}

extern "C" int pg_ballquery_fill(const float *xyz, int32_t n, float radius, const int32_t *start_len, int32_t *idx,
                                 int64_t idx_capacity, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(n >= 0 && idx_capacity >= 0, "negative size");
    if (n == 0 || idx_capacity == 0) return PG_OK;
    PG_CHECK_ARG(xyz && start_len && idx && ws, "null pointer");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_fill: workspace too small"); return PG_EWORKSPACE; }
    int bits = 0;
    while ((1ll << bits) < (long long)n) bits++;
    const int passes = (bits + 7) / 8 < 1 ? 1 : (bits + 7) / 8;
    const uint32_t *sorted_pt = (passes & 1) ? w.vA : w.vB;   // same ping-pong parity as the count phase
    const float r2 = radius * radius;
    k_bq_fill<<<kNumSM * 8, 256, 0, st>>>(xyz, sorted_pt, w.cell, w.cand_start, w.kc, w.cand, (const int2 *)start_len, r2, n, idx);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
