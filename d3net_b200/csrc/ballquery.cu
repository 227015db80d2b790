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
    int32_t *pslot, *cell, *ccnt, *cstart, *kc, *cand_start, *counts, *nbr, *mbase;
    uint32_t *cand_idx;
    uint32_t *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    int64_t *scalars;   // [0] nCells [1] total candidates [2] total neighbours [5] merge tiles [6] mask words
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
    w.cand_idx = a.take<uint32_t>(n * 27);
    w.mbase = a.take<int32_t>(n + 1);
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

// one warp per cell: lane j < 27 looks up neighbour j (-1 when absent; slot 13 is the cell itself),
// the warp sums the candidate count
__global__ void __launch_bounds__(256) k_bq_neighbours(const int4 *__restrict__ keys, GroupTable tab,
                                                       const uint32_t *__restrict__ sorted_pt,
                                                       const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                       const int64_t *__restrict__ nCells, int32_t *__restrict__ nbr,
                                                       int32_t *__restrict__ kc) {
    const int64_t nc = *nCells;
    const int lane = threadIdx.x & 31;
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nc; c += nWarps) {
        const int4 k = keys[sorted_pt[cstart[c]]];
        int cnt = 0;
        if (lane < 27) {
            const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
            int id;
            if (lane == 13) id = (int)c;
            else {
                // wrapping adds: far-coordinate cells may sit at the int32 limits
                const int4 q = make_int4(k.x, (int)((unsigned)k.y + (unsigned)dx), (int)((unsigned)k.z + (unsigned)dy),
                                         (int)((unsigned)k.w + (unsigned)dz));
                id = group_lookup(keys, tab, q);
            }
            nbr[c * 27 + lane] = id;
            if (id >= 0) cnt = ccnt[id];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) kc[c] = cnt;
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

// ---- merge: one candidate array per cell = its <= 27 neighbour lists merged by ascending index ----
// A work item is (cell, tile): cells with <= 32 candidates are one tile; larger cells are cut into
// ~24-element tiles by splitting the cell's index range evenly (point indices are a random shuffle
// with respect to space, so even splits balance).  For a tile [a, b) lane l binary-searches both
// bounds in neighbour list l: the sum of the lower bounds over the lists IS the tile's offset in the
// merged array, so tiles need no scan between them.  The <= 32 keys of a tile are gathered through
// shared memory, sorted with a warp bitonic network and written out with their coordinates.  A tile
// that turns out larger than 32 is halved on a small stack.
constexpr int kMergeTile = 24;

__global__ void k_bq_tiles(const int32_t *__restrict__ kc, const int64_t *__restrict__ nCells, int32_t n,
                           int32_t *__restrict__ tiles) {
    const int64_t nc = *nCells;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const int K = c < nc ? kc[c] : 0;
        tiles[c] = K <= 32 ? (K > 0) : (K + kMergeTile - 1) / kMergeTile;
    }
}

__device__ __forceinline__ uint32_t warp_bitonic_sort(uint32_t key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = ((lane & k) == 0);          // ascending block
            const bool lower = ((lane & j) == 0);       // this lane keeps the smaller of the pair
            const uint32_t mn = min(key, other), mx = max(key, other);
            key = (up == lower) ? mn : mx;
        }
    }
    return key;
}

__global__ void __launch_bounds__(256) k_bq_merge(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                  const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                  const int32_t *__restrict__ nbr, const int32_t *__restrict__ kc,
                                                  const int32_t *__restrict__ cand_start,
                                                  const int32_t *__restrict__ tile_start, const int64_t *__restrict__ scalars,
                                                  float4 *__restrict__ cand, uint32_t *__restrict__ cand_idx) {
    __shared__ uint32_t scratch_all[8][32];
    uint32_t *scratch = scratch_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nCells = scalars[0], nItems = scalars[5];
    // every warp takes a contiguous run of work items: one binary search for the first cell, then it
    // walks forward (every cell has at least one tile, so the walk never skips)
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t per = (nItems + nWarps - 1) / nWarps;
    const int64_t w0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * per;
    const int64_t w1 = w0 + per < nItems ? w0 + per : nItems;
    int c = 0, t = 0;
    if (w0 < w1) {
        int64_t lo_c = 0, hi_c = nCells;
        while (hi_c - lo_c > 1) {
            const int64_t mid = (lo_c + hi_c) >> 1;
            if (__ldg(tile_start + mid) <= w0) lo_c = mid; else hi_c = mid;
        }
        c = (int)lo_c;
        t = (int)(w0 - __ldg(tile_start + c));
    }
    for (int64_t w = w0; w < w1; w++, t++) {
        int K = __ldg(kc + c);
        int T = K <= 32 ? 1 : (K + kMergeTile - 1) / kMergeTile;
        if (t >= T) {
            ++c; t = 0;
            K = __ldg(kc + c);
            T = K <= 32 ? 1 : (K + kMergeTile - 1) / kMergeTile;
        }
        // lane l < 27 owns neighbour list l
        int len = 0;
        const uint32_t *L = sorted_pt;
        if (lane < 27) {
            const int src = __ldg(nbr + (int64_t)c * 27 + lane);
            if (src >= 0) { len = __ldg(ccnt + src); L = sorted_pt + __ldg(cstart + src); }
        }
        const int cbase = __ldg(cand_start + c);
        float4 *dst = cand + cbase;
        uint32_t *dst_idx = cand_idx + cbase;
        uint32_t sa[30], sb[30];      // interval stack (warp-uniform)
        int sp = 0;
        if (T == 1) {
            sa[0] = 0u; sb[0] = 0xffffffffu; sp = 1;
        } else {
            uint32_t head = len ? L[0] : 0xffffffffu, tail = len ? L[len - 1] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                head = min(head, __shfl_xor_sync(0xffffffffu, head, o));
                tail = max(tail, __shfl_xor_sync(0xffffffffu, tail, o));
            }
            const uint32_t span = tail - head + 1u;
            const uint32_t width = (span + (uint32_t)T - 1u) / (uint32_t)T;
            const uint64_t a64 = (uint64_t)head + (uint64_t)t * width;
            if (a64 <= tail) {
                const uint64_t b64 = a64 + width;
                sa[0] = (uint32_t)a64;
                sb[0] = b64 > (uint64_t)tail ? tail + 1u : (uint32_t)b64;      // indices < 2^26: no overflow
                sp = 1;
            }
        }
        while (sp > 0) {
            --sp;
            const uint32_t a = sa[sp], b = sb[sp];
            int lbA = 0, lbB = len;
            if (a != 0u) lbA = lower_bound_u32(L, len, a);
            if (b != 0xffffffffu) lbB = lower_bound_u32(L, len, b);
            const int cnt = lbB - lbA;
            int s = cnt, off = lbA;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                off += __shfl_xor_sync(0xffffffffu, off, o);
            }
            if (s == 0) continue;
            if (s > 32) {     // uneven tile: halve its index interval (indices are unique, so this ends)
                const uint32_t mid = a + ((b - a) >> 1);
                sa[sp] = mid; sb[sp] = b; ++sp;
                sa[sp] = a; sb[sp] = mid; ++sp;
                continue;
            }
            // exclusive prefix of cnt over lanes -> slot of each list's run inside the tile
            int pre = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += v;
            }
            pre -= cnt;
            for (int i = 0; i < cnt; i++) scratch[pre + i] = L[lbA + i];
            __syncwarp();
            uint32_t key = lane < s ? scratch[lane] : 0xffffffffu;
            __syncwarp();
            key = warp_bitonic_sort(key, lane);
            if (lane < s) {
                const float *p = xyz + 3 * (int64_t)key;
                dst[off + lane] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float((int)key));
                dst_idx[off + lane] = key;
            }
        }
        (void)lt;
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

// Hit masks: the count pass already evaluates every (query, candidate) predicate, so it records the
// outcome -- one bit per pair, 32 candidates per word, laid out per cell as [block of 32 candidates]
// [query of the cell] -- and the fill pass turns bits into indices without touching a coordinate.
__global__ void k_bq_mask_sizes(const int32_t *__restrict__ ccnt, const int32_t *__restrict__ kc,
                                const int64_t *__restrict__ nCells, int32_t n1, int32_t *__restrict__ words) {
    const int64_t nc = *nCells;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n1; c += (int64_t)gridDim.x * blockDim.x) {
        long long w = c < nc ? (long long)ccnt[c] * ((kc[c] + 31) >> 5) : 0;
        words[c] = w > 0x7fffffffLL ? 0x7fffffff : (int32_t)w;     // saturates: the int64 total then exceeds the cap
    }
}

// count: one thread per query, taken in cell-sorted order.  When all 32 lanes of a warp sit in the
// same cell (every dense cell), the warp stages the cell's candidate array through shared memory 32
// records at a time -- one coalesced 512-byte load, prefetched one tile ahead -- and every lane reads
// the records back as LDS.128 broadcasts: the loop runs at shared-memory latency instead of waiting
// on a 16-byte global load per candidate.  Mixed warps (sparse regions) walk their own short lists.
template <bool MASK>
__global__ void __launch_bounds__(256) k_bq_count(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                  const int32_t *__restrict__ cell, const int32_t *__restrict__ cstart,
                                                  const int32_t *__restrict__ ccnt, const int32_t *__restrict__ cand_start,
                                                  const int32_t *__restrict__ kc, const float4 *__restrict__ cand,
                                                  const int32_t *__restrict__ mbase, uint32_t *__restrict__ masks,
                                                  float r2, int32_t n, int32_t *__restrict__ counts) {
    __shared__ float4 tile_all[8][32];
    float4 *tile = tile_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t n_up = ((int64_t)n + 31) / 32 * 32;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_up; q += (int64_t)gridDim.x * blockDim.x) {
        const bool live = q < n;
        const uint32_t k = live ? sorted_pt[q] : 0u;
        const int c = live ? __ldg(cell + k) : -1;
        const bool uni = __all_sync(0xffffffffu, c == __shfl_sync(0xffffffffu, c, 0));
        if (!live && !uni) continue;
        if (c < 0) continue;                         // a whole warp past the end
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (live) { ox = __ldg(xyz + 3 * (int64_t)k); oy = __ldg(xyz + 3 * (int64_t)k + 1); oz = __ldg(xyz + 3 * (int64_t)k + 2); }
        const float4 *cl = cand + __ldg(cand_start + c);
        const int K = __ldg(kc + c);
        const int nq = MASK ? __ldg(ccnt + c) : 0;
        uint32_t *mrow = MASK ? masks + __ldg(mbase + c) + (int)(q - __ldg(cstart + c)) : nullptr;
        int cnt = 0;
        if (uni) {
            const float4 pad = make_float4(INFINITY, INFINITY, INFINITY, 0.f);   // never within any radius
            float4 nx = lane < K ? __ldg(cl + lane) : pad;
            for (int base = 0; base < K; base += 32) {
                tile[lane] = nx;
                __syncwarp();
                nx = (base + 32 + lane < K) ? __ldg(cl + base + 32 + lane) : pad;
                unsigned m = 0;
#pragma unroll
                for (int u = 0; u < 32; u++) m |= (unsigned)bq_hit(ox, oy, oz, tile[u], r2) << u;
                cnt += __popc(m);
                if (MASK) mrow[(int64_t)(base >> 5) * nq] = m;
                __syncwarp();
            }
        } else {
            for (int base = 0; base < K; base += 32) {
                const int lim = min(32, K - base);
                unsigned m = 0;
                for (int u = 0; u < lim; u++) m |= (unsigned)bq_hit(ox, oy, oz, __ldg(cl + base + u), r2) << u;
                cnt += __popc(m);
                if (MASK) mrow[(int64_t)(base >> 5) * nq] = m;
            }
        }
        if (live) counts[k] = min(cnt, kCap);
    }
}

__global__ void k_bq_start_len(const int32_t *__restrict__ counts, const int32_t *__restrict__ starts, int32_t n,
                               int2 *__restrict__ start_len) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) start_len[i] = make_int2(starts[i], counts[i]);
}

// fill: a warp streams a cell's candidate array lane-per-candidate (coalesced 512-byte loads), ballots
// the hits and writes them compacted -- ascending by construction -- until the query's count is
// reached.  Four queries that share a cell (the common case wherever it matters: dense cells hold
// hundreds of queries) ride on the same candidate loads.
constexpr int kFillQ = 4;

__device__ __forceinline__ void bq_fill_one(const float *__restrict__ xyz, uint32_t k, const float4 *__restrict__ cl, int K,
                                            const int2 *__restrict__ start_len, float r2, int32_t *__restrict__ idx,
                                            int lane, unsigned lt) {
    const int2 sl = __ldg(start_len + k);
    if (sl.y == 0) return;
    const float ox = __ldg(xyz + 3 * (int64_t)k), oy = __ldg(xyz + 3 * (int64_t)k + 1), oz = __ldg(xyz + 3 * (int64_t)k + 2);
    int32_t *out = idx + sl.x;
    int written = 0;
    for (int base = 0; base < K && written < sl.y; base += 32) {
        const int e = base + lane;
        bool hit = false;
        float4 cd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < K) {
            cd = __ldg(cl + e);
            hit = bq_hit(ox, oy, oz, cd, r2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int pos = written + __popc(m & lt);
        if (hit && pos < sl.y) out[pos] = __float_as_int(cd.w);
        written += __popc(m);
    }
}

__global__ void __launch_bounds__(256) k_bq_fill(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                 const int32_t *__restrict__ cell, const int32_t *__restrict__ cand_start,
                                                 const int32_t *__restrict__ kc, const float4 *__restrict__ cand,
                                                 const int2 *__restrict__ start_len, float r2, int32_t n,
                                                 int32_t *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFillQ; q0 < n; q0 += nWarps * kFillQ) {
        const int nq = (int)((n - q0) < kFillQ ? (n - q0) : kFillQ);
        uint32_t k[kFillQ];
        int c[kFillQ];
        bool same = nq == kFillQ;
#pragma unroll
        for (int u = 0; u < kFillQ; u++) {
            k[u] = u < nq ? sorted_pt[q0 + u] : 0u;
            c[u] = u < nq ? __ldg(cell + k[u]) : -1;
            same = same && c[u] == c[0];
        }
        if (!same) {
            for (int u = 0; u < nq; u++)
                bq_fill_one(xyz, k[u], cand + __ldg(cand_start + c[u]), __ldg(kc + c[u]), start_len, r2, idx, lane, lt);
            continue;
        }
        const int K = __ldg(kc + c[0]);
        const float4 *cp = cand + __ldg(cand_start + c[0]) + lane;
        float ox[kFillQ], oy[kFillQ], oz[kFillQ];
        int wpos[kFillQ], wend[kFillQ];
#pragma unroll
        for (int u = 0; u < kFillQ; u++) {
            const int2 sl = __ldg(start_len + k[u]);
            ox[u] = __ldg(xyz + 3 * (int64_t)k[u]); oy[u] = __ldg(xyz + 3 * (int64_t)k[u] + 1); oz[u] = __ldg(xyz + 3 * (int64_t)k[u] + 2);
            wpos[u] = sl.x; wend[u] = sl.x + sl.y;
        }
        const float4 pad = make_float4(INFINITY, INFINITY, INFINITY, 0.f);   // never within any radius
        float4 nx = lane < K ? __ldg(cp) : pad;
        for (int base = 0; base < K; base += 32) {
            const float4 cd = nx;
            cp += 32;
            nx = (base + 32 + lane < K) ? __ldg(cp) : pad;                   // next tile in flight
#pragma unroll
            for (int u = 0; u < kFillQ; u++) {
                const bool hit = bq_hit(ox[u], oy[u], oz[u], cd, r2);
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                const int pos = wpos[u] + __popc(m & lt);
                if (hit && pos < wend[u]) idx[pos] = __float_as_int(cd.w);
                wpos[u] += __popc(m);
            }
            if (wpos[0] >= wend[0] && wpos[1] >= wend[1] && wpos[2] >= wend[2] && wpos[3] >= wend[3]) break;
        }
    }
}

// fill from hit masks: no coordinates, no predicate -- a word of 32 outcomes per (query, block); the
// lane whose bit is set writes its candidate's index at (hits so far) + (set bits below it).
__device__ __forceinline__ void bq_fill_mask_one(uint32_t k, const uint32_t *__restrict__ ci, int K,
                                                 const uint32_t *__restrict__ mrow, int nq,
                                                 const int2 *__restrict__ start_len, int32_t *__restrict__ idx, int lane,
                                                 unsigned lt) {
    const int2 sl = __ldg(start_len + k);
    int32_t *out = idx + sl.x;
    int written = 0;
    for (int base = 0; base < K && written < sl.y; base += 32) {
        const unsigned m = __ldg(mrow + (int64_t)(base >> 5) * nq);
        if (m == 0u) continue;
        const int pos = written + __popc(m & lt);
        if (((m >> lane) & 1u) && pos < sl.y) out[pos] = (int)__ldg(ci + base + lane);
        written += __popc(m);
    }
}

__global__ void __launch_bounds__(256) k_bq_fill_mask(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                                      const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                      const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kc,
                                                      const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                                      const uint32_t *__restrict__ masks, const int2 *__restrict__ start_len,
                                                      int32_t n, int32_t *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFillQ; q0 < n; q0 += nWarps * kFillQ) {
        const int nqr = (int)((n - q0) < kFillQ ? (n - q0) : kFillQ);
        uint32_t k[kFillQ];
        int c[kFillQ];
        bool same = nqr == kFillQ;
#pragma unroll
        for (int u = 0; u < kFillQ; u++) {
            k[u] = u < nqr ? sorted_pt[q0 + u] : 0u;
            c[u] = u < nqr ? __ldg(cell + k[u]) : -1;
            same = same && c[u] == c[0];
        }
        if (!same) {
            for (int u = 0; u < nqr; u++) {
                const int cc = c[u];
                bq_fill_mask_one(k[u], cand_idx + __ldg(cand_start + cc), __ldg(kc + cc),
                                 masks + __ldg(mbase + cc) + (int)(q0 + u - __ldg(cstart + cc)), __ldg(ccnt + cc), start_len,
                                 idx, lane, lt);
            }
            continue;
        }
        const int cc = c[0];
        const int K = __ldg(kc + cc);
        const int nq = __ldg(ccnt + cc);
        // running pointers, 32-bit output positions: the loop body is ~10 instructions per (query, block)
        const uint32_t *cp = cand_idx + __ldg(cand_start + cc) + lane;
        int wpos[kFillQ], wend[kFillQ];
#pragma unroll
        for (int u = 0; u < kFillQ; u++) {
            const int2 sl = __ldg(start_len + k[u]);
            wpos[u] = sl.x; wend[u] = sl.x + sl.y;
        }
        // Eight blocks per trip: ONE load brings the 8 x 4 mask words (lane = block * 4 + query) and eight
        // independent loads the 256 candidate indices, so nine loads per lane are in flight together
        // and their latency is paid once per eight blocks.
        const int nb = (K + 31) >> 5;
        const uint32_t *mp8 = masks + __ldg(mbase + cc) + (int)(q0 - __ldg(cstart + cc)) + (int64_t)(lane >> 2) * nq + (lane & 3);
        bool done = false;
        for (int b0 = 0; b0 < nb && !done; b0 += 8) {
            const unsigned mw = (b0 + (lane >> 2) < nb) ? __ldg(mp8 + (int64_t)b0 * nq) : 0u;
            int cidr[8];
#pragma unroll
            for (int j = 0; j < 8; j++) cidr[j] = ((b0 + j) * 32 + lane < K) ? (int)__ldg(cp + (b0 + j) * 32) : 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int u = 0; u < kFillQ; u++) {
                    const unsigned m = __shfl_sync(0xffffffffu, mw, j * 4 + u);
                    const int pos = wpos[u] + __popc(m & lt);
                    if (((m >> lane) & 1u) && pos < wend[u]) idx[pos] = cidr[j];
                    wpos[u] += __popc(m);
                }
            }
            done = wpos[0] >= wend[0] && wpos[1] >= wend[1] && wpos[2] >= wend[2] && wpos[3] >= wend[3];
        }
    }
}

}  // namespace pg

using namespace pg;

extern "C" size_t pg_ballquery_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return bq_layout(nullptr, 0, n).used + 256;
}

// ping-pong parity of the radix sort (prepare / count / fill must agree on where sorted_pt landed)
static const uint32_t *bq_sorted(const BqWs &w, int32_t n) {
    int bits = 0;
    while ((1ll << bits) < (long long)n) bits++;
    const int passes = (bits + 7) / 8 < 1 ? 1 : (bits + 7) / 8;
    return (passes & 1) ? w.vA : w.vB;
}

extern "C" int pg_ballquery_prepare(const float *xyz, const int32_t *batch_idxs, const int32_t *batch_offsets, int32_t n,
                                    int32_t B, float radius, void *ws, size_t ws_bytes, int64_t *host_mask_words,
                                    void *stream) {
    (void)batch_offsets; (void)B;   // scene membership comes from batch_idxs
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_mask_words, "null host_mask_words");
    *host_mask_words = 0;
    PG_CHECK_ARG(n >= 0 && n <= (1 << 26), "n out of range (0 .. 2^26)");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(xyz && batch_idxs && ws, "null pointer");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_prepare: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }

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
    uint32_t *spare2 = res == 0 ? w.kA : w.kB;  // after the sort the key buffers are free: merge tile starts
    PG_CUDA(cudaMemsetAsync(w.ccnt + n, 0, sizeof(int32_t), st));
    PG_TRY(scan_exclusive_i32(w.ccnt, w.cstart, (int64_t)n + 1, nullptr, w.scan_tmp, st));
    const unsigned gsm = kNumSM * 8;
    k_bq_neighbours<<<kNumSM * 16, 256, 0, st>>>(w.keys, w.tab, sorted_pt, w.cstart, w.ccnt, w.scalars, w.nbr, w.kc);
    k_bq_clear_tail<<<gsm, 256, 0, st>>>(w.kc, w.scalars, n + 1);   // kc beyond nCells must scan as 0
    PG_TRY(scan_exclusive_i32(w.kc, w.cand_start, (int64_t)n + 1, w.scalars + 1, w.scan_tmp, st));
    k_bq_tiles<<<gsm, 256, 0, st>>>(w.kc, w.scalars, n, (int32_t *)spare2);
    PG_TRY(scan_exclusive_i32((int32_t *)spare2, (int32_t *)spare2, n, w.scalars + 5, w.scan_tmp, st));
    k_bq_merge<<<kNumSM * 8, 256, 0, st>>>(xyz, sorted_pt, w.cstart, w.ccnt, w.nbr, w.kc, w.cand_start, (int32_t *)spare2,
                                           w.scalars, w.cand, w.cand_idx);
    k_bq_mask_sizes<<<gsm, 256, 0, st>>>(w.ccnt, w.kc, w.scalars, n + 1, w.mbase);
    PG_TRY(scan_exclusive_i32(w.mbase, w.mbase, (int64_t)n + 1, w.scalars + 6, w.scan_tmp, st));
    PG_LAUNCH_CHECK();
    int64_t words = 0;
    PG_CUDA(cudaMemcpyAsync(&words, w.scalars + 6, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *host_mask_words = words >= 0x7fffffffLL ? -1 : words;   // -1: too many for int32 bases, run without masks
    return PG_OK;
}

extern "C" int pg_ballquery_count(const float *xyz, int32_t n, float radius, int32_t *start_len, uint32_t *masks,
                                  int64_t mask_words, void *ws, size_t ws_bytes, int64_t *host_total, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_total, "null host_total");
    *host_total = 0;
    PG_CHECK_ARG(n >= 0 && n <= (1 << 26), "n out of range (0 .. 2^26)");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(xyz && start_len && ws, "null pointer");
    PG_CHECK_ARG(masks == nullptr || mask_words >= 0, "negative mask_words");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_count: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    const uint32_t *sorted_pt = bq_sorted(w, n);
    const float r2 = radius * radius;
    const unsigned grid = (unsigned)div_up(n, 256);
    if (masks) k_bq_count<true><<<grid, 256, 0, st>>>(xyz, sorted_pt, w.cell, w.cstart, w.ccnt, w.cand_start, w.kc, w.cand, w.mbase, masks, r2, n, w.counts);
    else k_bq_count<false><<<grid, 256, 0, st>>>(xyz, sorted_pt, w.cell, w.cstart, w.ccnt, w.cand_start, w.kc, w.cand, w.mbase, nullptr, r2, n, w.counts);
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

extern "C" int pg_ballquery_fill(const float *xyz, int32_t n, float radius, const int32_t *start_len,
                                 const uint32_t *masks, int32_t *idx, int64_t idx_capacity, void *ws, size_t ws_bytes,
                                 void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(n >= 0 && idx_capacity >= 0, "negative size");
    if (n == 0 || idx_capacity == 0) return PG_OK;
    PG_CHECK_ARG(xyz && start_len && idx && ws, "null pointer");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_fill: workspace too small"); return PG_EWORKSPACE; }
    const uint32_t *sorted_pt = bq_sorted(w, n);
    const float r2 = radius * radius;
    if (masks)
        k_bq_fill_mask<<<kNumSM * 8, 256, 0, st>>>(sorted_pt, w.cell, w.cstart, w.ccnt, w.cand_start, w.kc, w.cand_idx, w.mbase,
                                                   masks, (const int2 *)start_len, n, idx);
    else
        k_bq_fill<<<kNumSM * 8, 256, 0, st>>>(xyz, sorted_pt, w.cell, w.cand_start, w.kc, w.cand, (const int2 *)start_len, r2, n, idx);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
