// ballquery_batch_p: uniform-grid radius search with the reference's exact fp32 predicate and its
// "first 1000 neighbours by ascending index" rule.  Reference behaviour:
// lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.cu:15-90 (brute-force scan of the whole scene).
//
// Pipeline (all on `stream`; the host only reads two sizes):
//   prepare
//     1. cell key per point: (scene, floor(x/s), floor(y/s), floor(z/s)), s = 1.0001 * |r| in fp64, so
//        every pair that can pass the predicate sits in adjacent cells;
//     2. hash-group the keys -> dense cell ids; a counting scatter puts every cell's points side by side (in no
//        particular order inside a cell: the count kernels order the candidates themselves); that order is also the
//        QUERY order;
//     3. per cell: its 27 neighbour cells (hash lookups), K = size of its candidate set, and the
//        cell's class: K <= kSmallK "sparse" (one warp), otherwise "dense" (one block);
//   count -- one pass over the cells, each cell builds its candidate set ONCE, merged in ascending
//     original index, and tests all of its queries against it:
//       sparse cell: the <= 128 candidate indices are sorted in registers (warp bitonic network);
//       dense cell:  rank-select through a shared-memory bitmap over the index range (set one bit per
//                    candidate, prefix-count the words, walk the bits in order), 1024 candidates at a
//                    time into a shared-memory tile of (x, y, z); queries are lanes, candidates are
//                    LDS.128 broadcasts; tiles stop as soon as every query of the cell holds 1000 hits;
//     every predicate outcome is recorded as one bit (32 candidates per word) and the sorted candidate
//     indices are kept (4 bytes each), so
//   fill turns bits into indices without touching a coordinate.
//   Segments of `idx` are laid out in query (= cell) order: neighbouring points own neighbouring segments, which is
//   what makes the consumer (bfs_cluster) cache-friendly.  Cells are numbered as they are claimed (roughly by first
//   appearance; PG_BQ_ORDERED_CELLS=1: exactly), points inside a cell in claim order: the PLACEMENT of the segments may
//   differ from run to run -- as the reference's does (atomicAdd, bfs_cluster.cu:47) -- every list's content is fixed.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "ballquery.cuh"
#include "scan.cuh"

namespace pg {

constexpr int kCap = PG_BALLQUERY_CAP;
constexpr int kSmallK = 256;             // sparse/dense class boundary (candidates per cell): one warp up to here
constexpr int kSmallSplit = 128;         // the warp kernel runs as two instantiations, K <= 128 (<= 4 keys per lane)
                                         // and 128 < K <= 256 (8 keys per lane, twice the registers)
constexpr int kMediumQ = 16;             // ... the latter only for cells with few queries: one warp tests them one
                                         // after the other, a block shares them out (surfaces of a 1M-point scene:
                                         // ~9 queries on ~150 candidates per cell, 100k such cells -> warp; the blobs of
                                         // shifted coordinates: 50+ queries per cell -> block)

// the class of a cell: served by the block-per-cell kernel?
__host__ __device__ __forceinline__ bool bq_is_dense(int K, int nq) { return K > kSmallK || (K > kSmallSplit && nq > kMediumQ); }
constexpr int kChunkBlocks = 4;          // 32-candidate blocks per (query group, chunk) work item

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
                          double inv_s, int4 *__restrict__ keys, Fill table, Fill cnt) {
    pdl_enter();
    grid_fill(table);      // the grouping's hash table and counts start clean
    grid_fill(cnt);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(xyz + 3 * (int64_t)i), y = __ldg(xyz + 3 * (int64_t)i + 1), z = __ldg(xyz + 3 * (int64_t)i + 2);
    keys[i] = make_int4(__ldg(batch_idxs + i), cell_coord(x, inv_s), cell_coord(y, inv_s), cell_coord(z, inv_s));
}

// points side by side per cell: slot claimed from the cell's cursor; the cell's smallest / largest point index on the way
__global__ void k_bq_scatter(const int32_t *__restrict__ cell, const int32_t *__restrict__ cstart, int32_t n,
                             int32_t *cursor, uint32_t *__restrict__ sorted_pt, uint32_t *cmin, uint32_t *cmax) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = __ldg(cell + i);
    sorted_pt[__ldg(cstart + c) + atomicAdd(cursor + c, 1)] = (uint32_t)i;
    // the smallest index is kept as the largest complement, so that all three arrays start from zero (one memset)
    if (cmin[c] < ~(uint32_t)i) atomicMax(cmin + c, ~(uint32_t)i);
    if (cmax[c] < (uint32_t)i) atomicMax(cmax + c, (uint32_t)i);
}

// (A thread per cell -- 5x fewer warp instructions, three first probes in flight per thread, ids staged in shared memory
// and written as one coalesced block per warp -- was measured at 0.42 ms against 0.23: 27 dependent probes per thread are
// a longer chain than the machine has threads to hide.)
// one warp per cell: lane j < 27 looks up neighbour j (-1 when absent; slot 13 is the cell itself),
// the warp sums the candidate count and files dense cells in the (unordered) dense list
__global__ void __launch_bounds__(256) k_bq_neighbours(const int4 *__restrict__ keys, GroupTable tab,
                                                       const uint32_t *__restrict__ sorted_pt,
                                                       const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                       int64_t *scalars, int32_t *__restrict__ nbr,
                                                       int32_t *__restrict__ kc, int32_t *__restrict__ dense,
                                                       const uint32_t *__restrict__ cmin, const uint32_t *__restrict__ cmax,
                                                       uint2 *__restrict__ crange) {
    pdl_enter();
    const int64_t nc = scalars[0];
    const int lane = threadIdx.x & 31;
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nc; c += nWarps) {
        const int4 k = keys[sorted_pt[cstart[c]]];
        int cnt = 0, id = -1;
        if (lane < 27) {
            const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
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
        if (lane == 0) {
            kc[c] = cnt;
            if (bq_is_dense(cnt, ccnt[c])) dense[atomicAdd((unsigned long long *)&scalars[3], 1ULL)] = (int32_t)c;
        }
        if (bq_is_dense(cnt, ccnt[c])) {          // a dense cell: the index range of its candidates
            uint32_t head = 0xffffffffu, tail = 0u;
            if (id >= 0) { head = ~__ldg(cmin + id); tail = __ldg(cmax + id); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                head = min(head, __shfl_xor_sync(0xffffffffu, head, o));
                tail = max(tail, __shfl_xor_sync(0xffffffffu, tail, o));
            }
            if (lane == 0) crange[c] = make_uint2(head, tail);
        }
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

// bfs_cluster.cu:36 as nvcc compiles it (-fmad=true): d2 = fma(dz, dz, fma(dx, dx, dy * dy)), with
// dx = o_x - x etc.; the compare is strict and false for NaN.
__device__ __forceinline__ bool bq_hit(float ox, float oy, float oz, float4 c, float r2) {
    const float dx = __fsub_rn(ox, c.x), dy = __fsub_rn(oy, c.y), dz = __fsub_rn(oz, c.z);
    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    return d2 < r2;
}

// ---- scans with their producers / consumers fused in (scan.cuh) -----------------------------------------------------
// candidate-list starts: the scan runs over n + 1 entries, kc beyond nCells counts (and is left) as 0
struct CandLoad {
    const int32_t *kc;
    const int64_t *nCells;
    __device__ int operator()(int64_t c) const { return c < *nCells ? kc[c] : 0; }
};
struct CandStore {
    int32_t *kc, *cand_start;
    const int64_t *nCells;
    __device__ void operator()(int64_t c, int start, int) const {
        cand_start[c] = start;
        if (c >= *nCells) kc[c] = 0;
    }
};
// Hit masks: one bit per (query, candidate) pair, 32 candidates per word, laid out per cell as
// [block of 32 candidates][query of the cell].
struct MaskLoad {
    const int32_t *ccnt, *kc;
    const int64_t *nCells;
    __device__ int operator()(int64_t c) const {
        const long long w = c < *nCells ? (long long)ccnt[c] * ((kc[c] + 31) >> 5) : 0;
        return w > 0x7fffffffLL ? 0x7fffffff : (int32_t)w;       // saturates: the int64 total then exceeds the cap
    }
};
struct MaskStore {
    int32_t *mbase;
    __device__ void operator()(int64_t c, int start, int) const { mbase[c] = start; }
};
// start_len rows from the per-query counts (query order) and their exclusive scan
struct CountLoad {
    const int32_t *counts;
    __device__ int operator()(int64_t q) const { return counts[q]; }
};
struct StartLenStore {
    const uint32_t *sorted_pt;
    int2 *start_len;
    int32_t *qpos;
    __device__ void operator()(int64_t q, int start, int len) const {
        const uint32_t k = sorted_pt[q];
        start_len[k] = make_int2(start, len);
        qpos[k] = (int32_t)q;
    }
};

// ---- sparse cells: one warp per cell, candidates sorted in registers ------------------------------
// Striped layout: register r of lane l holds position r * 32 + l, so after the sort register r IS
// 32-candidate block r and a ballot over it IS the mask word.
template <int KP>
__device__ __forceinline__ void warp_sort_striped(uint32_t (&key)[KP], int lane) {
    constexpr int N = 32 * KP;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < KP; r++) {
                    if ((r & jr) == 0) {
                        const int r2 = r | jr;
                        const bool up = (((r << 5) | lane) & k) == 0;
                        const uint32_t a = key[r], b = key[r2];
                        const uint32_t mn = min(a, b), mx = max(a, b);
                        key[r] = up ? mn : mx;
                        key[r2] = up ? mx : mn;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < KP; r++) {
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const bool up = (((r << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    const uint32_t mn = min(key[r], other), mx = max(key[r], other);
                    key[r] = (up == lower) ? mn : mx;
                }
            }
        }
    }
}

template <int KP>
__device__ __forceinline__ void bq_small_cell(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                              const uint32_t *scratch, int K, int nq, int qs, int cbase, int mb,
                                              uint32_t *__restrict__ masks, uint32_t *__restrict__ cand_idx,
                                              int32_t *__restrict__ counts, float r2, int lane) {
    uint32_t key[KP];
    float cx[KP], cy[KP], cz[KP];
#pragma unroll
    for (int r = 0; r < KP; r++) key[r] = (r * 32 + lane < K) ? scratch[r * 32 + lane] : 0xffffffffu;
    warp_sort_striped<KP>(key, lane);
#pragma unroll
    for (int r = 0; r < KP; r++) {
        const bool valid = r * 32 + lane < K;
        cx[r] = cy[r] = cz[r] = INFINITY;                        // padding never passes the predicate
        if (valid) {
            const float *p = xyz + 3 * (int64_t)key[r];
            cx[r] = __ldg(p); cy[r] = __ldg(p + 1); cz[r] = __ldg(p + 2);
            cand_idx[cbase + r * 32 + lane] = key[r];
        }
    }
    const int nblk = (K + 31) >> 5;
    for (int qi = 0; qi < nq; qi++) {
        const uint32_t k = __ldg(sorted_pt + qs + qi);            // warp-uniform
        const float ox = __ldg(xyz + 3 * (int64_t)k), oy = __ldg(xyz + 3 * (int64_t)k + 1), oz = __ldg(xyz + 3 * (int64_t)k + 2);
        int cnt = 0;
        unsigned mine = 0;
#pragma unroll
        for (int r = 0; r < KP; r++) {
            const unsigned m = __ballot_sync(0xffffffffu, bq_hit(ox, oy, oz, make_float4(cx[r], cy[r], cz[r], 0.f), r2));
            cnt += __popc(m);
            if (lane == r) mine = m;
        }
        if (masks && lane < nblk) masks[mb + lane * nq + qi] = mine;
        if (lane == 0) counts[qs + qi] = cnt;                     // <= kSmallK < kCap
    }
}

template <int LO, int HI>
__global__ void __launch_bounds__(256) k_bq_cells_small(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                        const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                        const int32_t *__restrict__ nbr, const int32_t *__restrict__ kc,
                                                        const int32_t *__restrict__ cand_start,
                                                        const int32_t *__restrict__ mbase, const int64_t *__restrict__ scalars,
                                                        uint32_t *__restrict__ masks, int64_t mask_capacity, float r2,
                                                        uint32_t *__restrict__ cand_idx, int32_t *__restrict__ counts,
                                                        int32_t *__restrict__ kb) {
    pdl_enter();
    __shared__ uint32_t scratch_all[8][HI];
    uint32_t *scratch = scratch_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t nCells = scalars[0];
    if (scalars[6] > mask_capacity) masks = nullptr;            // the caller's mask buffer is too small: run without
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    auto serve = [&](int64_t c, int K) {
        // lane j < 27 owns neighbour list j and copies it to its slot of the scratch row
        int len = 0;
        const uint32_t *L = sorted_pt;
        if (lane < 27) {
            const int src = __ldg(nbr + c * 27 + lane);
            if (src >= 0) { len = __ldg(ccnt + src); L = sorted_pt + __ldg(cstart + src); }
        }
        int pre = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += v;
        }
        pre -= len;
        for (int i = 0; i < len; i++) scratch[pre + i] = __ldg(L + i);
        __syncwarp();
        const int nq = __ldg(ccnt + c), qs = __ldg(cstart + c), cbase = __ldg(cand_start + c);
        const int mb = masks ? __ldg(mbase + c) : 0;
        if (HI > kSmallSplit) bq_small_cell<HI / 32>(xyz, sorted_pt, scratch, K, nq, qs, cbase, mb, masks, cand_idx, counts, r2, lane);
        else if (K <= 32) bq_small_cell<1>(xyz, sorted_pt, scratch, K, nq, qs, cbase, mb, masks, cand_idx, counts, r2, lane);
        else if (K <= 64) bq_small_cell<2>(xyz, sorted_pt, scratch, K, nq, qs, cbase, mb, masks, cand_idx, counts, r2, lane);
        else bq_small_cell<4>(xyz, sorted_pt, scratch, K, nq, qs, cbase, mb, masks, cand_idx, counts, r2, lane);
        if (lane == 0) kb[c] = K;
        __syncwarp();
    };
    if (LO == 0) {
        // nearly every cell of a sparse set is this instantiation's: one warp per cell, strided
        for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nCells; c += nWarps) {
            const int K = __ldg(kc + c);
            if (K > HI || bq_is_dense(K, __ldg(ccnt + c))) continue;
            serve(c, K);
        }
    } else {
        // few cells are: a warp looks at 32 consecutive cells at once (lane = cell) and serves the ones of its class
        // one by one -- skipping the others costs one coalesced load per 32 cells, not a dependent load per cell
        for (int64_t c0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; c0 < nCells; c0 += nWarps * 32) {
            const int Kl = (c0 + lane < nCells) ? __ldg(kc + c0 + lane) : 0;
            const bool wanted = Kl > LO && Kl <= HI && !bq_is_dense(Kl, __ldg(ccnt + c0 + lane));
            unsigned todo = __ballot_sync(0xffffffffu, wanted);
            while (todo) {
                const int bit = __ffs((int)todo) - 1;
                todo &= todo - 1u;
                serve(c0 + bit, __shfl_sync(0xffffffffu, Kl, bit));
            }
        }
    }
}

// ---- dense cells ----------------------------------------------------------------------------------------
// Two kernels.  MERGE (a warp per cell, no block barrier anywhere): the cell's 27 ascending point lists are merged
// through a small per-warp bitmap, window by window over the index range, into the cell's candidate list in
// ascending original index; the indices go to cand_idx (the fill phase reads them), the coordinates to cand_xy /
// cand_z in the pair layout of the packed predicate.  TEST (a block per cell, persistent, work claimed from a
// counter): candidate tiles arrive in shared memory by bulk asynchronous copy (cp.async.bulk, completion on an
// mbarrier; the next tile is in flight while the current one is tested), lanes are queries, candidates are
// shared-memory broadcasts, two candidates per FADD2 / FMUL2 / FFMA2.
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int kMergeThreads = 256;
constexpr int kWinBits = 160 * 1024;       // index window covered by the rank bitmap (a 150k-point scene in one)
constexpr int kWinWords = kWinBits / 32;
constexpr int kWordsPerThread = kWinWords / kMergeThreads;     // 20: every thread owns a run of consecutive words
static_assert(kWordsPerThread % 4 == 0 && kWordsPerThread * kMergeThreads == kWinWords, "window / block shape");
constexpr int kTestThreads = 128;
constexpr int kTestTile = 512;             // candidates per shared-memory tile
constexpr int kTestQ = 256;                // queries of one cell handled per pass

struct MergeSmem {
    uint32_t bm[kWinWords];          // rank bitmap over the index window [s0, s0 + kWinBits)
    const uint32_t *lptr[27];
    int32_t llen[27];
    int32_t wsum[kMergeThreads / 32];
    uint32_t lo, hi;
    int32_t cellslot;
    long long cb;
};

// MERGE: rank-select through a shared-memory bitmap.  One bit per candidate over the cell's index window, a block
// scan of the per-thread popcounts, then every thread walks its words and emits the set bits in order -- ascending
// candidates without a comparison.  The indices go straight to cand_idx; a second, lane-dense sweep fetches the
// coordinates (all loads of a warp in flight together) and writes them in the pair layout.
__global__ void __launch_bounds__(kMergeThreads) k_bq_merge_dense(
    const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cstart,
    const int32_t *__restrict__ ccnt, const int32_t *__restrict__ nbr, const int32_t *__restrict__ kc,
    const int32_t *__restrict__ cand_start, const int32_t *__restrict__ dense, const uint2 *__restrict__ crange,
    int64_t *scalars, uint32_t *cand_idx, float4 *__restrict__ cand_xy, float2 *__restrict__ cand_z,
    int32_t *__restrict__ dbase) {
    pdl_enter();
    __shared__ __align__(16) MergeSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nDense = scalars[3];
    float *xy_f = reinterpret_cast<float *>(cand_xy);
    float *z_f = reinterpret_cast<float *>(cand_z);
    if (tid == 0) S.cellslot = (int32_t)atomicAdd((unsigned long long *)&scalars[5], 1ULL);
    for (;;) {
        __syncthreads();
        const int64_t slot = S.cellslot;
        if (slot >= nDense) break;
        int32_t next_slot = 0;                                   // claimed now, published after this cell's work
        if (tid == 0) next_slot = (int32_t)atomicAdd((unsigned long long *)&scalars[5], 1ULL);
        const int c = __ldg(dense + slot);
        const int K = __ldg(kc + c), cbase = __ldg(cand_start + c);
        const int Kpad = (K + 31) & ~31;
        if (warp == 0) {
            int len = 0;
            const uint32_t *L = sorted_pt;
            if (lane < 27) {
                const int src = __ldg(nbr + (int64_t)c * 27 + lane);
                if (src >= 0) { len = __ldg(ccnt + src); L = sorted_pt + __ldg(cstart + src); }
                S.lptr[lane] = L;
                S.llen[lane] = len;
            }
            if (lane == 0) {                                     // where this cell's coordinates go (a multiple of 32)
                const uint2 rg = __ldg(crange + c);
                S.lo = rg.x;
                S.hi = rg.y;
                const long long cb = (long long)atomicAdd((unsigned long long *)&scalars[7], (unsigned long long)Kpad);
                dbase[slot] = (int32_t)(uint32_t)cb;             // < 34 * 2^26 < 2^32
                S.cb = cb;
            }
        }
        __syncthreads();
        const long long cb = S.cb;
        int rank_base = 0;                                       // candidates emitted by earlier windows
        // windows of kWinBits indices over [lo, hi] (one, unless a scene holds more than 160k points).  The lists are in
        // no particular order: a window takes the entries that fall into it.
        const uint32_t hi = S.hi;
        const bool one_window = hi - (S.lo & ~31u) < (uint32_t)kWinBits;
        for (uint32_t s0 = S.lo & ~31u;; s0 += (uint32_t)kWinBits) {
            // the bitmap is all zero here (cleared below by the threads that walked it)
            if (rank_base == 0 && s0 == (S.lo & ~31u)) {
                uint4 *z = reinterpret_cast<uint4 *>(S.bm + tid * kWordsPerThread);
#pragma unroll
                for (int k = 0; k < kWordsPerThread / 4; k++) z[k] = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
            }
            for (int j = warp; j < 27; j += kMergeThreads / 32) {
                const uint32_t *L = S.lptr[j];
                const int b1 = S.llen[j];
                for (int t = lane; t < b1; t += 32) {
                    const uint32_t v = __ldg(L + t) - s0;
                    if (one_window || v < (uint32_t)kWinBits) atomicOr(&S.bm[v >> 5], 1u << (v & 31u));
                }
            }
            __syncthreads();
            // ---- popcount of the thread's run of words, block scan -> rank of its first bit
            uint4 *run = reinterpret_cast<uint4 *>(S.bm + tid * kWordsPerThread);
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < kWordsPerThread / 4; k++) {
                const uint4 q = run[k];
                cnt += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) S.wsum[warp] = inc;
            __syncthreads();
            int before = 0, tw = 0;
#pragma unroll
            for (int i = 0; i < kMergeThreads / 32; i++) {
                const int t = S.wsum[i];
                if (i < warp) before += t;
                tw += t;
            }
            // ---- walk: emit the set bits of the run in order, clear the words for the next window / cell
            if (cnt) {
                uint32_t *out = cand_idx + cbase + rank_base + before + inc - cnt;
                const uint32_t base_id = s0 + (uint32_t)(tid * kWordsPerThread << 5);
#pragma unroll
                for (int k = 0; k < kWordsPerThread / 4; k++) {
                    const uint4 q = run[k];
                    const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        uint32_t word = wv[u];
                        while (word) {
                            const int b = __ffs((int)word) - 1;
                            word &= word - 1u;
                            *out++ = base_id + (uint32_t)((k * 4 + u) << 5) + (uint32_t)b;
                        }
                    }
                    run[k] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            __syncthreads();                                     // the window's indices are in cand_idx (same block: visible)
            // ---- coordinates, lane-dense
            for (int e = tid; e < tw; e += kMergeThreads) {
                const int pos = rank_base + e;
                const uint32_t id = __ldcg(cand_idx + cbase + pos);
                const float *p = xyz + 3 * (int64_t)id;
                const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
                const long long q = cb + pos;
                float *pxy = xy_f + ((q >> 1) << 2) + (q & 1);
                pxy[0] = x;
                pxy[2] = y;
                z_f[q] = z;
            }
            rank_base += tw;
            if (hi - s0 < (uint32_t)kWinBits) break;             // the window held the largest index
        }
        // padding up to a multiple of 32: +inf never passes the predicate
        if (tid < Kpad - K) {
            const long long q = cb + K + tid;
            float *pxy = xy_f + ((q >> 1) << 2) + (q & 1);
            pxy[0] = INFINITY;
            pxy[2] = INFINITY;
            z_f[q] = INFINITY;
        }
        if (tid == 0) S.cellslot = next_slot;
    }
}

// mbarrier / bulk-copy helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TestSmem {
    float4 txy[2][kTestTile / 2];     // two tiles in flight: pair p = candidates (2p, 2p+1): (x0, x1, y0, y1)
    float2 tz[2][kTestTile / 2];      //                                                        (z0, z1)
    float qx[kTestQ], qy[kTestQ], qz[kTestQ];
    int32_t qcnt[kTestQ];             // hits so far per query
    uint8_t qsat[kTestQ];             // snapshot at the last tile boundary: the query already holds kCap hits
    uint64_t full[2];                 // mbarriers: tile landed
    int32_t cellslot;
};

// One work item: query group g (lane l tests queries g*128 + j*32 + l, j < Q) against the 32-candidate
// blocks [b0, b1) of the tile in `buf`.  A candidate pair is read from shared memory ONCE (LDS.128 + LDS.64, broadcast)
// and tested against the lane's Q queries: the shared-memory pipe delivers 512 B per LDS.128 whatever the broadcast, so
// at Q = 1 it, not the FP32 pipe, bounds the kernel (6 cycles per pair against 2.25 issue cycles).
template <int Q>
__device__ __forceinline__ void bq_test_item(TestSmem &S, int buf, int g, int b0, int b1, int nqs, int nq, int sg0, int64_t mrow0,
                                             uint32_t *__restrict__ masks, int thr, bool can_saturate, int lane) {
    float ox[Q], oy[Q], oz[Q];
    bool live[Q];
    int cnt[Q];
    bool any_unsat = false;
#pragma unroll
    for (int j = 0; j < Q; j++) {
        const int qi = (g << 7) + j * 32 + lane;
        live[j] = qi < nqs;
        ox[j] = live[j] ? S.qx[qi] : NAN; oy[j] = live[j] ? S.qy[qi] : NAN; oz[j] = live[j] ? S.qz[qi] : NAN;
        cnt[j] = 0;
        any_unsat |= live[j] && !S.qsat[qi];
    }
    if (can_saturate && !__any_sync(0xffffffffu, any_unsat)) return;   // all of them already hold kCap hits
    for (int b = b0; b < b1; b++) {
        const float4 *pxy = S.txy[buf] + (b << 4);
        const float2 *pz = S.tz[buf] + (b << 4);
        unsigned m[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) m[j] = 0u;
        // Two candidates per instruction (sm_100 packed fp32: FADD2 / FMUL2 / FFMA2, each half rounded to nearest
        // exactly like the scalar op): d2 = fma(dz, dz, fma(dx, dx, dy * dy)) as compiled from bfs_cluster.cu:36.
        // hit <=> d2 < r2.  d2 is a sum of squares: +0 .. +inf or the canonical NaN (0x7fffffff), so the float compare
        // is the signed compare of the bit patterns and its outcome the sign bit of bits(d2) - bits(r2) -- taken on the
        // integer pipe, which the packed FP32 work leaves idle.
#pragma unroll
        for (int u = 15; u >= 0; u--) {           // candidate 31 first: it ends up in bit 31
            const float4 cxy = pxy[u];
            const float2 cz = pz[u];
            const float2 ncx = make_float2(-cxy.x, -cxy.y), ncy = make_float2(-cxy.z, -cxy.w), ncz = make_float2(-cz.x, -cz.y);
#pragma unroll
            for (int j = 0; j < Q; j++) {
                const float2 dx = __fadd2_rn(make_float2(ox[j], ox[j]), ncx);
                const float2 dy = __fadd2_rn(make_float2(oy[j], oy[j]), ncy);
                const float2 dz = __fadd2_rn(make_float2(oz[j], oz[j]), ncz);
                const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
                m[j] = __funnelshift_l((unsigned)(__float_as_int(d2.y) - thr), m[j], 1);
                m[j] = __funnelshift_l((unsigned)(__float_as_int(d2.x) - thr), m[j], 1);
            }
        }
#pragma unroll
        for (int j = 0; j < Q; j++) {
            cnt[j] += __popc(m[j]);
            if (masks && live[j]) masks[mrow0 + (int64_t)b * nq + sg0 + (g << 7) + j * 32 + lane] = m[j];
        }
    }
#pragma unroll
    for (int j = 0; j < Q; j++)
        if (live[j] && cnt[j]) atomicAdd(&S.qcnt[(g << 7) + j * 32 + lane], cnt[j]);
}

__global__ void __launch_bounds__(kTestThreads, 8) k_bq_test_dense(
    const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cstart,
    const int32_t *__restrict__ ccnt, const int32_t *__restrict__ kc, const int32_t *__restrict__ mbase,
    const int32_t *__restrict__ dense, const int32_t *__restrict__ dbase, const float4 *__restrict__ cand_xy,
    const float2 *__restrict__ cand_z, int64_t *scalars, uint32_t *__restrict__ masks, int64_t mask_capacity, float r2,
    int32_t *__restrict__ counts, int32_t *__restrict__ kb) {
    pdl_enter();
    __shared__ __align__(128) TestSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nDense = scalars[3];
    if (scalars[6] > mask_capacity) masks = nullptr;            // the caller's mask buffer is too small: run without
    const int thr = (r2 == r2) ? __float_as_int(r2) : 0;         // NaN radius: nothing is a neighbour
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.cellslot = (int32_t)atomicAdd((unsigned long long *)&scalars[4], 1ULL);
    }
    uint32_t uses0 = 0, uses1 = 0;                               // completed uses of each buffer (its mbarrier's phase)
    for (;;) {
        __syncthreads();
        const int64_t slot = S.cellslot;
        if (slot >= nDense) break;
        int32_t next_slot = 0;                                   // claimed now, published after this cell's work
        if (tid == 0) next_slot = (int32_t)atomicAdd((unsigned long long *)&scalars[4], 1ULL);
        const int c = __ldg(dense + slot);
        const int K = __ldg(kc + c), nq = __ldg(ccnt + c), qs = __ldg(cstart + c);
        const int64_t cb = (int64_t)(uint32_t)__ldg(dbase + slot);
        const int64_t mb = masks ? __ldg(mbase + c) : 0;
        const int Kpad = (K + 31) & ~31;
        const int ntiles = (Kpad + kTestTile - 1) / kTestTile;
        const bool can_saturate = K > kCap;
        auto issue = [&](int t, int buf) {                       // thread 0: start the copy of tile t into buffer buf
            const int n = min(kTestTile, Kpad - t * kTestTile);  // a multiple of 32 candidates
            const int64_t q0 = cb + (int64_t)t * kTestTile;      // a multiple of 32
            mbar_expect_tx(&S.full[buf], (uint32_t)n * 12u);
            bulk_g2s(S.txy[buf], cand_xy + (q0 >> 1), (uint32_t)n * 8u, &S.full[buf]);
            bulk_g2s(S.tz[buf], cand_z + (q0 >> 1), (uint32_t)n * 4u, &S.full[buf]);
        };
        int built_max = 0;
        for (int sg0 = 0; sg0 < nq; sg0 += kTestQ) {                 // passes over the cell's queries (one, normally)
            const int nqs = min(kTestQ, nq - sg0);
            if (tid == 0) issue(0, 0);
            for (int t = tid; t < nqs; t += kTestThreads) {
                const uint32_t k = __ldg(sorted_pt + qs + sg0 + t);
                S.qx[t] = __ldg(xyz + 3 * (int64_t)k); S.qy[t] = __ldg(xyz + 3 * (int64_t)k + 1); S.qz[t] = __ldg(xyz + 3 * (int64_t)k + 2);
                S.qcnt[t] = 0;
                S.qsat[t] = 0;
            }
            __syncthreads();
            // query groups: 128 queries each (four per lane); the last one holds the remainder with as few queries
            // per lane as it needs (1..4), so the padding is what groups of 32 would give
            const int G = (nqs + 127) >> 7;
            int built = 0;
            for (int t = 0; t < ntiles; t++) {
                const int buf = t & 1;
                // the other buffer was released by the barrier that ended tile t - 1
                if (tid == 0 && t + 1 < ntiles) issue(t + 1, buf ^ 1);
                mbar_wait(&S.full[buf], (buf ? uses1 : uses0) & 1u);
                if (buf) uses1++; else uses0++;
                const int nt = min(kTestTile, Kpad - t * kTestTile);
                const int nblk = nt >> 5;
                // chunks of blocks: at least one item per warp when the tile allows it
                const int G32 = (nqs + 31) >> 5;
                const int per = max(1, min(kChunkBlocks, (nblk * G32 + 4 * (kTestThreads / 32) - 1) / (4 * (kTestThreads / 32))));
                const int nch = (nblk + per - 1) / per;
                const int64_t mrow0 = mb + (int64_t)t * (kTestTile / 32) * nq;
                for (int item = warp; item < G * nch; item += kTestThreads / 32) {
                    const int g = item / nch, ch = item - g * nch;
                    const int b0 = ch * per, b1 = min(nblk, b0 + per);
                    const int Qg = min(4, (nqs - (g << 7) + 31) >> 5);
                    if (Qg == 4) bq_test_item<4>(S, buf, g, b0, b1, nqs, nq, sg0, mrow0, masks, thr, can_saturate, lane);
                    else if (Qg == 3) bq_test_item<3>(S, buf, g, b0, b1, nqs, nq, sg0, mrow0, masks, thr, can_saturate, lane);
                    else if (Qg == 2) bq_test_item<2>(S, buf, g, b0, b1, nqs, nq, sg0, mrow0, masks, thr, can_saturate, lane);
                    else bq_test_item<1>(S, buf, g, b0, b1, nqs, nq, sg0, mrow0, masks, thr, can_saturate, lane);
                }
                built = min(K, (t + 1) * kTestTile);
                if (can_saturate) {
                    __syncthreads();                              // the tile's counts are complete
                    int unsat = 0;
                    for (int q = tid; q < nqs; q += kTestThreads) {
                        const bool s = S.qcnt[q] >= kCap;
                        S.qsat[q] = s;
                        unsat |= !s;
                    }
                    if (!__syncthreads_or(unsat)) {               // nobody needs later candidates
                        if (t + 1 < ntiles) {                     // ... but tile t + 1 is already on its way: let it land
                            mbar_wait(&S.full[buf ^ 1], (buf ? uses0 : uses1) & 1u);
                            if (buf) uses0++; else uses1++;
                        }
                        break;
                    }
                } else {
                    __syncthreads();                              // everyone is done with this buffer
                }
            }
            for (int q = tid; q < nqs; q += kTestThreads) counts[qs + sg0 + q] = min(S.qcnt[q], kCap);
            built_max = max(built_max, built);
            __syncthreads();
        }
        if (tid == 0) { kb[c] = built_max; S.cellslot = next_slot; }
    }
}

// ---- fill -------------------------------------------------------------------------------------------
constexpr int kFillQ = 4;
constexpr int kFillShortK = kSmallK;     // the cells of the warp-per-cell count kernels: at most eight mask words per query

// without masks: a warp per query re-evaluates the predicate over the cell's sorted candidates
__global__ void __launch_bounds__(256) k_bq_fill(const float *__restrict__ xyz, const uint32_t *__restrict__ sorted_pt,
                                                 const int32_t *__restrict__ cell, const int32_t *__restrict__ cand_start,
                                                 const int32_t *__restrict__ kb, const uint32_t *__restrict__ cand_idx,
                                                 const int2 *__restrict__ start_len, float r2, int32_t n,
                                                 int32_t *__restrict__ idx) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n; q += nWarps) {
        const uint32_t k = sorted_pt[q];
        const int2 sl = __ldg(start_len + k);
        if (sl.y == 0) continue;
        const int c = __ldg(cell + k);
        const uint32_t *ci = cand_idx + __ldg(cand_start + c);
        const int K = __ldg(kb + c);
        const float ox = __ldg(xyz + 3 * (int64_t)k), oy = __ldg(xyz + 3 * (int64_t)k + 1), oz = __ldg(xyz + 3 * (int64_t)k + 2);
        int32_t *out = idx + sl.x;
        int written = 0;
        for (int base = 0; base < K && written < sl.y; base += 32) {
            const int e = base + lane;
            bool hit = false;
            uint32_t id = 0;
            if (e < K) {
                id = __ldg(ci + e);
                const float *p = xyz + 3 * (int64_t)id;
                hit = bq_hit(ox, oy, oz, make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f), r2);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const int pos = written + __popc(m & lt);
            if (hit && pos < sl.y) out[pos] = (int)id;
            written += __popc(m);
        }
    }
}

// fill from hit masks: no coordinates, no predicate -- a word of 32 outcomes per (query, block); the
// lane whose bit is set writes its candidate's index at (hits so far) + (set bits below it).
// One call handles a RUN of L <= 4 consecutive queries of the same cell (queries q_first .. q_first + L - 1 of
// the sorted order).  Eight blocks per trip: ONE load brings the 8 x 4 mask words (lane = block * 4 + query) and
// eight independent loads the 256 candidate indices, so nine loads per lane are in flight together and
// their latency is paid once per eight blocks.
__device__ __forceinline__ void bq_fill_run(int64_t q_first, int L, int cc, const uint32_t *__restrict__ sorted_pt,
                                            const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                            const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kb,
                                            const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                            const uint32_t *__restrict__ masks, const int2 *__restrict__ start_len,
                                            int32_t *__restrict__ idx, int lane, unsigned lt) {
    const int K = __ldg(kb + cc);
    const int nq = __ldg(ccnt + cc);
    const uint32_t *cp = cand_idx + __ldg(cand_start + cc) + lane;
    int wpos[kFillQ], wend[kFillQ];
#pragma unroll
    for (int u = 0; u < kFillQ; u++) {
        wpos[u] = wend[u] = 0;
        if (u < L) {
            const int2 sl = __ldg(start_len + __ldg(sorted_pt + q_first + u));
            wpos[u] = sl.x; wend[u] = sl.x + sl.y;
        }
    }
    const int nb = (K + 31) >> 5;
    const bool mine = (lane & 3) < L;              // this lane's (block, query) slot holds a query of the run
    const uint32_t *mp8 = masks + __ldg(mbase + cc) + (int)(q_first - __ldg(cstart + cc)) + (int64_t)(lane >> 2) * nq + (lane & 3);
    const unsigned lanebit = 1u << lane;
    // A list shorter than the cap holds every hit of its query, so the recorded bits ARE the list: no
    // position needs checking against the segment's end.  Only a run with a full list (kCap entries:
    // later hits are dropped, and words past the last one may not have been written) takes the checked loop.
    const bool capped = wend[0] - wpos[0] >= kCap || wend[1] - wpos[1] >= kCap || wend[2] - wpos[2] >= kCap ||
                        wend[3] - wpos[3] >= kCap;
    if (!capped) {
        for (int b0 = 0; b0 < nb; b0 += 8) {
            const unsigned mw = (mine && b0 + (lane >> 2) < nb) ? __ldg(mp8 + (int64_t)b0 * nq) : 0u;
            int cidr[8];
#pragma unroll
            for (int j = 0; j < 8; j++) cidr[j] = ((b0 + j) * 32 + lane < K) ? (int)__ldg(cp + (b0 + j) * 32) : 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int u = 0; u < kFillQ; u++) {
                    const unsigned m = __shfl_sync(0xffffffffu, mw, j * 4 + u);
                    if (m & lanebit) idx[wpos[u] + __popc(m & lt)] = cidr[j];
                    wpos[u] += __popc(m);
                }
            }
        }
        return;
    }
    bool done = false;
    for (int b0 = 0; b0 < nb && !done; b0 += 8) {
        const unsigned mw = (mine && b0 + (lane >> 2) < nb) ? __ldg(mp8 + (int64_t)b0 * nq) : 0u;
        int cidr[8];
#pragma unroll
        for (int j = 0; j < 8; j++) cidr[j] = ((b0 + j) * 32 + lane < K) ? (int)__ldg(cp + (b0 + j) * 32) : 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
#pragma unroll
            for (int u = 0; u < kFillQ; u++) {
                const unsigned m = __shfl_sync(0xffffffffu, mw, j * 4 + u);
                const int pos = wpos[u] + __popc(m & lt);
                if ((m & lanebit) && pos < wend[u]) idx[pos] = cidr[j];
                wpos[u] = min(wpos[u] + __popc(m), wend[u]);   // words past a full list may be unwritten: stay put
            }
        }
        done = wpos[0] >= wend[0] && wpos[1] >= wend[1] && wpos[2] >= wend[2] && wpos[3] >= wend[3];
    }
}

// Short candidate lists (K <= kFillShortK: at most three mask words per query, never a full list) -- the bulk of
// a sparse set such as raw scene coordinates, ~9 neighbours per point.  A warp per four queries spends its time
// waiting on one short dependent chain after another there (0.20 ms for 9 M outputs); with a THREAD per query a
// million chains are in flight at once.  k_bq_fill_mask skips the queries this kernel has written.
__global__ void __launch_bounds__(256) k_bq_fill_short(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                                       const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                       const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kb,
                                                       const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                                       const uint32_t *__restrict__ masks, const int2 *__restrict__ start_len,
                                                       int32_t n, int32_t *__restrict__ idx) {
    pdl_enter();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t k = __ldg(sorted_pt + q);
    const int c = __ldg(cell + k);
    const int K = __ldg(kb + c);
    if (K > kFillShortK) return;
    const int2 sl = __ldg(start_len + k);
    const int nq = __ldg(ccnt + c);
    const uint32_t *mrow = masks + __ldg(mbase + c) + (q - __ldg(cstart + c));
    const uint32_t *ci = cand_idx + __ldg(cand_start + c);
    unsigned m[(kFillShortK + 31) / 32];
#pragma unroll
    for (int w = 0; w < (kFillShortK + 31) / 32; w++) m[w] = (w * 32 < K) ? __ldg(mrow + (int64_t)w * nq) : 0u;
    int32_t *out = idx + sl.x;
    int pos = 0;
#pragma unroll
    for (int w = 0; w < (kFillShortK + 31) / 32; w++) {
        unsigned mm = m[w];
        while (mm) {
            const int b = __ffs((int)mm) - 1;
            mm &= mm - 1u;
            out[pos++] = (int)__ldg(ci + w * 32 + b);
        }
    }
}

// A warp takes four consecutive queries of the sorted order at a time and cuts them into runs of equal cell
// (one run when the cell has at least four queries left, which is the common case on dense data).
__global__ void __launch_bounds__(256, 5) k_bq_fill_mask(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                                      const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                      const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kb,
                                                      const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                                      const uint32_t *__restrict__ masks, const int2 *__restrict__ start_len,
                                                      int32_t n, int32_t *__restrict__ idx) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFillQ; q0 < n; q0 += nWarps * kFillQ) {
        const int nqr = (int)((n - q0) < kFillQ ? (n - q0) : kFillQ);
        int c[kFillQ];
#pragma unroll
        for (int u = 0; u < kFillQ; u++) c[u] = u < nqr ? __ldg(cell + __ldg(sorted_pt + q0 + u)) : -1;
        // bit u of `ends`: a run of equal cell ends with query u (warp-uniform)
        unsigned ends = 0;
#pragma unroll
        for (int u = 0; u < kFillQ; u++)
            if (u < nqr && (u == nqr - 1 || c[(u + 1) % kFillQ] != c[u])) ends |= 1u << u;
        int u0 = 0;
        while (u0 < nqr) {
            const int L = __ffs((int)(ends >> u0));
            const int cc = u0 == 0 ? c[0] : __ldg(cell + __ldg(sorted_pt + q0 + u0));
            if (__ldg(kb + cc) > kFillShortK)               // shorter candidate lists were decoded by k_bq_fill_short
                bq_fill_run(q0 + u0, L, cc, sorted_pt, cstart, ccnt, cand_start, kb, cand_idx, mbase, masks, start_len, idx, lane, lt);
            u0 += L;
        }
    }
}

// ---- lazy lists (see ballquery.cuh) ---------------------------------------------------------------------------------
// n-th (1-based) set bit of m
__device__ __forceinline__ int nth_set_bit(unsigned m, int nth) {
    for (int i = 1; i < nth; i++) m &= m - 1u;
    return __ffs((int)m) - 1;
}

// a thread per query (consecutive queries of a cell read consecutive mask words: coalesced)
__global__ void __launch_bounds__(256) k_bq_list_samples(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                                        const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                        const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kb,
                                                        const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                                        const uint32_t *__restrict__ masks, const int32_t *__restrict__ counts,
                                                        int32_t n, int4 *__restrict__ samples) {
    pdl_enter();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t k = __ldg(sorted_pt + q);
    const int len = __ldg(counts + q);
    int e0 = (int)k, e1 = (int)k, em = (int)k, el = (int)k;          // an empty list samples the point itself (a self edge)
    if (len > 0) {
        const int c = __ldg(cell + k);
        const int nq = __ldg(ccnt + c);
        const int nb = (__ldg(kb + c) + 31) >> 5;
        const uint32_t *mrow = masks + __ldg(mbase + c) + (q - __ldg(cstart + c));
        const uint32_t *cp = cand_idx + __ldg(cand_start + c);
        int seen = 0, p0 = -1, p1 = -1, pm = -1, pl = -1;
        for (int b = 0; b < nb; b++) {
            unsigned m = __ldg(mrow + (int64_t)b * nq);
            if (m == 0u) continue;
            const int pc = __popc(m);
            if (seen + pc >= len) {                                  // the list ends inside this word (later bits: hits beyond the cap)
                const int at = nth_set_bit(m, len - seen);
                pl = (b << 5) + at;
                m &= (at == 31) ? 0xffffffffu : ((2u << at) - 1u);
            }
            if (p0 < 0) {
                p0 = (b << 5) + __ffs((int)m) - 1;
                const unsigned r = m & (m - 1u);
                if (r) p1 = (b << 5) + __ffs((int)r) - 1;
            } else if (p1 < 0) p1 = (b << 5) + __ffs((int)m) - 1;
            if (pm < 0 && seen + pc >= (len + 1) / 2) pm = (b << 5) + nth_set_bit(m, max(1, min(pc, (len + 1) / 2 - seen)));
            seen += pc;
            if (pl >= 0) break;
        }
        e0 = (int)__ldg(cp + p0);
        e1 = p1 >= 0 ? (int)__ldg(cp + p1) : e0;
        em = pm >= 0 ? (int)__ldg(cp + pm) : e0;
        el = pl >= 0 ? (int)__ldg(cp + pl) : e0;
    }
    samples[k] = make_int4(e0, e1, em, el);
}

// a warp per listed point
__global__ void __launch_bounds__(256, 5) k_bq_fill_lists(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                                       const int32_t *__restrict__ cstart, const int32_t *__restrict__ ccnt,
                                                       const int32_t *__restrict__ cand_start, const int32_t *__restrict__ kb,
                                                       const uint32_t *__restrict__ cand_idx, const int32_t *__restrict__ mbase,
                                                       const uint32_t *__restrict__ masks, const int2 *__restrict__ start_len,
                                                       const int32_t *__restrict__ qpos, const uint32_t *__restrict__ worklist,
                                                       const unsigned long long *__restrict__ count, int32_t *__restrict__ idx) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const long long nl = (long long)*count;
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < nl; e += nWarps) {
        const uint32_t i = __ldg(worklist + e);
        if (__ldg(&start_len[i].y) == 0) continue;
        bq_fill_run(__ldg(qpos + i), 1, __ldg(cell + i), sorted_pt, cstart, ccnt, cand_start, kb, cand_idx, mbase, masks, start_len,
                    idx, lane, lt);
    }
}

int bq_list_samples(const BqWs &w, const uint32_t *masks, int32_t n, int4 *samples, cudaStream_t st) {
    PG_KTIME("k_bq_list_samples", st);
    launch(k_bq_list_samples, (unsigned)div_up(n, 256), 256, 0, st, bq_sorted(w, n), w.cell, w.cstart, w.ccnt, w.cand_start, w.kb, w.cand_idx,
                                                               w.mbase, masks, w.counts, n, samples);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

int bq_fill_lists(const BqWs &w, const uint32_t *masks, const int2 *start_len, const uint32_t *worklist,
                  const unsigned long long *count, int32_t n, int32_t *idx, cudaStream_t st) {
    PG_KTIME("k_bq_fill_lists", st);
    launch(k_bq_fill_lists, kNumSM * PG_RESIDENT(k_bq_fill_lists, 256, 0) * 4, 256, 0, st, bq_sorted(w, n), w.cell, w.cstart, w.ccnt, w.cand_start,
                                                                                   w.kb, w.cand_idx, w.mbase, masks, start_len, w.qpos,
                                                                                   worklist, count, idx);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" size_t pg_ballquery_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return bq_layout(nullptr, 0, n).used + 256;
}

extern "C" int pg_ballquery_prepare(const float *xyz, const int32_t *batch_idxs, const int32_t *batch_offsets, int32_t n,
                                    int32_t B, float radius, void *ws, size_t ws_bytes, void *stream) {
    (void)batch_offsets; (void)B;   // scene membership comes from batch_idxs
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(n >= 0 && n <= (1 << 26), "n out of range (0 .. 2^26)");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(xyz && batch_idxs && ws, "null pointer");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_prepare: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }

    const double s = fabs((double)radius) * 1.0001;
    const double inv_s = (s > 0.0 && isfinite(s)) ? 1.0 / s : 0.0;   // r = 0 / inf / NaN: one cell per scene
    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 8 * sizeof(int64_t), st));
    launch(k_bq_keys, (unsigned)div_up(n, 256), 256, 0, st, xyz, batch_idxs, n, inv_s, w.keys, group_table_fill(w.tab), Fill{(uint32_t *)w.ccnt, (size_t)n + 1, 0u});
    // scratch of the scatter, adjacent in the workspace (cleared by the grouping's last kernel): the cells' cursors, the
    // complement of their smallest and their largest point index
    uint32_t *sorted_pt = w.kA, *cmin = w.kB, *cmax = w.vB;
    int32_t *cursor = reinterpret_cast<int32_t *>(w.vA);
    PG_TRY(group_int4(w.keys, n, w.tab, w.pslot, w.cell, w.ccnt, w.scalars, w.scan_tmp, st, nullptr,
                      Fill{w.vA, ((size_t)((char *)w.vB - (char *)w.vA) + align_up((size_t)n * 4)) / 4, 0u},   // to the padded end of vB
                      getenv("PG_BQ_ORDERED_CELLS") != nullptr));         // cells numbered as they are claimed (scalars[0] is zero)
    PG_TRY(scan_exclusive_i32(w.ccnt, w.cstart, (int64_t)n + 1, nullptr, w.scan_tmp, st));
    launch(k_bq_scatter, (unsigned)div_up(n, 256), 256, 0, st, w.cell, w.cstart, n, cursor, sorted_pt, cmin, cmax);
    { PG_KTIME("k_bq_neighbours", st);
    launch(k_bq_neighbours, kNumSM * PG_RESIDENT(k_bq_neighbours, 256, 0) * 2, 256, 0, st, w.keys, w.tab, sorted_pt, w.cstart, w.ccnt, w.scalars, w.nbr, w.kc, w.dense, cmin, cmax, w.crange); }
    PG_TRY(scan_fused(CandLoad{w.kc, w.scalars}, CandStore{w.kc, w.cand_start, w.scalars}, (int64_t)n + 1, w.scalars + 1, w.scan_tmp, st));
    PG_TRY(scan_fused(MaskLoad{w.ccnt, w.kc, w.scalars}, MaskStore{w.mbase}, (int64_t)n + 1, w.scalars + 6, w.scan_tmp, st));
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_ballquery_count(const float *xyz, int32_t n, float radius, int32_t *start_len, uint32_t *masks,
                                  int64_t mask_words, void *ws, size_t ws_bytes, int64_t *host_total, int *host_masks_used,
                                  void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_total && host_masks_used, "null host_total / host_masks_used");
    *host_total = 0;
    *host_masks_used = 0;
    PG_CHECK_ARG(n >= 0 && n <= (1 << 26), "n out of range (0 .. 2^26)");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(xyz && start_len && ws, "null pointer");
    PG_CHECK_ARG(masks == nullptr || mask_words >= 0, "negative mask_words");
    BqWs w = bq_layout(ws, ws_bytes, n);
    if (!w.ok) { set_error("pg_ballquery_count: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    const uint32_t *sorted_pt = bq_sorted(w, n);
    const float r2 = radius * radius;
    PG_CUDA(cudaMemsetAsync(w.scalars + 4, 0, 2 * sizeof(int64_t), st));   // the dense kernels' work counters
    PG_CUDA(cudaMemsetAsync(w.scalars + 7, 0, sizeof(int64_t), st));       // ... and the coordinate cursor
    // masks are used when they fit the caller's buffer (and int32 bases): decided on the device, reported below
    const int64_t mask_cap = masks ? (mask_words < 0x7fffffffLL ? mask_words : 0x7ffffffeLL) : -1;
    const int64_t gsmall_want = div_up(n, 8);
    auto ksmall = k_bq_cells_small<0, kSmallSplit>;
    auto kmedium = k_bq_cells_small<kSmallSplit, kSmallK>;
    const int64_t gsmall_max = (int64_t)kNumSM * PG_RESIDENT(ksmall, 256, 0) * 4;
    const int64_t gmedium_max = (int64_t)kNumSM * PG_RESIDENT(kmedium, 256, 0) * 4;
    const unsigned gsmall = (unsigned)(gsmall_want < gsmall_max ? gsmall_want : gsmall_max);
    const unsigned gmedium = (unsigned)(gsmall_want < gmedium_max ? gsmall_want : gmedium_max);
    { PG_KTIME("k_bq_cells_small", st);
    launch(ksmall, gsmall, 256, 0, st, xyz, sorted_pt, w.cstart, w.ccnt, w.nbr, w.kc, w.cand_start, w.mbase, w.scalars, masks, mask_cap,
                                   r2, w.cand_idx, w.counts, w.kb); }
    { PG_KTIME("k_bq_cells_medium", st);
    launch(kmedium, gmedium, 256, 0, st, xyz, sorted_pt, w.cstart, w.ccnt, w.nbr, w.kc, w.cand_start, w.mbase, w.scalars, masks, mask_cap,
                                     r2, w.cand_idx, w.counts, w.kb); }
    { PG_KTIME("k_bq_merge_dense", st);
    launch(k_bq_merge_dense, kNumSM * PG_RESIDENT(k_bq_merge_dense, kMergeThreads, 0), kMergeThreads, 0, st, xyz, sorted_pt, w.cstart, w.ccnt, w.nbr, w.kc, w.cand_start, w.dense, w.crange, w.scalars, w.cand_idx, w.cand_xy, w.cand_z, w.dbase); }
    { PG_KTIME("k_bq_test_dense", st);
    launch(k_bq_test_dense, kNumSM * PG_RESIDENT(k_bq_test_dense, kTestThreads, 0), kTestThreads, 0, st, xyz, sorted_pt, w.cstart, w.ccnt, w.kc, w.mbase, w.dense, w.dbase, w.cand_xy, w.cand_z, w.scalars, masks, mask_cap, r2,
        w.counts, w.kb); }
    // starts in query order, written straight into the interleaved (start, len) rows per point
    PG_TRY(scan_fused(CountLoad{w.counts}, StartLenStore{sorted_pt, (int2 *)start_len, w.qpos}, n, w.scalars + 2, w.scan_tmp, st));
    PG_LAUNCH_CHECK();
    int64_t back[5] = {0, 0, 0, 0, 0};               // scalars [2] total neighbours ... [6] mask words
    PG_CUDA(cudaMemcpyAsync(back, w.scalars + 2, sizeof(back), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    const int64_t total = back[0];
    *host_total = total;
    *host_masks_used = (masks && back[4] <= mask_cap) ? 1 : 0;
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
    if (masks) {
        PG_KTIME("k_bq_fill_short", st);
        launch(k_bq_fill_short, (unsigned)div_up(n, 256), 256, 0, st, sorted_pt, w.cell, w.cstart, w.ccnt, w.cand_start, w.kb, w.cand_idx,
                                                                  w.mbase, masks, (const int2 *)start_len, n, idx);
    }
    PG_KTIME(masks ? "k_bq_fill_mask" : "k_bq_fill", st);
    if (masks)
        launch(k_bq_fill_mask, kNumSM * PG_RESIDENT(k_bq_fill_mask, 256, 0) * 4, 256, 0, st, sorted_pt, w.cell, w.cstart, w.ccnt, w.cand_start, w.kb, w.cand_idx, w.mbase,
                                                   masks, (const int2 *)start_len, n, idx);
    else
        launch(k_bq_fill, kNumSM * PG_RESIDENT(k_bq_fill, 256, 0) * 4, 256, 0, st, xyz, sorted_pt, w.cell, w.cand_start, w.kb, w.cand_idx, (const int2 *)start_len,
                                              r2, n, idx);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
