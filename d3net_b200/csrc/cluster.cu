// bfs_cluster as GPU connected components.  Reference behaviour:
// lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.cpp:28-112 (single-thread FIFO BFS on the CPU).
//
// What the reference computes (see DESIGN.md for the proof): with seeds visited in ascending order,
// point v ends in the cluster of m(v) = the smallest index that REACHES v along same-label edges
// i -> idx[start_i .. start_i+len_i).  Clusters are numbered by ascending m after dropping those
// smaller than `threshold`.
//
// The edge relation is symmetric except where a neighbour list was truncated at 1000 entries
// (bfs_cluster.cu:38-43): j in list(i) but i not in list(j)  <=>  list(j) is full and i > last(list(j)).
//   fast path    ONE sweep over the edges: union-find with atomic hooking (larger root under smaller,
//                so a root is its component's minimum) over the two-way edges, taken from their
//                higher endpoint; one-way edges whose endpoints are not yet connected are parked in a
//                pending list and settled afterwards by a min-label propagation over that (tiny) list.
//   generic path min-label propagation over ALL edges with no unions: exact for any directed graph.
// The fast path is only trusted when the lists check out as a truncated symmetric relation: in range,
// ascending, and a 64-bit checksum of the two-way edge set that cancels against its own reversal.
// Both checks ride along in the same sweep and need no extra memory traffic.
#include "common.cuh"

namespace pg {

constexpr int kCapC = PG_BALLQUERY_CAP;

struct ClWs {
    uint2 *pl;           // per point: x = union-find parent, y = semantic label (one 8-byte read per edge)
    uint32_t *trunc;     // bitmap: list is full (len >= 1000), i.e. may lack reverse edges
    int32_t *last;       // last entry of a full list: u -> v is two-way  <=>  u <= last[v]
    int32_t *root;       // flattened component root per point (identity on the generic path)
    int32_t *lab;        // min-ancestor label forest over roots
    int32_t *size;       // points per final label
    int32_t *cid;        // cluster id per kept label (exclusive scan of keep flags)
    int32_t *csize;      // sizes in cluster order -> offsets
    int2 *pend;          // parked one-way edges (i -> j)
    uint32_t *key0, *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    // [0] checksum [1] bad [2] pending count [3] changed [4] nCluster [5] sumNPoint
    unsigned long long *scalars;
    size_t pend_cap;
    bool ok;
    size_t used;
};

static ClWs cl_layout(void *ws, size_t ws_bytes, int64_t N_) {
    Arena a(ws, ws_bytes);
    ClWs w;
    const size_t n = (size_t)(N_ > 0 ? N_ : 1);
    w.pl = a.take<uint2>(n);
    w.trunc = a.take<uint32_t>(n / 32 + 2);
    w.last = a.take<int32_t>(n);
    w.root = a.take<int32_t>(n);
    w.lab = a.take<int32_t>(n);
    w.size = a.take<int32_t>(n + 1);
    w.cid = a.take<int32_t>(n + 1);
    w.csize = a.take<int32_t>(n + 1);
    w.pend_cap = n + 1024;
    w.pend = a.take<int2>(w.pend_cap);
    w.key0 = a.take<uint32_t>(n);
    w.kA = a.take<uint32_t>(n);
    w.vA = a.take<uint32_t>(n);
    w.kB = a.take<uint32_t>(n);
    w.vB = a.take<uint32_t>(n);
    w.hist = a.take<int32_t>(radix_tmp_count(N_));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(n + radix_tmp_count(N_))));
    w.scalars = a.take<unsigned long long>(8);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

// 64-bit tag of an ordered pair (hi, lo): the product of two independently scrambled 32-bit words.
// Cheap (three integer multiplies); it only has to make accidental cancellation of unmatched edges
// in the checksum a 2^-64-class event, not resist an adversary.
__device__ __forceinline__ unsigned long long mix64(unsigned a, unsigned b) {
    const unsigned x = (a ^ 0x5bd1e995u) * 0x9E3779B1u;
    const unsigned y = (b + 0x7f4a7c15u) * 0x85EBCA77u;
    return (unsigned long long)(x ^ (x >> 15)) * (unsigned long long)((y ^ (y >> 13)) | 1u) + a;
}

// one thread per point, launched over ceil(N / 32) * 32 threads so every bitmap word has a full warp
__global__ void k_cl_prep(const int32_t *__restrict__ label, const int32_t *__restrict__ idx,
                          const int2 *__restrict__ start_len, int32_t N, int64_t nActive, uint2 *__restrict__ pl,
                          uint32_t *__restrict__ trunc, int32_t *__restrict__ last, int32_t *__restrict__ root,
                          int32_t *__restrict__ lab, unsigned long long *scalars) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool full = false;
    if (v < N) {
        const int2 sl = start_len[v];
        int lst = 0x7fffffff;
        if (sl.y < 0 || sl.x < 0 || (int64_t)sl.x + sl.y > nActive) scalars[1] = 2;   // malformed row
        else if (sl.y >= kCapC) { lst = __ldg(idx + sl.x + sl.y - 1); full = true; }
        pl[v] = make_uint2((unsigned)v, (unsigned)__ldg(label + v));
        last[v] = lst;
        root[v] = v;
        lab[v] = v;
    }
    const unsigned m = __ballot_sync(0xffffffffu, full);
    if ((threadIdx.x & 31) == 0 && (v >> 5) <= (N >> 5)) trunc[v >> 5] = m;
}

// union-find on pl[].x
// (parent reads go to L2 with ld.cg: an L1-stale "x is still a root" could spin the CAS loop)
__device__ __forceinline__ int uf_find(uint2 *pl, int x) {
    int p = (int)__ldcg(&pl[x].x);
    while (p != x) {
        const int gp = (int)__ldcg(&pl[p].x);
        if (gp != p) pl[x].x = (unsigned)gp;   // path halving; pointers only ever move to smaller ancestors
        x = p;
        p = gp;
    }
    return x;
}

// a, b: roots as far as the caller knows; returns the root of the merged set as far as this thread can tell
__device__ __forceinline__ int uf_union_roots(uint2 *pl, int a, int b) {
    for (;;) {
        if (a == b) return a;
        const int hi = max(a, b), lo = min(a, b);
        const unsigned old = atomicCAS(&pl[hi].x, (unsigned)hi, (unsigned)lo);
        if (old == (unsigned)hi) return lo;
        a = uf_find(pl, hi);     // hi was hooked by somebody else meanwhile
        b = uf_find(pl, lo);
    }
}

// The single edge sweep of the fast path.  G lanes share one point's list (coalesced reads of idx).
template <int G>
__global__ void __launch_bounds__(256) k_cl_union(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                  uint2 *pl, const uint32_t *__restrict__ trunc,
                                                  const int32_t *__restrict__ last, int32_t N, int2 *__restrict__ pend,
                                                  unsigned pend_cap, unsigned long long *scalars) {
    const int sub = threadIdx.x % G;
    const int64_t groups = (int64_t)gridDim.x * (blockDim.x / G);
    unsigned long long chk = 0;
    bool bad = false;
    // the loop bound is rounded up to whole groups so that every lane of a warp takes part in the shuffle
    const int64_t n_up = ((int64_t)N + groups - 1) / groups * groups;
    for (int64_t i64 = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; i64 < n_up; i64 += groups) {
        const bool live = i64 < N;
        const int i = live ? (int)i64 : 0;
        // one lane of the group finds the point's root; the others get it by shuffle
        int ri = (live && sub == 0) ? uf_find(pl, i) : 0;
        ri = __shfl_sync(0xffffffffu, ri, 0, G);
        if (!live) continue;
        const int2 sl = start_len[i];
        const unsigned li = pl[i].y;
        // software pipeline, kU edges per lane per trip: all index loads first, then all bitmap words,
        // then all (parent, label) records -- three dependent round trips per trip instead of per edge
        constexpr int kU = 4;
        for (int e0 = sub; e0 < sl.y; e0 += kU * G) {
            int jj[kU], pv[kU];
            unsigned tw[kU];
            uint2 w[kU];
            bool use[kU], twoway[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int e = e0 + u * G;
                jj[u] = e < sl.y ? __ldg(idx + sl.x + e) : -1;
                pv[u] = (e < sl.y && e > 0) ? __ldg(idx + sl.x + e - 1) : -1;      // same lines as a neighbour lane's jj
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                use[u] = jj[u] >= 0 && jj[u] < N && jj[u] != i;
                if (e0 + u * G < sl.y && (jj[u] < 0 || jj[u] >= N)) bad = true;
                if (pv[u] >= jj[u] && e0 + u * G < sl.y && e0 + u * G > 0) bad = true;   // lists must ascend
                tw[u] = use[u] ? __ldg(trunc + (jj[u] >> 5)) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const bool jfull = (tw[u] >> (jj[u] & 31)) & 1u;
                twoway[u] = !jfull || i <= __ldg(last + jj[u]);
                if (use[u] && twoway[u]) {
                    // each two-way pair {a > b} is seen as a -> b and as b -> a: the two terms cancel
                    if (jj[u] < i) chk += mix64((unsigned)i, (unsigned)jj[u]);
                    else { chk -= mix64((unsigned)jj[u], (unsigned)i); use[u] = false; }   // taken from the higher endpoint
                }
                w[u] = use[u] ? __ldcg(pl + jj[u]) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (!use[u] || w[u].y != li) continue;
                const int j = jj[u];
                const int p = (int)w[u].x;
                if (p == ri) continue;                 // already under the same root
                // climb from j's parent to its root, stopping early at ri (spares the hot root line)
                int r = p;
                while (r != ri) {
                    const int g = (int)__ldcg(&pl[r].x);
                    if (g == r) break;
                    r = g;
                }
                if (r != p) pl[j].x = (unsigned)r;     // compress j straight onto the ancestor found
                if (r == ri) continue;
                if (twoway[u]) {
                    ri = uf_union_roots(pl, ri, r);
                } else {
                    // one-way edge i -> j whose ends are not (yet) connected: park it
                    if (uf_find(pl, ri) == uf_find(pl, r)) continue;
                    const unsigned long long slot = atomicAdd(&scalars[2], 1ULL);
                    if (slot < pend_cap) pend[slot] = make_int2(i, j);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(0xffffffffu, chk, o);
    if ((threadIdx.x & 31) == 0 && chk) atomicAdd(&scalars[0], chk);
    if (bad) atomicMax(&scalars[1], 1ULL);
}

// Roots go to their own array: writing them back into the forest would race with the path-halving
// stores of other threads' finds, which may re-install an intermediate ancestor after the root.
__global__ void k_cl_flatten(uint2 *pl, int32_t *__restrict__ root, int32_t N) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) root[v] = uf_find(pl, v);
}

// resolve() on the label forest.  Its shortcut writes must be atomicMin: lab[x] is also the target
// of the propagation's atomicMin, and a plain store could undo a concurrent lowering.
__device__ __forceinline__ int lab_resolve(int32_t *lab, int x) {
    int p = __ldcg(lab + x);
    while (p != x) {
        const int gp = __ldcg(lab + p);
        if (gp != p) atomicMin(&lab[x], gp);
        x = p;
        p = gp;
    }
    return x;
}

// Min-ancestor propagation along one edge i -> j: lab[root(j)] <- min(., resolve(root(i))).
// resolve() follows lab to its fixed point; every value it passes through reaches the start node, so
// shortcutting is sound on a directed graph too.  Only rj itself may be relabelled: `mine` reaches
// rj, but not necessarily rj's current label.
__device__ __forceinline__ bool propagate_edge(const int32_t *__restrict__ root, int32_t *lab, int i, int j) {
    const int mine = lab_resolve(lab, root[i]);
    const int rj = root[j];
    return mine < lab_resolve(lab, rj) && atomicMin(&lab[rj], mine) > mine;
}

__global__ void k_cl_pending(const int2 *__restrict__ pend, unsigned long long n_pend, const int32_t *__restrict__ root,
                             int32_t *lab, unsigned long long *scalars) {
    bool changed = false;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pend;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const int2 e = pend[t];
        if (root[e.x] != root[e.y]) changed |= propagate_edge(root, lab, e.x, e.y);
    }
    if (changed) scalars[3] = 1;
}

// Full sweep: the generic path (ALL same-label edges) or, with ONE_WAY_ONLY, the fall-back when the
// pending list overflowed.
template <int G, bool ONE_WAY_ONLY>
__global__ void __launch_bounds__(256) k_cl_propagate(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                      const uint2 *pl, const uint32_t *__restrict__ trunc,
                                                      const int32_t *__restrict__ last, int32_t N,
                                                      const int32_t *__restrict__ root, int32_t *lab,
                                                      unsigned long long *scalars) {
    const int sub = threadIdx.x % G;
    const int64_t groups = (int64_t)gridDim.x * (blockDim.x / G);
    bool changed = false;
    for (int64_t i64 = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; i64 < N; i64 += groups) {
        const int i = (int)i64;
        const int2 sl = start_len[i];
        const unsigned li = pl[i].y;
        for (int e = sub; e < sl.y; e += G) {
            const int j = __ldg(idx + sl.x + e);
            if ((unsigned)j >= (unsigned)N || j == i) continue;
            if (ONE_WAY_ONLY) {
                const bool jfull = (__ldg(trunc + (j >> 5)) >> (j & 31)) & 1u;
                if (!jfull || i <= __ldg(last + j)) continue;
            }
            if (pl[j].y != li) continue;
            if (ONE_WAY_ONLY && root[i] == root[j]) continue;
            changed |= propagate_edge(root, lab, i, j);
        }
    }
    if (changed) scalars[3] = 1;
}

__global__ void k_cl_reset(int32_t *__restrict__ root, int32_t *__restrict__ lab, int32_t N) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) { root[v] = v; lab[v] = v; }
}

// final label per point, sizes per label (warp-aggregated: a floor-sized component would otherwise
// serialise tens of thousands of atomics on one counter)
__global__ void k_cl_label(const int32_t *__restrict__ root, int32_t *lab, int32_t N, int32_t *__restrict__ size,
                           uint32_t *__restrict__ key0) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = v < N;
    int l = -1;
    if (on) {
        l = lab_resolve(lab, root[v]);
        key0[v] = (uint32_t)l;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    if (on && (peers & lanemask_lt()) == 0) atomicAdd(&size[l], __popc(peers));
}

// keep[l] = 1 when l is a label with >= threshold points (written into cid for the scan)
__global__ void k_cl_keep(const uint32_t *__restrict__ key0, const int32_t *__restrict__ size, int32_t N, int32_t threshold,
                          int32_t *__restrict__ cid) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int n = size[v];
    cid[v] = (key0[v] == (uint32_t)v && n > 0 && n >= threshold) ? 1 : 0;
}

// cluster sizes in cluster order + totals
__global__ void k_cl_sizes(const int32_t *__restrict__ size, const int32_t *__restrict__ cid, int32_t N,
                           int32_t *__restrict__ csize, unsigned long long *scalars) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int c = cid[v], next = cid[v + 1];   // cid has N+1 entries after the scan (last = total)
    if (next != c) {
        csize[c] = size[v];
        atomicAdd(&scalars[5], (unsigned long long)size[v]);
    }
}

// sort key per point: cluster id, or nCluster for points of dropped components (they sort last)
__global__ void k_cl_keys(const uint32_t *__restrict__ key0, const int32_t *__restrict__ cid, int32_t N, int32_t nCluster,
                          uint32_t *__restrict__ keys) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const uint32_t l = key0[v];
    const int c = cid[l], next = cid[l + 1];
    keys[v] = (next != c) ? (uint32_t)c : (uint32_t)nCluster;
}

__global__ void k_cl_emit(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int32_t S,
                          int2 *__restrict__ cluster_idxs) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < S) cluster_idxs[k] = make_int2((int)keys[k], (int)vals[k]);
}

}  // namespace pg

using namespace pg;

// diagnostics of the most recent pg_bfs_cluster_count on this thread: {checksum != 0, bad, pending, sweeps}
static thread_local long long g_cl_dbg[4] = {0, 0, 0, 0};
extern "C" void pg_bfs_cluster_debug(long long *out) { for (int i = 0; i < 4; i++) out[i] = g_cl_dbg[i]; }

extern "C" size_t pg_bfs_cluster_workspace_bytes(int64_t N) {
    if (N < 0) N = 0;
    return cl_layout(nullptr, 0, N).used + 256;
}

extern "C" int pg_bfs_cluster_count(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                                    const int32_t *start_len, int32_t N, int64_t nActive, int32_t threshold,
                                    int generic, void *ws, size_t ws_bytes, int32_t *host_sizes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_sizes, "null host_sizes");
    host_sizes[0] = host_sizes[1] = host_sizes[2] = 0;
    PG_CHECK_ARG(N >= 0 && nActive >= 0, "negative size");
    if (N == 0) return PG_OK;
    PG_CHECK_ARG(semantic_label && start_len && ws && (ball_query_idxs || nActive == 0), "null pointer");
    ClWs w = cl_layout(ws, ws_bytes, N);
    if (!w.ok) { set_error("pg_bfs_cluster_count: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    const int2 *sl = (const int2 *)start_len;
    const unsigned nb = (unsigned)div_up(N, 256);
    const bool wide = nActive / N >= 12;          // lanes per neighbour list: 32 for long lists, 8 for short
    const unsigned eg = kNumSM * 8;

    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 8 * sizeof(unsigned long long), st));
    PG_CUDA(cudaMemsetAsync(w.size, 0, ((size_t)N + 1) * sizeof(int32_t), st));
    k_cl_prep<<<(unsigned)div_up((int64_t)N + 32, 256), 256, 0, st>>>(semantic_label, ball_query_idxs, sl, N, nActive, w.pl,
                                                                     w.trunc, w.last, w.root, w.lab, w.scalars);
    unsigned long long h[4] = {0, 0, 0, 0};
    bool use_generic = generic != 0;
    if (!use_generic) {
        if (wide) k_cl_union<32><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.pend,
                                                     (unsigned)w.pend_cap, w.scalars);
        else k_cl_union<8><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.pend,
                                               (unsigned)w.pend_cap, w.scalars);
        k_cl_flatten<<<nb, 256, 0, st>>>(w.pl, w.root, N);
        PG_LAUNCH_CHECK();
    }
    PG_CUDA(cudaMemcpyAsync(h, w.scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    if (h[1] == 2) {   // the reference would read out of bounds here (bfs_cluster.cpp:40-42)
        set_error("pg_bfs_cluster_count: start_len has a row outside ball_query_idxs[0..%lld)", (long long)nActive);
        return PG_EINVAL;
    }
    if (!use_generic && (h[0] != 0 || h[1] != 0)) {   // not a truncated symmetric relation
        use_generic = true;
        k_cl_reset<<<nb, 256, 0, st>>>(w.root, w.lab, N);
    }
    const unsigned long long n_pend = use_generic ? 0 : h[2];
    g_cl_dbg[0] = h[0] != 0; g_cl_dbg[1] = (long long)h[1]; g_cl_dbg[2] = (long long)h[2]; g_cl_dbg[3] = 0;
    const bool sweep = use_generic || n_pend > w.pend_cap;
    if (sweep || n_pend > 0) {
        for (int it = 0; it < 1000000; it++) {
            PG_CUDA(cudaMemsetAsync(w.scalars + 3, 0, sizeof(unsigned long long), st));
            if (!sweep) {
                k_cl_pending<<<(unsigned)div_up((int64_t)n_pend, 256), 256, 0, st>>>(w.pend, n_pend, w.root, w.lab, w.scalars);
            } else if (use_generic) {
                if (wide) k_cl_propagate<32, false><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
                else k_cl_propagate<8, false><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
            } else {
                if (wide) k_cl_propagate<32, true><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
                else k_cl_propagate<8, true><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
            }
            PG_LAUNCH_CHECK();
            unsigned long long changed = 0;
            PG_CUDA(cudaMemcpyAsync(&changed, w.scalars + 3, sizeof(changed), cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            g_cl_dbg[3]++;
            if (!changed) break;
        }
    }
    k_cl_label<<<nb, 256, 0, st>>>(w.root, w.lab, N, w.size, w.key0);
    k_cl_keep<<<nb, 256, 0, st>>>(w.key0, w.size, N, threshold, w.cid);
    PG_CUDA(cudaMemsetAsync(w.cid + N, 0, sizeof(int32_t), st));
    PG_TRY(scan_exclusive_i32(w.cid, w.cid, (int64_t)N + 1, (int64_t *)(w.scalars + 4), w.scan_tmp, st));
    k_cl_sizes<<<nb, 256, 0, st>>>(w.size, w.cid, N, w.csize, w.scalars);
    PG_LAUNCH_CHECK();
    unsigned long long r[2];
    PG_CUDA(cudaMemcpyAsync(r, w.scalars + 4, sizeof(r), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    host_sizes[0] = (int32_t)r[0];
    host_sizes[1] = (int32_t)r[1];
    host_sizes[2] = use_generic ? 1 : 0;
    return PG_OK;
}

extern "C" int pg_bfs_cluster_fill(int32_t N, int32_t nCluster, int32_t sumNPoint, void *ws, size_t ws_bytes,
                                   int32_t *cluster_idxs, int32_t *cluster_offsets, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(N >= 0 && nCluster >= 0 && sumNPoint >= 0 && sumNPoint <= N, "bad sizes");
    PG_CHECK_ARG(cluster_offsets, "null cluster_offsets");
    if (N == 0 || nCluster == 0) {
        PG_CUDA(cudaMemsetAsync(cluster_offsets, 0, sizeof(int32_t), st));
        return PG_OK;
    }
    PG_CHECK_ARG(ws && cluster_idxs, "null pointer");
    ClWs w = cl_layout(ws, ws_bytes, N);
    if (!w.ok) { set_error("pg_bfs_cluster_fill: workspace too small"); return PG_EWORKSPACE; }
    const unsigned nb = (unsigned)div_up(N, 256);
    // offsets = exclusive scan of the cluster sizes (+ the total as the last entry)
    PG_CUDA(cudaMemsetAsync(w.csize + nCluster, 0, sizeof(int32_t), st));
    PG_TRY(scan_exclusive_i32(w.csize, cluster_offsets, (int64_t)nCluster + 1, nullptr, w.scan_tmp, st));
    // stable sort of (cluster id | dropped, point): members ascend inside every cluster
    k_cl_keys<<<nb, 256, 0, st>>>(w.key0, w.cid, N, nCluster, w.kB);
    int bits = 0;
    while ((1ll << bits) < (long long)nCluster + 1) bits++;
    int res = 0;
    PG_TRY(radix_sort_pairs(w.kB, nullptr, w.kA, w.vA, w.kB, w.vB, N, bits, w.hist, w.scan_tmp, st, &res));
    k_cl_emit<<<(unsigned)div_up(sumNPoint, 256), 256, 0, st>>>(res == 0 ? w.kA : w.kB, res == 0 ? w.vA : w.vB, sumNPoint,
                                                              (int2 *)cluster_idxs);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
