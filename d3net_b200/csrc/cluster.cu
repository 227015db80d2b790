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
//   fast path    union-find with atomic hooking (larger root under smaller, so a root is its
//                component's minimum) over the edges that test symmetric, then a min-label
//                propagation over the remaining one-way edges between components, iterated to a
//                fixed point (normally zero or one extra sweep);
//   generic path the same propagation over ALL edges with no unions: exact for any directed graph.
// The fast path is only trusted when the lists check out as a truncated symmetric relation: ascending,
// in range, and a 64-bit checksum of the symmetric edge set equal to that of its reversal.
#include "common.cuh"

namespace pg {

constexpr int kCapC = PG_BALLQUERY_CAP;

struct ClWs {
    int2 *info;          // per point: (label, last index that still has a reverse edge)
    int32_t *parent;     // union-find forest (fast path only)
    int32_t *root;       // flattened component root per point (identity on the generic path)
    int32_t *lab;        // min-ancestor label forest over roots
    int32_t *size;       // points per final label
    int32_t *cid;        // cluster id per kept label (exclusive scan of keep flags)
    int32_t *csize;      // sizes in cluster order -> offsets
    uint32_t *key0, *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    unsigned long long *scalars;  // [0] checksum [1] bad [2] residual [3] changed [4] nCluster [5] sumNPoint [6] sort buffer
    bool ok;
    size_t used;
};

static ClWs cl_layout(void *ws, size_t ws_bytes, int64_t N_) {
    Arena a(ws, ws_bytes);
    ClWs w;
    const size_t n = (size_t)(N_ > 0 ? N_ : 1);
    w.info = a.take<int2>(n);
    w.parent = a.take<int32_t>(n);
    w.root = a.take<int32_t>(n);
    w.lab = a.take<int32_t>(n);
    w.size = a.take<int32_t>(n + 1);
    w.cid = a.take<int32_t>(n + 1);
    w.csize = a.take<int32_t>(n + 1);
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

__device__ __forceinline__ unsigned long long mix64(unsigned a, unsigned b) {
    unsigned long long x = ((unsigned long long)a << 32) | b;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// info[v] = (label, last): an edge u -> v has its reverse v -> u  <=>  u <= last.  Untruncated lists
// give last = INT_MAX; a full list gives its final (largest) entry.
__global__ void k_cl_prep(const int32_t *__restrict__ label, const int32_t *__restrict__ idx,
                          const int2 *__restrict__ start_len, int32_t N, int64_t nActive, int2 *__restrict__ info,
                          int32_t *__restrict__ parent, int32_t *__restrict__ lab, unsigned long long *scalars) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int2 sl = start_len[v];
    int last = 0x7fffffff;
    if (sl.y < 0 || sl.x < 0 || (int64_t)sl.x + sl.y > nActive) scalars[1] = 2;   // malformed row
    else if (sl.y >= kCapC) last = __ldg(idx + sl.x + sl.y - 1);
    info[v] = make_int2(__ldg(label + v), last);
    parent[v] = v;
    lab[v] = v;
}

__device__ __forceinline__ int uf_find(int32_t *parent, int x) {
    int p = parent[x];
    while (p != x) {
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving; pointers only ever move to smaller ancestors
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        const int hi = max(a, b), lo = min(a, b);
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
    }
}

// Edge sweep of the fast path.  G lanes share one point's list (coalesced reads of idx).
template <int G>
__global__ void __launch_bounds__(256) k_cl_union(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                  const int2 *__restrict__ info, int32_t N, int32_t *parent,
                                                  unsigned long long *scalars) {
    const int sub = threadIdx.x % G;
    const int64_t groups = (int64_t)gridDim.x * (blockDim.x / G);
    unsigned long long chk = 0;
    bool bad = false, residual = false;
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; i < N; i += groups) {
        const int2 sl = start_len[i];
        const int li = info[i].x;
        for (int e = sub; e < sl.y; e += G) {
            const int j = __ldg(idx + sl.x + e);
            if ((unsigned)j >= (unsigned)N) { bad = true; continue; }
            if (e > 0 && __ldg(idx + sl.x + e - 1) >= j) bad = true;   // lists must ascend for the O(1) symmetry test
            const int2 fj = __ldg(info + j);
            const bool sym = (int)i <= fj.y;
            if (sym) chk += mix64((unsigned)i, (unsigned)j) - mix64((unsigned)j, (unsigned)i);
            if (fj.x != li) continue;
            if (!sym) residual = true;
            else if (j < (int)i) uf_union(parent, (int)i, j);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(0xffffffffu, chk, o);
    if ((threadIdx.x & 31) == 0 && chk) atomicAdd(&scalars[0], chk);
    if (bad) atomicMax(&scalars[1], 1ULL);
    if (residual) scalars[2] = 1;
}

// Roots go to their own array: writing them back into `parent` would race with the path-halving
// stores of other threads' finds, which may re-install an intermediate ancestor after the root.
__global__ void k_cl_flatten(int32_t *parent, int32_t *__restrict__ root, int32_t N) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) root[v] = uf_find(parent, v);
}

__global__ void k_cl_reset(int32_t *__restrict__ root, int32_t *__restrict__ lab, int32_t N) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) { root[v] = v; lab[v] = v; }
}

// resolve() on the label forest.  Unlike uf_find, its shortcut writes must be atomicMin: lab[x] is
// also the target of the propagation's atomicMin, and a plain store could undo a concurrent lowering.
__device__ __forceinline__ int lab_resolve(int32_t *lab, int x) {
    int p = lab[x];
    while (p != x) {
        const int gp = lab[p];
        if (gp != p) atomicMin(&lab[x], gp);
        x = p;
        p = gp;
    }
    return x;
}

// Min-ancestor propagation: for each one-way edge i -> j, lab[root(j)] <- min(., resolve(root(i))).
// resolve() follows lab to its fixed point; every value it passes through reaches the start node, so
// shortcutting (path halving) is sound on a directed graph too.
template <int G, bool ALL_EDGES>
__global__ void __launch_bounds__(256) k_cl_propagate(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                      const int2 *__restrict__ info, int32_t N,
                                                      const int32_t *__restrict__ parent, int32_t *lab,
                                                      unsigned long long *scalars) {
    const int sub = threadIdx.x % G;
    const int64_t groups = (int64_t)gridDim.x * (blockDim.x / G);
    bool changed = false;
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; i < N; i += groups) {
        const int2 sl = start_len[i];
        const int li = info[i].x;
        int mine = -1;
        for (int e = sub; e < sl.y; e += G) {
            const int j = __ldg(idx + sl.x + e);
            if ((unsigned)j >= (unsigned)N) continue;
            const int2 fj = __ldg(info + j);
            if (fj.x != li) continue;
            if (!ALL_EDGES && (int)i <= fj.y) continue;      // symmetric edge: already merged by the unions
            if (mine < 0) mine = lab_resolve(lab, parent[i]);   // lazily: most points have no one-way edge
            const int rj = parent[j];
            // only rj itself may be relabelled: `mine` reaches rj, but not necessarily rj's current label
            if (mine < lab_resolve(lab, rj) && atomicMin(&lab[rj], mine) > mine) changed = true;
        }
    }
    if (changed) scalars[3] = 1;
}

// final label per point, sizes per label
__global__ void k_cl_label(const int32_t *__restrict__ parent, int32_t *lab, int32_t N, int32_t *__restrict__ size,
                           uint32_t *__restrict__ key0) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int l = lab_resolve(lab, parent[v]);
    key0[v] = (uint32_t)l;
    atomicAdd(&size[l], 1);
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
                           int32_t threshold, int32_t *__restrict__ csize, unsigned long long *scalars) {
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
    k_cl_prep<<<nb, 256, 0, st>>>(semantic_label, ball_query_idxs, sl, N, nActive, w.info, w.parent, w.lab, w.scalars);
    unsigned long long h[4] = {0, 0, 0, 0};
    bool use_generic = generic != 0;
    if (!use_generic) {
        if (wide) k_cl_union<32><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.parent, w.scalars);
        else k_cl_union<8><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.parent, w.scalars);
        k_cl_flatten<<<nb, 256, 0, st>>>(w.parent, w.root, N);
        PG_LAUNCH_CHECK();
        PG_CUDA(cudaMemcpyAsync(h, w.scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        if (h[0] != 0 || h[1] != 0) use_generic = true;   // not a truncated symmetric relation
    } else {
        PG_CUDA(cudaMemcpyAsync(h, w.scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
    }
    if (h[1] == 2) {   // the reference would read out of bounds here (bfs_cluster.cpp:40-42)
        set_error("pg_bfs_cluster_count: start_len has a row outside ball_query_idxs[0..%lld)", (long long)nActive);
        return PG_EINVAL;
    }
    if (use_generic) k_cl_reset<<<nb, 256, 0, st>>>(w.root, w.lab, N);
    if (use_generic || h[2] != 0) {
        for (int it = 0; it < 100000; it++) {
            PG_CUDA(cudaMemsetAsync(w.scalars + 3, 0, sizeof(unsigned long long), st));
            if (use_generic) {
                if (wide) k_cl_propagate<32, true><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.root, w.lab, w.scalars);
                else k_cl_propagate<8, true><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.root, w.lab, w.scalars);
            } else {
                if (wide) k_cl_propagate<32, false><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.root, w.lab, w.scalars);
                else k_cl_propagate<8, false><<<eg, 256, 0, st>>>(ball_query_idxs, sl, w.info, N, w.root, w.lab, w.scalars);
            }
            PG_LAUNCH_CHECK();
            unsigned long long changed = 0;
            PG_CUDA(cudaMemcpyAsync(&changed, w.scalars + 3, sizeof(changed), cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            if (!changed) break;
        }
    }
    k_cl_label<<<nb, 256, 0, st>>>(w.root, w.lab, N, w.size, w.key0);
    k_cl_keep<<<nb, 256, 0, st>>>(w.key0, w.size, N, threshold, w.cid);
    PG_CUDA(cudaMemsetAsync(w.cid + N, 0, sizeof(int32_t), st));
    PG_TRY(scan_exclusive_i32(w.cid, w.cid, (int64_t)N + 1, (int64_t *)(w.scalars + 4), w.scan_tmp, st));
    k_cl_sizes<<<nb, 256, 0, st>>>(w.size, w.cid, N, threshold, w.csize, w.scalars);
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
