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
//                component's minimum) over the two-way edges, in three steps:
//                  sample   every point is united with three of its neighbours (first, middle, last);
//                           in a ball-query graph that already connects almost every component;
//                  flatten  root[v] = find(v): a read-only snapshot of the partition so far;
//                  verify   ONE sweep over all edges, lists taken in the order their segments sit in
//                           idx (= spatial order when the producer is this library's ball query), so
//                           the snapshot reads of a block hit L1: an edge whose ends share a snapshot
//                           root is done; the rare others go through the coherent union-find.  One-way
//                           edges between different components are parked in a pending list and
//                           settled afterwards by a min-label propagation over that (tiny) list.
//   generic path min-label propagation over ALL edges with no unions: exact for any directed graph.
// The fast path is only trusted when the lists check out as a truncated symmetric relation: in range,
// ascending, and a 64-bit checksum of the two-way edge set that cancels against its own reversal.
// Both checks ride along in the verify sweep and need no extra memory traffic.
#include "common.cuh"
#include "ballquery.cuh"
#include "scan.cuh"

namespace pg {

constexpr int kCapC = PG_BALLQUERY_CAP;

struct ClWs {
    uint2 *pl;           // per point: x = union-find parent, y = semantic label (one 8-byte read per edge)
    uint32_t *trunc;     // bitmap: list is full (len >= 1000), i.e. may lack reverse edges
    int32_t *last;       // last entry of a full list: u -> v is two-way  <=>  u <= last[v]
    int32_t *root;       // flattened component root per point (identity on the generic path)
    uint32_t *snap;      // root | label bits | full bit: what the sweep reads per edge
    int32_t *lab;        // min-ancestor label forest over roots
    int32_t *size;       // points per final label
    int32_t *cid;        // cluster id per kept label (exclusive scan of keep flags)
    int32_t *csize;      // sizes in cluster order -> offsets
    int2 *pend;          // parked one-way edges (i -> j)
    int4 *samples;       // lazy lists: (first, second, middle, last) entry of every point's list
    uint2 *cellsum;      // grid-assisted mode, per ball-query cell: (main label, snapshot root)
    int32_t *cstate;     //   skip threshold L per cell
    int32_t *cthr;       //   three minima of last(j) over the cell's full points (T1, T2, T3)
    uint32_t *key0, *kA, *vA, *kB, *vB;
    int32_t *pubA, *pubB;  // a tree's new root, published by the tree's old root for its other points (find_via_old_root)
    int32_t *hist;
    int64_t *scan_tmp;
    // [0] checksum [1] bad [2] pending count [3] changed [4] nCluster [5] sumNPoint [6] verify work counter
    // [7] some list is full [8] lists left to the sweep (grid-assisted mode) [9] [11] block tickets of the ordered passes
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
    w.snap = a.take<uint32_t>(n);
    w.lab = a.take<int32_t>(n);
    w.size = a.take<int32_t>(n + 1);
    w.cid = a.take<int32_t>(n + 1);
    w.csize = a.take<int32_t>(n + 1);
    w.cellsum = a.take<uint2>(n);
    w.cstate = a.take<int32_t>(n);
    w.cthr = a.take<int32_t>(3 * n);
    w.samples = a.take<int4>(n);
    w.key0 = a.take<uint32_t>(n);
    w.kA = a.take<uint32_t>(n);
    w.vA = a.take<uint32_t>(n);
    w.kB = a.take<uint32_t>(n);
    w.vB = a.take<uint32_t>(n);
    w.pubA = a.take<int32_t>(n);
    w.pubB = a.take<int32_t>(n);
    w.hist = a.take<int32_t>(radix_tmp_count(N_));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(n + radix_tmp_count(N_))));
    w.scalars = a.take<unsigned long long>(12);
    // the parking lot for one-way edges comes last and takes whatever the caller's buffer holds beyond the minimum:
    // a scene with hundreds of thousands of truncated lists parks millions of them
    w.pend_cap = n + 1024;
    if (ws && ws_bytes > a.used + (w.pend_cap + 64) * sizeof(int2)) w.pend_cap = (ws_bytes - a.used) / sizeof(int2) - 64;
    w.pend = a.take<int2>(w.pend_cap);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

// 64-bit tag of an ordered pair (hi, lo): the product of two independently scrambled 32-bit words.
// Cheap (three integer multiplies); it only has to make accidental cancellation of unmatched edges
// in the checksum a 2^-64-class event, not resist an adversary.
__device__ __forceinline__ unsigned long long mix64(unsigned a, unsigned b) {
    const unsigned x = a * 0x9E3779B1u + 0x5bd1e995u;
    const unsigned y = b * 0x85EBCA77u + 0x7f4a7c15u;
    return (unsigned long long)(x ^ (x >> 15)) * (unsigned long long)((y ^ (y >> 13)) | 1u);
}

// one thread per point, launched over ceil(N / 32) * 32 threads so every bitmap word has a full warp.
// Besides the per-point records it plants the first tree edges without a single atomic: a point
// hangs itself under the first entry of its list (the smallest index in its ball) when that is a
// smaller, equal-label point that lists it back -- parents only ever point to smaller indices, so the
// result is a forest, and in a ball-query graph every ball-sized neighbourhood is already one tree.
// `samples` (lazy lists): the entries this kernel would read from idx -- first, second and last of every list -- come from
// there instead; idx is not touched.
__global__ void k_cl_prep(const int32_t *__restrict__ label, const int32_t *__restrict__ idx,
                          const int2 *__restrict__ start_len, int32_t N, int64_t nActive, int hook,
                          uint2 *__restrict__ pl, uint32_t *__restrict__ trunc, int32_t *__restrict__ last,
                          int32_t *__restrict__ root, int32_t *__restrict__ lab, uint32_t *__restrict__ skey, int shift,
                          unsigned long long *scalars, const int4 *__restrict__ samples, Fill sizes) {
    pdl_enter();
    grid_fill(sizes);      // the label pass's counters
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool full = false;
    if (v < N) {
        const int2 sl = start_len[v];
        skey[v] = (uint32_t)(sl.x < 0 ? 0 : sl.x) >> shift;   // where the list sits in idx: the sweep order
        int lst = 0x7fffffff;
        int parent = v;
        const int lv = __ldg(label + v);
        if (sl.y < 0 || sl.x < 0 || (int64_t)sl.x + sl.y > nActive) scalars[1] = 2;   // malformed row
        else {
            int4 sm = make_int4(v, v, v, v);
            if (samples) sm = __ldg(samples + v);
            if (sl.y >= kCapC) { lst = samples ? sm.w : __ldg(idx + sl.x + sl.y - 1); full = true; }
            if (hook && sl.y > 0) {
                int j = samples ? sm.x : __ldg(idx + sl.x);
                if (j == v && sl.y > 1) j = samples ? sm.y : __ldg(idx + sl.x + 1);
                if (j >= 0 && j < v && __ldg(label + j) == lv) {
                    const int2 sj = __ldg(start_len + j);
                    bool back = true;                              // does j list v?  (only a full list may not)
                    if (sj.y >= kCapC)
                        back = sj.x >= 0 && (int64_t)sj.x + sj.y <= nActive &&
                               v <= (samples ? __ldg(&samples[j].w) : __ldg(idx + sj.x + sj.y - 1));
                    if (back) parent = j;
                }
            }
        }
        pl[v] = make_uint2((unsigned)parent, (unsigned)lv);
        last[v] = lst;
        root[v] = v;
        lab[v] = v;
    }
    const unsigned m = __ballot_sync(0xffffffffu, full);
    if ((threadIdx.x & 31) == 0 && (v >> 5) <= (N >> 5)) trunc[v >> 5] = m;
    if (m && (threadIdx.x & 31) == 0) scalars[7] = 1;
}

// union-find on pl[].x.  Parent reads are ordinary (L1-cached) loads: a component's root line is read
// by every find that ends there, and served from L2 alone it becomes a serialised hot spot.  A stale
// parent is still an ancestor (pointers only ever move towards the root), so stale reads cost steps,
// never correctness; the one place that needs fresh data takes it from the CAS itself (below).
__device__ __forceinline__ int uf_find(uint2 *pl, int x) {
    int p = (int)pl[x].x;
    while (p != x) {
        const int gp = (int)pl[p].x;
        if (gp != p) pl[x].x = (unsigned)gp;   // path halving; pointers only ever move to smaller ancestors
        x = p;
        p = gp;
    }
    return x;
}

// a, b: roots as far as the caller knows; returns the root of the merged set as far as this thread can tell.
// When the CAS fails its return value IS the fresh parent of `hi`: the retry continues from there, so a
// stale "still a root" in L1 cannot spin the loop -- every failure moves strictly up the tree.
__device__ __forceinline__ int uf_union_roots(uint2 *pl, int a, int b) {
    for (;;) {
        if (a == b) return a;
        const int hi = max(a, b), lo = min(a, b);
        const unsigned old = atomicCAS(&pl[hi].x, (unsigned)hi, (unsigned)lo);
        if (old == (unsigned)hi) return lo;
        a = uf_find(pl, (int)old);     // hi was hooked by somebody else meanwhile
        b = uf_find(pl, lo);
    }
}

// Between two flattens only ROOTS get new parents (k_cl_sample and the sweep hang a root under another root), so every
// point of a tree reaches the forest through its old root A = root[v] <= v, and all of them would walk the same chain of
// hooked roots from there -- thousands of threads in the same few lines.  Instead A's own thread walks the chain once and
// publishes the result in pub[A] (-1 before); the tree's other points wait for it.  Safe because the blocks of these passes
// take their index range from a ticket counter: A <= v sits in the same or an earlier ticket, whose block is running or
// done; inside a warp the roots go first (__syncwarp).  Every lane of the warp must call this (v >= N: pass v = -1).
// (After the SWEEP the chains of hooked roots are long and everybody's path halving helps: the label pass keeps the
// plain find -- see k_cl_label.)
__device__ __forceinline__ int find_via_old_root(uint2 *pl, const int32_t *root_old, volatile int32_t *pub, int v) {
    const int A = v >= 0 ? root_old[v] : -1;
    int r = -1;
    if (v >= 0 && A == v) {
        r = uf_find(pl, v);
        pub[v] = r;
    }
    __syncwarp();
    if (v >= 0 && A != v) {
        while ((r = pub[A]) < 0) {}
    }
    return r;
}

// the block's index range from a ticket (see find_via_old_root)
__device__ __forceinline__ int ticket_block(unsigned long long *ticket) {
    __shared__ int s_b;
    if (threadIdx.x == 0) s_b = (int)atomicAdd(ticket, 1ULL);
    __syncthreads();
    return s_b;
}

// Snapshot word per point, written by k_cl_flatten and read (through the read-only L1 path) once per edge
// by the sweep:  [31] the point's list is full   [30:26] low 5 bits of its label   [25:0] its root.
// Equal roots = already connected.  Different label bits = never connected.  Only a pair that differs in
// the root field alone needs the forest.
constexpr unsigned kSnapRoot = 0x03ffffffu, kSnapLabel = 0x7c000000u, kSnapFull = 0x80000000u;

// sample rounds: between two flattens every point probes two of its neighbours and, where the snapshot
// roots differ, tries to hang the larger root under the smaller with ONE compare-and-swap -- no find, no
// retry: if the larger root has been taken meanwhile the link is simply dropped (some other link took
// it; whatever stays unmerged is merged by the verify sweep).  Every root that sees a smaller
// neighbouring root is hooked by somebody, so each round shrinks the forest Boruvka-fashion.  Only
// two-way edges between equal labels are used, exactly the edges the sweep would unite along.
constexpr int kSampleRounds = 1;

template <int P>
__global__ void k_cl_sample(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len, uint2 *pl,
                            const int32_t *__restrict__ last, const uint32_t *__restrict__ snap, int32_t N, int64_t nActive,
                            int round, const int4 *__restrict__ samples) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int2 sl = start_len[i];
    if (sl.y <= 1) return;
    // a row outside idx[0, nActive) (foreign lists; k_cl_prep has flagged it, the host raises PG_EINVAL): never read
    if (sl.x < 0 || (int64_t)sl.x + sl.y > nActive) return;
    // round 0: last and middle entry (P = 2), plus the quarter points (P = 4); later rounds: the eighths, ...
    const int den = 2 << round;
    int pos[4];
    pos[0] = round == 0 ? sl.y - 1 : (int)(((long long)sl.y) / den);
    pos[1] = round == 0 ? sl.y >> 1 : (int)(((long long)sl.y * (den - 1)) / den);
    pos[2] = sl.y >> 2;
    pos[3] = (int)(((long long)sl.y * 3) >> 2);
    int jj[P];
    if (samples) {                                  // lazy lists: the last and a middle entry were sampled from the hit masks
        const int4 sm = __ldg(samples + i);
#pragma unroll
        for (int u = 0; u < P; u++) jj[u] = (u & 1) ? sm.z : sm.w;
    } else {
#pragma unroll
        for (int u = 0; u < P; u++) jj[u] = __ldg(idx + sl.x + pos[u]);
    }
    const unsigned si = __ldg(snap + i);
    unsigned sj[P];
#pragma unroll
    for (int u = 0; u < P; u++) sj[u] = ((unsigned)jj[u] < (unsigned)N) ? __ldg(snap + jj[u]) : si;
    int ri = (int)(si & kSnapRoot);
#pragma unroll
    for (int u = 0; u < P; u++) {
        const int j = jj[u];
        const unsigned x = (si ^ sj[u]) & ~kSnapFull;
        if (x == 0u || x > kSnapRoot) continue;                      // same tree, or different label bits
        if (pl[j].y != pl[i].y) continue;
        if ((sj[u] & kSnapFull) && i > __ldg(last + j)) continue;    // j does not list i back
        const int rj = (int)(sj[u] & kSnapRoot);
        if (rj == ri) continue;
        const int hi = max(ri, rj), lo = min(ri, rj);
        if (atomicCAS(&pl[hi].x, (unsigned)hi, (unsigned)lo) == (unsigned)hi && hi == ri) ri = lo;
    }
}

// verify: the single sweep over all edges.  G lanes share one point's list (coalesced, streaming reads of
// idx).  Lists are taken in `order` (ascending segment start): a block claims a chunk of consecutive
// lists from an atomic counter, so the lists it works on at any time belong to neighbouring points
// and their snapshot reads mostly hit L1.  Everything with latency is taken one step ahead: the next
// chunk is claimed and its list headers (point, start, length, snapshot) are fetched into shared
// memory while the current chunk is processed; inside a chunk every lane group pulls lists from a
// shared counter and always has the idx loads of its NEXT trip in flight, across list boundaries.
// TRUSTED: the caller vouches that the lists are this library's ball-query output (in range, a
// truncated symmetric relation) -- no range checks, no checksum.
constexpr int kVerThreads = 512;

template <int G, bool TRUSTED>
__global__ void __launch_bounds__(kVerThreads, 3) k_cl_verify(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                              const uint32_t *__restrict__ order, uint2 *pl,
                                                              const int32_t *__restrict__ last,
                                                              const uint32_t *__restrict__ snap, int32_t N, int64_t nActive,
                                                              int2 *__restrict__ pend, unsigned long long pend_cap,
                                                              unsigned long long *scalars, int use_list_count) {
    pdl_enter();
    constexpr int kGroups = kVerThreads / G;
    // lists to sweep: all N of them in `order`, or the scalars[8] the cell pass left over (grid-assisted mode)
    const long long NL = use_list_count ? (long long)scalars[8] : (long long)N;
    constexpr int kChunk = kGroups * 8 > kVerThreads ? kVerThreads : kGroups * 8;   // lists per claim (<= one per thread)
    constexpr int kU = 4;
    __shared__ int4 hdr[2][kChunk];              // (point, start, length, snapshot word) per list of a chunk
    __shared__ long long s_base[2];
    __shared__ int s_take[2];                    // next unclaimed list of the chunk
    const int tid = threadIdx.x, sub = tid % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((tid & 31) / G * G));
    unsigned long long chk = 0, chk_rev = 0;
    bool bad = false;
    int park_a = -1, park_b = -1;                // the pair of sets this lane parked last
    const bool any_full = scalars[7] != 0;       // some list holds kCap entries: reverse edges may be missing

    auto fetch_header = [&](long long base, int4 &h) {
        h = make_int4(0, 0, 0, 0);
        const long long p = base + tid;
        if (tid < kChunk && p < NL) {
            const int i = (int)__ldg(order + p);
            int2 sl = __ldg(start_len + i);
            // validating sweep: a row outside idx[0, nActive) is swept as an empty list (k_cl_prep has flagged it and
            // the host answers PG_EINVAL) -- the reference would read out of bounds there (bfs_cluster.cpp:40-42)
            if (!TRUSTED && (sl.x < 0 || sl.y < 0 || (int64_t)sl.x + sl.y > nActive)) sl.y = 0;
            h = make_int4(i, sl.x, sl.y, (int)__ldg(snap + i));
        }
    };
    if (tid == 0) {
        s_base[0] = (long long)atomicAdd(&scalars[6], (unsigned long long)kChunk);
        s_base[1] = (long long)atomicAdd(&scalars[6], (unsigned long long)kChunk);
        s_take[0] = 0;
        s_take[1] = 0;
    }
    __syncthreads();
    {
        int4 h;
        fetch_header(s_base[0], h);
        if (tid < kChunk) hdr[0][tid] = h;
    }
    __syncthreads();
    for (int round = 0;; round++) {
        const int cur = round & 1;
        const long long base = s_base[cur];
        if (base >= NL) break;
        const int nl = (int)(NL - base < kChunk ? NL - base : kChunk);
        // one step ahead: headers of the next chunk (its base was claimed a round ago), claim of the one after
        const long long base1 = s_base[cur ^ 1];
        int4 h1;
        fetch_header(base1, h1);
        long long claim = 0;
        if (tid == 0) claim = (long long)atomicAdd(&scalars[6], (unsigned long long)kChunk);

        auto grab = [&]() {                       // next list of the chunk for this lane group (-1: none left)
            int L = 0;
            if (sub == 0) L = atomicAdd(&s_take[cur], 1);
            L = __shfl_sync(gmask, L, 0, G);
            return L < nl ? L : -1;
        };
        auto load_trip = [&](const int4 &H, int e0, int (&nj)[kU]) {
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int e = e0 + u * G;
                nj[u] = e < H.z ? __ldcs(idx + H.y + e) : H.x;        // padding reads as the self edge
            }
        };
        int L = grab();
        if (L >= 0) {
            int4 H = hdr[cur][L];
            int e0 = sub;
            int nj[kU];
            load_trip(H, e0, nj);
            int ri = (int)((unsigned)H.w & kSnapRoot);   // current root of the list owner's set as far as this group knows
            for (;;) {
                int jj[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) jj[u] = nj[u];
                // the next trip (possibly of the next list) goes in flight before this one is looked at
                int4 H2 = H;
                int e2 = e0 + kU * G;
                bool more = true;
                if (e2 - sub >= H.z) {
                    const int L2 = grab();
                    more = L2 >= 0;
                    if (more) { H2 = hdr[cur][L2]; e2 = sub; }
                }
                if (more) load_trip(H2, e2, nj);

                const int i = H.x;
                const unsigned si = (unsigned)H.w;
                unsigned sj[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    if (!TRUSTED && (unsigned)jj[u] >= (unsigned)N) { bad = true; jj[u] = i; }
                    sj[u] = __ldg(snap + jj[u]);
                }
                if (!TRUSTED) {
#pragma unroll
                    for (int u = 0; u < kU; u++) {
                        if (jj[u] == i) continue;
                        const bool twoway = !any_full || !(sj[u] & kSnapFull) || i <= __ldg(last + jj[u]);
                        // each two-way pair {a > b} is seen as a -> b and as b -> a: the two sums must agree
                        const unsigned long long t = mix64((unsigned)max(i, jj[u]), (unsigned)min(i, jj[u]));
                        if (twoway && jj[u] < i) chk += t;
                        if (twoway && jj[u] > i) chk_rev += t;
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const unsigned x = (si ^ sj[u]) & ~kSnapFull;
                    if (x == 0u || x > kSnapRoot) continue;        // the common cases end here
                    const int j = jj[u];
                    if (pl[j].y != pl[i].y) continue;
                    const bool twoway = !(sj[u] & kSnapFull) || i <= __ldg(last + j);
                    const int a = uf_find(pl, ri), b = uf_find(pl, (int)(sj[u] & kSnapRoot));
                    ri = a;
                    if (a == b) continue;
                    if (twoway) {
                        ri = uf_union_roots(pl, a, b);
                    } else if (a != park_a || b != park_b) {
                        // one-way edge between two sets that are not connected (yet): park it as (a, b) -- the
                        // propagation only looks at the sets of its ends -- unless this lane has just parked the
                        // same pair (a truncated list points at the same few sets a thousand times)
                        park_a = a;
                        park_b = b;
                        const unsigned long long slot = atomicAdd(&scalars[2], 1ULL);
                        if (slot < pend_cap) pend[slot] = make_int2(a, b);
                    }
                }
                if (!more) break;
                if (H2.x != H.x) ri = (int)((unsigned)H2.w & kSnapRoot);
                H = H2;
                e0 = e2;
            }
        }
        __syncthreads();                          // everyone is done with hdr[cur], s_take[cur], s_base[cur]
        if (tid < kChunk) hdr[cur ^ 1][tid] = h1;   // (published for the next round by the barrier below)
        if (tid == 0) { s_base[cur] = claim; s_take[cur] = 0; }
        __syncthreads();
    }
    if (!TRUSTED) {
        chk -= chk_rev;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(0xffffffffu, chk, o);
        if ((threadIdx.x & 31) == 0 && chk) atomicAdd(&scalars[0], chk);
        if (bad) atomicMax(&scalars[1], 1ULL);
    }
}


// ---- grid-assisted sweep (trusted lists + the producing ball query's workspace) --------------------------
// The ball query's uniform grid is still in its workspace: cells, each cell's points, each cell's 27 neighbour
// cells; every list of a point of cell A is a subset of the points of those 27 cells.  After the sampling rounds
// almost every component is one tree plus a few stray twigs.  Each cell gets a main pair (M, R): the label and
// snapshot root most of its points carry.  Call a point with label M and another root a STRAY of that pair.
// A point i of cell A carrying (M_A, R_A) has nothing to tell the sweep -- its list is never read -- when
//     every stray j of (M_A, R_A) in A's 27 cells lists i back and is itself swept.
// Proof.  Take j in list(i) with label(j) = M_A (other labels are ignored by the sweep anyway).  Root R_A: already
// i's set -- equal snapshot roots stay equal, sets only merge -- so the edge neither unites nor, if it is one-way,
// needs parking.  A stray: it lists i, it is swept, and from its side the pair is classified exactly as from i's
// (two-way iff j <= last(i) when list(i) is full, which holds because j is IN list(i)).
// "Lists i back": d(i, j) < r, so a complete list of j holds i, and a full one (the first 1000 of j's ball by
// index) holds i iff i <= last(j).  Hence a per-cell THRESHOLD L_A = min last(j) over the full strays around A
// (INT_MAX without any): points of A carrying the main pair are skipped iff their index is <= L_A.
// "Is itself swept": a point is kept for the sweep unless it carries its own cell's main pair and passes its cell's
// threshold; a stray of (M_A, R_A) either has another root than its cell's main pair, or another label, or sits in
// a cell B with M_B = M_A and R_B != R_A -- and of two such adjacent cells only the one with the SMALLER root may
// lean on the other: the larger one gets L = -1, all its points are swept.
// Per neighbour B of A, from three per-cell minima of last(j) over B's full points -- T1: strays of B's own main
// pair, T2: points with label M_B, T3: all --
//     M_B = M_A, R_B = R_A :  L_A = min(L_A, T1_B)
//     M_B = M_A, R_B != R_A:  L_A = min(L_A, T2_B) if R_A < R_B, else L_A = -1
//     M_B != M_A           :  L_A = min(L_A, T3_B)      (label-M_A points are the minority in B and always swept)
// Four small passes, none with a long dependent chain: main per cell -> thresholds per (full) point -> L per cell
// -> work list per point in query order (lists sharing entries stay together).
// Cost: a few loads per point and 27 per cell, instead of one snapshot read per EDGE (~330 per point on
// shifted coordinates).
__global__ void k_cl_cell_main(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cstart,
                               const int32_t *__restrict__ ccnt, const int64_t *__restrict__ bq_scalars,
                               const uint2 *__restrict__ pl, const uint32_t *__restrict__ snap, uint2 *__restrict__ cellsum,
                               int32_t *__restrict__ cthr) {
    pdl_enter();
    const int64_t nCells = ld_after_wait(bq_scalars);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += (int64_t)gridDim.x * blockDim.x) {
        const int nq = __ldg(ccnt + c), qs = __ldg(cstart + c);
        const uint32_t p0 = __ldg(sorted_pt + qs);
        unsigned M = pl[p0].y, R = __ldg(snap + p0) & kSnapRoot;
        if (nq >= 3) {                       // majority of three: one noisy label or one stray does not name the cell
            const uint32_t p1 = __ldg(sorted_pt + qs + 1), p2 = __ldg(sorted_pt + qs + 2);
            const unsigned l1 = pl[p1].y, l2 = pl[p2].y;
            const unsigned r1 = __ldg(snap + p1) & kSnapRoot, r2 = __ldg(snap + p2) & kSnapRoot;
            if (l1 == l2 && r1 == r2 && (l1 != M || r1 != R)) { M = l1; R = r1; }
        }
        cellsum[c] = make_uint2(M, R);
        cthr[3 * c] = cthr[3 * c + 1] = cthr[3 * c + 2] = 0x7fffffff;
    }
}

__global__ void k_cl_cell_thresholds(const int32_t *__restrict__ cell, const uint2 *__restrict__ pl,
                                     const uint32_t *__restrict__ snap, const int32_t *__restrict__ last, int32_t N,
                                     const uint2 *__restrict__ cellsum, int32_t *cthr) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned sv = __ldg(snap + i);
    if (!(sv & kSnapFull)) return;           // only full lists can fail to list a neighbour back
    const int c = __ldg(cell + i);
    const uint2 cs = cellsum[c];
    const int lst = __ldg(last + i);
    int32_t *t = cthr + 3 * (int64_t)c;
    if (t[2] > lst) atomicMin(t + 2, lst);
    if (pl[i].y == cs.x) {
        if (t[1] > lst) atomicMin(t + 1, lst);
        if ((sv & kSnapRoot) != cs.y && t[0] > lst) atomicMin(t, lst);
    }
}

// cstate[c] = L_c: points of c carrying its main pair are skipped iff their index is <= L_c (-1: none)
__global__ void k_cl_cell_settle(const int32_t *__restrict__ nbr, const int64_t *__restrict__ bq_scalars,
                                 const uint2 *__restrict__ cellsum, const int32_t *__restrict__ cthr,
                                 int32_t *__restrict__ cstate) {
    pdl_enter();
    const int64_t nCells = ld_after_wait(bq_scalars);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint2 me = cellsum[c];
        int L = cthr[3 * c];
        const int32_t *nb = nbr + c * 27;
#pragma unroll 9
        for (int j = 0; j < 27; j++) {
            const int b = __ldg(nb + j);
            if (b < 0 || b == (int)c) continue;
            const uint2 o = cellsum[b];
            const int32_t *t = cthr + 3 * (int64_t)b;
            if (o.x != me.x) L = min(L, t[2]);
            else if (o.y == me.y) L = min(L, t[0]);
            else if (me.y < o.y) L = min(L, t[1]);
            else { L = -1; break; }
        }
        cstate[c] = L;
    }
}

__global__ void k_cl_cell_worklist(const uint32_t *__restrict__ sorted_pt, const int32_t *__restrict__ cell,
                                   const uint2 *__restrict__ pl, const uint32_t *__restrict__ snap,
                                   const uint2 *__restrict__ cellsum, const int32_t *__restrict__ cstate, int32_t N,
                                   uint32_t *__restrict__ worklist, unsigned long long *scalars) {
    pdl_enter();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    uint32_t i = 0;
    if (q < N) {
        i = __ldg(sorted_pt + q);
        const int c = __ldg(cell + i);
        const uint2 cs = cellsum[c];
        keep = !((int)i <= __ldg(cstate + c) && pl[i].y == cs.x && (__ldg(snap + i) & kSnapRoot) == cs.y);
    }
    // blocks append in whatever order they run; inside a block the query order is kept (a cell's lists share
    // their entries, and cells rarely straddle blocks)
    __shared__ unsigned long long s_base;
    __shared__ int s_warp[8];
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(km);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int t = s_warp[w]; s_warp[w] = tot; tot += t; }
        s_base = tot ? atomicAdd(&scalars[8], (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    if (keep) worklist[s_base + s_warp[warp] + __popc(km & lanemask_lt())] = i;
}

// Roots go to their own array: writing them back into the forest would race with the path-halving
// stores of other threads' finds, which may re-install an intermediate ancestor after the root.
// VIA: root[] holds the previous flatten's roots -- the find goes through them (find_via_old_root, pub_use); either way
// pub_reset is set to -1 for the next ordered pass.
template <bool VIA>
__global__ void __launch_bounds__(256) k_cl_flatten(uint2 *pl, const uint32_t *__restrict__ trunc, int32_t *root,
                                                    uint32_t *__restrict__ snap, int32_t N, int32_t *pub_use,
                                                    int32_t *__restrict__ pub_reset, unsigned long long *ticket) {
    pdl_enter();
    int v;
    int r;
    if (VIA) {
        v = ticket_block(ticket) * blockDim.x + threadIdx.x;
        r = find_via_old_root(pl, root, pub_use, v < N ? v : -1);
        if (v >= N) return;
    } else {
        v = blockIdx.x * blockDim.x + threadIdx.x;
        if (v >= N) return;
        r = uf_find(pl, v);
    }
    root[v] = r;
    pl[v].x = (unsigned)r;        // a hint only (another thread's halving store may replace it with another ancestor)
    if (pub_reset) pub_reset[v] = -1;
    const unsigned full = (trunc[v >> 5] >> (v & 31)) & 1u;
    snap[v] = (unsigned)r | ((pl[v].y & 31u) << 26) | (full << 31);
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
    pdl_enter();
    bool changed = false;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pend;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const int2 e = pend[t];
        if (root[e.x] != root[e.y]) changed |= propagate_edge(root, lab, e.x, e.y);
    }
    if (changed) scalars[3] = 1;
}

// The same fixed point without the host in the loop, for the trusted path: the parked edges are a few dozen to a
// few thousand, so ONE block settles them -- the count phase then synchronises the stream once (for the sizes)
// instead of twice.  More than kPendBlockMax parked edges: raise scalars[10]; the host reruns the two-sync path.
constexpr unsigned long long kPendBlockMax = 4096;

// FIND: the forest has not been flattened since the sweep (trusted path): the ends' roots come from uf_find (no unions
// run any more, so concurrent path halving only shortens paths) and are written back into the parked pair.
template <bool FIND>
__global__ void __launch_bounds__(1024) k_cl_pending_block(int2 *pend, unsigned long long pend_cap, uint2 *pl,
                                                           const int32_t *__restrict__ root, int32_t *lab,
                                                           unsigned long long *scalars) {
    pdl_enter();
    const unsigned long long n_pend = scalars[2];
    if (n_pend == 0) return;
    if (n_pend > pend_cap || n_pend > kPendBlockMax) {
        if (threadIdx.x == 0) scalars[10] = 1;
        return;
    }
    __shared__ int s_changed;
    if (FIND) {
        for (unsigned long long t = threadIdx.x; t < n_pend; t += blockDim.x) {
            const int2 e = pend[t];
            pend[t] = make_int2(uf_find(pl, e.x), uf_find(pl, e.y));
        }
        __syncthreads();
    }
    for (int it = 0; it < (1 << 24); it++) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        bool changed = false;
        for (unsigned long long t = threadIdx.x; t < n_pend; t += blockDim.x) {
            const int2 e = pend[t];
            if (FIND) {                                  // the pair holds roots: lab[rb] <- min(., resolve(ra))
                if (e.x != e.y) {
                    const int mine = lab_resolve(lab, e.x);
                    changed |= mine < lab_resolve(lab, e.y) && atomicMin(&lab[e.y], mine) > mine;
                }
            } else if (root[e.x] != root[e.y]) changed |= propagate_edge(root, lab, e.x, e.y);
        }
        if (changed) s_changed = 1;
        __syncthreads();
        const int again = s_changed;
        __syncthreads();
        if (!again) break;
    }
}

// Full sweep: the generic path (ALL same-label edges) or, with ONE_WAY_ONLY, the fall-back when the
// pending list overflowed.
template <int G, bool ONE_WAY_ONLY>
__global__ void __launch_bounds__(256) k_cl_propagate(const int32_t *__restrict__ idx, const int2 *__restrict__ start_len,
                                                      const uint2 *pl, const uint32_t *__restrict__ trunc,
                                                      const int32_t *__restrict__ last, int32_t N,
                                                      const int32_t *__restrict__ root, int32_t *lab,
                                                      unsigned long long *scalars) {
    pdl_enter();
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
    pdl_enter();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) { root[v] = v; lab[v] = v; }
}

// final label per point, sizes per label (warp-aggregated: a floor-sized component would otherwise
// serialise tens of thousands of atomics on one counter)
template <bool FIND>
__global__ void __launch_bounds__(256) k_cl_label(uint2 *pl, int32_t *root, int32_t *lab, int32_t N, int32_t *__restrict__ size,
                                                  uint32_t *__restrict__ key0) {
    pdl_enter();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = v < N;
    int l = -1;
    if (on) {
        int r;
        // the flatten pass folded in (trusted path).  Every thread walks: after the sweep the chains of hooked roots are
        // long and the walkers' path halving helps each other (through the old roots, find_via_old_root: 0.093 -> 0.134 ms
        // waiting, 0.114 ms with a walk where the root has not published yet)
        if (FIND) { r = uf_find(pl, v); root[v] = r; }
        else r = root[v];
        l = lab_resolve(lab, r);
        key0[v] = (uint32_t)l;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    if (on && (peers & lanemask_lt()) == 0) atomicAdd(&size[l], __popc(peers));
}

// keep -> scan -> sizes as one pass (scan.cuh): a label l with >= threshold points is a cluster; its exclusive prefix is
// the cluster's number; cid[0..N] keeps the prefixes (cid[N] = nCluster) for k_cl_keys, the kept labels write their size
// into csize in cluster order and add it to the member total.
struct KeepLoad {
    const uint32_t *key0;
    const int32_t *size;
    int32_t N, threshold;
    __device__ int operator()(int64_t v) const {
        if (v >= N) return 0;
        const int n = size[v];
        return (key0[v] == (uint32_t)v && n > 0 && n >= threshold) ? 1 : 0;
    }
};
struct SizesStore {
    const int32_t *size;
    int32_t *cid, *csize;
    unsigned long long *scalars;
    __device__ void operator()(int64_t v, int c, int keep) const {
        cid[v] = c;
        if (keep) {
            csize[c] = size[v];
            atomicAdd(&scalars[5], (unsigned long long)size[v]);
        }
    }
};

// sort key per point: cluster id, or nCluster for points of dropped components (they sort last).  One extra block
// turns the cluster sizes into the offsets (a few hundred to a few thousand entries: not worth a launch of its own).
__global__ void __launch_bounds__(256) k_cl_keys(const uint32_t *__restrict__ key0, const int32_t *__restrict__ cid, int32_t N,
                                                 int32_t nCluster, uint32_t *__restrict__ keys, const int32_t *__restrict__ csize,
                                                 int32_t *__restrict__ cluster_offsets) {
    pdl_enter();
    if (blockIdx.x == gridDim.x - 1) {                       // offsets[c] = sizes before c; offsets[nCluster] = the total
        __shared__ int warp_tot[32];
        int carry = 0;
        for (int c0 = 0; c0 <= nCluster; c0 += 256) {
            const int c = c0 + threadIdx.x;
            const int x = c < nCluster ? csize[c] : 0;
            int tot;
            const int incl = block_scan_incl(x, warp_tot, &tot);
            if (c <= nCluster) cluster_offsets[c] = carry + incl - x;
            carry += tot;
        }
        return;
    }
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const uint32_t l = key0[v];
    const int c = cid[l], next = cid[l + 1];
    keys[v] = (next != c) ? (uint32_t)c : (uint32_t)nCluster;
}

}  // namespace pg

using namespace pg;

// diagnostics of the most recent pg_bfs_cluster_count on this thread: {checksum != 0, bad, pending, sweeps}
// [4] lists the edge sweep read (N unless the grid-assisted mode settled whole cells first)
static thread_local long long g_cl_dbg[5] = {0, 0, 0, 0, 0};
extern "C" void pg_bfs_cluster_debug(long long *out) { for (int i = 0; i < 5; i++) out[i] = g_cl_dbg[i]; }

extern "C" size_t pg_bfs_cluster_workspace_bytes(int64_t N) {
    if (N < 0) N = 0;
    return cl_layout(nullptr, 0, N).used + 256;
}

// `bq_ws` (optional, trusted mode only): the workspace of the pg_ballquery_* calls that produced the lists,
// untouched since -- its grid lets whole cells skip the edge sweep (k_cl_cells)
// `lazy_masks` (with bq_ws, trusted): the lists exist only as the ball query's hit masks; `ball_query_idxs` is then a
// scratch buffer of nActive ints into which the lists the sweep reads are decoded (the others are never materialised).
// *need_lists is raised (and nothing else returned) when the parked one-way edges overflow their lot and the fall-back
// needs every list: the caller materialises them and calls the ordinary entry point.
static int bfs_count_impl(const int32_t *semantic_label, const int32_t *ball_query_idxs, const int32_t *start_len, int32_t N,
                          int64_t nActive, int32_t threshold, int mode, void *ws, size_t ws_bytes, void *bq_ws,
                          size_t bq_ws_bytes, int32_t *host_sizes, void *stream, const uint32_t *lazy_masks = nullptr,
                          int *need_lists = nullptr) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_sizes, "null host_sizes");
    host_sizes[0] = host_sizes[1] = host_sizes[2] = 0;
    PG_CHECK_ARG(N >= 0 && nActive >= 0, "negative size");
    PG_CHECK_ARG(N <= (1 << 26), "N out of range (0 .. 2^26)");
    PG_CHECK_ARG(mode == PG_BFS_AUTO || mode == PG_BFS_GENERIC || mode == PG_BFS_TRUSTED, "unknown mode");
    if (N == 0) return PG_OK;
    PG_CHECK_ARG(semantic_label && start_len && ws && (ball_query_idxs || nActive == 0), "null pointer");
    const int generic = mode == PG_BFS_GENERIC;
    ClWs w = cl_layout(ws, ws_bytes, N);
    if (!w.ok) { set_error("pg_bfs_cluster_count: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    const int2 *sl = (const int2 *)start_len;
    const unsigned nb = (unsigned)div_up(N, 256);
    const bool wide = nActive / N >= 12;          // lanes per neighbour list: 32 for long lists, 8 for short
    const unsigned eg = kNumSM * 8;
    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 12 * sizeof(unsigned long long), st));
    // sweep order = ascending segment start, on the top 16 bits of the position (two radix passes)
    int abits = 0;
    while ((1ll << abits) <= nActive) abits++;
    const int shift = abits > 16 ? abits - 16 : 0;
    unsigned long long h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool use_generic = generic != 0;
    const bool trusted = mode == PG_BFS_TRUSTED;
    const bool grid = trusted && bq_ws != nullptr;
    BqWs g{};
    if (grid) {
        g = bq_layout(bq_ws, bq_ws_bytes, N);
        if (!g.ok) { set_error("pg_bfs_cluster_count_grid: ball-query workspace too small for N = %d", N); return PG_EWORKSPACE; }
    }
    const bool lazy = lazy_masks != nullptr;
    const int4 *samples = nullptr;
    if (lazy) {
        PG_CHECK_ARG(grid && wide, "lazy lists need the producing ball query's workspace and long lists");
        PG_TRY(bq_list_samples(g, lazy_masks, N, w.samples, st));
        samples = w.samples;
    }
    launch(k_cl_prep, (unsigned)div_up((int64_t)N + 32, 256), 256, 0, st, semantic_label, ball_query_idxs, sl, N, nActive,
                                                                     generic ? 0 : 1, w.pl, w.trunc, w.last, w.root, w.lab,
                                                                     w.key0, shift, w.scalars, samples, Fill{(uint32_t *)w.size, (size_t)N + 1, 0u});
    if (!use_generic) {
        // the cell pass pays off on long lists only: it reads ~27 cells' worth of candidates per cell, which on
        // short lists (raw coordinates, ~5 neighbours per point) is more than the edges themselves
        const bool use_cells = grid && wide;
        const uint32_t *order;
        if (use_cells) {
            order = w.vA;               // the work list k_cl_cells leaves: query order minus the settled lists
        } else if (grid) {
            order = bq_sorted(g, N);    // the ball query's own query order IS the segment order
        } else {
            int res = 0;
            PG_TRY(radix_sort_pairs(w.key0, nullptr, w.kA, w.vA, w.kB, w.vB, N, abits - shift < 1 ? 1 : abits - shift, w.hist,
                                    w.scan_tmp, st, &res));
            order = res == 0 ? w.vA : w.vB;
        }
        launch(k_cl_flatten<false>, nb, 256, 0, st, w.pl, w.trunc, w.root, w.snap, N, nullptr, w.pubA, nullptr);
        static_assert(kSampleRounds == 1, "one ordered flatten per ticket / publish array");
        for (int round = 0; round < kSampleRounds; round++) {
            { PG_KTIME("k_cl_sample", st);
            launch(k_cl_sample<2>, nb, 256, 0, st, ball_query_idxs, sl, w.pl, w.last, w.snap, N, nActive, round, samples); }
            PG_KTIME("k_cl_flatten", st);
            launch(k_cl_flatten<true>, nb, 256, 0, st, w.pl, w.trunc, w.root, w.snap, N, w.pubA, w.pubB, w.scalars + 9);
        }
        if (use_cells) {
            PG_KTIME("k_cl_cells", st);      // the four passes of the cell pre-pass, timed as one
            const uint32_t *sp = bq_sorted(g, N);
            const unsigned cg = kNumSM * 8;
            launch(k_cl_cell_main, cg, 256, 0, st, sp, g.cstart, g.ccnt, g.scalars, w.pl, w.snap, w.cellsum, w.cthr);
            launch(k_cl_cell_thresholds, nb, 256, 0, st, g.cell, w.pl, w.snap, w.last, N, w.cellsum, w.cthr);
            launch(k_cl_cell_settle, cg, 256, 0, st, g.nbr, g.scalars, w.cellsum, w.cthr, w.cstate);
            launch(k_cl_cell_worklist, nb, 256, 0, st, sp, g.cell, w.pl, w.snap, w.cellsum, w.cstate, N, w.vA, w.scalars);
        }
        // lazy lists: only now, and only for the lists the sweep is about to read, do indices get written
        if (lazy) PG_TRY(bq_fill_lists(g, lazy_masks, sl, w.vA, w.scalars + 8, N, const_cast<int32_t *>(ball_query_idxs), st));
        const unsigned vg = kNumSM * 3;
#define PG_VERIFY(G, T)                                                                                              \
    launch(k_cl_verify<G, T>, vg, kVerThreads, 0, st, ball_query_idxs, sl, order, w.pl, w.last, w.snap, N, nActive, w.pend, \
                                                  (unsigned long long)w.pend_cap, w.scalars, use_cells ? 1 : 0)
        { PG_KTIME(trusted ? "k_cl_verify<trusted>" : "k_cl_verify<validating>", st);
        if (trusted) { if (wide) PG_VERIFY(32, true); else PG_VERIFY(8, true); }
        else { if (wide) PG_VERIFY(32, false); else PG_VERIFY(8, false); } }
#undef PG_VERIFY
        // trusted lists: the last flatten is folded into the label pass (the parked edges resolve their own roots)
        if (!trusted) launch(k_cl_flatten<true>, nb, 256, 0, st, w.pl, w.trunc, w.root, w.snap, N, w.pubB, nullptr, w.scalars + 11);
        PG_LAUNCH_CHECK();
    }
    // Trusted lists need no verdict from the sweep (no checksum, no range flags): the parked one-way edges are
    // settled on the device and the host reads everything back once, together with the sizes.
    const bool fast = trusted && !use_generic;
    if (fast) {
        launch(k_cl_pending_block<true>, 1, 1024, 0, st, w.pend, (unsigned long long)w.pend_cap, w.pl, w.root, w.lab, w.scalars);
    } else {
        PG_CUDA(cudaMemcpyAsync(h, w.scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
    }
    if (h[1] == 2) {   // the reference would read out of bounds here (bfs_cluster.cpp:40-42)
        set_error("pg_bfs_cluster_count: start_len has a row outside ball_query_idxs[0..%lld)", (long long)nActive);
        return PG_EINVAL;
    }
    if (!use_generic && (h[0] != 0 || h[1] != 0)) {   // not a truncated symmetric relation
        use_generic = true;
        launch(k_cl_reset, nb, 256, 0, st, w.root, w.lab, N);
    }
    const unsigned long long n_pend = use_generic ? 0 : h[2];
    g_cl_dbg[0] = h[0] != 0; g_cl_dbg[1] = (long long)h[1]; g_cl_dbg[2] = (long long)h[2]; g_cl_dbg[3] = 0;
    g_cl_dbg[4] = (grid && h[8]) ? (long long)h[8] : (long long)N;
    // min-label propagation to a fixed point, one launch + one readback per round (the parked edges, or -- on the
    // generic path / when the parking lot overflowed -- every edge)
    bool lists_needed = false;
    auto settle_host = [&](unsigned long long n_parked) -> int {
        const bool sweep = use_generic || n_parked > w.pend_cap;
        if (!sweep && n_parked == 0) return PG_OK;
        if (sweep && lazy) { lists_needed = true; return PG_OK; }     // every list would be read: not with lazy lists
        for (int it = 0; it < 1000000; it++) {
            PG_CUDA(cudaMemsetAsync(w.scalars + 3, 0, sizeof(unsigned long long), st));
            if (!sweep) {
                launch(k_cl_pending, (unsigned)div_up((int64_t)n_parked, 256), 256, 0, st, w.pend, n_parked, w.root, w.lab, w.scalars);
            } else if (use_generic) {
                if (wide) launch(k_cl_propagate<32, false>, eg, 256, 0, st, ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
                else launch(k_cl_propagate<8, false>, eg, 256, 0, st, ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
            } else {
                if (wide) launch(k_cl_propagate<32, true>, eg, 256, 0, st, ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
                else launch(k_cl_propagate<8, true>, eg, 256, 0, st, ball_query_idxs, sl, w.pl, w.trunc, w.last, N, w.root, w.lab, w.scalars);
            }
            PG_LAUNCH_CHECK();
            unsigned long long changed = 0;
            PG_CUDA(cudaMemcpyAsync(&changed, w.scalars + 3, sizeof(changed), cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            g_cl_dbg[3]++;
            if (!changed) break;
        }
        return PG_OK;
    };
    // final labels, sizes, kept clusters; everything the host needs comes back in one copy
    unsigned long long all[12];
    bool find_in_label = fast;
    auto finish = [&]() -> int {
        { PG_KTIME("k_cl_label", st);
        if (find_in_label) launch(k_cl_label<true>, nb, 256, 0, st, w.pl, w.root, w.lab, N, w.size, w.key0);
        else launch(k_cl_label<false>, nb, 256, 0, st, w.pl, w.root, w.lab, N, w.size, w.key0); }
        PG_TRY(scan_fused(KeepLoad{w.key0, w.size, N, threshold}, SizesStore{w.size, w.cid, w.csize, (unsigned long long *)w.scalars},
                          (int64_t)N + 1, (int64_t *)(w.scalars + 4), w.scan_tmp, st));
        PG_LAUNCH_CHECK();
        PG_CUDA(cudaMemcpyAsync(all, w.scalars, sizeof(all), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        return PG_OK;
    };
    PG_TRY(settle_host(n_pend));
    PG_TRY(finish());
    if (fast) {
        if (all[1] == 2) {
            set_error("pg_bfs_cluster_count: start_len has a row outside ball_query_idxs[0..%lld)", (long long)nActive);
            return PG_EINVAL;
        }
        g_cl_dbg[0] = 0; g_cl_dbg[1] = (long long)all[1]; g_cl_dbg[2] = (long long)all[2]; g_cl_dbg[3] = 0;
        g_cl_dbg[4] = (grid && all[8]) ? (long long)all[8] : (long long)N;
        if (all[10] != 0) {
            // more parked edges than one block settles quickly: the host-driven rounds after all, then the labels again
            // (root[] is complete: the label pass above wrote it)
            find_in_label = false;
            PG_TRY(settle_host(all[2]));
            if (lists_needed) { *need_lists = 1; return PG_OK; }
            PG_TRY(fill_u32(w.size, 0u, (size_t)N + 1, st));
            PG_CUDA(cudaMemsetAsync(w.scalars + 4, 0, 2 * sizeof(unsigned long long), st));
            PG_TRY(finish());
        }
    }
    host_sizes[0] = (int32_t)all[4];
    host_sizes[1] = (int32_t)all[5];
    host_sizes[2] = use_generic ? 1 : 0;
    return PG_OK;
}

extern "C" int pg_bfs_cluster_count(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                                    const int32_t *start_len, int32_t N, int64_t nActive, int32_t threshold,
                                    int mode, void *ws, size_t ws_bytes, int32_t *host_sizes, void *stream) {
    return bfs_count_impl(semantic_label, ball_query_idxs, start_len, N, nActive, threshold, mode, ws, ws_bytes, nullptr, 0,
                          host_sizes, stream);
}

extern "C" int pg_bfs_cluster_count_grid(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                                         const int32_t *start_len, int32_t N, int64_t nActive, int32_t threshold,
                                         void *ws, size_t ws_bytes, void *ballquery_ws, size_t ballquery_ws_bytes,
                                         int32_t *host_sizes, void *stream) {
    PG_CHECK_ARG(ballquery_ws != nullptr, "null ballquery_ws");
    return bfs_count_impl(semantic_label, ball_query_idxs, start_len, N, nActive, threshold, PG_BFS_TRUSTED, ws, ws_bytes,
                          ballquery_ws, ballquery_ws_bytes, host_sizes, stream);
}

extern "C" int pg_bfs_cluster_count_lazy(const int32_t *semantic_label, const int32_t *start_len, int32_t N, int64_t nActive,
                                         int32_t threshold, void *ws, size_t ws_bytes, void *ballquery_ws,
                                         size_t ballquery_ws_bytes, const uint32_t *masks, int32_t *idx_scratch,
                                         int32_t *host_sizes, int *host_need_lists, void *stream) {
    PG_CHECK_ARG(ballquery_ws != nullptr && masks != nullptr && host_need_lists != nullptr, "null pointer");
    PG_CHECK_ARG(idx_scratch != nullptr || nActive == 0, "null idx_scratch");
    *host_need_lists = 0;
    return bfs_count_impl(semantic_label, idx_scratch, start_len, N, nActive, threshold, PG_BFS_TRUSTED, ws, ws_bytes,
                          ballquery_ws, ballquery_ws_bytes, host_sizes, stream, masks, host_need_lists);
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
    // stable sort of (cluster id | dropped, point): members ascend inside every cluster; the key kernel's extra block writes
    // the offsets (exclusive scan of the cluster sizes + the total), the last radix pass the (cluster, point) rows
    launch(k_cl_keys, nb + 1, 256, 0, st, w.key0, w.cid, N, nCluster, w.kB, w.csize, cluster_offsets);
    int bits = 0;
    while ((1ll << bits) < (long long)nCluster + 1) bits++;
    int res = 0;
    PG_TRY(radix_sort_pairs(w.kB, nullptr, w.kA, w.vA, w.kB, w.vB, N, bits, w.hist, w.scan_tmp, st, &res, (int2 *)cluster_idxs, sumNPoint));
    PG_LAUNCH_CHECK();
    return PG_OK;
}
