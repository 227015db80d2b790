// Instance NMS of PointGroup.test.  Reference behaviour:
//   model/pointgroup.py:577-590  proposals_mask [nP, N] (dense, int) scattered from proposals_idx;
//                                intersection = mask_f @ mask_f.T; cross_ious = inter / (n_p + n_q - inter)
//   lib/utils/eval.py:75-97      get_nms_instances: greedy suppression in descending score order (on the CPU,
//                                after copying the [nP, nP] matrix to the host).
// The dense mask costs 4 * nP * N bytes (2.4 GB for 500 proposals on 1.2 M points) and its product is a
// [nP, N] x [N, nP] GEMM over a matrix that holds at most a few ones per column.  Here the (proposal, point)
// pairs themselves are the sparse matrix:
//   1. sort the pairs by (point, proposal)  -> every point's proposals, ascending, duplicates adjacent
//   2. sort them again by proposal (stable) -> every proposal's points, duplicates adjacent
//   3. one block per proposal a: walk its distinct points, and for each the distinct proposals b holding
//      that point; count into a shared-memory histogram (or, past 16 Ki proposals, into the output row);
//      then out[a][b] = fl(i / fl(fl(n_a + n_b) - i)) -- torch's three fp32 operations, so the result is
//      bit-identical (counts are exact in fp32 below 2^24, like the 0/1 matmul).
// No tensor cores: the contraction has ~2 nonzeros per column; the sparse count is O(sumNPoint).
#include "common.cuh"

namespace pg {

struct NmsWs {
    uint32_t *k0, *kA, *vA, *kB, *vB;    // sort buffers (k0: source keys of the pass being run)
    uint32_t *pt_sorted, *pp_sorted;     // (point, proposal) pairs ordered by (point, proposal)
    uint32_t *gp_sorted, *gpt_sorted;    // the same pairs ordered by (proposal, point)
    int32_t *row_start;                  // [N + 1]  first pair of every point in the (point, proposal) order
    int32_t *prop_start;                 // [nP + 1] first pair of every proposal in the (proposal, point) order
    int32_t *npoint;                     // [nP]     distinct points per proposal
    int32_t *hist;
    int64_t *scan_tmp;
    unsigned long long *scalars;         // [0] bad pair seen
    bool ok;
    size_t used;
};

static NmsWs nms_layout(void *ws, size_t ws_bytes, int64_t S_, int64_t N_, int64_t nP_) {
    Arena a(ws, ws_bytes);
    NmsWs w;
    const size_t S = (size_t)(S_ > 0 ? S_ : 1), N = (size_t)(N_ > 0 ? N_ : 1), nP = (size_t)(nP_ > 0 ? nP_ : 1);
    w.k0 = a.take<uint32_t>(S);
    w.kA = a.take<uint32_t>(S);
    w.vA = a.take<uint32_t>(S);
    w.kB = a.take<uint32_t>(S);
    w.vB = a.take<uint32_t>(S);
    w.pt_sorted = a.take<uint32_t>(S);
    w.pp_sorted = a.take<uint32_t>(S);
    w.gp_sorted = a.take<uint32_t>(S);
    w.gpt_sorted = a.take<uint32_t>(S);
    w.row_start = a.take<int32_t>(N + 1);
    w.prop_start = a.take<int32_t>(nP + 1);
    w.npoint = a.take<int32_t>(nP);
    w.hist = a.take<int32_t>(radix_tmp_count(S_));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(S + radix_tmp_count(S_))));
    w.scalars = a.take<unsigned long long>(2);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

static int bits_for(int64_t n) {           // bits needed for keys 0 .. n-1 (at least 1)
    int b = 1;
    while ((1ll << b) < n) b++;
    return b;
}

// column `col` of the [S, 2] pair list as sort keys; flags pairs outside [0, nP) x [0, N)
__global__ void k_nms_keys(const int2 *__restrict__ pairs, int32_t S, int32_t nP, int32_t N, uint32_t *__restrict__ keys,
                           unsigned long long *scalars) {
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int2 r = __ldg(pairs + s);
    if ((unsigned)r.x >= (unsigned)nP || (unsigned)r.y >= (unsigned)N) { scalars[0] = 1; keys[s] = 0; return; }
    keys[s] = (uint32_t)r.x;
}

// after the sort by proposal: key = point of the row, value = its proposal
__global__ void k_nms_second_keys(const int2 *__restrict__ pairs, const uint32_t *__restrict__ order, int32_t S, int32_t N,
                                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int2 r = __ldg(pairs + __ldg(order + s));
    keys[s] = (unsigned)r.y < (unsigned)N ? (uint32_t)r.y : 0u;
    vals[s] = (uint32_t)r.x;
}

// first[v] = lower bound of v in the ascending keys, v = 0 .. nV (first[nV] = S)
__global__ void k_nms_bounds(const uint32_t *__restrict__ keys, int32_t S, int32_t nV, int32_t *__restrict__ first) {
    pdl_enter();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    int lo = 0, hi = S;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < (uint32_t)v) lo = mid + 1; else hi = mid;
    }
    first[v] = lo;
}

// distinct points per proposal (the reference's proposals_mask.sum(1), model/pointgroup.py:582,589)
__global__ void k_nms_npoint(const uint32_t *__restrict__ gpt, const int32_t *__restrict__ prop_start, int32_t nP,
                             int32_t *__restrict__ npoint) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < nP; a += nWarps) {
        const int s0 = prop_start[a], s1 = prop_start[a + 1];
        int cnt = 0;
        for (int e = s0 + lane; e < s1; e += 32) cnt += (e == s0 || __ldg(gpt + e) != __ldg(gpt + e - 1)) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) npoint[a] = cnt;
    }
}

constexpr int kCrossThreads = 256;
constexpr int kCrossSmemBins = 16 * 1024;

// one block per proposal a (grid-stride): histogram of the proposals sharing a's points, then the IoU row
__global__ void __launch_bounds__(kCrossThreads) k_cross_iou(const uint32_t *__restrict__ gpt, const int32_t *__restrict__ prop_start,
                                                             const uint32_t *__restrict__ pp, const int32_t *__restrict__ row_start,
                                                             const int32_t *__restrict__ npoint, int32_t nP, int use_smem,
                                                             float *__restrict__ out) {
    pdl_enter();
    extern __shared__ int32_t bins_smem[];
    for (int a = blockIdx.x; a < nP; a += gridDim.x) {
        int32_t *bins = use_smem ? bins_smem : reinterpret_cast<int32_t *>(out + (int64_t)a * nP);
        for (int b = threadIdx.x; b < nP; b += kCrossThreads) bins[b] = 0;
        __syncthreads();
        const int s0 = prop_start[a], s1 = prop_start[a + 1];
        for (int e = s0 + threadIdx.x; e < s1; e += kCrossThreads) {
            const uint32_t pt = __ldg(gpt + e);
            if (e > s0 && __ldg(gpt + e - 1) == pt) continue;                 // the same (proposal, point) pair again
            const int r0 = __ldg(row_start + pt), r1 = __ldg(row_start + pt + 1);
            uint32_t prev = 0xffffffffu;
            for (int t = r0; t < r1; t++) {
                const uint32_t b = __ldg(pp + t);
                if (b != prev) atomicAdd(&bins[b], 1);
                prev = b;
            }
        }
        __syncthreads();
        const float na = (float)npoint[a];
        for (int b = threadIdx.x; b < nP; b += kCrossThreads) {
            const float inter = (float)bins[b];
            // model/pointgroup.py:590: intersection / (h + v - intersection), evaluated left to right in fp32
            out[(int64_t)a * nP + b] = __fdiv_rn(inter, __fsub_rn(__fadd_rn(na, (float)npoint[b]), inter));
        }
        __syncthreads();
    }
}

// ---- greedy suppression (lib/utils/eval.py:75-97) ------------------------------------------------------
// keys that sort ascending = scores descending; NaN scores last, like numpy's argsort of -scores
__global__ void k_nms_score_keys(const float *__restrict__ scores, int32_t n, uint32_t *__restrict__ keys) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = scores[i];
    keys[i] = (s != s) ? 0xffffffffu : f2ord(-s);
}

// One block.  `order` = proposal ids by descending score.  The next survivor is found by thread 0 walking
// the suppression bitmap; its row then suppresses, in parallel, every later proposal with IoU > threshold.
__global__ void __launch_bounds__(1024) k_nms_greedy(const float *__restrict__ cross, const uint32_t *__restrict__ order, int32_t n,
                                                     float threshold, int32_t *__restrict__ pick, int32_t *__restrict__ n_pick) {
    pdl_enter();
    extern __shared__ uint32_t dead[];             // bit k: position k of `order` is suppressed
    __shared__ int s_next, s_count;
    const int words = (n + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += blockDim.x) dead[w] = 0u;
    if (threadIdx.x == 0) { s_next = 0; s_count = 0; }
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) {
            int k = s_next;
            while (k < n) {
                const uint32_t alive = ~dead[k >> 5] & (0xffffffffu << (k & 31));
                if (alive) { k = (k & ~31) + __ffs((int)alive) - 1; break; }
                k = (k & ~31) + 32;
            }
            if (k >= n) k = n;
            s_next = k;
            if (k < n) pick[s_count++] = (int32_t)order[k];
        }
        __syncthreads();
        const int k = s_next;
        if (k >= n) break;
        const float *row = cross + (int64_t)order[k] * n;
        for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x)
            if (__ldg(row + order[j]) > threshold) atomicOr(&dead[j >> 5], 1u << (j & 31));
        __syncthreads();
        if (threadIdx.x == 0) s_next = k + 1;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_pick = s_count;
}

// clusters_mask = proposals_mask[pick_idxs] (model/pointgroup.py:593) without the dense [nProposal, N] mask: only
// the picked rows are materialised.  One block per picked row scatters the members of its proposal.
__global__ void __launch_bounds__(256) k_pick_masks(const int2 *__restrict__ pairs, const int32_t *__restrict__ offsets,
                                                    const int32_t *__restrict__ pick, int32_t nPick, int32_t nP, int32_t N,
                                                    int32_t *__restrict__ out, unsigned long long *bad) {
    pdl_enter();
    for (int k = blockIdx.x; k < nPick; k += gridDim.x) {
        const int p = __ldg(pick + k);
        if ((unsigned)p >= (unsigned)nP) { if (threadIdx.x == 0) *bad = 1; continue; }
        const int s0 = __ldg(offsets + p), s1 = __ldg(offsets + p + 1);
        int32_t *row = out + (int64_t)k * N;
        for (int e = s0 + threadIdx.x; e < s1; e += blockDim.x) {
            const int pt = __ldg(&pairs[e].y);
            if ((unsigned)pt < (unsigned)N) row[pt] = 1; else *bad = 1;
        }
    }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_pick_masks(const int32_t *proposals_idx, const int32_t *proposals_offset, int32_t nProposal,
                             const int32_t *pick, int32_t nPick, int32_t N, void *ws, int32_t *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(nProposal >= 0 && nPick >= 0 && N >= 0, "negative size");
    if (nPick == 0 || N == 0) return PG_OK;
    PG_CHECK_ARG(proposals_idx && proposals_offset && pick && out && ws, "null pointer");
    unsigned long long *d_bad = (unsigned long long *)ws;          // 8 bytes: out-of-range flag
    PG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st));
    PG_TRY(fill_u32(out, 0u, (size_t)nPick * (size_t)N, st));
    launch(k_pick_masks, (unsigned)(nPick < kNumSM * 8 ? nPick : kNumSM * 8), 256, 0, st, (const int2 *)proposals_idx, proposals_offset, pick,
                                                                                   nPick, nProposal, N, out, d_bad);
    PG_LAUNCH_CHECK();
    unsigned long long bad = 0;
    PG_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    if (bad) { set_error("pg_pick_masks: a picked proposal or one of its points is out of range"); return PG_EINVAL; }
    return PG_OK;
}

extern "C" size_t pg_cross_iou_workspace_bytes(int64_t nPairs, int64_t nProposal, int64_t N) {
    return nms_layout(nullptr, 0, nPairs, N, nProposal).used + 256;
}

extern "C" int pg_cross_iou(const int32_t *proposals_idx, int32_t nPairs, int32_t nProposal, int32_t N, void *ws,
                            size_t ws_bytes, float *cross_ious, int32_t *npoint, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(nPairs >= 0 && nProposal >= 0 && N >= 0, "negative size");
    if (nProposal == 0) return PG_OK;
    PG_CHECK_ARG(cross_ious && ws && (proposals_idx || nPairs == 0), "null pointer");
    NmsWs w = nms_layout(ws, ws_bytes, nPairs, N, nProposal);
    if (!w.ok) { set_error("pg_cross_iou: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    const int S = nPairs;
    const int2 *pairs = (const int2 *)proposals_idx;
    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 2 * sizeof(unsigned long long), st));
    const uint32_t *pt_sorted = w.pt_sorted, *pp_sorted = w.pp_sorted, *gp_sorted = w.gp_sorted, *gpt_sorted = w.gpt_sorted;
    if (S > 0) {
        const unsigned sb = (unsigned)div_up(S, 256);
        // (1) by proposal, (2) by point: LSD order, so the result is sorted by (point, proposal)
        launch(k_nms_keys, sb, 256, 0, st, pairs, S, nProposal, N, w.k0, w.scalars);
        int res = 0;
        PG_TRY(radix_sort_pairs(w.k0, nullptr, w.kA, w.vA, w.kB, w.vB, S, bits_for(nProposal), w.hist, w.scan_tmp, st, &res));
        launch(k_nms_second_keys, sb, 256, 0, st, pairs, res == 0 ? w.vA : w.vB, S, N, w.k0, w.gp_sorted);
        PG_TRY(radix_sort_pairs(w.k0, w.gp_sorted, w.kA, w.vA, w.kB, w.vB, S, bits_for(N), w.hist, w.scan_tmp, st, &res));
        PG_CUDA(cudaMemcpyAsync(w.pt_sorted, res == 0 ? w.kA : w.kB, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
        PG_CUDA(cudaMemcpyAsync(w.pp_sorted, res == 0 ? w.vA : w.vB, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
        // (3) stable by proposal again: sorted by (proposal, point)
        PG_TRY(radix_sort_pairs(w.pp_sorted, w.pt_sorted, w.kA, w.vA, w.kB, w.vB, S, bits_for(nProposal), w.hist, w.scan_tmp, st,
                                &res));
        gp_sorted = res == 0 ? w.kA : w.kB;
        gpt_sorted = res == 0 ? w.vA : w.vB;
    }
    launch(k_nms_bounds, (unsigned)div_up((int64_t)N + 1, 256), 256, 0, st, pt_sorted, S, N, w.row_start);
    launch(k_nms_bounds, (unsigned)div_up((int64_t)nProposal + 1, 256), 256, 0, st, gp_sorted, S, nProposal, w.prop_start);
    launch(k_nms_npoint, (unsigned)(div_up(nProposal, 8) < kNumSM * 8 ? div_up(nProposal, 8) : kNumSM * 8), 256, 0, st, gpt_sorted, w.prop_start, nProposal, w.npoint);
    const int use_smem = nProposal <= kCrossSmemBins;
    const size_t smem = use_smem ? (size_t)nProposal * sizeof(int32_t) : 0;
    if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(k_cross_iou, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PG_KTIME("k_cross_iou", st);
    launch(k_cross_iou, (unsigned)(nProposal < kNumSM * 8 ? nProposal : kNumSM * 8), kCrossThreads, smem, st, gpt_sorted, w.prop_start, pp_sorted, w.row_start, w.npoint, nProposal, use_smem, cross_ious); }
    if (npoint) PG_CUDA(cudaMemcpyAsync(npoint, w.npoint, (size_t)nProposal * 4, cudaMemcpyDeviceToDevice, st));
    PG_LAUNCH_CHECK();
    unsigned long long bad = 0;
    PG_CUDA(cudaMemcpyAsync(&bad, w.scalars, sizeof(bad), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    if (bad) { set_error("pg_cross_iou: a (proposal, point) pair lies outside [0, %d) x [0, %d)", nProposal, N); return PG_EINVAL; }
    return PG_OK;
}

extern "C" size_t pg_nms_instances_workspace_bytes(int32_t n) {
    const size_t m = (size_t)(n > 0 ? n : 1);
    return 5 * align_up(m * 4) + align_up(radix_tmp_count(n) * 4) + align_up(scan_tmp_count((int64_t)(m + radix_tmp_count(n))) * 8) + 512;
}

extern "C" int pg_nms_instances(const float *cross_ious, const float *scores, int32_t n, float threshold, void *ws,
                                size_t ws_bytes, int32_t *pick, int32_t *host_n_pick, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(host_n_pick, "null host_n_pick");
    *host_n_pick = 0;
    PG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(cross_ious && scores && ws && pick, "null pointer");
    PG_CHECK_ARG(n <= (1 << 20), "more than 2^20 proposals");
    Arena a(ws, ws_bytes);
    uint32_t *k0 = a.take<uint32_t>(n), *kA = a.take<uint32_t>(n), *vA = a.take<uint32_t>(n), *kB = a.take<uint32_t>(n),
             *vB = a.take<uint32_t>(n);
    int32_t *hist = a.take<int32_t>(radix_tmp_count(n));
    int64_t *scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)n + radix_tmp_count(n)));
    int32_t *d_count = a.take<int32_t>(1);
    if (!a.ok) { set_error("pg_nms_instances: workspace too small (%zu < %zu)", ws_bytes, a.used); return PG_EWORKSPACE; }
    launch(k_nms_score_keys, (unsigned)div_up(n, 256), 256, 0, st, scores, n, k0);
    int res = 0;
    PG_TRY(radix_sort_pairs(k0, nullptr, kA, vA, kB, vB, n, 32, hist, scan_tmp, st, &res));
    const size_t smem = (size_t)((n + 31) / 32) * 4;
    if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(k_nms_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PG_KTIME("k_nms_greedy", st);
    launch(k_nms_greedy, 1, 1024, smem, st, cross_ious, res == 0 ? vA : vB, n, threshold, pick, d_count); }
    PG_LAUNCH_CHECK();
    int32_t cnt = 0;
    PG_CUDA(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *host_n_pick = cnt;
    return PG_OK;
}
