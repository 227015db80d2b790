// Caller-side glue of the proposal path, fused: SURVEY.md section 8(f) row 1.
//
// clusters_voxelization (model/pointgroup.py:125-178) turns every proposal's points into integer
// coordinates of a fullscale^3 grid: gather the points, sec_mean, recentre, sec_min / sec_max, a
// per-proposal scale and offset, `.long()`.  The reference does it with ~40 elementwise torch
// kernels, three segment reductions and four gathers over [sumNPoint, 3] tensors.  Here it is two
// kernels that produce bit-identical results: every fp32 operation below is the operation the torch
// kernels perform, in their order, with explicit round-to-nearest intrinsics (no fma contraction).
//   k_glue_cluster_stats   one block per proposal: the mean with sec_mean's exact semantics (a serial
//                          add chain over fl(x / count), sec_mean.cu:17-25) while the other threads
//                          keep per-channel min / max; then the proposal's centre, size, scale, offset
//   k_glue_cluster_coords  one thread per proposal point: fl(fl(fl(x - mean) * scale) + offset) -> int64
#include "common.cuh"

namespace pg {

constexpr int kGsThreads = 128;
constexpr int kGsRows = 512;                       // points per shared tile

struct GlueParams {          // per proposal, written by the stats kernel, read by the coords kernel
    float mean[3];
    float scale;
    float offset[3];
    float pad;
};

__global__ void __launch_bounds__(kGsThreads) k_glue_cluster_stats(const float *__restrict__ coords,
                                                                   const int2 *__restrict__ cluster_idxs,
                                                                   const int32_t *__restrict__ offsets, int32_t nC,
                                                                   float inv_fullscale, float fullscale, float max_scale,
                                                                   const float *__restrict__ rand6,
                                                                   GlueParams *__restrict__ params,
                                                                   float *__restrict__ center, float *__restrict__ size) {
    pdl_enter();
    __shared__ float buf[2][kGsRows * 3];
    __shared__ float red[2][3][kGsThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int p = blockIdx.x; p < nC; p += gridDim.x) {
        const int start = __ldg(offsets + p), end = __ldg(offsets + p + 1);
        const int len = end - start;
        const float count = (float)len;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        float acc = 0.f;                            // threads 0..2: the add chain of channel tid
        float reg[4][3];
        // prefetch tile 0: thread t owns rows t, t + 128, ...
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int r = k * kGsThreads + tid;
            reg[k][0] = reg[k][1] = reg[k][2] = 0.f;
            if (r < len) {
                const float *q = coords + 3 * (int64_t)__ldg(&cluster_idxs[start + r].y);
                reg[k][0] = __ldg(q); reg[k][1] = __ldg(q + 1); reg[k][2] = __ldg(q + 2);
            }
        }
        int cur = 0;
        for (int pos = 0; pos < len; pos += kGsRows) {
            const int rows = min(kGsRows, len - pos);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = k * kGsThreads + tid;
                if (r < rows) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float x = reg[k][c];
                        buf[cur][r * 3 + c] = __fdiv_rn(x, count);
                        if (x < mn[c]) mn[c] = x;            // strict compares: NaN is never selected (sec_mean.cu:44-50,70-76)
                        if (x > mx[c]) mx[c] = x;
                    }
                }
            }
            __syncthreads();
            const int npos = pos + kGsRows;
            if (npos < len) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = npos + k * kGsThreads + tid;
                    if (r < len) {
                        const float *q = coords + 3 * (int64_t)__ldg(&cluster_idxs[start + r].y);
                        reg[k][0] = __ldg(q); reg[k][1] = __ldg(q + 1); reg[k][2] = __ldg(q + 2);
                    }
                }
            }
            if (tid < 3) {
                const float *b = &buf[cur][tid];
                float a = acc;
                int r = 0;
                for (; r + 8 <= rows; r += 8) {
                    const float q0 = b[(r + 0) * 3], q1 = b[(r + 1) * 3], q2 = b[(r + 2) * 3], q3 = b[(r + 3) * 3];
                    const float q4 = b[(r + 4) * 3], q5 = b[(r + 5) * 3], q6 = b[(r + 6) * 3], q7 = b[(r + 7) * 3];
                    a = __fadd_rn(a, q0); a = __fadd_rn(a, q1); a = __fadd_rn(a, q2); a = __fadd_rn(a, q3);
                    a = __fadd_rn(a, q4); a = __fadd_rn(a, q5); a = __fadd_rn(a, q6); a = __fadd_rn(a, q7);
                }
                for (; r < rows; r++) a = __fadd_rn(a, b[r * 3]);
                acc = a;
            }
            cur ^= 1;
        }
        // block min / max per channel
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float a = mn[c], b = mx[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            if (lane == 0) { red[0][c][warp] = a; red[1][c][warp] = b; }
        }
        __shared__ float s_mean[3];
        if (tid < 3) s_mean[tid] = acc;
        __syncthreads();
        if (tid < 3) {
            const int c = tid;
            float lo = red[0][c][0], hi = red[1][c][0];
            for (int w = 1; w < kGsThreads / 32; w++) { lo = fminf(lo, red[0][c][w]); hi = fmaxf(hi, red[1][c][w]); }
            const float mean = s_mean[c];
            // recentring is monotone, so min(fl(x - mean)) = fl(min(x) - mean)  (pointgroup.py:139-142)
            const float cmin = __fsub_rn(lo, mean), cmax = __fsub_rn(hi, mean);
            const float extent = __fsub_rn(cmax, cmin);
            size[p * 3 + c] = extent;                                                    // :144
            center[p * 3 + c] = __fadd_rn(__fmul_rn(__fadd_rn(cmax, cmin), 0.5f), mean);  // :145
            // scale = min(1 / max_c(extent / fullscale) - 0.01, max_scale)                 :147-148
            // (torch divides by a Python scalar as a multiplication by its fp32 reciprocal)
            float t = __fmul_rn(extent, inv_fullscale);
            const float t1 = __shfl_sync(0x7u, t, 1), t2 = __shfl_sync(0x7u, t, 2), t0 = __shfl_sync(0x7u, t, 0);
            float tm = t0;
            if (t1 > tm || t1 != t1) tm = t1;       // torch.max propagates NaN
            if (t2 > tm || t2 != t2) tm = t2;
            float sc = __fsub_rn(__fdiv_rn(1.f, tm), 0.01f);
            sc = (sc != sc) ? sc : fminf(sc, max_scale);                                  // clamp(max=scale) keeps NaN
            const float min_xyz = __fmul_rn(cmin, sc), max_xyz = __fmul_rn(cmax, sc);      // :149-150
            const float rng = __fsub_rn(max_xyz, min_xyz);                                // :153
            float a = __fsub_rn(__fsub_rn(fullscale, rng), 0.001f);                       // clamp(fullscale - range - 0.001, min=0)
            a = (a != a) ? a : fmaxf(a, 0.f);
            float b = __fadd_rn(__fsub_rn(fullscale, rng), 0.001f);                       // clamp(fullscale - range + 0.001, max=0)
            b = (b != b) ? b : fminf(b, 0.f);
            const float off = __fadd_rn(__fadd_rn(-min_xyz, __fmul_rn(a, __ldg(rand6 + c))), __fmul_rn(b, __ldg(rand6 + 3 + c)));   // :154-155
            params[p].mean[c] = mean;
            params[p].offset[c] = off;
            if (c == 0) params[p].scale = sc;
        }
        __syncthreads();
    }
}

__global__ void k_glue_cluster_coords(const float *__restrict__ coords, const int2 *__restrict__ cluster_idxs,
                                      const GlueParams *__restrict__ params, int32_t S, int64_t *__restrict__ out) {
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int2 ci = __ldg(cluster_idxs + s);
    const float4 m = __ldg(reinterpret_cast<const float4 *>(params + ci.x));          // mean.xyz, scale
    const float4 o = __ldg(reinterpret_cast<const float4 *>(params + ci.x) + 1);      // offset.xyz
    const float *q = coords + 3 * (int64_t)ci.y;
    const float x = __fadd_rn(__fmul_rn(__fsub_rn(__ldg(q), m.x), m.w), o.x);
    const float y = __fadd_rn(__fmul_rn(__fsub_rn(__ldg(q + 1), m.y), m.w), o.y);
    const float z = __fadd_rn(__fmul_rn(__fsub_rn(__ldg(q + 2), m.z), m.w), o.z);
    longlong4 *dst = reinterpret_cast<longlong4 *>(out) + s;
    // .long(): truncation toward zero, the conversion torch's cast kernel compiles to
    *dst = make_longlong4((long long)ci.x, (long long)x, (long long)y, (long long)z);
}

// ---- pack_proposals: the padded per-scene proposal block that is all-gathered (d3net_b200/dist.py) ----
// Row layout (46 floats): score feats 16 | 8 box corners x 3 | centre 3 | semantic class | score | mask.
// Scene s keeps its first P proposals in proposal order (convert_stack_to_batch without the randperm,
// model/pointgroup.py:237-257).  One block ranks the proposals (a proposal's slot = number of earlier
// proposals of its scene), then the rows are written one thread per float.
constexpr int kPackWidth = 46;

__global__ void __launch_bounds__(1024) k_pack_rank(const int2 *__restrict__ proposals_idx, const int32_t *__restrict__ offsets,
                                                    const int64_t *__restrict__ locs_scaled, int32_t nP,
                                                    int32_t *__restrict__ scene_of, int32_t *__restrict__ slot_of) {
    pdl_enter();
    extern __shared__ int32_t sc[];                 // scene id per proposal
    for (int p = threadIdx.x; p < nP; p += blockDim.x) {
        const int first = __ldg(&proposals_idx[__ldg(offsets + p)].y);
        sc[p] = (int32_t)__ldg(locs_scaled + 4 * (int64_t)first);      // proposals_batchId (:349)
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nP; p += blockDim.x) {
        const int mine = sc[p];
        int r = 0;
        for (int q = 0; q < p; q++) r += sc[q] == mine;
        scene_of[p] = mine;
        slot_of[p] = r;
    }
}

__global__ void k_pack_rows(const int2 *__restrict__ proposals_idx, const int32_t *__restrict__ offsets,
                            const int64_t *__restrict__ semantic_preds, const float *__restrict__ center,
                            const float *__restrict__ size, const float *__restrict__ feats, const float *__restrict__ score,
                            const int32_t *__restrict__ scene_of, const int32_t *__restrict__ slot_of, int32_t nP, int32_t C,
                            int32_t B, int32_t P, float *__restrict__ out) {
    pdl_enter();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nP * kPackWidth) return;
    const int p = (int)(t / kPackWidth), j = (int)(t - (int64_t)p * kPackWidth);
    const int b = scene_of[p], s = slot_of[p];
    if (s >= P || b < 0 || b >= B) return;
    float v;
    if (j < 16) v = j < C ? feats[(int64_t)p * C + j] : 0.f;
    else if (j < 40) {
        const int k = (j - 16) / 3, a = (j - 16) - 3 * k;              // corner k (sign bits x, y, z = 4, 2, 1), axis a
        const float sign = ((k >> (2 - a)) & 1) ? 1.f : -1.f;
        v = __fadd_rn(center[p * 3 + a], __fmul_rn(__fmul_rn(0.5f, size[p * 3 + a]), sign));
    } else if (j < 43) v = center[p * 3 + (j - 40)];
    else if (j == 43) v = (float)semantic_preds[__ldg(&proposals_idx[__ldg(offsets + p)].y)];     // sem_cls (:359)
    else if (j == 44) v = score[p];
    else v = 1.f;
    out[((int64_t)b * P + s) * kPackWidth + j] = v;
}

// ---- collate_points: the per-point part of sparse_collate_fn (lib/dataset/pipeline.py:937-985) ----------
// One thread per point of the concatenated batch: its scene b (binary search in batch_offsets) becomes column 0
// of locs_scaled (:939-943, the float coordinates truncated toward zero like .long()), sem_labels widen to
// int64 (:982) and instance ids other than -1 move up by the instances of the scenes before (:963-964,983).
__global__ void k_collate_points(const float *__restrict__ locs_scaled, const int32_t *__restrict__ sem_labels,
                                 const int32_t *__restrict__ instance_ids, const int32_t *__restrict__ batch_offsets,
                                 const int32_t *__restrict__ instance_offsets, int32_t N, int32_t B,
                                 int64_t *__restrict__ out_locs, int64_t *__restrict__ out_sem, int64_t *__restrict__ out_inst) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int lo = 0, hi = B;                       // last b with batch_offsets[b] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(batch_offsets + mid) <= i) lo = mid; else hi = mid;
    }
    const float *q = locs_scaled + 3 * (int64_t)i;
    reinterpret_cast<longlong4 *>(out_locs)[i] =
        make_longlong4((long long)lo, (long long)__ldg(q), (long long)__ldg(q + 1), (long long)__ldg(q + 2));
    if (out_sem) out_sem[i] = (int64_t)__ldg(sem_labels + i);
    if (out_inst) {
        const int id = __ldg(instance_ids + i);
        out_inst[i] = id == -1 ? -1 : (int64_t)id + (int64_t)__ldg(instance_offsets + lo);
    }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_pack_proposals(const int32_t *proposals_idx, const int32_t *proposals_offset, const int64_t *locs_scaled,
                                 const int64_t *semantic_preds, const float *center, const float *size, const float *feats,
                                 const float *score, int32_t nProposal, int32_t C, int32_t B, int32_t P, int32_t *ws,
                                 float *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(nProposal >= 0 && B >= 0 && P >= 0 && C >= 0, "negative size");
    PG_CHECK_ARG(nProposal <= 49152, "more proposals than one ranking block holds (49152)");
    if ((int64_t)B * P > 0) {
        PG_CHECK_ARG(out, "null pointer");
        PG_CUDA(cudaMemsetAsync(out, 0, (size_t)B * P * kPackWidth * sizeof(float), st));
    }
    if (nProposal == 0 || (int64_t)B * P == 0) return PG_OK;
    PG_CHECK_ARG(proposals_idx && proposals_offset && locs_scaled && semantic_preds && center && size && feats && score && ws,
                 "null pointer");
    const size_t smem = (size_t)nProposal * sizeof(int32_t);
    if (smem > 48 * 1024)
        PG_CUDA(cudaFuncSetAttribute(k_pack_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch(k_pack_rank, 1, 1024, smem, st, (const int2 *)proposals_idx, proposals_offset, locs_scaled, nProposal, ws, ws + nProposal);
    const int64_t total = (int64_t)nProposal * kPackWidth;
    launch(k_pack_rows, (unsigned)div_up(total, 256), 256, 0, st, (const int2 *)proposals_idx, proposals_offset, semantic_preds, center,
                                                              size, feats, score, ws, ws + nProposal, nProposal, C, B, P, out);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" size_t pg_cluster_coords_workspace_bytes(int32_t nCluster) {
    return (size_t)(nCluster > 0 ? nCluster : 1) * sizeof(GlueParams) + 256;
}

extern "C" int pg_cluster_coords(const float *coords, const int32_t *cluster_idxs, const int32_t *cluster_offsets,
                                 int32_t sumNPoint, int32_t nCluster, int32_t fullscale, float scale, const float *rand6,
                                 void *ws, size_t ws_bytes, int64_t *out_coords, float *center, float *size, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(sumNPoint >= 0 && nCluster >= 0, "negative size");
    if (nCluster == 0 || sumNPoint == 0) return PG_OK;
    PG_CHECK_ARG(coords && cluster_idxs && cluster_offsets && rand6 && ws && out_coords && center && size, "null pointer");
    PG_CHECK_ARG(ws_bytes >= (size_t)nCluster * sizeof(GlueParams), "workspace too small");
    PG_CHECK_ARG(((uintptr_t)ws & 15u) == 0 && ((uintptr_t)out_coords & 31u) == 0, "workspace / output not aligned");
    GlueParams *params = (GlueParams *)ws;
    const float fs = (float)fullscale;
    const float inv_fs = 1.0f / fs;
    const unsigned grid = (unsigned)(nCluster < kNumSM * 8 ? nCluster : kNumSM * 8);
    { PG_KTIME("k_glue_cluster_stats", st);
    launch(k_glue_cluster_stats, grid, kGsThreads, 0, st, coords, (const int2 *)cluster_idxs, cluster_offsets, nCluster, inv_fs, fs, scale,
                                                     rand6, params, center, size); }
    launch(k_glue_cluster_coords, (unsigned)div_up(sumNPoint, 256), 256, 0, st, coords, (const int2 *)cluster_idxs, params, sumNPoint,
                                                                            out_coords);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_collate_points(const float *locs_scaled, const int32_t *sem_labels, const int32_t *instance_ids,
                                 const int32_t *batch_offsets, const int32_t *instance_offsets, int32_t N, int32_t B,
                                 int64_t *out_locs_scaled, int64_t *out_sem_labels, int64_t *out_instance_ids, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(N >= 0 && B >= 0, "negative size");
    if (N == 0) return PG_OK;
    PG_CHECK_ARG(B >= 1, "points without a scene");
    PG_CHECK_ARG(locs_scaled && batch_offsets && out_locs_scaled, "null pointer");
    PG_CHECK_ARG((out_sem_labels == nullptr) == (sem_labels == nullptr), "sem_labels in / out must come together");
    PG_CHECK_ARG((out_instance_ids == nullptr) == (instance_ids == nullptr) && (instance_ids == nullptr || instance_offsets),
                 "instance_ids in / out / offsets must come together");
    PG_CHECK_ARG(((uintptr_t)out_locs_scaled & 31u) == 0, "out_locs_scaled not 32-byte aligned");
    launch(k_collate_points, (unsigned)div_up(N, 256), 256, 0, st, locs_scaled, sem_labels, instance_ids, batch_offsets,
                                                               instance_offsets, N, B, out_locs_scaled, out_sem_labels,
                                                               out_instance_ids);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
