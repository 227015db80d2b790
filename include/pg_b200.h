/*
 * pg_b200.h -- C ABI of libpg_b200.so: the PointGroup proposal ops of D3Net (lib/pointgroup_ops)
 * as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary.  Every entry point replaces one function of the reference's native
 * module PG_OP (lib/pointgroup_ops/src/pointgroup_ops_api.cpp:6-24); the citation on each
 * declaration names the reference interface it stands in for.  Signatures carry plain pointers and
 * sizes only -- no torch types.  All data pointers are DEVICE pointers on the current CUDA device
 * unless the name starts with `host_`.  `stream` is a cudaStream_t passed as void*.
 *
 * Conventions
 *   - return value: 0 (PG_OK) on success, otherwise a negative PG_E* code or a positive cudaError_t;
 *     pg_last_error() returns a thread-local message for the last failure.  Nothing calls exit()
 *     (the reference's ball query does, bfs_cluster.cu:82-86).
 *   - fixed-size outputs are written in place.  Ops that ACCUMULATE (voxelize_bp, roipool_bp) add
 *     onto what the buffer holds, exactly like the reference's atomicAdd kernels, so the caller
 *     zero-fills them (functions/pointgroup_ops.py:57,70,93,106,215).  All other outputs are fully
 *     overwritten and need no zero fill.
 *   - variable-size outputs (voxelize_idx, ballquery, bfs_cluster) are two-phase: `*_count`/`*_map`
 *     computes the sizes, synchronises `stream` once and stores them through a host pointer; the
 *     caller allocates exactly and calls `*_fill`.  The workspace passed to both phases must be the
 *     same untouched buffer.  This replaces the reference's resize_() inside native code
 *     (voxelize.cpp:22-26, bfs_cluster.cpp:103-106) and its ball-query retry loop
 *     (functions/pointgroup_ops.py:135-142).
 *   - the library keeps no state between calls, owns no device memory and never allocates:
 *     scratch space is a caller-provided workspace sized by the matching `*_workspace_bytes`.
 *   - kernels are launched on `stream`; only the `*_count`/`*_map` phases synchronise it.
 */
#ifndef PG_B200_H_
#define PG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_EINVAL (-1)     /* bad argument (null pointer, negative size, unsupported mode) */
#define PG_EWORKSPACE (-2) /* workspace too small */
#define PG_EOVERFLOW (-3)  /* a count does not fit the int32 the reference's tensors use */

#define PG_BALLQUERY_CAP 1000 /* bfs_cluster.cu:20,38 -- at most the first 1000 neighbours by index */

/* pg_bfs_cluster_count `mode` */
#define PG_BFS_AUTO 0    /* validate the lists in the sweep; fall back to the any-digraph path if they fail */
#define PG_BFS_GENERIC 1 /* force the any-digraph propagation path */
#define PG_BFS_TRUSTED 2 /* the caller vouches that (idx, start_len) are pg_ballquery_* output, unmodified:
                            in range and a truncated symmetric relation -- the sweep skips its validation */

const char *pg_last_error(void);
int pg_abi_version(void);

/* Measurement aid (no reference counterpart): with timing enabled the library brackets its main kernels
 * with CUDA events on the launching stream.  pg_kernel_timing(1) clears and starts, (0) clears and stops;
 * the report waits for the recorded events and writes one "name<TAB>launches<TAB>total_ms" line per
 * kernel into buf (NUL-terminated, truncated to cap) and returns the untruncated length. */
void pg_kernel_timing(int enable);
size_t pg_kernel_timing_report(char *buf, size_t cap);

/* ------------------------------------------------------------------------------------------------
 * voxelize_idx     replaces PG_OP.voxelize_idx (src/pointgroup_ops.cpp:13, voxelize.cpp:11-152)
 * coords: int64 [N,4] (batch, x, y, z); every column is narrowed to int32 like the reference's
 * Point<3>/Int (datatype.h:9-11).  Voxel ids follow first occurrence in input order.
 * phase 1 writes input_map[N] and host_sizes = {M, maxActive}; phase 2 writes
 * output_coords int64 [M,4] and output_map int32 [M, maxActive+1] = [cnt, p0 < p1 < ..., 0-pad].
 * mode: 0/1 keep the first point, 2 the last, 3 sum, 4 mean (maxActive = 1 unless mode is 3 or 4).
 * ---------------------------------------------------------------------------------------------- */
size_t pg_voxelize_idx_workspace_bytes(int64_t N);
int pg_voxelize_idx_map(const int64_t *coords, int64_t N, int mode, int32_t *input_map, void *ws,
                        size_t ws_bytes, int32_t *host_sizes, void *stream);
int pg_voxelize_idx_fill(const int64_t *coords, const int32_t *input_map, int64_t N, int32_t M,
                         int32_t maxActive, int mode, void *ws, size_t ws_bytes, int64_t *output_coords,
                         int32_t *output_map, void *stream);

/* ------------------------------------------------------------------------------------------------
 * voxelize_fp / voxelize_bp     replace PG_OP.voxelize_fp / voxelize_bp
 *                               (src/pointgroup_ops.cpp:18,25; voxelize.cu:10-53)
 * point_recover_fp / _bp        replace PG_OP.point_recover_fp / _bp (pointgroup_ops.cpp:30,35;
 *                               voxelize.cpp:182-202): the same two kernels with average = 0.
 * fp: out[v][c] = sum_i fl(mult * feats[r_i][c]) left to right from 0, mult = fl(1/cnt) if average.
 * bp: d_feats[r_i][c] += fl(mult * d_out[v][c]).
 * ---------------------------------------------------------------------------------------------- */
int pg_voxelize_fp(const float *feats, float *out, const int32_t *rules, int32_t M, int32_t maxActive,
                   int32_t C, int average, void *stream);
int pg_voxelize_bp(const float *d_out, float *d_feats, const int32_t *rules, int32_t M,
                   int32_t maxActive, int32_t C, int average, void *stream);
int pg_point_recover_fp(const float *feats, float *out, const int32_t *rules, int32_t M,
                        int32_t maxActive, int32_t C, void *stream);
int pg_point_recover_bp(const float *d_out, float *d_feats, const int32_t *rules, int32_t M,
                        int32_t maxActive, int32_t C, void *stream);

/* ------------------------------------------------------------------------------------------------
 * ballquery_batch_p     replaces PG_OP.ballquery_batch_p (bfs_cluster.h:15, bfs_cluster.cu:15-90)
 * Per point i: every k of the same scene with fma(dz,dz,fma(dx,dx,dy*dy)) < fl(r*r), ascending k,
 * at most the first PG_BALLQUERY_CAP.  Three phases over one workspace:
 *   prepare  builds the uniform grid (cells, sorted point lists, neighbour cells); asynchronous;
 *   count    writes start_len int32 [n,2] = (start, cnt), segments laid out in query (cell) order
 *            (deterministic; the reference's atomicAdd placement is not), host_total = sum cnt.  `masks`
 *            (device, capacity mask_words uint32; may be NULL) is an optional buffer in which the pass
 *            records every predicate outcome; it is used when the batch needs no more words than it
 *            holds (about 25 per point at 330 neighbours per point) -- host_masks_used says whether;
 *   fill     writes idx int32 [total].  With masks it only turns recorded bits into indices; pass NULL
 *            when host_masks_used was 0 and the predicates are evaluated again.
 * ---------------------------------------------------------------------------------------------- */
size_t pg_ballquery_workspace_bytes(int64_t n);
int pg_ballquery_prepare(const float *xyz, const int32_t *batch_idxs, const int32_t *batch_offsets, int32_t n,
                         int32_t B, float radius, void *ws, size_t ws_bytes, void *stream);
int pg_ballquery_count(const float *xyz, int32_t n, float radius, int32_t *start_len, uint32_t *masks,
                       int64_t mask_words, void *ws, size_t ws_bytes, int64_t *host_total, int *host_masks_used,
                       void *stream);
int pg_ballquery_fill(const float *xyz, int32_t n, float radius, const int32_t *start_len, const uint32_t *masks,
                      int32_t *idx, int64_t idx_capacity, void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * bfs_cluster     replaces PG_OP.bfs_cluster (bfs_cluster.h:18, bfs_cluster.cpp:28-112)
 * Connected components of the neighbour graph restricted to equal semantic labels; components with
 * >= threshold points are kept and numbered by ascending smallest member (= the reference's seed
 * order).  Members are emitted in ascending point order (the reference emits BFS order; membership
 * and cluster order are identical).  nActive = length of ball_query_idxs.  phase 1 stores
 * host_sizes = {nCluster, sumNPoint, 1 if the generic path ran}; phase 2
 * writes cluster_idxs int32 [sumNPoint,2] = (cluster_id, point) and cluster_offsets int32 [nCluster+1].
 * `mode`: PG_BFS_AUTO checks, in the same sweep, that the lists are a truncated symmetric relation and
 * switches to the any-digraph propagation path when they are not; PG_BFS_GENERIC forces
 * that path; PG_BFS_TRUSTED skips the check (only for lists that came from pg_ballquery_* untouched).
 * ---------------------------------------------------------------------------------------------- */
size_t pg_bfs_cluster_workspace_bytes(int64_t N);
int pg_bfs_cluster_count(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                         const int32_t *start_len, int32_t N, int64_t nActive, int32_t threshold, int mode,
                         void *ws, size_t ws_bytes, int32_t *host_sizes, void *stream);
/* Trusted lists together with the workspace of the pg_ballquery_prepare/count/fill calls that produced them
 * (same n = N, not written since): the ball query's uniform grid is still in there, and a cell whose whole
 * 27-cell neighbourhood already sits in one component after the sampling rounds has nothing left to tell the
 * edge sweep -- its lists are not read at all.  Results are identical to
 * pg_bfs_cluster_count; phase 2 is the same pg_bfs_cluster_fill. */
int pg_bfs_cluster_count_grid(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                              const int32_t *start_len, int32_t N, int64_t nActive, int32_t threshold, void *ws,
                              size_t ws_bytes, void *ballquery_ws, size_t ballquery_ws_bytes, int32_t *host_sizes,
                              void *stream);
/* Lazy lists -- the fused form of ballquery_batch_p + bfs_cluster for callers that need the clusters but not the lists
 * (model/pointgroup.py:296-297 hands idx to bfs_cluster and never looks at it again): after pg_ballquery_prepare/count
 * WITH hit masks, the lists are clustered straight from (masks, merged candidates); only the lists the edge sweep has
 * to read are decoded, into idx_scratch (device, nActive ints, otherwise untouched).  Same clusters as
 * pg_bfs_cluster_count on the materialised lists.  *host_need_lists = 1 (sizes not valid): so many one-way edges were
 * parked that the fall-back needs every list -- call pg_ballquery_fill and pg_bfs_cluster_count_grid instead. */
int pg_bfs_cluster_count_lazy(const int32_t *semantic_label, const int32_t *start_len, int32_t N, int64_t nActive,
                              int32_t threshold, void *ws, size_t ws_bytes, void *ballquery_ws, size_t ballquery_ws_bytes,
                              const uint32_t *masks, int32_t *idx_scratch, int32_t *host_sizes, int *host_need_lists,
                              void *stream);
/* Diagnostics of this thread's last count phase (no reference counterpart): out[5] = {list checksum failed,
 * malformed lists, parked one-way edges, propagation sweeps, neighbour lists the edge sweep read}. */
void pg_bfs_cluster_debug(long long *out);
int pg_bfs_cluster_fill(int32_t N, int32_t nCluster, int32_t sumNPoint, void *ws, size_t ws_bytes,
                        int32_t *cluster_idxs, int32_t *cluster_offsets, void *stream);

/* ------------------------------------------------------------------------------------------------
 * roipool_fp / roipool_bp     replace PG_OP.roipool_fp / roipool_bp (roipool.h:15,21; roipool.cu:12-57)
 * fp: per proposal and channel, max over rows [offsets[p], offsets[p+1]) with the LOWEST row on ties,
 * -inf / argmax -1 when nothing compares greater than -inf (empty proposal, NaN, -inf).
 * nRows = rows of feats (bounds the tiling; rows outside [offsets[0], offsets[nProposal]) are ignored).
 * ws: nProposal * C * 8 bytes.
 * bp: d_feats[maxidx[p][c]][c] += d_out[p][c] (entries with maxidx < 0 are skipped; the reference
 * writes out of bounds there, roipool.cu:45-46).
 * ---------------------------------------------------------------------------------------------- */
size_t pg_roipool_workspace_bytes(int32_t nProposal, int32_t C);
int pg_roipool_fp(const float *feats, const int32_t *offsets, float *out, int32_t *maxidx, int32_t nRows,
                  int32_t nProposal, int32_t C, void *ws, size_t ws_bytes, void *stream);
int pg_roipool_bp(float *d_feats, const int32_t *offsets, const int32_t *maxidx, const float *d_out,
                  int32_t nProposal, int32_t C, void *stream);

/* ------------------------------------------------------------------------------------------------
 * sec_mean / sec_min / sec_max     replace PG_OP.sec_mean / sec_min / sec_max
 *                                  (sec_mean.h:14,17,20; sec_mean.cu:12-86)
 * mean = sum_i fl(x_i / (float)len) accumulated left to right (bit-exact with the reference's order);
 * min / max from +inf / -inf with strict compares.
 * ---------------------------------------------------------------------------------------------- */
int pg_sec_mean(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
                int32_t C, void *stream);
int pg_sec_min(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
               int32_t C, void *stream);
int pg_sec_max(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
               int32_t C, void *stream);

/* ------------------------------------------------------------------------------------------------
 * get_iou     replaces PG_OP.get_iou (get_iou.h:16, get_iou.cu:12-38)
 * iou[p][g] = (float)((double)(float)inter / ((double)(float)(|P| + pointnum[g] - inter) + 1e-5)).
 * ---------------------------------------------------------------------------------------------- */
int pg_get_iou(const int32_t *proposals_idx, const int32_t *proposals_offset,
               const int64_t *instance_labels, const int32_t *instance_pointnum, float *proposals_iou,
               int32_t nInstance, int32_t nProposal, void *stream);

/* ------------------------------------------------------------------------------------------------
 * gather_rows     an ADDITION, not a PG_OP function: dst[i][:] = src[idx[i]][:], the row gather the
 * reference's caller performs in torch between the ops (clusters_feats = feats[c_idxs],
 * model/pointgroup.py:133-134; score_feats.features[p2v_map], :333).  idx is int32 or int64.
 * ---------------------------------------------------------------------------------------------- */
int pg_gather_rows(const float *src, const void *idx, int idx_is_int64, float *dst, int64_t nIdx, int32_t C,
                   void *stream);

/* ------------------------------------------------------------------------------------------------
 * cluster_coords     (no native counterpart: the caller-side glue of clusters_voxelization,
 *                    model/pointgroup.py:125-167, ~40 torch kernels + sec_mean / sec_min / sec_max there)
 * For proposals given as bfs_cluster output (cluster_idxs int32 [S,2], cluster_offsets int32 [nC+1]) over
 * point coordinates fp32 [N,3]: out_coords int64 [S,4] = (cluster id, voxel x, y, z) on the fullscale^3
 * grid, center / size fp32 [nC,3] = the proposals' axis-aligned boxes.  rand6 (device, 6 floats) stands
 * for the two torch.rand(3) draws of :161.  Bit-identical to the torch sequence, operation by operation.
 * ---------------------------------------------------------------------------------------------- */
size_t pg_cluster_coords_workspace_bytes(int32_t nCluster);
int pg_cluster_coords(const float *coords, const int32_t *cluster_idxs, const int32_t *cluster_offsets,
                      int32_t sumNPoint, int32_t nCluster, int32_t fullscale, float scale, const float *rand6,
                      void *ws, size_t ws_bytes, int64_t *out_coords, float *center, float *size, void *stream);

/* ------------------------------------------------------------------------------------------------
 * pack_proposals     (multi-GPU plumbing, no native counterpart: the padding step of convert_stack_to_batch,
 *                    model/pointgroup.py:237-257, for the block that is all-gathered between ranks)
 * out fp32 [B, P, 46] (zero-filled here): scene b keeps its first P proposals in proposal order; row =
 * score feats 16 | 8 corners x 3 | centre 3 | semantic class | score | mask.  ws: 2 * nProposal int32.
 * ---------------------------------------------------------------------------------------------- */
int pg_pack_proposals(const int32_t *proposals_idx, const int32_t *proposals_offset, const int64_t *locs_scaled,
                      const int64_t *semantic_preds, const float *center, const float *size, const float *feats,
                      const float *score, int32_t nProposal, int32_t C, int32_t B, int32_t P, int32_t *ws, float *out,
                      void *stream);

/* ------------------------------------------------------------------------------------------------
 * cross_iou / nms_instances     (no native counterpart: the instance NMS of PointGroup.test)
 * cross_iou replaces model/pointgroup.py:577-590 -- a dense int mask [nProposal, N] scattered from the
 * (proposal, point) rows of proposals_idx, its product with its own transpose (torch.mm) and
 * inter / (n_p + n_q - inter) -- by a sparse count over the rows themselves.  proposals_idx int32
 * [nPairs, 2] in any order, repeated rows count once (as in the mask); cross_ious fp32 [nProposal, nProposal],
 * bit-identical to the torch sequence (0/0 = NaN for proposals without points); npoint int32 [nProposal]
 * (optional) = distinct points per proposal (proposals_mask.sum(1), :582).  Synchronises `stream` once
 * (rows outside [0, nProposal) x [0, N) are reported as PG_EINVAL).
 * nms_instances replaces lib/utils/eval.py:75-97 (get_nms_instances, numpy on the host after a D2H copy of
 * the matrix): proposals in descending score order (ties: lower index first; NaN scores last), a proposal
 * is dropped when a kept one has cross_iou > threshold with it.  pick int32 [n] receives the kept
 * proposals in pick order, *host_n_pick their number.
 * ---------------------------------------------------------------------------------------------- */
size_t pg_cross_iou_workspace_bytes(int64_t nPairs, int64_t nProposal, int64_t N);
int pg_cross_iou(const int32_t *proposals_idx, int32_t nPairs, int32_t nProposal, int32_t N, void *ws,
                 size_t ws_bytes, float *cross_ious, int32_t *npoint, void *stream);
/* pick_masks replaces clusters_mask = proposals_mask[pick_idxs] (model/pointgroup.py:593): out int32 [nPick, N], row k
 * = 0/1 membership of proposal pick[k], whose members are rows proposals_offset[p] .. proposals_offset[p+1] of
 * proposals_idx (bfs_cluster's layout).  ws: 8 bytes.  Synchronises `stream` once (out-of-range ids: PG_EINVAL). */
int pg_pick_masks(const int32_t *proposals_idx, const int32_t *proposals_offset, int32_t nProposal,
                  const int32_t *pick, int32_t nPick, int32_t N, void *ws, int32_t *out, void *stream);
size_t pg_nms_instances_workspace_bytes(int32_t n);
int pg_nms_instances(const float *cross_ious, const float *scores, int32_t n, float threshold, void *ws,
                     size_t ws_bytes, int32_t *pick, int32_t *host_n_pick, void *stream);

/* ------------------------------------------------------------------------------------------------
 * collate_points     (no native counterpart: the per-point part of sparse_collate_fn,
 *                    lib/dataset/pipeline.py:937-985, which the reference runs in forked CPU workers)
 * For the concatenated points of B scenes (batch_offsets int32 [B+1]): out_locs_scaled int64 [N,4] =
 * (scene, trunc(x), trunc(y), trunc(z)) of the fp32 voxel-scale coordinates (:939-943) -- the input of
 * voxelize_idx --, out_sem_labels = int64 copy (:982), out_instance_ids = int64 ids with every id other
 * than -1 moved up by instance_offsets[scene] (:963-964,983).  The sem / instance pairs may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int pg_collate_points(const float *locs_scaled, const int32_t *sem_labels, const int32_t *instance_ids,
                      const int32_t *batch_offsets, const int32_t *instance_offsets, int32_t N, int32_t B,
                      int64_t *out_locs_scaled, int64_t *out_sem_labels, int64_t *out_instance_ids, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PG_B200_H_ */
