/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the PointGroup proposal ops of
 * daveredrum/D3Net, lib/pointgroup_ops.  It is the checker for the CUDA path and the CPU leg of
 * bench.py; nothing under d3net_b200/ may import, link or execute it.
 *
 * Every function cites the reference file:line (relative to /root/reference/lib/pointgroup_ops/)
 * whose arithmetic and ordering it restates.  It is written from the reference's behaviour, not
 * copied: plain C, flat arrays, no torch, no sparsehash.
 *
 * Parity status: PINNED against the reference's own compiled ops (oracle/_ref/PG_OP.so, built by
 * oracle/build_ref.py from the unmodified sources) -- on CPU for voxelize_idx and bfs_cluster (the
 * only ops the reference implements on CPU), and on the GPU box for the nine CUDA kernels
 * (tests/test_gpu_reference.py).  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so the committed fixtures in tests/golden/ were generated from that binary
 * by tests/golden/make_golden.py.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp oracle/pg_oracle.c -o oracle/libpg_oracle.so -lm
 * -ffp-contract=off is REQUIRED: the float ops below must round exactly where the reference's do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API void orc_free(void *p) { free(p); }

/* ============================================================================================
 * voxelize_idx   (src/voxelize/voxelize.cpp:11-152)
 *
 * Points are grouped by (batch, x, y, z).  The reference narrows every coordinate to Int = int32
 * when it builds Point<3> (voxelize.cpp:96-97, datatype.h:9,11) and keeps one hash map per batch
 * index but ONE global voxel counter (voxelize.cpp:98), so voxel ids follow first occurrence in
 * input order across the whole batch.  Which hash is used is irrelevant (only find / operator[]).
 * ============================================================================================ */
typedef struct { int32_t b, x, y, z; int32_t vid; } orc_slot;

static inline uint64_t orc_mix(int32_t b, int32_t x, int32_t y, int32_t z) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    uint32_t k[4] = {(uint32_t)b, (uint32_t)x, (uint32_t)y, (uint32_t)z};
    for (int i = 0; i < 4; i++) { h ^= k[i]; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 32; }
    return h;
}

/* Phase 1: input_map (voxelize.cpp:82,103), number of voxels, maxActive (voxelize.cpp:139-142;
 * fixed to 1 for modes 0/1/2, :111).  ncols must be 4 (batch column first): the reference's
 * 3-column branch groups correctly but then reads coords with stride 4 (voxelize.cpp:43). */
ORC_API int orc_voxelize_idx_map(const int64_t *coords, int N, int mode, int32_t *input_map,
                                 int32_t *nActive_out, int32_t *maxActive_out) {
    uint64_t cap = 16;
    while (cap < (uint64_t)N * 2 + 2) cap <<= 1;
    orc_slot *tab = (orc_slot *)malloc(cap * sizeof(orc_slot));
    int32_t *cnt = (int32_t *)calloc((size_t)N + 1, sizeof(int32_t));
    if (!tab || !cnt) { free(tab); free(cnt); return -1; }
    for (uint64_t i = 0; i < cap; i++) tab[i].vid = -1;
    int32_t nActive = 0;
    for (int i = 0; i < N; i++) {
        const int64_t *c = coords + (size_t)i * 4;
        int32_t b = (int32_t)c[0], x = (int32_t)c[1], y = (int32_t)c[2], z = (int32_t)c[3];
        uint64_t s = orc_mix(b, x, y, z) & (cap - 1);
        for (;;) {
            orc_slot *t = &tab[s];
            if (t->vid < 0) { t->b = b; t->x = x; t->y = y; t->z = z; t->vid = nActive++; break; }
            if (t->b == b && t->x == x && t->y == y && t->z == z) break;
            s = (s + 1) & (cap - 1);
        }
        input_map[i] = tab[s].vid;
        cnt[tab[s].vid]++;
    }
    int32_t maxActive = 1;
    if (mode == 3 || mode == 4)
        for (int v = 0; v < nActive; v++) if (cnt[v] > maxActive) maxActive = cnt[v];
    *nActive_out = nActive;
    *maxActive_out = maxActive;
    free(tab); free(cnt);
    return 0;
}

/* Phase 2: output_map rows [cnt, p0 < p1 < ..., 0-pad] (voxelize.cpp:143-149) and output_coords =
 * the coords row of rule[1], the voxel's first listed point (voxelize.cpp:39-47).  Mode 1 keeps the
 * FIRST point (front(), :130), mode 2 the LAST (back(), :136), mode 0 asserts uniqueness (:121-125).
 * Both outputs must be pre-zeroed by the caller (voxelize.cpp:22-26). */
ORC_API void orc_voxelize_idx_fill(const int64_t *coords, const int32_t *input_map, int N, int M,
                                   int maxActive, int mode, int64_t *output_coords, int32_t *output_map) {
    const int W = maxActive + 1;
    for (int i = 0; i < N; i++) {
        int32_t *row = output_map + (size_t)input_map[i] * W;
        if (mode == 3 || mode == 4) { row[0]++; row[row[0]] = i; }
        else if (mode == 2) { row[0] = 1; row[1] = i; }
        else if (row[0] == 0) { row[0] = 1; row[1] = i; }            /* modes 0, 1: first point */
    }
    for (int v = 0; v < M; v++) {
        const int64_t *c = coords + (size_t)output_map[(size_t)v * W + 1] * 4;
        for (int j = 0; j < 4; j++) output_coords[(size_t)v * 4 + j] = c[j];
    }
}

/* ============================================================================================
 * voxelize_fp / voxelize_bp  (src/voxelize/voxelize.cu:10-23, 35-48)
 * point_recover_fp = voxelize_bp(average=false), point_recover_bp = voxelize_fp(average=false)
 * (src/voxelize/voxelize.cpp:189,201).
 * One thread owns one (row, plane), so the atomicAdd sequence is a plain left-to-right sum of
 * fl(mult * x) starting from the pre-zeroed output; mult = fl(1 / cnt) (IEEE division).
 * ============================================================================================ */
ORC_API void orc_voxelize_fp(const float *feats, float *out, const int32_t *rules, int M, int maxActive,
                             int C, int average) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < M; row++) {
        const int32_t *r = rules + (size_t)row * (maxActive + 1);
        int n = r[0];
        float mult = (average && n > 0) ? 1.0f / (float)n : 1.0f;
        float *o = out + (size_t)row * C;
        for (int i = 1; i <= n; i++) {
            const float *inp = feats + (size_t)r[i] * C;
            for (int c = 0; c < C; c++) { float t = mult * inp[c]; o[c] = o[c] + t; }
        }
    }
}

ORC_API void orc_voxelize_bp(const float *d_out, float *d_feats, const int32_t *rules, int M, int maxActive,
                             int C, int average) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < M; row++) {
        const int32_t *r = rules + (size_t)row * (maxActive + 1);
        int n = r[0];
        float mult = (average && n > 0) ? 1.0f / (float)n : 1.0f;
        const float *o = d_out + (size_t)row * C;
        for (int i = 1; i <= n; i++) {
            float *inp = d_feats + (size_t)r[i] * C;
            for (int c = 0; c < C; c++) { float t = mult * o[c]; inp[c] = inp[c] + t; }
        }
    }
}

/* ============================================================================================
 * ballquery_batch_p  (src/bfs_cluster/bfs_cluster.cu:15-60)
 *
 * Per point i: scan k over its scene [batch_offsets[b], batch_offsets[b+1]) in ascending order, keep
 * k when d2 < r2 (strict; self included; NaN never matches), stop after the first 1000 hits (:37-45).
 * As compiled by nvcc (default -fmad=true) the distance is
 *     d2 = fma(dz, dz, fma(dx, dx, fl(dy * dy))),  dx = ox - x ...,  r2 = fl(r * r)
 * (verified in the SASS of oracle/_ref).  Segment placement in the reference comes
 * from an atomicAdd race (:48); here segments are laid out in point order, which is one of the
 * reference's possible outcomes.  use_grid=0 is the literal O(n * n_scene) scan; use_grid=1 finds
 * the same candidates through a uniform grid and evaluates the identical predicate (checked equal
 * to the literal scan in tests/test_oracle.py).
 * ============================================================================================ */
#define ORC_BQ_CAP 1000

static inline int orc_bq_hit(const float *xyz, int i, int k, float r2) {
    float dx = xyz[i * 3 + 0] - xyz[k * 3 + 0];
    float dy = xyz[i * 3 + 1] - xyz[k * 3 + 1];
    float dz = xyz[i * 3 + 2] - xyz[k * 3 + 2];
    float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    return d2 < r2;
}

static int orc_cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

typedef struct { int *v; int n, cap; } orc_ivec;
static void orc_push(orc_ivec *a, int x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (int *)realloc(a->v, (size_t)a->cap * sizeof(int)); }
    a->v[a->n++] = x;
}

/* Returns the total neighbour count; *idx_out is malloc'ed (release with orc_free). */
ORC_API int64_t orc_ballquery(const float *xyz, const int32_t *batch_idxs, const int32_t *batch_offsets,
                              int n, int B, float radius, int use_grid, int32_t *start_len, int32_t **idx_out) {
    const float r2 = radius * radius;
    orc_ivec *lists = (orc_ivec *)calloc((size_t)n > 0 ? n : 1, sizeof(orc_ivec));
    if (!use_grid) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            int b = batch_idxs[i];
            int cnt = 0;
            for (int k = batch_offsets[b]; k < batch_offsets[b + 1]; k++) {
                if (orc_bq_hit(xyz, i, k, r2)) {
                    if (cnt < ORC_BQ_CAP) orc_push(&lists[i], k); else break;
                    ++cnt;
                }
            }
        }
    } else {
        /* per-scene uniform grid with cell edge slightly above r; candidates = 27 surrounding cells */
        const double cell0 = fabs((double)radius) * 1.001;   /* r enters the predicate only as r*r */
        for (int b = 0; b < B && r2 > 0.f; b++) {
            int s = batch_offsets[b], e = batch_offsets[b + 1];
            if (e <= s) continue;
            double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = s; i < e; i++) for (int d = 0; d < 3; d++) {
                double v = xyz[i * 3 + d];
                if (isfinite(v)) { if (v < mn[d]) mn[d] = v; if (v > mx[d]) mx[d] = v; }
            }
            int64_t dim[3];
            double cell[3];
            int ok = 1;
            for (int d = 0; d < 3; d++) {
                if (!(mx[d] >= mn[d])) { ok = 0; break; }
                cell[d] = cell0;                     /* cells may only grow (>= r keeps the 27-cell search exact) */
                if ((mx[d] - mn[d]) / cell[d] > 1048576.0) cell[d] = (mx[d] - mn[d]) / 1048576.0;
                dim[d] = (int64_t)floor((mx[d] - mn[d]) / cell[d]) + 1;
            }
            if (!ok) continue;                       /* no finite point in this scene */
            /* hash cells: open addressing on the linear cell id */
            uint64_t cap = 16; while (cap < (uint64_t)(e - s) * 2 + 2) cap <<= 1;
            int64_t *key = (int64_t *)malloc(cap * sizeof(int64_t));
            int *head = (int *)malloc(cap * sizeof(int));
            int *next = (int *)malloc((size_t)(e - s) * sizeof(int));
            int64_t *pc = (int64_t *)malloc((size_t)(e - s) * 3 * sizeof(int64_t));
            for (uint64_t i = 0; i < cap; i++) { key[i] = -1; head[i] = -1; }
            for (int i = e - 1; i >= s; i--) {      /* reverse so each chain is ascending */
                const float *p = xyz + i * 3;
                next[i - s] = -2;
                if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
                int64_t c[3];
                for (int d = 0; d < 3; d++) { c[d] = (int64_t)floor(((double)p[d] - mn[d]) / cell[d]); pc[(size_t)(i - s) * 3 + d] = c[d]; }
                int64_t id = (c[2] * dim[1] + c[1]) * dim[0] + c[0];
                uint64_t h = ((uint64_t)id * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
                while (key[h] != -1 && key[h] != id) h = (h + 1) & (cap - 1);
                key[h] = id; next[i - s] = head[h]; head[h] = i;
            }
#pragma omp parallel for schedule(dynamic, 256)
            for (int i = s; i < e; i++) {
                if (next[i - s] == -2) continue;     /* non-finite point: matches nothing, not even itself */
                orc_ivec cand = {0, 0, 0};
                const int64_t *c = pc + (size_t)(i - s) * 3;
                for (int64_t dz = -1; dz <= 1; dz++) for (int64_t dy = -1; dy <= 1; dy++) for (int64_t dx = -1; dx <= 1; dx++) {
                    int64_t x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                    if (x < 0 || y < 0 || z < 0 || x >= dim[0] || y >= dim[1] || z >= dim[2]) continue;
                    int64_t id = (z * dim[1] + y) * dim[0] + x;
                    uint64_t h = ((uint64_t)id * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
                    while (key[h] != -1 && key[h] != id) h = (h + 1) & (cap - 1);
                    if (key[h] == -1) continue;
                    for (int k = head[h]; k >= 0; k = next[k - s]) if (orc_bq_hit(xyz, i, k, r2)) orc_push(&cand, k);
                }
                qsort(cand.v, (size_t)cand.n, sizeof(int), orc_cmp_int);
                if (cand.n > ORC_BQ_CAP) cand.n = ORC_BQ_CAP;
                lists[i] = cand;
            }
            free(key); free(head); free(next); free(pc);
        }
    }
    int64_t total = 0;
    for (int i = 0; i < n; i++) { start_len[i * 2] = (int32_t)total; start_len[i * 2 + 1] = lists[i].n; total += lists[i].n; }
    int32_t *idx = (int32_t *)malloc((size_t)(total > 0 ? total : 1) * sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        if (lists[i].n) memcpy(idx + start_len[i * 2], lists[i].v, (size_t)lists[i].n * sizeof(int));
        free(lists[i].v);
    }
    free(lists);
    *idx_out = idx;
    return total;
}

/* ============================================================================================
 * bfs_cluster  (src/bfs_cluster/bfs_cluster.cpp:28-112)
 * For i ascending, every unvisited i seeds a FIFO BFS along its neighbour list to unvisited points
 * with an EQUAL semantic label (:44-45); components with size >= threshold are kept (:67) and
 * numbered in seed order; members are emitted in BFS order (:77-86).
 * Returns sumNPoint; both outputs are malloc'ed (release with orc_free).
 * ============================================================================================ */
ORC_API int64_t orc_bfs_cluster(const int32_t *semantic_label, const int32_t *ball_query_idxs,
                                const int32_t *start_len, int N, int threshold,
                                int32_t **cluster_idxs_out, int32_t **cluster_offsets_out, int32_t *nCluster_out) {
    unsigned char *visited = (unsigned char *)calloc((size_t)N + 1, 1);
    int32_t *queue = (int32_t *)malloc(((size_t)N + 1) * sizeof(int32_t));
    int32_t *members = (int32_t *)malloc(((size_t)N + 1) * sizeof(int32_t));      /* kept clusters, concatenated */
    int32_t *offsets = (int32_t *)malloc(((size_t)N + 2) * sizeof(int32_t));
    int32_t nC = 0;
    int64_t sum = 0;
    offsets[0] = 0;
    for (int i = 0; i < N; i++) {
        if (visited[i]) continue;
        int qh = 0, qt = 0;
        queue[qt++] = i; visited[i] = 1;
        while (qh < qt) {
            int cur = queue[qh++];
            int start = start_len[cur * 2], len = start_len[cur * 2 + 1];
            int label = semantic_label[cur];
            for (int e = start; e < start + len; e++) {
                int j = ball_query_idxs[e];
                if (semantic_label[j] != label) continue;
                if (visited[j]) continue;
                visited[j] = 1; queue[qt++] = j;
            }
        }
        if (qt >= threshold) {
            memcpy(members + sum, queue, (size_t)qt * sizeof(int32_t));
            sum += qt; nC++; offsets[nC] = (int32_t)sum;
        }
    }
    int32_t *ci = (int32_t *)malloc((size_t)(sum > 0 ? sum : 1) * 2 * sizeof(int32_t));
    for (int c = 0; c < nC; c++)
        for (int k = offsets[c]; k < offsets[c + 1]; k++) { ci[(size_t)k * 2] = c; ci[(size_t)k * 2 + 1] = members[k]; }
    *cluster_idxs_out = ci;
    *cluster_offsets_out = (int32_t *)realloc(offsets, ((size_t)nC + 1) * sizeof(int32_t));
    *nCluster_out = nC;
    free(visited); free(queue); free(members);
    return sum;
}

/* ============================================================================================
 * roipool_fp / roipool_bp  (src/roipool/roipool.cu:12-31, 42-49)
 * max_val starts at (float)-1e50 = -inf; strict '>' keeps the lowest row on ties and never selects
 * NaN; argmax is the GLOBAL row index into feats (-1 for an empty or all -inf/NaN proposal).
 * ============================================================================================ */
ORC_API void orc_roipool_fp(const float *feats, const int32_t *offsets, float *out, int32_t *maxidx, int nP, int C) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < nP; p++)
        for (int c = 0; c < C; c++) {
            int arg = -1; float mv = -INFINITY;
            for (int i = offsets[p]; i < offsets[p + 1]; i++) {
                float v = feats[(size_t)i * C + c];
                if (v > mv) { arg = i; mv = v; }
            }
            maxidx[(size_t)p * C + c] = arg; out[(size_t)p * C + c] = mv;
        }
}

ORC_API void orc_roipool_bp(float *d_feats, const int32_t *maxidx, const float *d_out, int nP, int C) {
    for (int p = 0; p < nP; p++)
        for (int c = 0; c < C; c++) {
            int a = maxidx[(size_t)p * C + c];
            if (a >= 0) d_feats[(size_t)a * C + c] += d_out[(size_t)p * C + c];  /* a < 0 is out of bounds in the reference */
        }
}

/* ============================================================================================
 * sec_mean / sec_min / sec_max  (src/sec_mean/sec_mean.cu:12-27, 38-53, 64-79)
 * mean = sum of fl(x / count) left to right, count = (float)(end - start); min/max strict compares
 * from +inf / -inf.
 * ============================================================================================ */
ORC_API void orc_sec_mean(const float *inp, const int32_t *offsets, float *out, int nP, int C) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < nP; p++) {
        float count = (float)(offsets[p + 1] - offsets[p]);
        for (int c = 0; c < C; c++) {
            float mean = 0.f;
            for (int i = offsets[p]; i < offsets[p + 1]; i++) { float q = inp[(size_t)i * C + c] / count; mean = mean + q; }
            out[(size_t)p * C + c] = mean;
        }
    }
}

ORC_API void orc_sec_min(const float *inp, const int32_t *offsets, float *out, int nP, int C) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < nP; p++)
        for (int c = 0; c < C; c++) {
            float mv = INFINITY;
            for (int i = offsets[p]; i < offsets[p + 1]; i++) { float v = inp[(size_t)i * C + c]; if (v < mv) mv = v; }
            out[(size_t)p * C + c] = mv;
        }
}

ORC_API void orc_sec_max(const float *inp, const int32_t *offsets, float *out, int nP, int C) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < nP; p++)
        for (int c = 0; c < C; c++) {
            float mv = -INFINITY;
            for (int i = offsets[p]; i < offsets[p + 1]; i++) { float v = inp[(size_t)i * C + c]; if (v > mv) mv = v; }
            out[(size_t)p * C + c] = mv;
        }
}

/* ============================================================================================
 * get_iou  (src/get_iou/get_iou.cu:12-29)
 * inter = #{i in proposal : (int)instance_labels[proposals_idx[i]] == g};
 * iou = (float)( (double)(float)inter / ( (double)(float)(P + I_g - inter) + 1e-5 ) )   (:26, the
 * 1e-5 literal is a double, so the add and the divide run in fp64 and the store rounds to fp32).
 * ============================================================================================ */
ORC_API void orc_get_iou(const int32_t *proposals_idx, const int32_t *proposals_offset,
                         const int64_t *instance_labels, const int32_t *instance_pointnum,
                         float *iou, int nInstance, int nProposal) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < nProposal; p++) {
        int start = proposals_offset[p], end = proposals_offset[p + 1];
        int32_t *hist = (int32_t *)calloc((size_t)nInstance + 1, sizeof(int32_t));
        for (int i = start; i < end; i++) {
            int l = (int)instance_labels[proposals_idx[i]];
            if (l >= 0 && l < nInstance) hist[l]++;
        }
        for (int g = 0; g < nInstance; g++) {
            int inter = hist[g];
            double den = (double)(float)(end - start + instance_pointnum[g] - inter) + 1e-5;
            iou[(size_t)p * nInstance + g] = (float)((double)(float)inter / den);
        }
        free(hist);
    }
}
