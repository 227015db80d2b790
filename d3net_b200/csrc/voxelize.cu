// voxelize_idx (GPU hash grouping + stable sort), voxelize_fp / voxelize_bp (segment mean scatter /
// gather), point_recover.  Reference behaviour: lib/pointgroup_ops/src/voxelize/voxelize.{cpp,cu}.
#include <stdlib.h>

#include "common.cuh"

namespace pg {

// =================================================================================================
// voxelize_idx
// =================================================================================================
struct VoxWs {
    int4 *keys;
    GroupTable tab;
    int32_t *pslot, *cnt, *voff;
    uint32_t *kA, *vA, *kB, *vB;
    int32_t *hist;
    int64_t *scan_tmp;
    int64_t *scalars;  // [0] nGroups, [1] maxActive
    bool ok;
    size_t used;
};

static VoxWs vox_layout(void *ws, size_t ws_bytes, int64_t N) {
    Arena a(ws, ws_bytes);
    VoxWs w;
    const size_t n = (size_t)(N > 0 ? N : 1);
    w.tab.cap = group_table_cap(N);
    w.keys = a.take<int4>(n);
    w.tab.slot_rep = a.take<int32_t>(w.tab.cap);
    w.tab.slot_gid = a.take<int32_t>(w.tab.cap);
    w.tab.slot_key = nullptr;                     // nobody looks voxels up by key afterwards
    w.pslot = a.take<int32_t>(n);
    w.cnt = a.take<int32_t>(n + 1);
    w.voff = a.take<int32_t>(n + 1);
    w.kA = a.take<uint32_t>(n);
    w.vA = a.take<uint32_t>(n);
    w.kB = a.take<uint32_t>(n);
    w.vB = a.take<uint32_t>(n);
    w.hist = a.take<int32_t>(radix_tmp_count(N));
    w.scan_tmp = a.take<int64_t>(scan_tmp_count((int64_t)(n + radix_tmp_count(N))));
    w.scalars = a.take<int64_t>(4);
    w.ok = a.ok;
    w.used = a.used;
    return w;
}

// int64 [N,4] -> int4 keys; every column narrowed to int32 like Point<3>/Int (voxelize.cpp:95-97)
__global__ void k_vox_keys(const int64_t *__restrict__ coords, int64_t N, int4 *__restrict__ keys, Fill table, Fill cnt) {
    pdl_enter();
    grid_fill(table);      // the grouping's hash table and counts start clean
    grid_fill(cnt);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const longlong2 *p = reinterpret_cast<const longlong2 *>(coords) + i * 2;
    longlong2 a = __ldg(p), b = __ldg(p + 1);
    keys[i] = make_int4((int)a.x, (int)a.y, (int)b.x, (int)b.y);
}

// The largest voxel, as a pass of its own: taking it from the counting atomics' return values made every one of them a
// round trip instead of a fire-and-forget reduction (k_group_assign 14 -> 37 us per call).  Plain pointers on purpose: the
// counts are written by the kernel just before and the group count by the one before that -- read-only loads could be
// moved above the dependency wait (common.cuh, HAZARD).
__global__ void k_max_i32(const int32_t *v, const int64_t *n_dev, int64_t *out) {
    pdl_enter();
    const int64_t n = *n_dev;
    int m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, v[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax((unsigned long long *)out, (unsigned long long)m);
}

constexpr int kVoxRankMax = 32;        // largest voxel (points) served by the sort-free fill: a thread per point counts
                                       // its voxel's segment, so long segments (cluster grids: hundreds of points per voxel,
                                       // different ones in every lane) diverge and scatter -- those keep the sort

// The points of every voxel side by side (voxel v: grouped[voff[v] .. voff[v] + cnt[v]), in claim order).
__global__ void k_vox_scatter(const int32_t *__restrict__ input_map, const int32_t *__restrict__ voff, int64_t N,
                              int32_t *cursor, uint32_t *__restrict__ grouped, Fill rows) {
    pdl_enter();
    grid_fill(rows);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int v = __ldg(input_map + i);
    grouped[__ldg(voff + v) + atomicAdd(cursor + v, 1)] = (uint32_t)i;
}

// output_map rows [cnt, p0 < p1 < ..., 0-pad] (voxelize.cpp:139-149; modes 0/1/2: :121-138) without a sort: a point's
// column is its RANK among the points of its voxel (how many of them have a smaller index), counted against the voxel's
// segment of `grouped` -- voxels hold one or two points on a scene grid and tens to hundreds on a cluster grid, and the
// segment is contiguous, so the count runs out of L1.  The map was zero-filled; the rank-0 (mode 2: last) point also
// writes the voxel's row of output_coords (voxelize.cpp:39-47) and the count column.
__global__ void __launch_bounds__(256) k_vox_rank(const int64_t *__restrict__ coords, const int32_t *__restrict__ input_map,
                                                  const int32_t *__restrict__ cnt, const int32_t *__restrict__ voff,
                                                  const uint32_t *__restrict__ grouped, int64_t N, int32_t W, int mode,
                                                  int64_t *__restrict__ out_coords, int32_t *__restrict__ out_map) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int v = __ldg(input_map + i);
    const int n = __ldg(cnt + v);
    const uint32_t *seg = grouped + __ldg(voff + v);
    int rank = 0;
    for (int t = 0; t < n; t++) rank += (__ldg(seg + t) < (uint32_t)i) ? 1 : 0;
    int32_t *row = out_map + (int64_t)v * W;
    const bool listed = mode == 3 || mode == 4;
    if (listed) row[1 + rank] = (int32_t)i;
    const bool rep = mode == 2 ? rank == n - 1 : rank == 0;           // the point whose coordinates name the voxel
    if (rep) {
        row[0] = listed ? n : 1;
        if (!listed) row[1] = (int32_t)i;
        const longlong2 *p = reinterpret_cast<const longlong2 *>(coords) + i * 2;
        longlong2 *q = reinterpret_cast<longlong2 *>(out_coords) + (int64_t)v * 2;
        q[0] = __ldg(p);
        q[1] = __ldg(p + 1);
    }
}

// flat index -> (row, column): a 32-bit division when the index fits (the 64-bit one is a ~80-instruction
// subroutine, and these kernels do little else per element)
__device__ __forceinline__ void row_col(int64_t t, int32_t width, int64_t &row, int &col) {
    if (t <= 0xffffffffLL) {
        const unsigned q = (unsigned)t / (unsigned)width;
        row = q;
        col = (int)((unsigned)t - q * (unsigned)width);
    } else {
        row = t / width;
        col = (int)(t - row * width);
    }
}

// output_map rows [cnt, p0 < p1 < ..., 0-pad] (voxelize.cpp:139-149; modes 0/1/2: :121-138) and
// output_coords = coords row of rule[1] (voxelize.cpp:39-47)
__device__ __forceinline__ int vox_map_entry(const uint32_t *__restrict__ sorted, int mode, int j, int n, int off) {
    if (mode == 3 || mode == 4) return (j == 0) ? n : (j <= n ? (int)sorted[off + j - 1] : 0);
    return (j == 0) ? 1 : (int)sorted[mode == 2 ? off + n - 1 : off];
}

// VEC: every thread produces four consecutive entries of the flat [M * W] map and stores them as one
// int4 -- one division per four entries, fully coalesced 16-byte stores (most of the map is padding).
template <bool VEC>
__global__ void k_vox_fill(const int64_t *__restrict__ coords, const int32_t *__restrict__ cnt,
                           const int32_t *__restrict__ voff, const uint32_t *__restrict__ sorted, int32_t M,
                           int32_t W, int mode, int64_t *__restrict__ out_coords, int32_t *__restrict__ out_map) {
    pdl_enter();
    const int64_t total = (int64_t)M * W;
    if (VEC) {
        const int64_t quads = total >> 2;
        {   // the (up to three) entries past the last full quad
            const int64_t t = (quads << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (t < total) {
                const int v = (int)(t / W), j = (int)(t - (int64_t)v * W);
                out_map[t] = vox_map_entry(sorted, mode, j, cnt[v], voff[v]);
            }
        }
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
            const int64_t t = q << 2;
            int64_t v64;
            int j;
            row_col(t, W, v64, j);
            int v = (int)v64;
            int n = cnt[v], off = voff[v];
            int val[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                val[u] = vox_map_entry(sorted, mode, j, n, off);
                if (++j == W && u < 3) {
                    j = 0;
                    ++v;                                 // v < M: entry t + u + 1 exists
                    n = cnt[v];
                    off = voff[v];
                }
            }
            __stcs(reinterpret_cast<int4 *>(out_map) + q, make_int4(val[0], val[1], val[2], val[3]));
        }
    } else {
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
            const int v = (int)(t / W), j = (int)(t - (int64_t)v * W);
            out_map[t] = vox_map_entry(sorted, mode, j, cnt[v], voff[v]);
        }
    }
    const int64_t total_c = (int64_t)M * 4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_c;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t >> 2), c = (int)(t & 3);
        const int n = cnt[v], off = voff[v];
        const int first = (int)sorted[mode == 2 ? off + n - 1 : off];
        out_coords[t] = coords[(int64_t)first * 4 + c];
    }
}

// =================================================================================================
// voxelize_fp / voxelize_bp
// One thread owns one (voxel row, vector column): the per-element operation order is the
// reference's -- fl(mult * x) added left to right starting from +0 (voxelize.cu:15-19) -- so the
// result is bit-identical, while consecutive threads touch consecutive addresses of both the voxel
// row and the gathered point row.
// =================================================================================================
template <int V> struct Vec;
template <> struct Vec<1> { using T = float; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

template <int V> __device__ __forceinline__ void vzero(typename Vec<V>::T &a);
template <> __device__ __forceinline__ void vzero<1>(float &a) { a = 0.f; }
template <> __device__ __forceinline__ void vzero<2>(float2 &a) { a = make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ void vzero<4>(float4 &a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ void vmuladd(float &a, float m, float x) { a = __fadd_rn(a, __fmul_rn(m, x)); }
__device__ __forceinline__ void vmuladd(float2 &a, float m, float2 x) {
    vmuladd(a.x, m, x.x); vmuladd(a.y, m, x.y);
}
__device__ __forceinline__ void vmuladd(float4 &a, float m, float4 x) {
    vmuladd(a.x, m, x.x); vmuladd(a.y, m, x.y); vmuladd(a.z, m, x.z); vmuladd(a.w, m, x.w);
}
__device__ __forceinline__ float vscale(float m, float x) { return __fmul_rn(m, x); }
__device__ __forceinline__ float2 vscale(float m, float2 x) { return make_float2(__fmul_rn(m, x.x), __fmul_rn(m, x.y)); }
__device__ __forceinline__ float4 vscale(float m, float4 x) {
    return make_float4(__fmul_rn(m, x.x), __fmul_rn(m, x.y), __fmul_rn(m, x.z), __fmul_rn(m, x.w));
}
__device__ __forceinline__ void vred(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void vred(float2 *p, float2 v) { atomicAdd(p, v); }
__device__ __forceinline__ void vred(float4 *p, float4 v) { atomicAdd(p, v); }

template <int V>
__global__ void __launch_bounds__(256) k_voxelize_fp(const float *__restrict__ feats, float *__restrict__ out,
                                                     const int32_t *__restrict__ rules, int32_t M, int32_t W,
                                                     int32_t Cv, int average, int skip_above) {
    pdl_enter();
    using T = typename Vec<V>::T;
    const T *__restrict__ f = reinterpret_cast<const T *>(feats);
    T *__restrict__ o = reinterpret_cast<T *>(out);
    const int64_t total = (int64_t)M * Cv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t v;
        int c;
        row_col(t, Cv, v, c);
        const int32_t *r = rules + v * W;
        const int n = __ldg(r);
        if (n > skip_above) continue;                    // a long list: k_voxelize_fp_long's row
        const float mult = (average && n > 0) ? __fdiv_rn(1.0f, (float)n) : 1.0f;
        T acc;
        vzero<V>(acc);
        int i = 1;
        for (; i + 3 <= n; i += 4) {   // four independent gathers in flight, then the ordered adds
            const int r0 = __ldg(r + i), r1 = __ldg(r + i + 1), r2 = __ldg(r + i + 2), r3 = __ldg(r + i + 3);
            const T x0 = __ldg(f + (int64_t)r0 * Cv + c), x1 = __ldg(f + (int64_t)r1 * Cv + c);
            const T x2 = __ldg(f + (int64_t)r2 * Cv + c), x3 = __ldg(f + (int64_t)r3 * Cv + c);
            vmuladd(acc, mult, x0); vmuladd(acc, mult, x1); vmuladd(acc, mult, x2); vmuladd(acc, mult, x3);
        }
        for (; i <= n; i++) {
            const int r0 = __ldg(r + i);
            vmuladd(acc, mult, __ldg(f + (int64_t)r0 * Cv + c));
        }
        __stcs(o + t, acc);
    }
}

// Long lists on narrow rows (the 14^3 cluster grids: C = 16, a voxel of the floor-sized proposal holds hundreds of points).
// The adds of one (voxel, column) are a serial chain in list order -- that is the semantics -- but a thread that also
// FETCHES its operands four at a time spends a memory round trip per four points: the longest list (~400 points, ~100 round
// trips) was the kernel's run time while the rest of the machine idled.  Here a WARP takes a long row: 32 lanes gather
// batches of 32 points' rows straight into an eight-stage shared-memory ring with cp.async (seven batches in flight while
// lanes 0 .. Cv-1 add the oldest), the point indices run two batches ahead of the gathers.  Rows of at most kLongMin points
// stay with the flat kernel, which skips the rows this one takes.  (One batch in flight through registers: 76 us for the
// floor proposal's ~1500-point voxels.)
constexpr int kLongMinDefault = 24;
constexpr int kLongCv = 4;                             // vectors per row this kernel is built for (C = 16 as float4)

constexpr int kLongStages = 8;                         // batches of 32 points in flight per warp: most of a 300-point row at once
constexpr int kLongWarps = 2;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// float4 rows only (cp.async moves 16 bytes per lane and vector): C = 16
__global__ void __launch_bounds__(kLongWarps * 32) k_voxelize_fp_long(const float *__restrict__ feats, float *__restrict__ out,
                                                                      const int32_t *__restrict__ rules, int32_t M, int32_t W,
                                                                      int average, int kLongMin) {
    pdl_enter();
    using T = float4;
    constexpr int Cv = kLongCv, S = kLongStages;
    __shared__ T stage[kLongWarps][S][32 * Cv];
    const T *__restrict__ f = reinterpret_cast<const T *>(feats);
    T *__restrict__ o = reinterpret_cast<T *>(out);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nWarps = (int64_t)gridDim.x * kLongWarps;
    const int64_t w0 = (int64_t)blockIdx.x * kLongWarps + warp;
    // A warp's rows are w0, w0 + nWarps, ... (long rows come in runs -- the voxels of one big proposal -- so consecutive rows
    // must go to different warps); it looks at the lengths of 32 of them at once (lane = row: one memory round trip per
    // 32 rows instead of one per row -- 350 k rows, 3 k of them long), the next 32 lengths already in flight, and serves the
    // long ones.
    auto len_of = [&](int64_t k0) {
        const int64_t vm = w0 + (k0 + lane) * nWarps;
        return vm < M ? __ldg(rules + vm * W) : 0;
    };
    int nnext = len_of(0);
    for (int64_t k0 = 0; w0 + k0 * nWarps < M; k0 += 32) {
      const int nmine = nnext;
      nnext = len_of(k0 + 32);
      unsigned todo = __ballot_sync(0xffffffffu, nmine > kLongMin);
      while (todo) {
        const int bit = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const int64_t v = w0 + (k0 + bit) * nWarps;
        const int n = __shfl_sync(0xffffffffu, nmine, bit);
        const int32_t *r = rules + v * W;
        const float mult = average ? __fdiv_rn(1.0f, (float)n) : 1.0f;
        const int nb = (n + 31) >> 5;
        T acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // the point indices run two batches ahead of the gathers, the gathers S - 1 batches ahead of the adds
        auto idx_of = [&](int bq) { return (bq < nb && bq * 32 + lane < n) ? __ldg(r + 1 + bq * 32 + lane) : 0; };
        // lane l issues vectors l, l + 32, l + 64, l + 96 of the batch's 32 x Cv: point (item >> 2), column (item & 3)
        auto issue = [&](int bq, int pidx) {
            if (bq < nb) {
                T *dst = stage[warp][bq % S];
#pragma unroll
                for (int k = 0; k < Cv; k++) {
                    const int item = k * 32 + lane;
                    const int src = __shfl_sync(0xffffffffu, pidx, item >> 2);
                    if (bq * 32 + (item >> 2) < n) cp_async16(dst + item, f + (int64_t)src * Cv + (item & 3));
                }
            }
            cp_async_commit();                         // (an empty group keeps the count uniform)
        };
        int p0 = idx_of(0), p1 = idx_of(1);
#pragma unroll
        for (int bq = 0; bq < S - 1; bq++) {
            issue(bq, p0);
            p0 = p1;
            p1 = idx_of(bq + 2);
        }
        for (int bq = 0; bq < nb; bq++) {
            cp_async_wait<S - 2>();                     // batch bq has landed (this lane's part)
            __syncwarp();                               // ... and everybody else's
            if (lane < Cv) {
                const T *st = stage[warp][bq % S];
                const int cnt = n - bq * 32 < 32 ? n - bq * 32 : 32;
                // eight operands out of shared memory at a time, then their ordered adds
                int p = 0;
                for (; p + 8 <= cnt; p += 8) {
                    T x[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) x[u] = st[(p + u) * Cv + lane];
#pragma unroll
                    for (int u = 0; u < 8; u++) vmuladd(acc, mult, x[u]);
                }
                for (; p < cnt; p++) vmuladd(acc, mult, st[p * Cv + lane]);
            }
            __syncwarp();
            issue(bq + S - 1, p0);                      // into stage (bq - 1) % S, consumed one round ago
            p0 = p1;
            p1 = idx_of(bq + S + 1);
        }
        cp_async_wait<0>();
        if (lane < Cv) __stcs(o + v * Cv + lane, acc);
      }
    }
}

// The same sums with G lanes per output row (G a power of two >= min(Cv, 32)): a row's lanes sit side by side, so
// the gathers of one point's feature row are one coalesced request, the map row is read once per lane group
// instead of once per output element, and no thread divides by Cv (the flat kernel above spends a 64-bit
// division per element on t / Cv).  NS (<= 4) column strips per lane are loaded before the ordered adds; full
// occupancy (32 registers) measured faster than more strips in flight (4.6 vs 4.0 TB/s at C = 134).
template <int V, int G, int NS>
__global__ void __launch_bounds__(256, (V * NS <= 6 ? 8 : (V * NS <= 8 ? 6 : 4))) k_voxelize_fp_rows(const float *__restrict__ feats, float *__restrict__ out,
                                                          const int32_t *__restrict__ rules, int32_t M, int32_t W,
                                                          int32_t Cv, int average) {
    pdl_enter();
    using T = typename Vec<V>::T;
    constexpr int kRows = 32 / G;                     // rows per warp
    const T *__restrict__ f = reinterpret_cast<const T *>(feats);
    T *__restrict__ o = reinterpret_cast<T *>(out);
    const int lane = threadIdx.x & 31, sub = lane & (G - 1), rw = lane / G;
    const int64_t nWarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t v0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRows; v0 < M; v0 += nWarps * kRows) {
        const int64_t v = v0 + rw;
        if (v >= M) continue;
        const int32_t *r = rules + v * W;
        const int n = __ldg(r);
        const float mult = (average && n > 0) ? __fdiv_rn(1.0f, (float)n) : 1.0f;
        for (int c0 = sub; c0 < Cv; c0 += NS * G) {
            T acc[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) vzero<V>(acc[k]);
            for (int i = 1; i <= n; i++) {
                const T *row = f + (int64_t)__ldg(r + i) * Cv;
                T x[NS];
#pragma unroll
                for (int k = 0; k < NS; k++)
                    if (c0 + k * G < Cv) x[k] = __ldg(row + c0 + k * G);
#pragma unroll
                for (int k = 0; k < NS; k++)
                    if (c0 + k * G < Cv) vmuladd(acc[k], mult, x[k]);
            }
#pragma unroll
            for (int k = 0; k < NS; k++)
                if (c0 + k * G < Cv) __stcs(o + v * Cv + c0 + k * G, acc[k]);
        }
    }
}

// dst[i] = src[idx[i]] for whole rows (the caller-side feats[idx] gathers of the proposal path)
template <int V, typename I>
__global__ void __launch_bounds__(256) k_gather_rows(const float *__restrict__ src, const I *__restrict__ idx,
                                                     float *__restrict__ dst, int64_t nIdx, int32_t Cv) {
    pdl_enter();
    using T = typename Vec<V>::T;
    const T *__restrict__ s = reinterpret_cast<const T *>(src);
    T *__restrict__ d = reinterpret_cast<T *>(dst);
    const int64_t total = nIdx * Cv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row;
        int c;
        row_col(t, Cv, row, c);
        __stcs(d + t, __ldg(s + (int64_t)__ldg(idx + row) * Cv + c));
    }
}

template <int V>
__global__ void __launch_bounds__(256) k_voxelize_bp(const float *__restrict__ d_out, float *d_feats,
                                                     const int32_t *__restrict__ rules, int32_t M, int32_t W,
                                                     int32_t Cv, int average) {
    pdl_enter();
    using T = typename Vec<V>::T;
    const T *__restrict__ g = reinterpret_cast<const T *>(d_out);
    T *df = reinterpret_cast<T *>(d_feats);
    const int64_t total = (int64_t)M * Cv;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(t / Cv), c = (int)(t - (int64_t)v * Cv);
        const int32_t *r = rules + (int64_t)v * W;
        const int n = __ldg(r);
        if (n <= 0) continue;
        const float mult = average ? __fdiv_rn(1.0f, (float)n) : 1.0f;
        const T val = vscale(mult, __ldcs(g + t));
        for (int i = 1; i <= n; i++) vred(df + (int64_t)__ldg(r + i) * Cv + c, val);
    }
}

static int pick_vec(const void *a, const void *b, int C) {
    const uintptr_t pa = (uintptr_t)a, pb = (uintptr_t)b;
    if (C % 4 == 0 && pa % 16 == 0 && pb % 16 == 0) return 4;
    if (C % 2 == 0 && pa % 8 == 0 && pb % 8 == 0) return 2;
    return 1;
}

static int voxelize_launch(bool fp, const float *src, float *dst, const int32_t *rules, int32_t M,
                           int32_t maxActive, int32_t C, int average, cudaStream_t st) {
    PG_CHECK_ARG(M >= 0 && maxActive >= 0 && C >= 0, "negative size");
    if (M == 0 || C == 0) return PG_OK;
    PG_CHECK_ARG(src && dst && rules, "null pointer");
    const int V = pick_vec(src, dst, C);
    const int Cv = C / V, W = maxActive + 1;
    const int64_t total = (int64_t)M * Cv;
    const unsigned grid = (unsigned)(div_up(total, 256) < (int64_t)kNumSM * 64 ? div_up(total, 256) : (int64_t)kNumSM * 64);
    // wide rows (Cv >= 32: the scene features, C = 134) take the warp-per-row kernel; narrow rows with long
    // point lists (the 14^3 cluster grids, C = 16, up to hundreds of points per voxel) keep the flat kernel, whose
    // four-deep gather pipeline along the point list is what matters there (measured: 0.17 vs 0.23 ms)
    const bool rows = fp && Cv >= 32;
    // narrow rows whose lists CAN be long: a warp per long row first, the flat kernel for the rest
    static const int kLongMin = []() { const char *e = getenv("PG_VOX_LONGMIN"); return e ? atoi(e) : kLongMinDefault; }();
    const bool long_rows = fp && !rows && V == 4 && Cv == kLongCv && maxActive > kLongMin;
    const int64_t lblocks = div_up(M, kLongWarps);
    const int64_t lwave = (int64_t)kNumSM * PG_RESIDENT(k_voxelize_fp_long, kLongWarps * 32, 0);     // one wave: no late blocks
    const unsigned lgrid = (unsigned)(lblocks < lwave ? lblocks : lwave);
    const int64_t row_blocks = div_up(M, 8);
    const unsigned rgrid = (unsigned)(row_blocks < (int64_t)kNumSM * 64 ? row_blocks : (int64_t)kNumSM * 64);
#define PG_VOX(VV)                                                                                     \
    if (rows) {                                                                                        \
        if (Cv <= 64) launch(k_voxelize_fp_rows<VV, 32, 2>, rgrid, 256, 0, st, src, dst, rules, M, W, Cv, average);      \
        else if (Cv <= 96) launch(k_voxelize_fp_rows<VV, 32, 3>, rgrid, 256, 0, st, src, dst, rules, M, W, Cv, average); \
        else launch(k_voxelize_fp_rows<VV, 32, 4>, rgrid, 256, 0, st, src, dst, rules, M, W, Cv, average);               \
    }                                                                                                  \
    else if (fp) {                                                                                     \
        if (long_rows) launch(k_voxelize_fp_long, lgrid, kLongWarps * 32, 0, st, src, dst, rules, M, W, average, kLongMin);  \
        launch(k_voxelize_fp<VV>, grid, 256, 0, st, src, dst, rules, M, W, Cv, average, long_rows ? kLongMin : 0x7fffffff); \
    }                                                                                                  \
    else launch(k_voxelize_bp<VV>, grid, 256, 0, st, src, dst, rules, M, W, Cv, average)
    { PG_KTIME(fp ? "k_voxelize_fp" : "k_voxelize_bp", st);
    if (V == 4) { PG_VOX(4); } else if (V == 2) { PG_VOX(2); } else { PG_VOX(1); } }
#undef PG_VOX
    PG_LAUNCH_CHECK();
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" size_t pg_voxelize_idx_workspace_bytes(int64_t N) {
    if (N < 0) N = 0;
    return vox_layout(nullptr, 0, N).used + 256;
}

extern "C" int pg_voxelize_idx_map(const int64_t *coords, int64_t N, int mode, int32_t *input_map, void *ws,
                                   size_t ws_bytes, int32_t *host_sizes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(N >= 0 && N < 0x7fffffff, "N out of range");
    PG_CHECK_ARG(mode >= 0 && mode <= 4, "mode must be 0..4");
    PG_CHECK_ARG(host_sizes, "null host_sizes");
    host_sizes[0] = 0;
    host_sizes[1] = 1;   // voxelize.cpp:139 -- maxActive starts at 1, also for N == 0
    if (N == 0) return PG_OK;
    PG_CHECK_ARG(coords && input_map && ws, "null pointer");
    VoxWs w = vox_layout(ws, ws_bytes, N);
    if (!w.ok) { set_error("pg_voxelize_idx_map: workspace too small (%zu < %zu)", ws_bytes, w.used); return PG_EWORKSPACE; }
    PG_CUDA(cudaMemsetAsync(w.scalars, 0, 4 * sizeof(int64_t), st));
    launch(k_vox_keys, (unsigned)div_up(N, 256), 256, 0, st, coords, N, w.keys, group_table_fill(w.tab), Fill{(uint32_t *)w.cnt, (size_t)N, 0u});
    // the fill phase's cursors (kA) are cleared on the way; then the largest voxel
    PG_TRY(group_int4(w.keys, N, w.tab, w.pslot, input_map, w.cnt, w.scalars, w.scan_tmp, st, nullptr, Fill{w.kA, (size_t)N, 0u}));
    launch(k_max_i32, kNumSM * 4, 256, 0, st, w.cnt, w.scalars, w.scalars + 1);
    PG_LAUNCH_CHECK();
    int64_t h[2];
    PG_CUDA(cudaMemcpyAsync(h, w.scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    host_sizes[0] = (int32_t)h[0];
    host_sizes[1] = (mode == 3 || mode == 4) ? (int32_t)(h[1] > 1 ? h[1] : 1) : 1;
    return PG_OK;
}

extern "C" int pg_voxelize_idx_fill(const int64_t *coords, const int32_t *input_map, int64_t N, int32_t M,
                                    int32_t maxActive, int mode, void *ws, size_t ws_bytes, int64_t *output_coords, int32_t *output_map,
                                    void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(N >= 0 && M >= 0 && M <= N && maxActive >= 1, "bad sizes");
    PG_CHECK_ARG(mode >= 0 && mode <= 4, "mode must be 0..4");
    if (N == 0 || M == 0) return PG_OK;
    PG_CHECK_ARG(coords && input_map && ws && output_coords && output_map, "null pointer");
    VoxWs w = vox_layout(ws, ws_bytes, N);
    if (!w.ok) { set_error("pg_voxelize_idx_fill: workspace too small"); return PG_EWORKSPACE; }
    PG_TRY(scan_exclusive_i32(w.cnt, w.voff, M, nullptr, w.scan_tmp, st));
    const int W = maxActive + 1;
    if (maxActive <= kVoxRankMax) {
        // no sort: points side by side per voxel, then every point finds its column by counting (see k_vox_rank)
        int32_t *cursor = reinterpret_cast<int32_t *>(w.kA);
        // the cursors were cleared by the map phase; the scatter clears the rows (their zero padding) for the rank pass
        launch(k_vox_scatter, (unsigned)div_up(N, 256), 256, 0, st, input_map, w.voff, N, cursor, w.vA, Fill{(uint32_t *)output_map, (size_t)M * W, 0u});
        PG_KTIME("k_vox_rank", st);
        launch(k_vox_rank, (unsigned)div_up(N, 256), 256, 0, st, coords, input_map, w.cnt, w.voff, w.vA, N, W, mode, output_coords, output_map);
        PG_LAUNCH_CHECK();
        return PG_OK;
    }
    // voxels with thousands of points (the counting pass is quadratic in the voxel's size): stable sort of
    // (voxel id, point) by voxel id -- points end up ascending inside every voxel -- and a coalesced fill
    int bits = 0;
    while ((1ll << bits) < (long long)M) bits++;
    int res = 0;
    PG_TRY(radix_sort_pairs(reinterpret_cast<const uint32_t *>(input_map), nullptr, w.kA, w.vA, w.kB, w.vB, N, bits, w.hist, w.scan_tmp, st, &res));
    const uint32_t *sorted = res == 0 ? w.vA : w.vB;
    const int64_t total = (int64_t)M * W;
    const bool vec = ((uintptr_t)output_map & 15u) == 0;
    const int64_t units = vec ? (total >> 2 > (int64_t)M * 4 ? total >> 2 : (int64_t)M * 4) : total;
    const unsigned grid = (unsigned)(div_up(units, 256) < (int64_t)kNumSM * 32 ? div_up(units, 256) : (int64_t)kNumSM * 32);
    { PG_KTIME("k_vox_fill", st);
    if (vec) launch(k_vox_fill<true>, grid, 256, 0, st, coords, w.cnt, w.voff, sorted, M, W, mode, output_coords, output_map);
    else launch(k_vox_fill<false>, grid, 256, 0, st, coords, w.cnt, w.voff, sorted, M, W, mode, output_coords, output_map); }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_voxelize_fp(const float *feats, float *out, const int32_t *rules, int32_t M, int32_t maxActive,
                              int32_t C, int average, void *stream) {
    return voxelize_launch(true, feats, out, rules, M, maxActive, C, average, (cudaStream_t)stream);
}
extern "C" int pg_voxelize_bp(const float *d_out, float *d_feats, const int32_t *rules, int32_t M,
                              int32_t maxActive, int32_t C, int average, void *stream) {
    return voxelize_launch(false, d_out, d_feats, rules, M, maxActive, C, average, (cudaStream_t)stream);
}
// point_recover_fp = voxelize_bp(average = false), point_recover_bp = voxelize_fp(average = false)
// (voxelize.cpp:189,201)
extern "C" int pg_point_recover_fp(const float *feats, float *out, const int32_t *rules, int32_t M,
                                   int32_t maxActive, int32_t C, void *stream) {
    return voxelize_launch(false, feats, out, rules, M, maxActive, C, 0, (cudaStream_t)stream);
}
extern "C" int pg_point_recover_bp(const float *d_out, float *d_feats, const int32_t *rules, int32_t M,
                                   int32_t maxActive, int32_t C, void *stream) {
    return voxelize_launch(true, d_out, d_feats, rules, M, maxActive, C, 0, (cudaStream_t)stream);
}

// Not part of the reference's native module: the row gather its caller does in torch
// (clusters_feats = feats[c_idxs], model/pointgroup.py:133-134,333).  idx is int32 or int64.
extern "C" int pg_gather_rows(const float *src, const void *idx, int idx_is_int64, float *dst, int64_t nIdx, int32_t C,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PG_CHECK_ARG(nIdx >= 0 && C >= 0, "negative size");
    if (nIdx == 0 || C == 0) return PG_OK;
    PG_CHECK_ARG(src && idx && dst, "null pointer");
    const int V = pick_vec(src, dst, C);
    const int Cv = C / V;
    const int64_t total = nIdx * Cv;
    const unsigned grid = (unsigned)(div_up(total, 256) < (int64_t)kNumSM * 64 ? div_up(total, 256) : (int64_t)kNumSM * 64);
#define PG_GATHER(VV)                                                                                              \
    if (idx_is_int64) launch(k_gather_rows<VV, int64_t>, grid, 256, 0, st, src, (const int64_t *)idx, dst, nIdx, Cv); \
    else launch(k_gather_rows<VV, int32_t>, grid, 256, 0, st, src, (const int32_t *)idx, dst, nIdx, Cv)
    if (V == 4) { PG_GATHER(4); } else if (V == 2) { PG_GATHER(2); } else { PG_GATHER(1); }
#undef PG_GATHER
    PG_LAUNCH_CHECK();
    return PG_OK;
}
