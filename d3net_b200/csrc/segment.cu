// Segment reductions over proposals: roipool_fp / roipool_bp, sec_mean / sec_min / sec_max, get_iou.
// Reference behaviour: lib/pointgroup_ops/src/{roipool/roipool.cu, sec_mean/sec_mean.cu, get_iou/get_iou.cu}.
#include "common.cuh"

namespace pg {

// =================================================================================================
// max / min (+ argmax) over row segments.
//
// Work is cut by ROWS, not by proposals: one warp owns a tile of R consecutive rows and walks the
// pieces of the proposals that overlap it, so a 50-point proposal and a 300k-point floor cost the
// same per row and no warp is left with a giant segment.  A lane owns V consecutive channels of
// every G-th row (G = 32 / (C/V) row groups per warp-wide load, fully coalesced); the row groups are
// folded with shuffles and the piece result is merged into the proposal's slot with one atomic per
// channel.  Order-independence is what makes this legal: max with "lowest row wins ties"
// (roipool.cu:22-26) is a max over the key (value, -row), min/max over values are plain lattices.
// =================================================================================================
enum SegMode { kMaxArg = 0, kMax = 1, kMin = 2 };

__device__ __forceinline__ unsigned long long pack_key(float v, int row) {
    return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)row);
}

template <int MODE>
__global__ void k_seg_init(void *slots, int64_t n) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == kMaxArg) ((unsigned long long *)slots)[i] = pack_key(-INFINITY, -1);
        else if (MODE == kMax) ((unsigned *)slots)[i] = f2ord(-INFINITY);
        else ((unsigned *)slots)[i] = f2ord(INFINITY);
    }
}

template <int MODE>
__global__ void k_seg_decode(void *slots, float *__restrict__ out, int32_t *__restrict__ maxidx, int64_t n) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == kMaxArg) {
            unsigned long long k = ((const unsigned long long *)slots)[i];
            out[i] = ord2f((unsigned)(k >> 32));
            maxidx[i] = (int)(0xffffffffu - (unsigned)(k & 0xffffffffu));
        } else {
            out[i] = ord2f(((const unsigned *)slots)[i]);   // in place: slots == out
        }
    }
}

template <int MODE>
__device__ __forceinline__ void seg_acc(float x, int row, float &best, int &arg) {
    if (MODE == kMin) { if (x < best) best = x; }
    else if (x > best) { best = x; arg = row; }
}
// merge (v2, a2) into (best, arg) where the two come from different rows
template <int MODE>
__device__ __forceinline__ void seg_merge(float v2, int a2, float &best, int &arg) {
    if (MODE == kMin) { if (v2 < best) best = v2; }
    else if (v2 > best || (MODE == kMaxArg && v2 == best && (unsigned)a2 < (unsigned)arg)) { best = v2; arg = a2; }
}

template <int MODE, int V>
__global__ void __launch_bounds__(256) k_seg_reduce(const float *__restrict__ inp, const int32_t *__restrict__ offsets,
                                                    void *slots, int32_t nRows, int32_t nP, int32_t C, int32_t R) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t r0 = tile * R, r1 = r0 + R;
    const int first = ld_after_wait(offsets), last = ld_after_wait(offsets + nP);
    if (r0 < first) r0 = first;
    if (r1 > last) r1 = last;
    if (r1 > nRows) r1 = nRows;
    if (r0 >= r1) return;
    // proposal containing r0: largest p with offsets[p] <= r0
    int lo = 0, hi = nP;   // invariant: offsets[lo] <= r0 < offsets[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= r0) lo = mid; else hi = mid;
    }
    const int CV = C / V;                       // vector columns per row
    const int G = CV >= 32 ? 1 : 32 / CV;       // row groups per warp-wide load
    const int g = CV >= 32 ? 0 : lane / CV;     // this lane's row group
    const int c0 = CV >= 32 ? lane : lane - g * CV;
    const bool lane_on = CV >= 32 ? true : (g < G);
    for (int p = lo; p < nP; p++) {
        const int64_t ps = __ldg(offsets + p), pe = __ldg(offsets + p + 1);
        if (ps >= r1) break;
        const int64_t a = ps > r0 ? ps : r0, b = pe < r1 ? pe : r1;
        if (a >= b) continue;
        for (int cv = c0; cv < CV; cv += 32) {   // one trip unless C/V > 32
            float best[V];
            int arg[V];
#pragma unroll
            for (int k = 0; k < V; k++) { best[k] = (MODE == kMin) ? INFINITY : -INFINITY; arg[k] = -1; }
            if (lane_on) {
                int64_t row = a + g;
                for (; row + 3 * G < b; row += 4 * G) {   // four loads in flight per lane
                    float x[4][V];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float *q = inp + (row + (int64_t)u * G) * C + (int64_t)cv * V;
                        if constexpr (V == 4) { float4 t = __ldg((const float4 *)q); x[u][0] = t.x; x[u][1] = t.y; x[u][2] = t.z; x[u][3] = t.w; }
                        else if constexpr (V == 2) { float2 t = __ldg((const float2 *)q); x[u][0] = t.x; x[u][1] = t.y; }
                        else x[u][0] = __ldg(q);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int k = 0; k < V; k++) seg_acc<MODE>(x[u][k], (int)(row + (int64_t)u * G), best[k], arg[k]);
                }
                for (; row < b; row += G) {
                    const float *q = inp + row * C + (int64_t)cv * V;
#pragma unroll
                    for (int k = 0; k < V; k++) seg_acc<MODE>(__ldg(q + k), (int)row, best[k], arg[k]);
                }
            }
            // fold the row groups: lane (g, c0) <- lanes (g + s, c0)
            if (CV < 32) {
                for (int s = 1; s < G; s <<= 1) {
#pragma unroll
                    for (int k = 0; k < V; k++) {
                        const int src = lane + s * CV;
                        float v2 = __shfl_sync(0xffffffffu, best[k], src & 31);
                        int a2 = __shfl_sync(0xffffffffu, arg[k], src & 31);
                        if (lane_on && src < G * CV && (g % (2 * s)) == 0 && g + s < G) seg_merge<MODE>(v2, a2, best[k], arg[k]);
                    }
                }
            }
            if (lane_on && g == 0) {
#pragma unroll
                for (int k = 0; k < V; k++) {
                    const int64_t slot = (int64_t)p * C + (int64_t)cv * V + k;
                    if (MODE == kMaxArg) {
                        if (arg[k] >= 0) atomicMax((unsigned long long *)slots + slot, pack_key(best[k], arg[k]));
                    } else if (MODE == kMax) {
                        atomicMax((unsigned *)slots + slot, f2ord(best[k]));
                    } else {
                        atomicMin((unsigned *)slots + slot, f2ord(best[k]));
                    }
                }
            }
        }
    }
}

template <int MODE>
static int seg_reduce_launch(const float *inp, const int32_t *offsets, float *out, int32_t *maxidx, void *slots,
                             int32_t nRows, int32_t nP, int32_t C, cudaStream_t st) {
    PG_CHECK_ARG(nRows >= 0 && nP >= 0 && C >= 0, "negative size");
    const int64_t nslot = (int64_t)nP * C;
    if (nslot == 0) return PG_OK;
    PG_CHECK_ARG(offsets && out && slots && (inp || nRows == 0), "null pointer");
    const unsigned ig = (unsigned)(div_up(nslot, 256) < kNumSM * 8 ? div_up(nslot, 256) : kNumSM * 8);
    launch(k_seg_init<MODE>, ig, 256, 0, st, slots, nslot);
    if (nRows > 0) {
        int V = 1;
        if (C % 4 == 0 && (uintptr_t)inp % 16 == 0) V = 4;
        else if (C % 2 == 0 && (uintptr_t)inp % 8 == 0) V = 2;
        int R = 4096 / (C > 0 ? C : 1);
        if (R < 32) R = 32;
        // ... but enough tiles for ~32 warps per SM: a lane's loop is a chain of memory round trips (four rows in flight), and
        // on narrow rows (C = 3: 1365 rows per tile, 1.4 k tiles for 1.9 M rows) 34 of them in a row on 8 warps per SM were
        // the kernel's run time (sec_max 0.070 ms for 23 MB)
        const int64_t spread = nRows / ((int64_t)kNumSM * 32);
        if (R > spread) R = (int)(spread > 128 ? spread : 128);
        const int64_t tiles = div_up(nRows, R);
        const unsigned grid = (unsigned)div_up(tiles, 8);
        if (V == 4) launch(k_seg_reduce<MODE, 4>, grid, 256, 0, st, inp, offsets, slots, nRows, nP, C, R);
        else if (V == 2) launch(k_seg_reduce<MODE, 2>, grid, 256, 0, st, inp, offsets, slots, nRows, nP, C, R);
        else launch(k_seg_reduce<MODE, 1>, grid, 256, 0, st, inp, offsets, slots, nRows, nP, C, R);
    }
    launch(k_seg_decode<MODE>, ig, 256, 0, st, slots, out, maxidx, nslot);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// roipool_bp: d_feats[argmax][c] += d_out[p][c]  (roipool.cu:42-49)
__global__ void k_roipool_bp(float *d_feats, const int32_t *__restrict__ maxidx, const float *__restrict__ d_out,
                             int64_t n, int32_t C) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = __ldg(maxidx + i);
        if (a >= 0) atomicAdd(d_feats + (int64_t)a * C + (i % C), __ldg(d_out + i));
    }
}

// =================================================================================================
// sec_mean: mean = sum_i fl(x_i / count), accumulated strictly left to right (sec_mean.cu:17-25).
// The serial fp32 add chain is the semantics, so the chain is all that stays serial: the block
// streams the segment through a double-buffered shared tile (coalesced loads in flight while the
// previous tile is consumed), applies the IEEE division in parallel on the way in, and threads
// 0..C-1 run one add chain per channel out of shared memory.
// =================================================================================================
constexpr int kMeanThreads = 128;
constexpr int kMeanPer = 16;                          // floats prefetched per thread
constexpr int kMeanTile = kMeanThreads * kMeanPer;    // 2048 floats per buffer

__global__ void __launch_bounds__(kMeanThreads) k_sec_mean(const float *__restrict__ inp,
                                                           const int32_t *__restrict__ offsets,
                                                           float *__restrict__ out, int32_t nP, int32_t C) {
    pdl_enter();
    __shared__ float buf[2][kMeanTile];
    const int tid = threadIdx.x;
    const int rowsPerTile = kMeanTile / C;            // C <= kMeanTile guaranteed by the launcher
    const int tileFloats = rowsPerTile * C;
    for (int p = blockIdx.x; p < nP; p += gridDim.x) {
        const int64_t start = __ldg(offsets + p), end = __ldg(offsets + p + 1);
        const int64_t len = end - start;
        const float count = (float)(int)len;
        const float *base = inp + start * C;
        const int64_t total = len > 0 ? len * C : 0;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};          // channels tid, tid+128, ... (C <= 512 on this path)
        float reg[kMeanPer];
        int64_t pos = 0;
        int cur = 0;
        // prefetch tile 0
#pragma unroll
        for (int k = 0; k < kMeanPer; k++) {
            const int64_t e = pos + k * kMeanThreads + tid;
            reg[k] = (k * kMeanThreads + tid < tileFloats && e < total) ? __ldg(base + e) : 0.f;
        }
        while (pos < total) {
            const int64_t nflt = (total - pos < tileFloats) ? (total - pos) : tileFloats;
#pragma unroll
            for (int k = 0; k < kMeanPer; k++) buf[cur][k * kMeanThreads + tid] = __fdiv_rn(reg[k], count);
            __syncthreads();
            const int64_t npos = pos + tileFloats;
            if (npos < total) {
#pragma unroll
                for (int k = 0; k < kMeanPer; k++) {
                    const int64_t e = npos + k * kMeanThreads + tid;
                    reg[k] = (k * kMeanThreads + tid < tileFloats && e < total) ? __ldg(base + e) : 0.f;
                }
            }
            const int rows = (int)(nflt / C);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = tid + j * kMeanThreads;
                if (c < C) {
                    const float *b = &buf[cur][c];
                    float a = acc[j];
                    // the add chain is the critical path (4 cycles per row): the next eight operands are
                    // fetched from shared memory while the current eight are being added
                    int r = 0;
                    float q[8], nx[8];
                    if (rows >= 8) {
#pragma unroll
                        for (int u = 0; u < 8; u++) q[u] = b[u * C];
                        for (; r + 16 <= rows; r += 8) {
#pragma unroll
                            for (int u = 0; u < 8; u++) nx[u] = b[(r + 8 + u) * C];
#pragma unroll
                            for (int u = 0; u < 8; u++) a = __fadd_rn(a, q[u]);
#pragma unroll
                            for (int u = 0; u < 8; u++) q[u] = nx[u];
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) a = __fadd_rn(a, q[u]);
                        r += 8;
                    }
                    for (; r < rows; r++) a = __fadd_rn(a, b[r * C]);
                    acc[j] = a;
                }
            }
            pos = npos;
            cur ^= 1;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = tid + j * kMeanThreads;
            if (c < C) out[(int64_t)p * C + c] = acc[j];
        }
        __syncthreads();   // the next proposal's first store must not overtake this one's last chain
    }
}

// Narrow rows (C <= 32: the xyz means of clusters_voxelization): warp-specialised.  Warp 0 does nothing but the add
// chains (lane c = channel c) -- the serial chain of the floor-sized proposal IS the op's run time, 4 cycles per row at
// best -- while warps 1..3 fetch, divide and stage the NEXT tile in the other buffer; the block meets once per tile.
// (With everybody dividing and then waiting for the chains: 5.5 cycles per row.)
constexpr int kMeanWsProducers = kMeanThreads - 32;               // 96 threads
constexpr int kMeanWsPer = 16;
constexpr int kMeanWsTile = kMeanWsProducers * kMeanWsPer;       // 1536 floats per buffer

// CT: the row width when it is known at compile time (3: the model's call; shared-memory offsets become immediates), 0 = C.
template <int CT>
__global__ void __launch_bounds__(kMeanThreads) k_sec_mean_narrow(const float *__restrict__ inp, const int32_t *__restrict__ offsets,
                                                                  float *__restrict__ out, int32_t nP, int32_t C_) {
    pdl_enter();
    const int C = CT ? CT : C_;
    __shared__ float buf[2][kMeanWsTile];
    const int tid = threadIdx.x;
    const bool chain = tid < 32;
    const int ptid = tid - 32;                           // producer index 0..95
    const int rowsPerTile = kMeanWsTile / C;
    const int tileFloats = rowsPerTile * C;
    for (int p = blockIdx.x; p < nP; p += gridDim.x) {
        const int64_t start = __ldg(offsets + p), end = __ldg(offsets + p + 1);
        const int64_t len = end - start;
        const float count = (float)(int)len;
        const float *base = inp + start * C;
        const int64_t total = len > 0 ? len * C : 0;
        float a = 0.f;
        float reg[kMeanWsPer];
        auto fetch = [&](int64_t pos) {
#pragma unroll
            for (int k = 0; k < kMeanWsPer; k++) {
                const int64_t e = pos + k * kMeanWsProducers + ptid;
                reg[k] = (k * kMeanWsProducers + ptid < tileFloats && e < total) ? __ldg(base + e) : 0.f;
            }
        };
        auto stage = [&](int b) {
#pragma unroll
            for (int k = 0; k < kMeanWsPer; k++) buf[b][k * kMeanWsProducers + ptid] = __fdiv_rn(reg[k], count);
        };
        // tile 0 into buffer 0, tile 1 into the producers' registers
        if (!chain && total > 0) {
            fetch(0);
            stage(0);
            if (tileFloats < total) fetch(tileFloats);
        }
        __syncthreads();
        int cur = 0;
        for (int64_t pos = 0; pos < total; pos += tileFloats) {
            const int64_t nflt = (total - pos < tileFloats) ? (total - pos) : tileFloats;
            if (chain) {
                if (tid < C) {
                    const int rows = (int)(nflt / C);
                    const float *b = &buf[cur][tid];
                    // A warp issues in order and every add waits 4 cycles for the one before it: the loop holds nothing but
                    // the adds and, in their shadow, the shared-memory loads of the operands 8 rows ahead (two register
                    // sets, no moves).
                    int r = 0;
                    float q[8], nx[8];
                    if (rows >= 8) {
#pragma unroll
                        for (int u = 0; u < 8; u++) q[u] = b[u * C];
                        for (; r + 24 <= rows; r += 16) {
                            const float *bb = b + r * C;
#pragma unroll
                            for (int u = 0; u < 8; u++) { a = __fadd_rn(a, q[u]); nx[u] = bb[(8 + u) * C]; }
#pragma unroll
                            for (int u = 0; u < 8; u++) { a = __fadd_rn(a, nx[u]); q[u] = bb[(16 + u) * C]; }
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) a = __fadd_rn(a, q[u]);
                        r += 8;
                    }
                    for (; r < rows; r++) a = __fadd_rn(a, b[r * C]);
                }
            } else if (pos + tileFloats < total) {
                stage(cur ^ 1);                                              // tile t + 1, fetched one round ago
                if (pos + 2 * (int64_t)tileFloats < total) fetch(pos + 2 * (int64_t)tileFloats);
            }
            __syncthreads();
            cur ^= 1;
        }
        if (chain && tid < C) out[(int64_t)p * C + tid] = a;
    }
}

// any C: one thread per (proposal, channel), straight from global memory
__global__ void k_sec_mean_wide(const float *__restrict__ inp, const int32_t *__restrict__ offsets,
                                float *__restrict__ out, int32_t nP, int32_t C) {
    pdl_enter();
    const int64_t n = (int64_t)nP * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(t / C), c = (int)(t - (int64_t)p * C);
        const int64_t start = __ldg(offsets + p), end = __ldg(offsets + p + 1);
        const float count = (float)(int)(end - start);
        float a = 0.f;
        for (int64_t i = start; i < end; i++) a = __fadd_rn(a, __fdiv_rn(__ldg(inp + i * C + c), count));
        out[t] = a;
    }
}

// =================================================================================================
// get_iou: the intersection counts are a sparse histogram over (proposal, instance) -- one pass over the proposals' points
// (the reference rescans every proposal once per instance, get_iou.cu:19-25).  Work is cut by ROWS, not by proposals, so
// the floor-sized proposal is spread over the whole grid; a warp merges equal (proposal, instance) keys before it touches
// memory (a proposal's points mostly carry one instance).  The counts are accumulated as int32 in the output buffer
// itself (same 4 bytes per entry) and turned into IoUs in place with the reference's mixed fp32/fp64 formula (:26).
// =================================================================================================
__global__ void __launch_bounds__(256) k_iou_count(const int32_t *__restrict__ pidx, const int32_t *__restrict__ poff,
                                                   const int64_t *__restrict__ labels, int32_t *counts, int32_t nInst,
                                                   int32_t nP) {
    pdl_enter();
    const int S = __ldg(poff + nP);                                   // rows of proposals_idx (known on the device only)
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < S; base += stride) {   // warp-uniform trips
        const int64_t i = base + lane;
        long long key = -1;
        if (i < S) {
            const int l = (int)__ldg(labels + __ldg(pidx + i));        // (int) narrowing as in :22
            if (l >= 0 && l < nInst) {
                int lo = 0, hi = nP;                                  // the proposal that owns row i: poff[p] <= i < poff[p + 1]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(poff + mid) <= (int)i) lo = mid; else hi = mid;
                }
                key = (long long)lo * nInst + l;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (peers & lanemask_lt()) == 0) atomicAdd(counts + key, __popc(peers));
    }
}

__global__ void __launch_bounds__(256) k_iou_final(const int32_t *__restrict__ poff, const int32_t *__restrict__ pointnum,
                                                   float *iou, int32_t nInst, int64_t total) {
    pdl_enter();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int p = (int)(t / nInst), g = (int)(t - (int64_t)p * nInst);
    const int inter = reinterpret_cast<const int32_t *>(iou)[t];
    const int ptotal = __ldg(poff + p + 1) - __ldg(poff + p);
    const int itotal = __ldg(pointnum + g);
    const double den = (double)(float)(ptotal + itotal - inter) + 1e-5;
    iou[t] = (float)((double)(float)inter / den);
}

}  // namespace pg

using namespace pg;

extern "C" size_t pg_roipool_workspace_bytes(int32_t nProposal, int32_t C) {
    if (nProposal < 0 || C < 0) return 0;
    return (size_t)nProposal * (size_t)C * 8 + 256;
}

extern "C" int pg_roipool_fp(const float *feats, const int32_t *offsets, float *out, int32_t *maxidx, int32_t nRows,
                             int32_t nProposal, int32_t C, void *ws, size_t ws_bytes, void *stream) {
    if ((int64_t)nProposal * C > 0) {
        PG_CHECK_ARG(maxidx, "null maxidx");
        if (!ws || ws_bytes < (size_t)nProposal * C * 8) { set_error("pg_roipool_fp: workspace too small"); return PG_EWORKSPACE; }
        PG_CHECK_ARG((uintptr_t)ws % 8 == 0, "workspace must be 8-byte aligned");
    }
    return seg_reduce_launch<kMaxArg>(feats, offsets, out, maxidx, ws, nRows, nProposal, C, (cudaStream_t)stream);
}

extern "C" int pg_roipool_bp(float *d_feats, const int32_t *offsets, const int32_t *maxidx, const float *d_out,
                             int32_t nProposal, int32_t C, void *stream) {
    (void)offsets;
    PG_CHECK_ARG(nProposal >= 0 && C >= 0, "negative size");
    const int64_t n = (int64_t)nProposal * C;
    if (n == 0) return PG_OK;
    PG_CHECK_ARG(d_feats && maxidx && d_out, "null pointer");
    const unsigned grid = (unsigned)(div_up(n, 256) < kNumSM * 16 ? div_up(n, 256) : kNumSM * 16);
    launch(k_roipool_bp, grid, 256, 0, (cudaStream_t)stream, d_feats, maxidx, d_out, n, C);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_sec_max(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
                          int32_t C, void *stream) {
    return seg_reduce_launch<kMax>(inp, offsets, out, nullptr, out, nRows, nProposal, C, (cudaStream_t)stream);
}
extern "C" int pg_sec_min(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
                          int32_t C, void *stream) {
    return seg_reduce_launch<kMin>(inp, offsets, out, nullptr, out, nRows, nProposal, C, (cudaStream_t)stream);
}

extern "C" int pg_sec_mean(const float *inp, const int32_t *offsets, float *out, int32_t nRows, int32_t nProposal,
                           int32_t C, void *stream) {
    (void)nRows;
    PG_CHECK_ARG(nProposal >= 0 && C >= 0, "negative size");
    if ((int64_t)nProposal * C == 0) return PG_OK;
    PG_CHECK_ARG(offsets && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    PG_KTIME("k_sec_mean", st);
    if (C <= 32) {
        const unsigned grid = (unsigned)(nProposal < kNumSM * 8 ? nProposal : kNumSM * 8);
        if (C == 3) launch(k_sec_mean_narrow<3>, grid, kMeanThreads, 0, st, inp, offsets, out, nProposal, C);
        else launch(k_sec_mean_narrow<0>, grid, kMeanThreads, 0, st, inp, offsets, out, nProposal, C);
    } else if (C <= 4 * kMeanThreads) {
        const unsigned grid = (unsigned)(nProposal < kNumSM * 8 ? nProposal : kNumSM * 8);
        launch(k_sec_mean, grid, kMeanThreads, 0, st, inp, offsets, out, nProposal, C);
    } else {
        const int64_t n = (int64_t)nProposal * C;
        launch(k_sec_mean_wide, (unsigned)div_up(n, 128), 128, 0, st, inp, offsets, out, nProposal, C);
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_get_iou(const int32_t *proposals_idx, const int32_t *proposals_offset,
                          const int64_t *instance_labels, const int32_t *instance_pointnum, float *proposals_iou,
                          int32_t nInstance, int32_t nProposal, void *stream) {
    PG_CHECK_ARG(nInstance >= 0 && nProposal >= 0, "negative size");
    if ((int64_t)nInstance * nProposal == 0) return PG_OK;
    PG_CHECK_ARG(proposals_offset && instance_pointnum && proposals_iou, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)nInstance * nProposal;
    PG_CHECK_ARG(proposals_idx && instance_labels, "null pointer");
    PG_TRY(fill_u32(proposals_iou, 0u, (size_t)total, st));
    { PG_KTIME("k_iou_count", st);
    launch(k_iou_count, kNumSM * 8, 256, 0, st, proposals_idx, proposals_offset, instance_labels,
                                           reinterpret_cast<int32_t *>(proposals_iou), nInstance, nProposal); }
    launch(k_iou_final, (unsigned)div_up(total, 256), 256, 0, st, proposals_offset, instance_pointnum, proposals_iou, nInstance, total);
    PG_LAUNCH_CHECK();
    return PG_OK;
}
