// Single-pass exclusive scan (decoupled look-back) with the producer and the consumer of the scanned values fused in:
//   load(i)                    -> the int32 value of element i (computed on the fly: a flag, a size, ...)
//   store(i, exclusive, value) -> whatever the caller does with element i's prefix (write it, publish an id, ...)
// so that "compute flags -> scan -> act on the ranks" is ONE launch instead of three and the flags never touch memory.
// prim.cu's scan_exclusive_i32 is this kernel with plain array functors.
#pragma once
#include "common.cuh"

namespace pg {

constexpr int kScanThreads = 256;
constexpr int kScanRounds = 16;
constexpr int kScanTile = kScanThreads * kScanRounds;

// inclusive block scan of one int per thread; returns inclusive value, *block_total = sum over block
__device__ __forceinline__ int block_scan_incl(int v, int *warp_tot /*[32] smem*/, int *block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    int before = 0, all = 0;
    for (int i = 0; i < nw; i++) {
        int t = warp_tot[i];
        if (i < w) before += t;
        all += t;
    }
    __syncthreads();
    *block_total = all;
    return v + before;
}

// One 64-bit word per tile: [63:62] 0 = nothing yet, 1 = aggregate, 2 = inclusive prefix; [61:0] value.
// tmp[0] is the ticket counter, tmp[1 + t] tile t's word; both zero at launch.
constexpr unsigned long long kSpValueMask = (1ULL << 62) - 1;

// A block takes the next tile (ticket from an atomic counter, so every earlier tile is already running or done), scans
// it, publishes its aggregate, and warp 0 walks back over the predecessors' published words -- 32 at a time -- until it
// meets an inclusive prefix.  Returns the tile's exclusive prefix to every thread; *tile_out = the tile index.
__device__ __forceinline__ long long scan_tile_prefix(unsigned long long *tmp, int tot, int64_t *tile_out, long long *s_tile,
                                                      long long *s_prefix) {
    const int64_t t = *s_tile;
    volatile unsigned long long *state = tmp + 1;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        long long prefix = 0;
        if (t == 0) {
            if (lane == 0) state[0] = (2ULL << 62) | (unsigned long long)(long long)tot;
        } else {
            if (lane == 0) state[t] = (1ULL << 62) | ((unsigned long long)(long long)tot & kSpValueMask);
            int64_t hi = t - 1;                       // the newest predecessor not yet added
            for (;;) {
                const int64_t k = hi - lane;
                unsigned long long w = (2ULL << 62);  // below tile 0: an inclusive prefix of 0
                if (k >= 0) w = state[k];
                const unsigned flag = (unsigned)(w >> 62);
                const unsigned empty = __ballot_sync(0xffffffffu, flag == 0u);
                const unsigned incl_m = __ballot_sync(0xffffffffu, flag == 2u);
                // usable lanes: from lane 0 up to the first inclusive word, all of them published
                const int stop = incl_m ? __ffs((int)incl_m) - 1 : 31;
                if (empty & ((stop == 31 ? 0xffffffffu : ((2u << stop) - 1u)))) continue;   // somebody in range is not there yet
                long long val = (lane <= stop) ? (long long)(w & kSpValueMask) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (incl_m) break;
                hi -= 32;
            }
            if (lane == 0) state[t] = (2ULL << 62) | ((unsigned long long)(prefix + tot) & kSpValueMask);
        }
        if (lane == 0) *s_prefix = prefix;
    }
    __syncthreads();
    *tile_out = t;
    return *s_prefix;
}

// Elements are loaded and stored STRIPED (consecutive threads, consecutive elements: the functors' first-level accesses
// coalesce) and scanned BLOCKED (16 consecutive elements per thread); the two layouts meet in shared memory, padded by
// one word per 32 so that neither access pattern has bank conflicts.
__device__ __forceinline__ int scan_pad(int i) { return i + (i >> 5); }

// Tiles of 2048 (8 per thread): the functors' gathers want many blocks in flight more than they want long tiles
// (4096-element tiles: 22 % occupancy at 33 KB of shared memory and 74-104 registers, ~21 us per million elements).
constexpr int kFusedRounds = 8;
constexpr int kFusedTile = kScanThreads * kFusedRounds;

template <typename Load, typename Store>
__global__ void __launch_bounds__(kScanThreads) k_scan_fused(Load load, Store store, int64_t n, unsigned long long *tmp,
                                                             int64_t *__restrict__ total) {
    pdl_enter();
    __shared__ int warp_tot[32];
    __shared__ long long s_tile, s_prefix;
    __shared__ int s_val[kFusedTile + kFusedTile / 32], s_pre[kFusedTile + kFusedTile / 32];
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(tmp, 1ULL);
    __syncthreads();
    const int64_t tile_base = (int64_t)s_tile * kFusedTile;
#pragma unroll
    for (int k = 0; k < kFusedRounds; k++) {
        const int e = k * kScanThreads + threadIdx.x;
        s_val[scan_pad(e)] = (tile_base + e < n) ? load(tile_base + e) : 0;
    }
    __syncthreads();
    int v[kFusedRounds];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kFusedRounds; k++) { v[k] = tsum; tsum += s_val[scan_pad(threadIdx.x * kFusedRounds + k)]; }
    int tot;
    const int incl = block_scan_incl(tsum, warp_tot, &tot);
    int64_t t;
    const long long prefix = scan_tile_prefix(tmp, tot, &t, &s_tile, &s_prefix);
    if (total && threadIdx.x == 0 && (t + 1) * (int64_t)kFusedTile >= n) *total = prefix + tot;
    const int off = (int)prefix + incl - tsum;
#pragma unroll
    for (int k = 0; k < kFusedRounds; k++) s_pre[scan_pad(threadIdx.x * kFusedRounds + k)] = off + v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFusedRounds; k++) {
        const int e = k * kScanThreads + threadIdx.x;
        if (tile_base + e < n) store(tile_base + e, s_pre[scan_pad(e)], s_val[scan_pad(e)]);
    }
}

// `tmp` holds scan_tmp_count(n) int64 values; `tmp_is_zero`: the caller has zeroed them already on this stream (one
// memset for all the scans of an op instead of one each).
template <typename Load, typename Store>
int scan_fused(Load load, Store store, int64_t n, int64_t *total, int64_t *tmp, cudaStream_t st, bool tmp_is_zero = false) {
    if (n <= 0) {
        if (total) PG_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    const int64_t nb = div_up(n, kFusedTile);
    if (!tmp_is_zero) PG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)(nb + 1) * sizeof(int64_t), st));
    launch(k_scan_fused<Load, Store>, (unsigned)nb, kScanThreads, 0, st, load, store, n, reinterpret_cast<unsigned long long *>(tmp), total);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

}  // namespace pg
