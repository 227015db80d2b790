// Device-wide primitives used by several ops: exclusive scan, stable LSD radix sort of pairs, and
// hash grouping of int4 keys (first-occurrence numbering).  Hand-written; no CUB/Thrust.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace pg {

// ---------------------------------------------------------------------------------------------
// error string
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

bool pdl_enabled() {
    static const bool on = []() { const char *e = getenv("PG_B200_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of a plain array (single pass, decoupled look-back: scan.cuh), 16 items per thread through int4 loads
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) k_scan_onepass(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                               int64_t n, unsigned long long *tmp, int64_t *__restrict__ total,
                                                               int aligned) {
    pdl_enter();
    __shared__ int warp_tot[32];
    __shared__ long long s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(tmp, 1ULL);
    __syncthreads();
    const int64_t base = (int64_t)s_tile * kScanTile + (int64_t)threadIdx.x * kScanRounds;
    int v[kScanRounds];
    const bool full = aligned && base + kScanRounds <= n;
    if (full) {
        const int4 *p = reinterpret_cast<const int4 *>(in + base);
#pragma unroll
        for (int k = 0; k < kScanRounds / 4; k++) {
            const int4 q = __ldg(p + k);
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanRounds; k++) v[k] = (base + k < n) ? in[base + k] : 0;
    }
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanRounds; k++) { const int q = v[k]; v[k] = tsum; tsum += q; }
    int tot;
    const int incl = block_scan_incl(tsum, warp_tot, &tot);
    int64_t t;
    const long long prefix = scan_tile_prefix(tmp, tot, &t, &s_tile, &s_prefix);
    if (total && threadIdx.x == 0 && (t + 1) * (int64_t)kScanTile >= n) *total = prefix + tot;
    const int off = (int)prefix + incl - tsum;
    if (full) {
        int4 *q = reinterpret_cast<int4 *>(out + base);
#pragma unroll
        for (int k = 0; k < kScanRounds / 4; k++)
            q[k] = make_int4(off + v[4 * k], off + v[4 * k + 1], off + v[4 * k + 2], off + v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanRounds; k++)
            if (base + k < n) out[base + k] = off + v[k];
    }
}

size_t scan_tmp_count(int64_t n) { return (size_t)div_up(n > 0 ? n : 1, kFusedTile) + 2; }   // the smaller of the two tile sizes

int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, int64_t *total, int64_t *tmp,
                       cudaStream_t st) {
    if (n <= 0) {
        if (total) PG_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    const int64_t nb = div_up(n, kScanTile);
    const bool aligned = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    PG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)(nb + 1) * sizeof(int64_t), st));
    launch(k_scan_onepass, (unsigned)nb, kScanThreads, 0, st, in, out, n, reinterpret_cast<unsigned long long *>(tmp), total,
                                                         aligned ? 1 : 0);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, one kernel per pass ("onesweep"): the digit histograms of ALL passes are taken in
// one sweep over the keys up front; a pass then ranks its tile, publishes the tile's 256 digit counts, and finds its
// global offsets by decoupled look-back over the earlier tiles' published counts (chained scan, one thread per digit)
// instead of a histogram kernel + a device-wide scan per pass.
// ---------------------------------------------------------------------------------------------
constexpr int kRadixThreads = 256;
constexpr int kRadixRounds = 8;
constexpr int kRadixTile = kRadixThreads * kRadixRounds;
constexpr int kRadixMaxPasses = 4;
constexpr int kRadixMaxBins = 1024;         // digits of 8, 9 or 10 bits: 9 .. 10 / 17 .. 20 / 25 .. 30 key bits take one pass less
constexpr int kRadixPassBins = 3072;        // most bins over all passes of one sort: 3 x 1024 (4 passes only happen at 8 bits)
// scratch layout (int32 words): [0, 3072) global digit counts, pass after pass; [3072, 3072 + 16) tile tickets per pass;
// then per pass nb * bins look-back words: [31:30] 0 = nothing yet, 1 = the tile's count, 2 = inclusive prefix; [29:0] value
constexpr int kRadixHead = kRadixPassBins + 16;

__global__ void __launch_bounds__(kRadixThreads) k_radix_hist_all(const uint32_t *__restrict__ keys, int64_t n, int passes, int dbits,
                                                                  int32_t *__restrict__ scratch, int64_t state_words) {
    pdl_enter();
    __shared__ int h[kRadixPassBins];
    const int bins = 1 << dbits;
    for (int i = threadIdx.x; i < passes * bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&h[p * bins + ((k >> (dbits * p)) & (bins - 1))], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * bins; i += blockDim.x)
        if (h[i]) atomicAdd(&scratch[i], h[i]);
    // the passes' look-back words start from zero
    int32_t *state = scratch + kRadixHead;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < state_words; i += (int64_t)gridDim.x * blockDim.x) state[i] = 0;
}

// Each warp owns a contiguous 256-element slice of the tile (8 rounds of 32 consecutive elements), so the stable
// order inside the tile is (warp, round, lane): a warp ranks its own slice with nothing but warp-level
// primitives -- `__match_any_sync` groups equal digits, a per-warp counter row in shared memory carries the
// running count from round to round.  Keys, values and ranks stay in registers.  Thread t owns digits
// t * PER .. t * PER + PER - 1 in the publish / look-back step.
template <int DBITS>
__global__ void __launch_bounds__(kRadixThreads)    // (capped at 64 registers for 4 blocks per SM: spills, 35 -> 48 us per pass)
k_radix_onesweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                 uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                 const int32_t *__restrict__ ghist, unsigned *ticket, unsigned *state, int2 *__restrict__ pairs_out,
                 int64_t pairs_n) {
    pdl_enter();
    constexpr int kWarps = kRadixThreads / 32;
    constexpr int BINS = 1 << DBITS, PER = BINS / kRadixThreads;
    __shared__ int wcnt[kWarps][BINS];                // per-warp digit counts, then global bases per warp
    __shared__ int s_scan[32];
    __shared__ unsigned s_tile;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kWarps * BINS; i += kRadixThreads) (&wcnt[0][0])[i] = 0;
    // exclusive scan of the pass's global digit counts: where each digit's run starts in the output
    int gcount[PER], gsum = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) { gcount[q] = ghist[threadIdx.x * PER + q]; gsum += gcount[q]; }
    int unused;
    int gstart = block_scan_incl(gsum, s_scan, &unused) - gsum;          // (two block barriers: s_tile is visible after them)
    const unsigned tile = s_tile;
    const int64_t slice = (int64_t)tile * kRadixTile + (int64_t)w * (kRadixRounds * 32);
    uint32_t key[kRadixRounds], val[kRadixRounds];
    int rank[kRadixRounds];                           // position inside the warp's slice among equal digits
    unsigned dig[kRadixRounds];
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        const int64_t i = slice + r * 32 + lane;
        const bool live = i < n;
        key[r] = live ? keys_in[i] : 0u;
        val[r] = live ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
        dig[r] = live ? ((key[r] >> shift) & (unsigned)(BINS - 1)) : (unsigned)BINS;   // dead lanes never match a real digit
    }
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        const int before = __popc(peers & lt);
        int run = 0;
        if (dig[r] < (unsigned)BINS) {
            run = wcnt[w][dig[r]];                    // equal digits read the same counter, then the leader bumps it
        }
        __syncwarp();
        if (dig[r] < (unsigned)BINS && before == 0) wcnt[w][dig[r]] = run + __popc(peers);
        __syncwarp();
        rank[r] = run + before;
    }
    __syncthreads();
    // per digit: the tile's count goes out, the earlier tiles' counts come in, the eight warp counts become bases
    int c[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int d = threadIdx.x * PER + q;
        c[q] = 0;
#pragma unroll
        for (int k = 0; k < kWarps; k++) c[q] += wcnt[k][d];
        reinterpret_cast<volatile unsigned *>(state)[(size_t)tile * BINS + d] = (tile == 0 ? (2u << 30) : (1u << 30)) | (unsigned)c[q];
    }
    int before[PER];
    if (tile != 0) {
        // the thread's PER look-back chains advance together: their loads are in flight at the same time
        volatile unsigned *st = state + threadIdx.x * PER;
        int64_t at[PER];
        bool open = false;
#pragma unroll
        for (int q = 0; q < PER; q++) { before[q] = 0; at[q] = (int64_t)tile - 1; open = true; }
        while (open) {
            unsigned wv[PER];
            bool together = true;                    // the usual case: all chains at the same tile -> one vector load
#pragma unroll
            for (int q = 1; q < PER; q++) together &= at[q] == at[0];
            if (PER == 4 && together && at[0] >= 0) {
                const unsigned *p = state + (size_t)at[0] * BINS + threadIdx.x * PER;
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3]) : "l"(p));
            } else if (PER == 2 && together && at[0] >= 0) {
                const unsigned *p = state + (size_t)at[0] * BINS + threadIdx.x * PER;
                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(wv[0]), "=r"(wv[1]) : "l"(p));
            } else {
#pragma unroll
                for (int q = 0; q < PER; q++) wv[q] = at[q] >= 0 ? st[(size_t)at[q] * BINS + q] : 0u;
            }
            open = false;
#pragma unroll
            for (int q = 0; q < PER; q++) {
                if (at[q] < 0) continue;
                const unsigned flag = wv[q] >> 30;
                if (flag != 0u) {
                    before[q] += (int)(wv[q] & 0x3fffffffu);
                    at[q] = flag == 2u ? -1 : at[q] - 1;
                    if (at[q] < 0) st[(size_t)tile * BINS + q] = (2u << 30) | (unsigned)(before[q] + c[q]);
                }
                open |= at[q] >= 0;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < PER; q++) before[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int d = threadIdx.x * PER + q;
        int acc = gstart + before[q];
        gstart += gcount[q];
#pragma unroll
        for (int k = 0; k < kWarps; k++) {
            const int cw = wcnt[k][d];
            wcnt[k][d] = acc;
            acc += cw;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        if (dig[r] < (unsigned)BINS) {
            const int pos = wcnt[w][dig[r]] + rank[r];
            if (pairs_out) {                          // last pass of a sort whose caller wants (key, value) rows
                if (pos < pairs_n) pairs_out[pos] = make_int2((int)key[r], (int)val[r]);
            } else {
                keys_out[pos] = key[r];
                vals_out[pos] = val[r];
            }
        }
    }
}

size_t radix_tmp_count(int64_t n) { return (size_t)kRadixHead + (size_t)kRadixPassBins * (size_t)div_up(n > 0 ? n : 1, kRadixTile); }

int radix_sort_pairs(const uint32_t *keys_src, const uint32_t *vals_src, uint32_t *keysA, uint32_t *valsA,
                     uint32_t *keysB, uint32_t *valsB, int64_t n, int bits, int32_t *hist, int64_t *scan_tmp,
                     cudaStream_t st, int *result_buf, int2 *pairs_out, int64_t pairs_n) {
    (void)scan_tmp;
    *result_buf = 0;
    if (n <= 0) return PG_OK;
    if (bits < 1) bits = 1;
    if (bits > 32) bits = 32;
    // the fewest passes with digits of at most 10 bits, then the narrowest digit (>= 8 bits) that still makes it
    static const int max_dbits = []() { const char *e = getenv("PG_RADIX_MAXBITS"); const int v = e ? atoi(e) : 9; return v < 8 ? 8 : v > 10 ? 10 : v; }();
    const int passes = (bits + max_dbits - 1) / max_dbits;
    int dbits = (bits + passes - 1) / passes;
    if (dbits < 8) dbits = 8;
    const int bins = 1 << dbits;
    const int nb = (int)div_up(n, kRadixTile);
    uint32_t *k[2] = {keysA, keysB}, *v[2] = {valsA, valsB};
    const uint32_t *kin = keys_src, *vin = vals_src;
    PG_CUDA(cudaMemsetAsync(hist, 0, (size_t)kRadixHead * sizeof(int32_t), st));
    const int64_t state_words = (int64_t)passes * bins * nb;
    const int hgrid = nb < kNumSM * 8 ? nb : kNumSM * 8;
    launch(k_radix_hist_all, hgrid, kRadixThreads, 0, st, keys_src, n, passes, dbits, hist, state_words);
    int dst = 0;
    for (int p = 0; p < passes; p++) {
        unsigned *ticket = reinterpret_cast<unsigned *>(hist) + kRadixPassBins + p;
        unsigned *state = reinterpret_cast<unsigned *>(hist) + kRadixHead + (size_t)p * bins * nb;
        const int32_t *gh = hist + p * bins;
        int2 *po = p == passes - 1 ? pairs_out : nullptr;      // the last pass may write (key, value) rows instead
        if (dbits == 8) launch(k_radix_onesweep<8>, nb, kRadixThreads, 0, st, kin, vin, k[dst], v[dst], n, dbits * p, gh, ticket, state, po, pairs_n);
        else if (dbits == 9) launch(k_radix_onesweep<9>, nb, kRadixThreads, 0, st, kin, vin, k[dst], v[dst], n, dbits * p, gh, ticket, state, po, pairs_n);
        else launch(k_radix_onesweep<10>, nb, kRadixThreads, 0, st, kin, vin, k[dst], v[dst], n, dbits * p, gh, ticket, state, po, pairs_n);
        kin = k[dst];
        vin = v[dst];
        *result_buf = dst;
        dst ^= 1;
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------
// hash grouping of int4 keys
// ---------------------------------------------------------------------------------------------
uint32_t group_table_cap(int64_t n) {
    uint64_t c = 1024;
    while (c < (uint64_t)(n > 0 ? n : 1) * 2) c <<= 1;
    return (uint32_t)c;
}

// Large fills run as a kernel: the driver may hand a big cudaMemsetAsync to a copy engine, where it queues behind an
// upload in flight on another stream (measured: the end-to-end loop lost its copy / compute overlap, 12.6 -> 15.3 ms).
__global__ void k_fill_u32(uint32_t *__restrict__ p, uint32_t v, size_t count) {
    pdl_enter();
    const size_t n16 = count / 4;
    const uint4 q = make_uint4(v, v, v, v);
    uint4 *p16 = reinterpret_cast<uint4 *>(p);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p16[i] = q;
    if (blockIdx.x == 0 && threadIdx.x < (count & 3)) p[n16 * 4 + threadIdx.x] = v;
}

int fill_u32(void *ptr, uint32_t value, size_t count, cudaStream_t st) {
    if (count == 0) return PG_OK;
    const bool bytewise = ((value & 0xff) * 0x01010101u) == value;
    if (bytewise && (count * 4 <= (1u << 20) || ((uintptr_t)ptr & 15u))) {
        PG_CUDA(cudaMemsetAsync(ptr, (int)(value & 0xff), count * 4, st));
        return PG_OK;
    }
    if ((uintptr_t)ptr & 15u) { set_error("fill_u32: unaligned fill of a non-byte pattern"); return PG_EINVAL; }
    const size_t want = (count / 4 + 255) / 256 + 1;
    launch(k_fill_u32, (unsigned)(want < (size_t)kNumSM * 16 ? want : (size_t)kNumSM * 16), 256, 0, st, (uint32_t *)ptr, value, count);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// Empty slots: slot_rep = -1, slot_gid = 0xffffffff -- one byte pattern, so one fill initialises the table.  While the
// table is built slot_gid holds the smallest point index of the slot's group (unsigned atomicMin), afterwards its group id.
__global__ void k_group_insert(const int4 *__restrict__ keys, int64_t n, int32_t *slot_rep, int32_t *slot_min,
                               uint32_t cap, int32_t *__restrict__ pslot) {
    pdl_enter();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 k = keys[i];
    unsigned h = hash4(k.x, k.y, k.z, k.w) & (cap - 1);
    for (;;) {
        int rep = slot_rep[h];
        if (rep < 0) {
            int prev = atomicCAS(&slot_rep[h], -1, (int)i);
            rep = (prev < 0) ? (int)i : prev;
        }
        if (rep == (int)i) break;
        const int4 o = keys[rep];
        if (o.x == k.x && o.y == k.y && o.z == k.z && o.w == k.w) break;
        h = (h + 1) & (cap - 1);
    }
    atomicMin(reinterpret_cast<unsigned *>(&slot_min[h]), (unsigned)i);
    pslot[i] = (int)h;
}

// Grouping WITHOUT first-occurrence numbering (the ball query's grid: any dense numbering of the cells will do): the thread
// that claims a slot numbers its group on the spot -- claimed slots are counted per block (ballots + one shared-memory
// pass), one atomicAdd per block on the group counter -- so the ids come out in roughly ascending order of first
// appearance (blocks start in index order), which keeps neighbouring cells neighbours in memory, but not exactly, and not
// the same from run to run.  No minimum per slot, no flag / scan / publish pass.
__global__ void __launch_bounds__(256) k_group_insert_claim(const int4 *__restrict__ keys, int64_t n, int32_t *slot_rep,
                                                            int32_t *slot_gid, uint32_t cap, int32_t *__restrict__ pslot,
                                                            int4 *__restrict__ slot_key, unsigned long long *nGroups) {
    pdl_enter();
    __shared__ int s_warp[8];
    __shared__ unsigned long long s_base;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool won = false;
    unsigned h = 0;
    int4 k = make_int4(0, 0, 0, 0);
    if (i < n) {
        k = keys[i];
        h = hash4(k.x, k.y, k.z, k.w) & (cap - 1);
        for (;;) {
            int rep = slot_rep[h];
            if (rep < 0) {
                const int prev = atomicCAS(&slot_rep[h], -1, (int)i);
                if (prev < 0) { won = true; break; }
                rep = prev;
            }
            const int4 o = keys[rep];
            if (o.x == k.x && o.y == k.y && o.z == k.z && o.w == k.w) break;
            h = (h + 1) & (cap - 1);
        }
        pslot[i] = (int)h;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned wm = __ballot_sync(0xffffffffu, won);
    if (lane == 0) s_warp[warp] = __popc(wm);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int t = s_warp[w]; s_warp[w] = tot; tot += t; }
        s_base = tot ? atomicAdd(nGroups, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    if (won) {
        slot_gid[h] = (int)s_base + s_warp[warp] + __popc(wm & lanemask_lt());
        if (slot_key) slot_key[h] = k;
    }
}

// flag -> scan -> publish as ONE pass (scan.cuh): element i's value is "i is the first (lowest-index) point of its
// group", its exclusive prefix is then the group's id, which the first point writes into the slot.  Reading the flag and
// overwriting the slot's minimum with the id in the same pass is safe: only point i compares the slot with i, and it
// does so before it publishes; every other point j > i of the group sees either i or the id r <= i, never j.
struct GroupFlagLoad {
    const int32_t *pslot, *slot_min;
    __device__ int operator()(int64_t i) const { return slot_min[pslot[i]] == (int)i ? 1 : 0; }
};
struct GroupPublishStore {
    const int32_t *pslot;
    int32_t *slot_min;
    const int4 *keys;
    int4 *slot_key;
    __device__ void operator()(int64_t i, int rank, int flag) const {
        if (flag) {
            const int s = pslot[i];
            slot_min[s] = rank;
            if (slot_key) slot_key[s] = keys[i];
        }
    }
};

// every point takes its group's id and counts itself; WITH_MAX: *cnt_max receives the largest group size -- the
// increment that completes the fullest group returns its final size, so the maximum over all increments is it
// (without it the increment is a fire-and-forget reduction)
template <bool WITH_MAX>
__global__ void __launch_bounds__(256) k_group_assign(const int32_t *__restrict__ pslot, const int32_t *__restrict__ slot_gid,
                                                      int64_t n, int32_t *__restrict__ gid, int32_t *__restrict__ cnt,
                                                      int64_t *cnt_max, Fill extra) {
    pdl_enter();
    grid_fill(extra);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int mine = 0;
    if (i < n) {
        int g = slot_gid[pslot[i]];
        gid[i] = g;
        if (WITH_MAX) mine = atomicAdd(&cnt[g], 1) + 1;
        else atomicAdd(&cnt[g], 1);
    }
    if (WITH_MAX) {
        __shared__ int s_max;
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
        mine = __reduce_max_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine > s_max) atomicMax(&s_max, mine);
        __syncthreads();
        if (threadIdx.x == 0 && s_max > 0) atomicMax(reinterpret_cast<long long *>(cnt_max), (long long)s_max);
    }
}

int group_int4(const int4 *keys, int64_t n, GroupTable tab, int32_t *pslot, int32_t *gid, int32_t *cnt,
               int64_t *nGroups, int64_t *scan_tmp, cudaStream_t st, int64_t *cnt_max, Fill extra, bool first_occurrence) {
    if (n <= 0) {
        PG_CUDA(cudaMemsetAsync(nGroups, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    if (tab.slot_gid != tab.slot_rep + tab.cap) { set_error("group_int4: slot_rep and slot_gid must be adjacent"); return PG_EINVAL; }
    const int T = 256;
    const unsigned nb = (unsigned)div_up(n, T);
    if (!first_occurrence) {
        launch(k_group_insert_claim, nb, T, 0, st, keys, n, tab.slot_rep, tab.slot_gid, tab.cap, pslot, tab.slot_key,
               reinterpret_cast<unsigned long long *>(nGroups));            // *nGroups: zeroed by the caller
    } else {
        launch(k_group_insert, nb, T, 0, st, keys, n, tab.slot_rep, tab.slot_gid, tab.cap, pslot);
        PG_TRY(scan_fused(GroupFlagLoad{pslot, tab.slot_gid}, GroupPublishStore{pslot, tab.slot_gid, keys, tab.slot_key}, n, nGroups,
                          scan_tmp, st));
    }
    if (cnt_max) launch(k_group_assign<true>, nb, T, 0, st, pslot, tab.slot_gid, n, gid, cnt, cnt_max, extra);
    else launch(k_group_assign<false>, nb, T, 0, st, pslot, tab.slot_gid, n, gid, cnt, cnt_max, extra);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

}  // namespace pg

// ---------------------------------------------------------------------------------------------
// kernel timing
// ---------------------------------------------------------------------------------------------
namespace pg {
namespace {
struct KtRec { const char *name; cudaEvent_t a, b; };
std::mutex g_kt_mu;
std::vector<KtRec> g_kt;
bool g_kt_on = false;
}  // namespace

KTimer::KTimer(const char *name, cudaStream_t stream) : slot(-1), st(stream) {
    if (!g_kt_on) return;
    KtRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_kt_mu);
    slot = (int)g_kt.size();
    g_kt.push_back(r);
}
KTimer::~KTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_kt_mu);
    cudaEventRecord(g_kt[slot].b, st);
}
}  // namespace pg

extern "C" void pg_kernel_timing(int enable) {
    std::lock_guard<std::mutex> lk(pg::g_kt_mu);
    for (auto &r : pg::g_kt) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    pg::g_kt.clear();
    pg::g_kt_on = enable != 0;
}

extern "C" size_t pg_kernel_timing_report(char *buf, size_t cap) {
    std::lock_guard<std::mutex> lk(pg::g_kt_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto &r : pg::g_kt) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &e = agg[r.name];
            e.first += 1;
            e.second += ms;
        }
    }
    std::string out;
    char line[256];
    for (auto &kv : agg) {
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap) {
        const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}

extern "C" const char *pg_last_error(void) { return pg::last_error(); }
extern "C" int pg_abi_version(void) { return 3; }
