// Device-wide primitives used by several ops: exclusive scan, stable LSD radix sort of pairs, and
// hash grouping of int4 keys (first-occurrence numbering).  Hand-written; no CUB/Thrust.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

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

// ---------------------------------------------------------------------------------------------
// exclusive scan (single pass, decoupled look-back)
// ---------------------------------------------------------------------------------------------
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

// Single pass with decoupled look-back: a block takes the next tile (ticket from an atomic counter, so
// every earlier tile is already running or done), scans it, publishes its aggregate, and warp 0 walks
// back over the predecessors' published words -- 32 at a time -- until it meets an inclusive prefix.
// One 64-bit word per tile: [63:62] 0 = nothing yet, 1 = aggregate, 2 = inclusive prefix; [61:0] value.
// tmp[0] is the ticket counter, tmp[1 + t] tile t's word; the launcher zeroes them.
constexpr unsigned long long kSpValueMask = (1ULL << 62) - 1;

__global__ void __launch_bounds__(kScanThreads) k_scan_onepass(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                               int64_t n, unsigned long long *tmp, int64_t *__restrict__ total,
                                                               int aligned) {
    __shared__ int warp_tot[32];
    __shared__ long long s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(tmp, 1ULL);
    __syncthreads();
    const int64_t t = s_tile;
    volatile unsigned long long *state = tmp + 1;
    const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanRounds;
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
        if (lane == 0) {
            s_prefix = prefix;
            if (total && (t + 1) * (int64_t)kScanTile >= n) *total = prefix + tot;
        }
    }
    __syncthreads();
    const int off = (int)s_prefix + incl - tsum;
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

size_t scan_tmp_count(int64_t n) { return (size_t)div_up(n > 0 ? n : 1, kScanTile) + 2; }

int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, int64_t *total, int64_t *tmp,
                       cudaStream_t st) {
    if (n <= 0) {
        if (total) PG_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    const int64_t nb = div_up(n, kScanTile);
    const bool aligned = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    PG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)(nb + 1) * sizeof(int64_t), st));
    k_scan_onepass<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, reinterpret_cast<unsigned long long *>(tmp), total,
                                                         aligned ? 1 : 0);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits: histogram -> scan -> ranked scatter per pass
// ---------------------------------------------------------------------------------------------
constexpr int kRadixThreads = 256;
constexpr int kRadixRounds = 8;
constexpr int kRadixTile = kRadixThreads * kRadixRounds;

__global__ void __launch_bounds__(kRadixThreads) k_radix_hist(const uint32_t *__restrict__ keys, int64_t n,
                                                              int shift, int32_t *__restrict__ hist, int nb) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRadixTile;
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        int64_t i = base + r * kRadixThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// Each warp owns a contiguous 256-element slice of the tile (8 rounds of 32 consecutive elements), so the stable
// order inside the tile is (warp, round, lane): a warp ranks its own slice with nothing but warp-level
// primitives -- `__match_any_sync` groups equal digits, a per-warp counter row in shared memory carries the
// running count from round to round -- and the block meets only twice per tile: once to turn the eight counter
// rows into per-warp bases, once before the rows are reused.  Keys, values and ranks stay in registers.
__global__ void __launch_bounds__(kRadixThreads)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                const int32_t *__restrict__ hist, int nb) {
    constexpr int kWarps = kRadixThreads / 32;
    __shared__ int wcnt[kWarps][256];                 // per-warp digit counts, then exclusive bases across warps
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const int gbase = hist[(int64_t)threadIdx.x * nb + blockIdx.x];   // global base of this tile's run of digit tid
#pragma unroll
    for (int k = 0; k < kWarps; k++) wcnt[k][threadIdx.x] = 0;
    __syncthreads();
    const int64_t slice = (int64_t)blockIdx.x * kRadixTile + (int64_t)w * (kRadixRounds * 32);
    uint32_t key[kRadixRounds], val[kRadixRounds];
    int rank[kRadixRounds];                           // position inside the warp's slice among equal digits
    unsigned dig[kRadixRounds];
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        const int64_t i = slice + r * 32 + lane;
        const bool live = i < n;
        key[r] = live ? keys_in[i] : 0u;
        val[r] = live ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
        dig[r] = live ? ((key[r] >> shift) & 255u) : 256u;   // dead lanes never match a real digit
    }
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        const int before = __popc(peers & lt);
        int run = 0;
        if (dig[r] < 256u) {
            run = wcnt[w][dig[r]];                    // equal digits read the same counter, then the leader bumps it
        }
        __syncwarp();
        if (dig[r] < 256u && before == 0) wcnt[w][dig[r]] = run + __popc(peers);
        __syncwarp();
        rank[r] = run + before;
    }
    __syncthreads();
    {   // digit `tid`: exclusive prefix of the eight warp counts, shifted by the tile's global base
        int acc = gbase;
#pragma unroll
        for (int k = 0; k < kWarps; k++) {
            const int c = wcnt[k][threadIdx.x];
            wcnt[k][threadIdx.x] = acc;
            acc += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRadixRounds; r++) {
        if (dig[r] < 256u) {
            const int pos = wcnt[w][dig[r]] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

size_t radix_tmp_count(int64_t n) { return (size_t)256 * (size_t)div_up(n > 0 ? n : 1, kRadixTile) + 16; }

int radix_sort_pairs(const uint32_t *keys_src, const uint32_t *vals_src, uint32_t *keysA, uint32_t *valsA,
                     uint32_t *keysB, uint32_t *valsB, int64_t n, int bits, int32_t *hist, int64_t *scan_tmp,
                     cudaStream_t st, int *result_buf) {
    *result_buf = 0;
    if (n <= 0) return PG_OK;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int nb = (int)div_up(n, kRadixTile);
    uint32_t *k[2] = {keysA, keysB}, *v[2] = {valsA, valsB};
    const uint32_t *kin = keys_src, *vin = vals_src;
    int dst = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        k_radix_hist<<<nb, kRadixThreads, 0, st>>>(kin, n, shift, hist, nb);
        PG_TRY(scan_exclusive_i32(hist, hist, (int64_t)256 * nb, nullptr, scan_tmp, st));
        k_radix_scatter<<<nb, kRadixThreads, 0, st>>>(kin, vin, k[dst], v[dst], n, shift, hist, nb);
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

__global__ void k_group_init(int32_t *__restrict__ slot_rep, int32_t *__restrict__ slot_min, uint32_t cap) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        slot_rep[i] = -1;
        slot_min[i] = 0x7fffffff;
    }
}

__global__ void k_group_insert(const int4 *__restrict__ keys, int64_t n, int32_t *slot_rep, int32_t *slot_min,
                               uint32_t cap, int32_t *__restrict__ pslot) {
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
    atomicMin(&slot_min[h], (int)i);
    pslot[i] = (int)h;
}

// flag[i] = 1 when i is the first (lowest-index) point of its group
__global__ void k_group_flag(const int32_t *__restrict__ pslot, const int32_t *__restrict__ slot_min, int64_t n,
                             int32_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (slot_min[pslot[i]] == (int)i) ? 1 : 0;
}

// first points publish their rank (= group id) into the slot; needs the pre-scan flags, which after
// the in-place scan are recovered as rank[i+1] - rank[i] (or total - rank[n-1])
__global__ void k_group_publish(const int32_t *__restrict__ pslot, int32_t *slot_min, const int32_t *__restrict__ rank,
                                const int64_t *__restrict__ total, int64_t n, const int4 *__restrict__ keys,
                                int4 *__restrict__ slot_key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = rank[i];
    int next = (i + 1 < n) ? rank[i + 1] : (int)*total;
    if (next != r) {
        slot_min[pslot[i]] = r;
        if (slot_key) slot_key[pslot[i]] = keys[i];
    }
}

__global__ void k_group_assign(const int32_t *__restrict__ pslot, const int32_t *__restrict__ slot_gid, int64_t n,
                               int32_t *__restrict__ gid, int32_t *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = slot_gid[pslot[i]];
    gid[i] = g;
    atomicAdd(&cnt[g], 1);
}

int group_int4(const int4 *keys, int64_t n, GroupTable tab, int32_t *pslot, int32_t *gid, int32_t *cnt,
               int64_t *nGroups, int64_t *scan_tmp, cudaStream_t st) {
    if (n <= 0) {
        PG_CUDA(cudaMemsetAsync(nGroups, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    const int T = 256;
    const unsigned nb = (unsigned)div_up(n, T);
    k_group_init<<<kNumSM * 8, T, 0, st>>>(tab.slot_rep, tab.slot_gid, tab.cap);
    PG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(int32_t), st));
    k_group_insert<<<nb, T, 0, st>>>(keys, n, tab.slot_rep, tab.slot_gid, tab.cap, pslot);
    k_group_flag<<<nb, T, 0, st>>>(pslot, tab.slot_gid, n, gid);
    PG_TRY(scan_exclusive_i32(gid, gid, n, nGroups, scan_tmp, st));
    k_group_publish<<<nb, T, 0, st>>>(pslot, tab.slot_gid, gid, nGroups, n, keys, tab.slot_key);
    k_group_assign<<<nb, T, 0, st>>>(pslot, tab.slot_gid, n, gid, cnt);
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
