// Shared helpers for the sm_100a PointGroup kernels (error plumbing, launch math, small device utils).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pg_b200.h"

namespace pg {

constexpr int kWarp = 32;
constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids of persistent kernels are sized from this

void set_error(const char *fmt, ...);

#define PG_CHECK_ARG(cond, msg)                           \
    do {                                                  \
        if (!(cond)) {                                    \
            pg::set_error("%s: %s", __func__, msg);       \
            return PG_EINVAL;                             \
        }                                                 \
    } while (0)

#define PG_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            pg::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

#define PG_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            pg::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e));  \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

#define PG_TRY(expr)             \
    do {                         \
        int _rc = (expr);        \
        if (_rc != PG_OK) return _rc; \
    } while (0)

// Programmatic dependent launch (sm_90+): every kernel of this library starts with pdl_enter() and is launched through
// launch(), which sets cudaLaunchAttributeProgrammaticStreamSerialization.  A kernel's blocks then become resident while
// the kernel before it on the stream is still draining (its blocks have all STARTED -- that is when launch_dependents has
// been executed by every one of them) and sit at griddepcontrol.wait until that kernel has completed and its writes are
// visible: stream order as before, but the launch latency between the ~100 short kernels of a step overlaps the tail of
// the predecessor.  A kernel that runs through launch() MUST call pdl_enter() before it touches memory.
// HAZARD (seen once, in an experiment): with the trigger at the kernel's start a kernel's blocks can be resident while the
// kernel TWO before it is still running (its predecessor's blocks have all started and sit at their own wait), and ptxas
// moves loads it knows to be read-only -- `const __restrict__` parameters, __ldg: LDG.E.CONSTANT -- ABOVE the wait (the PTX
// order is right; neither a "memory" clobber nor a branch on a value the asm produces stops it).  Such a load reads what
// the last two kernels of this library on the stream have not written yet.  So: data produced by the previous two kernels
// must not be the first thing a kernel reads through a read-only pointer (ld_after_wait below is the plain load for such
// scalars); tests/test_abi.py disassembles the library and fails on any memory instruction ahead of ACQBULK.
// PG_B200_NO_PDL=1 in the environment falls back to plain launches (for A/B timing).
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// A kernel's FIRST look at a scalar another kernel may have produced: a plain (coherent) load in a volatile asm, which
// ptxas keeps behind the wait (it moves only loads it knows to be read-only: see HAZARD above).
__device__ __forceinline__ long long ld_after_wait(const int64_t *p) {
    long long v;
    asm volatile("ld.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_after_wait(const int32_t *p) {
    int v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through PG_LAUNCH_CHECK
}

// Kernel timing (pg_kernel_timing): when switched on, a launch wrapped in a KTimer scope is bracketed by
// two CUDA events on its own stream; pg_kernel_timing_report() sums them per kernel name.  Off by
// default (one predictable branch per launch); bench.py uses it for the per-kernel roofline numbers.
struct KTimer {
    int slot;
    cudaStream_t st;
    KTimer(const char *name, cudaStream_t stream);
    ~KTimer();
};
#define PG_KT_CAT2(a, b) a##b
#define PG_KT_CAT(a, b) PG_KT_CAT2(a, b)
#define PG_KTIME(name, st) pg::KTimer PG_KT_CAT(_pg_kt_, __LINE__)(name, st)

// Resident blocks per SM of `kernel` at (threads, dynamic smem), asked of the runtime once per call site.
// Statically partitioned grid-stride kernels size their grid as kNumSM * resident * k: a grid that is not a
// multiple of what fits leaves a second, mostly empty wave (k_bq_fill_mask: 40 registers -> 6 blocks of 256
// per SM, so a grid of 8 per SM ran as one full wave plus a third of one, 51 % achieved occupancy).
template <typename K>
inline int resident_blocks(K kernel, int threads, size_t smem = 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
#define PG_RESIDENT(kernel, threads, smem) ([&]() { static const int _n = pg::resident_blocks(kernel, threads, smem); return _n; }())

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b = 256) { return (a + b - 1) / b * b; }

// Bump allocator over the caller-provided workspace; both phases of a two-phase op replay the same
// sequence of take() calls so they see the same layout.
struct Arena {
    char *base;
    size_t size, used;
    bool ok;
    Arena(void *p, size_t n) : base((char *)p), size(n), used(0), ok(p != nullptr || n == 0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (used + bytes > size) { ok = false; used += bytes; return nullptr; }
        T *r = (T *)(base + used);
        used += bytes;
        return r;
    }
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// A second job for a kernel that runs before a scratch buffer's first use anyway: clear it, grid-stride, instead of a
// fill launch of its own (a large cudaMemsetAsync may be handed to a copy engine and queue behind an upload -- prim.cu).
struct Fill {
    uint32_t *p;
    size_t n;       // 32-bit words
    uint32_t v;
};
__device__ __forceinline__ void grid_fill(const Fill &f) {
    if (f.n == 0) return;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    if (((uintptr_t)f.p & 15u) == 0) {
        const size_t n16 = f.n >> 2;
        const uint4 q = make_uint4(f.v, f.v, f.v, f.v);
        uint4 *p16 = reinterpret_cast<uint4 *>(f.p);
        for (size_t i = t; i < n16; i += nt) p16[i] = q;
        if (t < (f.n & 3)) f.p[(n16 << 2) + t] = f.v;
    } else {
        for (size_t i = t; i < f.n; i += nt) f.p[i] = f.v;
    }
}

// float <-> order-preserving uint32 (total order on non-NaN floats)
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ unsigned hash4(int a, int b, int c, int d) {
    unsigned h = 0x9E3779B9u;
    h ^= (unsigned)a; h *= 0x85EBCA6Bu; h ^= h >> 15;
    h ^= (unsigned)b; h *= 0xC2B2AE35u; h ^= h >> 13;
    h ^= (unsigned)c; h *= 0x27D4EB2Fu; h ^= h >> 16;
    h ^= (unsigned)d; h *= 0x165667B1u; h ^= h >> 15;
    return h;
}

// ---- host-side primitives (prim.cu) ------------------------------------------------------------
// Exclusive prefix sum of int32 `in[0..n)` into `out` (may alias `in`).  `total` (device int64, may
// be null) receives the grand total.  `tmp` must hold scan_tmp_count(n) int64 values.
size_t scan_tmp_count(int64_t n);
int scan_exclusive_i32(const int32_t *in, int32_t *out, int64_t n, int64_t *total, int64_t *tmp,
                       cudaStream_t st);

// count 32-bit words of `value` at ptr: a kernel above 1 MB (see prim.cu), cudaMemsetAsync below
int fill_u32(void *ptr, uint32_t value, size_t count, cudaStream_t st);

// Stable LSD radix sort of (key, value) pairs on the low `bits` bits of the key (8-bit digits).
// Pass 0 reads (keys_src, vals_src) -- vals_src == nullptr means "values are 0..n-1" -- and the
// passes ping-pong between the A and B buffers, so the source arrays are never written.
// *result_buf = 0 when the sorted pairs end in (keysA, valsA), 1 for (keysB, valsB).
size_t radix_tmp_count(int64_t n);  // int32 count for the histogram scratch
int radix_sort_pairs(const uint32_t *keys_src, const uint32_t *vals_src, uint32_t *keysA, uint32_t *valsA,
                     uint32_t *keysB, uint32_t *valsB, int64_t n, int bits, int32_t *hist, int64_t *scan_tmp,
                     cudaStream_t st, int *result_buf, int2 *pairs_out = nullptr, int64_t pairs_n = 0);
// pairs_out: the LAST pass writes its first pairs_n sorted (key, value) pairs there as int2 rows instead of the A / B
// buffers (bfs_cluster's (cluster, point) rows: one launch less than sorting and then interleaving).

// Group n int4 keys: gid[i] = group id numbered by first occurrence in input order; *nGroups (device
// int64) = number of groups; cnt[g] = group size.  The open-addressing table (cap slots, cap a power
// of two >= 2n) stays usable afterwards through group_lookup(): slot_rep[s] = a point holding the
// slot's key (-1 empty), slot_gid[s] = its group id.
struct GroupTable {
    int32_t *slot_rep;  // [cap]
    int32_t *slot_gid;  // [cap]  (min point index while building, group id afterwards; -1 = empty slot)
    uint32_t cap;
    int4 *slot_key;     // [cap] or null: the slot's key itself, for lookups that should not chase slot_rep -> keys[]
};
uint32_t group_table_cap(int64_t n);
// cnt_max (optional, device, zeroed by the caller): the largest group size.
// The CALLER's key kernel clears the table and the counts on its way (grid_fill of group_table_fill() and of cnt, which
// no kernel has touched yet at that point); `extra` is cleared by the last kernel of the grouping for whoever comes next.
int group_int4(const int4 *keys, int64_t n, GroupTable tab, int32_t *pslot /*[n] scratch*/,
               int32_t *gid /*[n]*/, int32_t *cnt /*[n]*/, int64_t *nGroups, int64_t *scan_tmp,
               cudaStream_t st, int64_t *cnt_max = nullptr, Fill extra = Fill{nullptr, 0, 0}, bool first_occurrence = true);
// first_occurrence = false: any dense numbering (roughly by first appearance, not reproducible); *nGroups must be zero on entry.
// slot_rep and slot_gid are adjacent in every workspace layout: one range of 0xffffffff
inline Fill group_table_fill(const GroupTable &tab) { return Fill{reinterpret_cast<uint32_t *>(tab.slot_rep), (size_t)tab.cap * 2, 0xffffffffu}; }

__device__ __forceinline__ int group_lookup(const int4 *keys, const GroupTable &tab, int4 k) {
    unsigned h = hash4(k.x, k.y, k.z, k.w) & (tab.cap - 1);
    if (tab.slot_key) {                 // key and id of a slot are two independent loads: one memory round trip per probe
        for (;;) {
            const int gid = __ldg(tab.slot_gid + h);
            const int4 o = __ldg(tab.slot_key + h);
            if (gid < 0) return -1;
            if (o.x == k.x && o.y == k.y && o.z == k.z && o.w == k.w) return gid;
            h = (h + 1) & (tab.cap - 1);
        }
    }
    for (;;) {
        int rep = tab.slot_rep[h];
        if (rep < 0) return -1;
        int4 o = keys[rep];
        if (o.x == k.x && o.y == k.y && o.z == k.z && o.w == k.w) return tab.slot_gid[h];
        h = (h + 1) & (tab.cap - 1);
    }
}

}  // namespace pg
