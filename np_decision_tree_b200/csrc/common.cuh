// Shared infrastructure of libb200tree: context, scratch buffers, per-kernel profiling,
// warp helpers and the flag words of the single-pass (decoupled look-back) scans.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/b200dt.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// ---------------------------------------------------------------------------------------
// kernel classes for the profiler (order = b200dt_profile_name ids)
// ---------------------------------------------------------------------------------------
enum KernelClass {
    KC_SORT_PREPARE = 0,  // X (row-major f64) -> keys (feature-major u64) + digit histograms
    KC_SORT_PASS,         // one 8-bit onesweep radix pass
    KC_L2_SCAN,           // tiled segmented scan + gain + argmax (large nodes)
    KC_L2_EXACT,          // sequential-order scan (small / near-tie nodes)
    KC_FINALIZE,          // per-node arg-reduce of the partial candidates
    KC_MARK,              // left-row side mask
    KC_PARTITION,         // stable segmented partition
    KC_PREDICT,           // batched traversal
    KC_L1_RANK,           // node-local y ranks for the order-statistic descent
    KC_L1_DESCENT,        // one bit level of the prefix/suffix median descent
    KC_L1_COST,           // L1 cost scans + argmin
    KC_MISC,              // small helpers (histogram scan, reductions, layout converts)
    KC_LEVEL_FUSED,       // fused per-level partition + search of the children (streaming y payload)
    KC_PAYLOAD,           // y payload gather, once after the sort
    KC_SORT_FIXUP,        // 32-bit sort path: gather full keys, repair runs tied on the top 32 bits
    KC_CLASS_HIST,        // classification: per-tile class histograms
    KC_CLASS_SCORE,       // classification: gini / precision-at-1 scores + argmax
    KC_ROWS,              // random-threshold splitters: column min/max, threshold counts, row-list split
    KC_CHILD_SCAN,        // two-level search: both children scored from the parent's order (side-bit look-up)
    KC_RELABEL,           // once per fit: contiguous row ids per node (relabel.cu)
    KC_COUNT
};

static const char *const kKernelClassNames[KC_COUNT] = {
    "sort_prepare", "sort_radix_pass", "l2_scan_search", "l2_exact_search", "finalize_argbest",
    "mark_left_rows", "partition", "predict", "l1_rank", "l1_median_descent", "l1_cost_search",
    "misc", "level_fused_partition_search", "gather_y_payload", "sort_fixup_ties", "class_tile_histogram",
    "class_score_search", "row_list_ops", "l2_children_scan_search", "relabel_rows"};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ProfSlot {
    cudaEvent_t a, b;
    int kc;
};

struct Session;       // fit.cu
struct ClassSession;  // classes.cu
struct RowSession;    // rows.cu

struct b200dt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;        // host->device copies that overlap the sort
    std::vector<cudaEvent_t> copy_events;
    std::string err;
    int sm_count = 148;

    // profiling
    bool prof_on = false;
    std::vector<ProfSlot> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[KC_COUNT] = {0};
    int64_t prof_launches[KC_COUNT] = {0};
    double prof_bytes[KC_COUNT] = {0};
    int64_t launches = 0;
    // epoch of the look-back flag words: monotonic per context because the status buffers are
    // reused across sessions (a stale flag of an old launch must never look current)
    u32 epoch = 0;
    u32 attr_mask = 0;  // which kernels already had their dynamic shared-memory limit raised on this device

    // named grow-only scratch buffers (no allocation in steady state)
    std::vector<DevBuf> scratch;
    std::vector<void *> pinned;  // pinned host staging (index-matched with pinned_cap)
    std::vector<size_t> pinned_cap;

    Session *session = nullptr;
    ClassSession *class_session = nullptr;
    RowSession *row_session = nullptr;

    int fail(const char *what, cudaError_t e, const char *file, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        err = buf;
        return 1;
    }
    int fail_msg(const std::string &m) {
        err = m;
        return 1;
    }
};

#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return ctx->fail(#call, e_, __FILE__, __LINE__); \
    } while (0)

#define CKR(call)               \
    do {                        \
        int r_ = (call);        \
        if (r_ != 0) return r_; \
    } while (0)

// scratch slot ids
enum ScratchId {
    SB_KEYS0 = 0, SB_KEYS1, SB_SORT_HIST, SB_SORT_STATUS, SB_ORD_A, SB_ORD_B, SB_X, SB_Y, SB_MASK,
    SB_SEG, SB_TILES, SB_PARTIAL, SB_RESULT, SB_STATUS_F, SB_STATUS_VAL, SB_STATUS_P, SB_MISC,
    SB_TMP_I64, SB_TMP2, SB_TREE, SB_OUT, SB_LEAF, SB_L1_A, SB_L1_B, SB_L1_C, SB_L1_D, SB_L1_E,
    SB_L1_F, SB_L1_G, SB_L1_H, SB_LASTCOL, SB_STATUS_VAL2, SB_YPAY0, SB_YPAY1, SB_PLAN, SB_SORT_RANGE, SB_W, SB_YW, SB_WPAY0, SB_WPAY1, SB_TW, SB_MASK_A, SB_CODES, SB_PLAN4, SB_NEWID, SB_ROWMAP, SB_RL_BINS, SB_RL_COUNTS, SB_NODE_BEST, SB_COUNT
};

static inline int scratch_get(b200dt_ctx *ctx, int id, size_t bytes, void **out, bool zero_new = false) {
    if (ctx->scratch.size() < (size_t)SB_COUNT) ctx->scratch.resize(SB_COUNT);
    DevBuf &b = ctx->scratch[id];
    if (b.cap < bytes) {
        if (b.p) CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            want = bytes;
            CK(cudaMalloc(&b.p, want));
        }
        b.cap = want;
        if (zero_new) CK(cudaMemsetAsync(b.p, 0, want, ctx->stream));
    }
    *out = b.p;
    return 0;
}

static inline int scratch_free(b200dt_ctx *ctx, int id) {
    if (ctx->scratch.size() <= (size_t)id) return 0;
    DevBuf &b = ctx->scratch[id];
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    return 0;
}

// pinned host staging buffers, slot-indexed
static inline int pinned_get(b200dt_ctx *ctx, int slot, size_t bytes, void **out) {
    if (ctx->pinned.size() <= (size_t)slot) {
        ctx->pinned.resize(slot + 1, nullptr);
        ctx->pinned_cap.resize(slot + 1, 0);
    }
    if (ctx->pinned_cap[slot] < bytes) {
        // an earlier async copy may still be reading the old buffer
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned[slot]) CK(cudaFreeHost(ctx->pinned[slot]));
        ctx->pinned[slot] = nullptr;
        size_t want = bytes * 2 + 4096;
        CK(cudaMallocHost(&ctx->pinned[slot], want));
        ctx->pinned_cap[slot] = want;
    }
    *out = ctx->pinned[slot];
    return 0;
}

// NVTX range around a host-side phase (sort, level search / decide / partition, predict): shows up in Nsight
// timelines; header-only, a no-op without a profiler attached
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---------------------------------------------------------------------------------------
// launch bracket: counts launches; with profiling on, times them with events
// ---------------------------------------------------------------------------------------
struct LaunchScope {
    b200dt_ctx *ctx;
    int kc;
    cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(b200dt_ctx *c, int kclass, double alg_bytes, int n_launches = 1) : ctx(c), kc(kclass) {
        ctx->launches += n_launches;
        if (!ctx->prof_on) return;
        ctx->prof_launches[kc] += n_launches;
        ctx->prof_bytes[kc] += alg_bytes;
        auto take = [&]() {
            cudaEvent_t e;
            if (!ctx->prof_free.empty()) {
                e = ctx->prof_free.back();
                ctx->prof_free.pop_back();
            } else {
                cudaEventCreate(&e);
            }
            return e;
        };
        a = take();
        b = take();
        cudaEventRecord(a, ctx->stream);
    }
    ~LaunchScope() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->prof_pending.push_back({a, b, kc});
    }
};

static inline void prof_drain(b200dt_ctx *ctx) {
    static const bool trace = getenv("B200DT_TRACE_LAUNCHES") != nullptr;   // one line per bracket on stderr
    for (auto &s : ctx->prof_pending) {
        cudaEventSynchronize(s.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        if (trace) fprintf(stderr, "b200dt launch class %d: %.3f ms\n", s.kc, ms);
        ctx->prof_ms[s.kc] += ms;
        ctx->prof_free.push_back(s.a);
        ctx->prof_free.push_back(s.b);
    }
    ctx->prof_pending.clear();
}

// Persistent kernels whose tiles wait on lower-numbered tiles need every block resident at once.
// A cooperative launch makes the runtime REFUSE a grid that cannot be co-resident (e.g. SMs taken by
// another client) instead of letting it dead-lock on the device.
static inline cudaError_t launch_resident(const void *kernel, unsigned grid, unsigned block, void **args,
                                          size_t smem, cudaStream_t stream) {
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(block), args, smem, stream);
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Device-side index assertions of the bounds-checked build (-DB200DT_BOUNDS_CHECK, libb200tree_checked.so): the
// substitute for compute-sanitizer memcheck where that tool is not available.  A failing check prints and traps.
#ifdef B200DT_BOUNDS_CHECK
#define DCHECK(cond)                                                                         \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            printf("B200DT_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                       \
            __trap();                                                                        \
        }                                                                                    \
    } while (0)
#else
#define DCHECK(cond) \
    do {             \
    } while (0)
#endif

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(FULL_MASK, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }

// streaming loads / stores: the order arrays are read once per level
__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ---- bulk asynchronous copies (TMA, 1-D) into shared memory, completion on an mbarrier ------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival + `bytes` more transaction bytes to wait for in the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE_%=;\n\t"
        "bra MBAR_WAIT_%=;\n\t"
        "MBAR_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; signals `bar` with complete_tx
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// order this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// relaxed / acquire-release accesses for the look-back flag words
__device__ __forceinline__ u32 ld_relaxed_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32 *p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// fast reciprocal, ~1e-15 relative: hardware seed + 2 Newton steps (approximate search only)
__device__ __forceinline__ double rcp_fast(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// node-level records shared by search / finalize / session
// ---------------------------------------------------------------------------------------
// Partial candidate written by a search block for one (tile or small node, column).
struct Partial {
    double score;   // L2 gain (maximise) or L1 cost (minimise)
    double aux;     // L2: |S-(k+1)mean|; L1: unused
    double lsum;    // L2: prefix sum at k
    double score2;  // best score among the other candidates seen
    int k;          // position inside the node, -1 = none
    int f;          // local column
    int ppos;       // two-level search: position of that row inside the PARENT's segment (data not yet moved); else -1
    int pad_;
};

// cross-file entry points (host side)
int sort_columns_device(b200dt_ctx *ctx, const double *dX, int64_t N, int64_t F, int64_t ldx,
                        int *d_orders_out, int *d_orders_alt, int64_t col_stride, u64 *ext_keys0 = nullptr,
                        u64 *ext_keys1 = nullptr, int64_t ext_key_capacity = 0, const double *y = nullptr,
                        double *ypay_out = nullptr);
