// Device-side records and host launchers of the per-level tree kernels (K2, K4, finalize).
#pragma once
#include "common.cuh"

// Tiles are owned by single warps (no block barriers anywhere in K2 / K4):
constexpr int WT_ITEMS = 16;
constexpr int WT_TILE = 32 * WT_ITEMS;   // K2: 512 sorted positions per warp tile
#ifndef PT_ITEMS_CFG
#define PT_ITEMS_CFG 16
#endif
#ifndef PT_WARPS_CFG
#define PT_WARPS_CFG 8
#endif
constexpr int PT_ITEMS = PT_ITEMS_CFG;
constexpr int PT_WARPS = PT_WARPS_CFG;
constexpr int PT_WARP = 32 * PT_ITEMS;   // K4: 512 sorted positions per warp ...
constexpr int PT_TILE = PT_WARPS * PT_WARP;     // ... 4096 per block tile (one look-back chain entry)
constexpr int P4_ITEMS = 16;
constexpr int P4_TILE = 32 * P4_ITEMS;   // 4-way partition of the two-level mode (children.cu): one warp per tile
constexpr int ST_ITEMS = 32;
constexpr int ST_TILE = 32 * ST_ITEMS;   // K2 streaming (y payload): 1024 positions per warp tile
constexpr int EXACT_ROWS = 2048;         // nodes up to this size take the sequential-order search

// One node of the current level = one contiguous segment [start, start+n) of every column of
// the order arrays.
struct SegDev {
    double total;   // sum of y over the node (approximate path); exact path recomputes it
    double total_w; // weighted fits: sum of the weights of the node
    int start;
    int n;
    int pbase;      // first slot of this node's partial candidates
    int np;         // number of slots (tiles, or 1 for a sequentially scanned node)
    int nleft;      // partition: rows going left
    int exact;      // 1: searched by the sequential-order kernel
    int pstep;      // candidate slot j of this node is pbase + j * pstep + pside: (1, 0) for slots of
    int pside;      // its own; (2, side) when the fused level kernel scored it from its parent's tiles
    int vstart;     // >= 0: the node's rows still lie inside its PARENT's segment, which starts here (two-level
    int cbase;      // search: candidates then carry Partial::ppos); -1: [start, start+n) is physical.  cbase: first
                    // look-back chain slot of this node (its tiles own slots cbase .. cbase + np - 1 of every column,
                    // whatever order the tiles are processed in)
};

// One split node handed to the fused level kernel (partition + search of the children).
struct PlanSeg {
    double total_l, total_r;  // sums of y over the children (their means are total / n)
    int start, n, nleft;
    int fbase;                // candidate slot of (tile tin, side) = fbase + 2 * tin + side
    int score_l, score_r;     // 1: that child is searched by the parallel scan at the next level
};

// One parent of a level pair handed to the 4-way partition (children.cu)
struct Plan4 {
    int start, n;
    int n_l;    // rows of the left child
    int n_ll;   // rows of the left child's left child (0 if it was not split)
    int n_rl;   // rows of the right child's left child
    int pad;
};

struct TileDesc {
    int seg;  // index into the SegDev array
    int tin;  // tile index inside the segment
};

// Look-back state of the tiled scans.  flag word = (epoch << 2) | state, state 1 = aggregate
// published, 2 = inclusive prefix published; epochs make clearing between launches unnecessary.
struct ScanStatus {
    // one 16-byte slot per (tile, column): word j = [63:34] epoch, [33:32] state, [31:0] half j of the
    // fp64 value.  A slot is valid when both words carry the current epoch and the SAME state, so a
    // reader needs no fence and a torn read (aggregate half + inclusive half) is simply retried.
    u64 *slot;
};

struct LevelArgs {
    const int *orders;      // [F_store][col_stride] row ids, feature-major
    int64_t col_stride;
    int F_search;           // columns searched (local feature slice)
    const double *y;        // (N,) targets by row id (gather path)
    const double *ypay;     // [F_store][col_stride] y travelling with each row id (streaming path), or null
    const double *wpay;     // weighted fits: the sample weights travelling the same way (ypay then holds y*w)
    const SegDev *segs;
    int n_segs;
    const int *row_map;     // non-null after the rows were relabelled (relabel.cu): row id -> ORIGINAL row
    long long *node_best;   // per live node: bits of the best gain found so far (streaming search), zeroed per level
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// Decoupled look-back over the earlier tiles of one (node, column) chain, run by ONE WARP.
// Chain entries are consecutive status indices idx-tin .. idx.  Each step inspects a window
// of 32 predecessors: it needs every entry up to the first one that already holds an
// inclusive prefix to have at least its aggregate published, then adds them up (fixed lane
// order).  Slots are self-validating 16-byte records polled with relaxed loads: no fences.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void status_publish(u64 *slot, u32 epoch, u32 state, double v) {
    const u64 bits = (u64)__double_as_longlong(v);
    const u64 tag = ((u64)epoch << 34) | ((u64)state << 32);
    // one 16-byte store; each half is self-describing, so ordering between them does not matter
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(tag | (bits & 0xffffffffull)),
                 "l"(tag | (bits >> 32))
                 : "memory");
}

// state of a slot for this epoch (0 = not ready / torn) and its value
__device__ __forceinline__ u32 status_read(const u64 *slot, u32 epoch, double &v) {
    u64 w0, w1;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    const u32 t0 = (u32)(w0 >> 32), t1 = (u32)(w1 >> 32);
    v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    return (t0 == t1 && (t0 >> 2) == epoch) ? (t0 & 3u) : 0u;
}

__device__ __forceinline__ double lookback_f64(const ScanStatus &st, int64_t idx, int tin, u32 epoch,
                                               double tile_total, int lane) {
    if (tin == 0) {
        if (lane == 0) status_publish(st.slot + 2 * idx, epoch, 2u, tile_total);
        return 0.0;
    }
    if (lane == 0) status_publish(st.slot + 2 * idx, epoch, 1u, tile_total);
    double excl = 0.0;
    int remaining = tin;
    int64_t look = idx - 1;
    while (true) {
        const bool in_range = lane < remaining;
        u32 state;
        int first;
        double val = 0.0;
        while (true) {
            state = 2u;  // past the chain head: empty prefix
            val = 0.0;
            if (in_range) state = status_read(st.slot + 2 * (look - lane), epoch, val);
            const u32 ready = __ballot_sync(FULL_MASK, state != 0u);
            const u32 incl = __ballot_sync(FULL_MASK, state == 2u);
            first = __ffs(incl) - 1;
            const u32 need = first < 0 || first == 31 ? 0xffffffffu : ((2u << first) - 1u);
            if ((ready & need) == need) break;
        }
        double v = (in_range && (first < 0 || lane <= first)) ? val : 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
        excl += v;
        if (first >= 0) break;
        look -= 32;
        remaining -= 32;
    }
    if (lane == 0) status_publish(st.slot + 2 * idx, epoch, 2u, excl + tile_total);
    return excl;
}

// Same for integer counts packed with their flag in one 64-bit word:
// [63:62] state, [61:32] epoch, [31:0] count -- the word itself is the message, no fence needed.
__device__ __forceinline__ int lookback_count(u64 *status, int64_t idx, int tin, u32 epoch, int tile_count, int lane) {
    const u64 tag = (u64)(epoch & 0x3fffffffu) << 32;
    if (tin == 0) {
        if (lane == 0) st_relaxed_u64(status + idx, (2ull << 62) | tag | (u32)tile_count);
        return 0;
    }
    if (lane == 0) st_relaxed_u64(status + idx, (1ull << 62) | tag | (u32)tile_count);
    int excl = 0;
    int remaining = tin;
    int64_t look = idx - 1;
    while (true) {
        const bool in_range = lane < remaining;
        u32 state;
        int first, val;
        while (true) {
            const u64 w = in_range ? ld_relaxed_u64(status + (look - lane)) : 0ull;
            state = in_range ? (((w & 0x3fffffff00000000ull) == tag) ? (u32)(w >> 62) : 0u) : 2u;
            val = in_range ? (int)(u32)w : 0;
            const u32 ready = __ballot_sync(FULL_MASK, state != 0u);
            const u32 incl = __ballot_sync(FULL_MASK, state == 2u);
            first = __ffs(incl) - 1;
            const u32 need = first < 0 || first == 31 ? 0xffffffffu : ((2u << first) - 1u);
            if ((ready & need) == need) break;
        }
        int v = (in_range && (first < 0 || lane <= first)) ? val : 0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL_MASK, v, m);
        excl += v;
        if (first >= 0) break;
        look -= 32;
        remaining -= 32;
    }
    if (lane == 0) st_relaxed_u64(status + idx, (2ull << 62) | tag | (u32)(excl + tile_count));
    return excl;
}
// Look-back of the fused level kernel: (left count, left sum, right sum) in one 48-byte slot of six
// tagged words: [count | sumL.lo | sumL.hi | sumR.lo | sumR.hi | spare], each [63:34] epoch,
// [33:32] state, [31:0] payload.  Valid when the five used words agree on (epoch, state).
struct Carry3 {
    int c;
    double l, r;
};

__device__ __forceinline__ void status3_publish(u64 *slot, u32 epoch, u32 state, int c, double l, double r) {
    const u64 tag = ((u64)epoch << 34) | ((u64)state << 32);
    const u64 lb = (u64)__double_as_longlong(l), rb = (u64)__double_as_longlong(r);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(tag | (u64)(u32)c),
                 "l"(tag | (lb & 0xffffffffull)) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 2), "l"(tag | (lb >> 32)),
                 "l"(tag | (rb & 0xffffffffull)) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 4), "l"(tag | (rb >> 32)), "l"(tag)
                 : "memory");
}

__device__ __forceinline__ u32 status3_read(const u64 *slot, u32 epoch, Carry3 &v) {
    u64 w0, w1, w2, w3, w4, w5;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(slot + 2) : "memory");
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w4), "=l"(w5) : "l"(slot + 4) : "memory");
    const u32 t0 = (u32)(w0 >> 32);
    const bool same = t0 == (u32)(w1 >> 32) && t0 == (u32)(w2 >> 32) && t0 == (u32)(w3 >> 32) &&
                      t0 == (u32)(w4 >> 32);
    v.c = (int)(u32)w0;
    v.l = __longlong_as_double((long long)((w1 & 0xffffffffull) | (w2 << 32)));
    v.r = __longlong_as_double((long long)((w3 & 0xffffffffull) | (w4 << 32)));
    return (same && (t0 >> 2) == epoch) ? (t0 & 3u) : 0u;
}

__device__ __forceinline__ Carry3 lookback3(u64 *status, int64_t idx, int tin, u32 epoch, Carry3 tile, int lane) {
    Carry3 excl{0, 0.0, 0.0};
    if (tin == 0) {
        if (lane == 0) status3_publish(status + 6 * idx, epoch, 2u, tile.c, tile.l, tile.r);
        return excl;
    }
    if (lane == 0) status3_publish(status + 6 * idx, epoch, 1u, tile.c, tile.l, tile.r);
    int remaining = tin;
    int64_t look = idx - 1;
    while (true) {
        const bool in_range = lane < remaining;
        u32 state;
        int first;
        Carry3 val{0, 0.0, 0.0};
        while (true) {
            state = 2u;
            if (in_range) state = status3_read(status + 6 * (look - lane), epoch, val);
            const u32 ready = __ballot_sync(FULL_MASK, state != 0u);
            const u32 incl = __ballot_sync(FULL_MASK, state == 2u);
            first = __ffs(incl) - 1;
            const u32 need = first < 0 || first == 31 ? 0xffffffffu : ((2u << first) - 1u);
            if ((ready & need) == need) break;
        }
        const bool use = in_range && (first < 0 || lane <= first);
        int c = use ? val.c : 0;
        double l = use ? val.l : 0.0, r = use ? val.r : 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            c += __shfl_xor_sync(FULL_MASK, c, m);
            l += shfl_xor_d(l, m);
            r += shfl_xor_d(r, m);
        }
        excl.c += c;
        excl.l += l;
        excl.r += r;
        if (first >= 0) break;
        look -= 32;
        remaining -= 32;
    }
    if (lane == 0) status3_publish(status + 6 * idx, epoch, 2u, excl.c + tile.c, excl.l + tile.l, excl.r + tile.r);
    return excl;
}
#endif  // __CUDACC__

int launch_children_scan(b200dt_ctx *ctx, const int *ord, const double *ypay, int64_t col_stride, int F_search,
                         const PlanSeg *plan, const TileDesc *tiles, int n_tiles, int64_t n_elems, const u32 *mask,
                         u64 *status, u32 epoch, Partial *partials, int variant);
int launch_mark_left_virtual(b200dt_ctx *ctx, const int *orders, int64_t col_stride, const int *marks /* [n][4] */,
                             int n_marks, int max_span, const u32 *parent_mask, u32 *mask);
int launch_partition4(b200dt_ctx *ctx, const int *orders_in, int *orders_out, const double *y_in, double *y_out,
                      int64_t col_stride, int F_store, const void *d_plan4, const TileDesc *tiles, int n_tiles,
                      int64_t n_elems, const u32 *mask_a, const u32 *mask_b, int64_t N, u32 *codes, u64 *status,
                      u32 epoch);
int launch_mul(b200dt_ctx *ctx, const double *a, const double *b, int64_t n, double *out);
int launch_l2w_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles, int64_t n_elems,
                    u64 *status, u32 epoch, Partial *partials);
int launch_l2w_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small, int64_t n_elems,
                     int total_col, double *ty, double *tw, Partial *partials);
int launch_gather_payload(b200dt_ctx *ctx, const int *orders, const double *y, double *ypay, int64_t col_stride,
                          int F_store, int64_t N);

int launch_l2_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles,
                   int64_t n_elems, ScanStatus st, u32 epoch, u32 *ticket, Partial *partials, int variant,
                   bool bulk = false);
// tile size the search of this session uses (the streaming kernel takes larger tiles)
static inline int l2_scan_tile(bool payload) { return payload ? ST_TILE : WT_TILE; }
int launch_l2_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small,
                    int64_t n_elems, int total_col, double *exact_total, Partial *partials, int variant,
                    bool totals_given = false);
int launch_l2_exact_totals(b200dt_ctx *ctx, const LevelArgs &a, const int *list, int n_list, int64_t n_elems,
                           int total_col, double *exact_total);
int launch_clear_partials(b200dt_ctx *ctx, const LevelArgs &a, const int *list, int n_list, Partial *partials);
int launch_emit_totals(b200dt_ctx *ctx, const int *list, int n_list, const double *exact_total, int owner_col,
                       b200dt_candidate *out);
constexpr int FIN_SLICES_HOST = 64;  // == FIN_SLICES in l2.cu
int launch_finalize(b200dt_ctx *ctx, const LevelArgs &a, const Partial *partials, int64_t max_records,
                    Partial *sliced, const double *X, int64_t ldx, int col_offset, b200dt_candidate *out,
                    bool flat_rule = false);
int launch_mark_left(b200dt_ctx *ctx, const int *orders, int64_t col_stride, const int *marks /* [n][3]: f,start,k */,
                     int n_marks, int max_left, u32 *mask);
int launch_partition(b200dt_ctx *ctx, const int *orders_in, int *orders_out, int64_t col_stride, int F_store,
                     const SegDev *segs, const TileDesc *tiles, int n_tiles, int64_t n_elems, const u32 *mask,
                     u64 *status, u32 epoch, u32 *ticket, const double *y_in = nullptr, double *y_out = nullptr,
                     const double *w_in = nullptr, double *w_out = nullptr, int packed_row_mask = 0,
                     const int *newid = nullptr);
int launch_relabel(b200dt_ctx *ctx, const int *col, const SegDev *d_segs, int n_segs, const TileDesc *d_tiles,
                   int n_tiles, const u32 *mask, int64_t N, const int *d_bin_start, unsigned short *d_bins,
                   int *d_counts, int *d_newid, const int *d_map_old, int *d_map_new);
int relabel_chunks(int64_t N);
