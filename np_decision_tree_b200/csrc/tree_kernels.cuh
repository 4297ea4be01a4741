// Device-side records and host launchers of the per-level tree kernels (K2, K4, finalize).
#pragma once
#include "common.cuh"

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;  // 2048 sorted positions per tile

// One node of the current level = one contiguous segment [start, start+n) of every column of
// the order arrays.
struct SegDev {
    double total;   // sum of y over the node (approximate path); exact path recomputes it
    int start;
    int n;
    int pbase;      // first slot of this node's partial candidates
    int np;         // number of slots (tiles, or 1 for a sequentially scanned node)
    int nleft;      // partition: rows going left
    int exact;      // 1: searched by the sequential-order kernel
};

struct TileDesc {
    int seg;  // index into the SegDev array
    int tin;  // tile index inside the segment
};

// Look-back state of the tiled scans.  flag word = (epoch << 2) | state, state 1 = aggregate
// published, 2 = inclusive prefix published; epochs make clearing between launches unnecessary.
struct ScanStatus {
    u32 *flag;
    double *agg;
    double *incl;
};

struct LevelArgs {
    const int *orders;      // [F_store][col_stride] row ids, feature-major
    int64_t col_stride;
    int F_search;           // columns searched (local feature slice)
    const double *y;
    const SegDev *segs;
    int n_segs;
};

int launch_l2_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles,
                   int64_t n_elems, ScanStatus st, u32 epoch, u32 *ticket, Partial *partials, int variant);
int launch_l2_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small,
                    int64_t n_elems, int total_col, double *exact_total, Partial *partials, int variant);
int launch_finalize(b200dt_ctx *ctx, const LevelArgs &a, const Partial *partials, const double *X,
                    int64_t ldx, int col_offset, b200dt_candidate *out);
int launch_mark_left(b200dt_ctx *ctx, const int *orders, int64_t col_stride, const int *marks /* [n][3]: f,start,k */,
                     int n_marks, int max_left, u32 *mask);
int launch_partition(b200dt_ctx *ctx, const int *orders_in, int *orders_out, int64_t col_stride, int F_store,
                     const SegDev *segs, const TileDesc *tiles, int n_tiles, int64_t n_elems, const u32 *mask,
                     u64 *status, u32 epoch, u32 *ticket);
