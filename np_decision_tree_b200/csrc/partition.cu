// K4: stable partition of every column of every split node, and the left-row side mask.
//
// replaces split_orders(orders, index, n_samples)      (ref decision_tree/one.py:110-118)
//
// The reference builds a boolean (N,) array with the left rows set, gathers it through
// `orders` and compacts each column with boolean indexing (stable).  Here the mask is an
// N-bit array (L1/L2 resident), and one warp per (tile, column) computes the destination
// of 512 row ids from ballot counts plus a decoupled look-back over the left-counts of the
// earlier tiles of the same (node, column).  The children are the contiguous ranges
// [start, start+nleft) and [start+nleft, start+n) of the output array, so every node of the
// next level is again one segment shared by all columns.
// Algorithmic bytes: 4 (row id in) + 1 (side flag) + 4 (row id out) = 9 per element.
#include <algorithm>
#include <cstdlib>

#include "tree_kernels.cuh"

namespace {

__global__ void __launch_bounds__(256) mark_left_kernel(const int *__restrict__ orders, int64_t col_stride,
                                                        const int *__restrict__ marks, u32 *mask) {
    const int m = blockIdx.y;
    const int f = marks[3 * m], start = marks[3 * m + 1], k = marks[3 * m + 2];
    const int *col = orders + (int64_t)f * col_stride + start;  // one.py:106: first k+1 rows of column f
    for (int j = blockIdx.x * 256 + threadIdx.x; j <= k; j += gridDim.x * 256) {
        const int row = col[j];
        atomicOr(&mask[row >> 5], 1u << (row & 31));
    }
}


// One BLOCK owns one tile of PT_TILE = 8 x PT_WARP consecutive positions of one (node, column); warp w moves the
// PT_WARP positions [w * PT_WARP, (w+1) * PT_WARP) of it.  Every lane sees every ballot, so a warp's left-count
// lives in registers; the eight counts meet in shared memory and ONE decoupled look-back per block (warp 0)
// supplies the left-count of the earlier tiles of the same (node, column).  With one look-back chain entry per
// 4096 positions the number of in-flight entries per column stays small (resident blocks / columns) even when a
// rank owns few columns (feature-sharded fits: F_local = 32 at 8 GPUs), where per-warp entries made every tile
// poll ~140 predecessors.  Persistent grid, static tile-major order (see l2.cu).
// RELABEL: this level also hands out the new, node-contiguous row ids (relabel.cu): the look-up reads the
// row's new id instead of its side bit, the side is `new id < first id of the right child`, and the new id
// is what gets written.
#ifndef PT_YG
#define PT_YG 8
#endif
constexpr int PT_THREADS = 32 * PT_WARPS;
// y (PAY == 1) is requested PT_YG items at a time AFTER the look-back.  Measured (profiles/r02_partition_variants.md):
// 8 at a time beats 4 (13.4 vs 14.2 ms per level, 16 spill bytes); anything that requests y EARLIER is slower -- the
// first group before the barrier (15.1 ms), all of it with the row ids (8 items x 16 warps: 16.5 ms), or a bulk
// asynchronous copy into shared memory at the start of the tile (17 ms).

template <int PAY, bool RELABEL>
__global__ void __launch_bounds__(PT_THREADS, 1024 / PT_THREADS) partition_kernel(
    const int *__restrict__ in, int *__restrict__ out, const double *__restrict__ y_in, double *__restrict__ y_out,
    const double *__restrict__ w_in, double *__restrict__ w_out, int64_t col_stride, const SegDev *__restrict__ segs,
    const TileDesc *__restrict__ tiles, int n_tiles, const u32 *__restrict__ mask, u64 *status, u32 epoch, int n_cols,
    int row_mask, const int *__restrict__ newid) {
    __shared__ int s_cnt[2][PT_WARPS];
    __shared__ int s_excl[2];
    constexpr int YG = PAY == 1 ? PT_YG : 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    const int64_t total_tiles = (int64_t)n_tiles * n_cols;
    int par = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, par ^= 1) {
        const int f = (int)(t % n_cols), ti = (int)(t / n_cols);
        // (fetching the next tile's descriptors early costs registers the kernel does not have: 14.9 vs 12.8 ms)
        const TileDesc td = tiles[ti];
        const SegDev sg = segs[td.seg];
        const int off = td.tin * PT_TILE + warp * PT_WARP;
        const int cnt = max(0, min(PT_WARP, sg.n - off));     // 0: this warp's part lies beyond the node's end
        const int64_t cbase = (int64_t)f * col_stride + sg.start;
        const int *col = in + cbase + off;
        const double *ycol = (PAY == 1 || PAY == 2) ? y_in + cbase + off : nullptr;
        int rows[PT_ITEMS];
        if (cnt == PT_WARP) {
#pragma unroll
            for (int i = 0; i < PT_ITEMS; ++i) rows[i] = ld_stream_i32(col + i * 32 + lane);
        } else {
#pragma unroll
            for (int i = 0; i < PT_ITEMS; ++i) rows[i] = (i * 32 + lane < cnt) ? ld_stream_i32(col + i * 32 + lane) : -1;
        }
        u32 leftbits = 0;  // bit i: item i of this lane goes left
#pragma unroll
        for (int i = 0; i < PT_ITEMS; ++i) {
            // PAY == 3: the class id rides in the bits above row_mask (classes.cu), so the sign bit is data
            const int row = PAY == 3 ? (rows[i] & row_mask) : rows[i];
            const bool ok = PAY == 3 ? (i * 32 + lane < cnt) : (row >= 0);
            bool is_left;
            if (RELABEL) {
                DCHECK(!ok || row >= 0);
                const int nid = ok ? __ldg(newid + row) : 0x7fffffff;
                DCHECK(!ok || (nid >= sg.start && nid < sg.start + sg.n));
                is_left = nid < sg.start + sg.nleft;   // the left child takes the ids [start, start + nleft)
                rows[i] = nid;
            } else {
                is_left = ok && ((__ldg(mask + (row >> 5)) >> (row & 31)) & 1u);
            }
            leftbits |= (is_left ? 1u : 0u) << i;
        }
        // every lane sees every ballot, so the warp's left count is register arithmetic; the ballots
        // are recomputed in the write loop instead of being kept (registers -> occupancy)
        int run = 0;
#pragma unroll
        for (int i = 0; i < PT_ITEMS; ++i) run += __popc(__ballot_sync(FULL_MASK, (leftbits >> i) & 1u));
        if (lane == 0) s_cnt[par][warp] = run;
        __syncthreads();
        if (warp == 0) {
            int total = lane < PT_WARPS ? s_cnt[par][lane] : 0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) total += __shfl_xor_sync(FULL_MASK, total, m);
            const int excl = lookback_count(status, (int64_t)f * n_tiles + ti, td.tin, epoch, total, lane);
            if (lane == 0) s_excl[par] = excl;
        }
        __syncthreads();
        int lb_row = s_excl[par];     // left rows of the node before this warp's first position
#pragma unroll
        for (int w = 0; w < PT_WARPS - 1; ++w) lb_row += w < warp ? s_cnt[par][w] : 0;
        // (the slots of this parity are rewritten two tiles later, after every warp has passed the barriers of the
        // next tile, i.e. after these reads)
        int *ocol = out + cbase;
        double *oy = (PAY == 1 || PAY == 2) ? y_out + cbase : nullptr;
        const double *wcol = PAY == 2 ? w_in + cbase + off : nullptr;
        double *ow = PAY == 2 ? w_out + cbase : nullptr;
        if (cnt > 0) {
#pragma unroll
            for (int i0 = 0; i0 < PT_ITEMS; i0 += YG) {
                double yv[YG], wv[YG];
                if (PAY == 1 || PAY == 2) {
#pragma unroll
                    for (int u = 0; u < YG; ++u) {
                        const int e = (i0 + u) * 32 + lane;
                        yv[u] = e < cnt ? ld_stream_f64(ycol + e) : 0.0;
                        if (PAY == 2) wv[u] = e < cnt ? __ldcs(wcol + e) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < YG; ++u) {
                    const int i = i0 + u;
                    const u32 b = __ballot_sync(FULL_MASK, (leftbits >> i) & 1u);
                    const int e = i * 32 + lane;
                    if (e < cnt) {
                        const int lb = lb_row + __popc(b & lt);  // left rows before this element
                        const int dest = ((leftbits >> i) & 1u) ? lb : sg.nleft + (off + e - lb);
                        DCHECK(dest >= 0 && dest < sg.n && lb <= sg.nleft && (((leftbits >> i) & 1u) ? lb < sg.nleft : true));
                        // streaming stores: the children are read by the NEXT kernel; keep L1 for the side mask
                        __stcs(ocol + dest, rows[i]);
                        if (PAY == 1 || PAY == 2) __stcs(oy + dest, yv[u]);
                        if (PAY == 2) __stcs(ow + dest, wv[u]);
                    }
                    lb_row += __popc(b);
                }
            }
        }
    }
}

// y travelling with the row ids: ypay[f][j] = y[orders[f][j]] (once per fit, after the sort)
__global__ void __launch_bounds__(256) gather_payload_kernel(const int *__restrict__ orders,
                                                             const double *__restrict__ y, double *__restrict__ ypay,
                                                             int64_t col_stride, int64_t N) {
    const int64_t cbase = (int64_t)blockIdx.y * col_stride;
    const int64_t step = (int64_t)gridDim.x * 256;
    // four independent (row id -> y) chains in flight per thread
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < N; j += 4 * step) {
        int r[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = (j + u * step < N) ? ld_stream_i32(orders + cbase + j + u * step) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = r[u] >= 0 ? __ldg(y + r[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r[u] >= 0) __stcs(ypay + cbase + j + u * step, v[u]);
    }
}

}  // namespace

int launch_gather_payload(b200dt_ctx *ctx, const int *orders, const double *y, double *ypay, int64_t col_stride,
                          int F_store, int64_t N) {
    if (N <= 0 || F_store <= 0) return 0;
    int gx = (int)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->sm_count * 16 / std::max(1, std::min(F_store, 16)) + 1);
    LaunchScope ls(ctx, KC_PAYLOAD, 20.0 * (double)N * F_store);
    gather_payload_kernel<<<dim3((unsigned)gx, (unsigned)F_store), 256, 0, ctx->stream>>>(orders, y, ypay, col_stride, N);
    CK(cudaGetLastError());
    return 0;
}

int launch_mark_left(b200dt_ctx *ctx, const int *orders, int64_t col_stride, const int *marks, int n_marks,
                     int max_left, u32 *mask) {
    if (n_marks <= 0) return 0;
    int gx = (max_left + 255) / 256;
    if (gx > 2048) gx = 2048;
    if (gx < 1) gx = 1;
    for (int m0 = 0; m0 < n_marks; m0 += 65535) {
        const int nm = n_marks - m0 < 65535 ? n_marks - m0 : 65535;
        LaunchScope ls(ctx, KC_MARK, 0.0);
        mark_left_kernel<<<dim3(gx, nm), 256, 0, ctx->stream>>>(orders, col_stride, marks + 3 * m0, mask);
    }
    CK(cudaGetLastError());
    return 0;
}

int launch_partition(b200dt_ctx *ctx, const int *orders_in, int *orders_out, int64_t col_stride, int F_store,
                     const SegDev *segs, const TileDesc *tiles, int n_tiles, int64_t n_elems, const u32 *mask,
                     u64 *status, u32 epoch, u32 *ticket, const double *y_in, double *y_out, const double *w_in,
                     double *w_out, int packed_row_mask, const int *newid) {
    (void)ticket;
    if (n_tiles <= 0) return 0;
    const int pay = packed_row_mask ? 3 : (y_in && y_out) ? ((w_in && w_out) ? 2 : 1) : 0;
    const int64_t total = (int64_t)n_tiles * F_store;
    int per_sm = 0;
    if (newid && pay != 1 && pay != 2) return ctx->fail_msg("partition: row relabelling needs a payload session");
    const void *kern = pay == 3 ? (const void *)partition_kernel<3, false>
                       : pay == 2 ? (newid ? (const void *)partition_kernel<2, true> : (const void *)partition_kernel<2, false>)
                       : pay == 1 ? (newid ? (const void *)partition_kernel<1, true> : (const void *)partition_kernel<1, false>)
                                  : (const void *)partition_kernel<0, false>;
    // the kernel's only use of the unified L1 / shared memory is the side-mask window: ask for all of it as L1
    // (12.8 instead of 13.3 ms per level)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * ctx->sm_count;  // all blocks resident: tiles wait on earlier tiles
    if (grid > total) grid = total;
    LaunchScope ls(ctx, KC_PARTITION, 9.0 * (double)n_elems * F_store);
    const double *yi = (pay == 1 || pay == 2) ? y_in : nullptr, *wi = pay == 2 ? w_in : nullptr;
    double *yo = (pay == 1 || pay == 2) ? y_out : nullptr, *wo = pay == 2 ? w_out : nullptr;
    void *args[] = {&orders_in, &orders_out, &yi, &yo, &wi, &wo, &col_stride, &segs, &tiles, &n_tiles, &mask, &status,
                    &epoch, &F_store, &packed_row_mask, &newid};
    CK(launch_resident(kern, (unsigned)grid, PT_THREADS, args, 0, ctx->stream));
    return 0;
}
