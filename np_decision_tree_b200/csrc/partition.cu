// K4: stable partition of every column of every split node, and the left-row side mask.
//
// replaces split_orders(orders, index, n_samples)      (ref decision_tree/one.py:110-118)
//
// The reference builds a boolean (N,) array with the left rows set, gathers it through
// `orders` and compacts each column with boolean indexing (stable).  Here the mask is an
// N-bit array (L1/L2 resident), and one block per (tile, column) computes the destination
// of 2048 row ids from ballot counts plus a decoupled look-back over the left-counts of the
// earlier tiles of the same (node, column).  The children are the contiguous ranges
// [start, start+nleft) and [start+nleft, start+n) of the output array, so every node of the
// next level is again one segment shared by all columns.
// Algorithmic bytes: 4 (row id in) + 1 (side flag) + 4 (row id out) = 9 per element.
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


__global__ void __launch_bounds__(SC_THREADS) partition_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                               int64_t col_stride, const SegDev *__restrict__ segs,
                                                               const TileDesc *__restrict__ tiles, int n_tiles,
                                                               const u32 *__restrict__ mask, u64 *status, u32 epoch,
                                                               u32 *ticket, int n_cols) {
    __shared__ int s_cnt[SC_ITEMS * 8];   // left count per (item row, warp), element order
    __shared__ int s_pref[SC_ITEMS * 8];
    __shared__ int s_before;
    __shared__ u32 s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 t = s_ticket;
    const int f = t % n_cols, ti = t / n_cols;  // tile-major order (see sort.cu)
    const TileDesc td = tiles[ti];
    const SegDev sg = segs[td.seg];
    const int off = td.tin * SC_TILE;
    const int cnt = min(SC_TILE, sg.n - off);
    const int *col = in + (int64_t)f * col_stride + sg.start + off;
    const u32 lt = lanemask_lt();

    int rows[SC_ITEMS];
    int myrank[SC_ITEMS];
    u32 leftbits = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const int e = i * SC_THREADS + tid;
        bool is_left = false;
        rows[i] = 0;
        if (e < cnt) {
            const int row = ld_stream_i32(col + e);
            rows[i] = row;
            is_left = (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
        }
        const u32 bal = __ballot_sync(FULL_MASK, is_left);
        if (lane == 0) s_cnt[i * 8 + warp] = __popc(bal);
        myrank[i] = __popc(bal & lt);
        leftbits |= (is_left ? 1u : 0u) << i;
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the 64 counts (element order = item row major, warp minor)
        const int c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
        const int sum = c0 + c1;
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        s_pref[2 * lane] = incl - sum;
        s_pref[2 * lane + 1] = incl - sum + c0;
        const int tile_left = __shfl_sync(FULL_MASK, incl, 31);
        const int before = lookback_count(status, (int64_t)f * n_tiles + ti, td.tin, epoch, tile_left, lane);
        if (lane == 0) s_before = before;
    }
    __syncthreads();
    const int before = s_before;
    int *ocol = out + (int64_t)f * col_stride + sg.start;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const int e = i * SC_THREADS + tid;
        if (e < cnt) {
            const int lb = before + s_pref[i * 8 + warp] + myrank[i];  // left rows before this element
            const int dest = ((leftbits >> i) & 1u) ? lb : sg.nleft + (off + e - lb);
            ocol[dest] = rows[i];
        }
    }
}

}  // namespace

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
                     u64 *status, u32 epoch, u32 *ticket) {
    if (n_tiles <= 0) return 0;
    CK(cudaMemsetAsync(ticket, 0, 4, ctx->stream));
    const int64_t grid = (int64_t)n_tiles * F_store;
    if (grid > 0x7fffffffLL) return ctx->fail_msg("partition: grid too large");
    LaunchScope ls(ctx, KC_PARTITION, 9.0 * (double)n_elems * F_store);
    partition_kernel<<<(unsigned)grid, SC_THREADS, 0, ctx->stream>>>(orders_in, orders_out, col_stride, segs, tiles,
                                                                      n_tiles, mask, status, epoch, ticket, F_store);
    CK(cudaGetLastError());
    return 0;
}
