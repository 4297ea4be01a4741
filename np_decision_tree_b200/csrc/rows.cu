// K7: node-at-a-time primitives for the random-threshold (ExtraTree-style) classification splitters.
//
// replaces the array work of random_classify / random_classify_p_at_k        (ref decision_tree/strategy.py:149-172)
//          and of build_tree's row masks                                        (ref tree_builder.py:24-41, 56-96)
//
// The reference draws its thresholds from numpy's GLOBAL generator, per node, in depth-first order:
//     thresholds = np.random.uniform(np.min(X, axis=0), np.max(X, axis=0));  masks = X <= thresholds
//     left_sums = masks.T.dot(y);  left_counts = masks.sum(axis=0)
// A level-synchronous builder would consume the generator in another order, so this path keeps the
// reference's node order: the host (Python, the same np.random calls) asks the GPU, per node, for
//   1. the column minima / maxima over the node's rows            (exact selections),
//   2. the class counts of the rows with x <= threshold, per column  (whole numbers),
//   3. the stable split of the node's row list by the chosen column.
// A node is a contiguous segment [start, start + n) of a row-id array; the children are written to
// the same segment of the alternate array (left rows first), so no allocation happens per node.
// Rows are read whole (row-major X: one coalesced 8 * F-byte read per row).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include <math_constants.h>

#include "session.cuh"

struct RowSession {
    int64_t N = 0, ldx = 0;
    int F = 0, C = 0;
    const double *dX = nullptr;
    const int *dlab = nullptr;
    const double *dy = nullptr;   // regression targets (b200dt_rows_set_targets)
    int *rows[2] = {nullptr, nullptr};
};

void row_session_free(b200dt_ctx *ctx) {
    delete ctx->row_session;
    ctx->row_session = nullptr;
}

namespace {

// order-preserving 64-bit image of a double (for atomicMin / atomicMax)
__device__ __forceinline__ u64 ordered_key(double x) {
    const u64 b = (u64)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double ordered_value(u64 k) {
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double x;
#ifdef __CUDA_ARCH__
    x = __longlong_as_double((long long)b);
#else
    memcpy(&x, &b, 8);
#endif
    return x;
}

__global__ void iota_kernel(int *rows, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rows[i] = (int)i;
}

__global__ void __launch_bounds__(256) label_hist_kernel(const int *__restrict__ labels, int64_t N, int C,
                                                         unsigned long long *hist, int *bad) {
    extern __shared__ int sh[];
    for (int c = threadIdx.x; c < C; c += 256) sh[c] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (int64_t)gridDim.x * 256) {
        const int c = labels[i];
        if (c < 0 || c >= C)
            *bad = 1;
        else
            atomicAdd(&sh[c], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256)
        if (sh[c]) atomicAdd(hist + c, (unsigned long long)sh[c]);
}

// thread (ty, tx): rows ty, ty + RPB, ... of the block's share; column cols[c0 + tx]
template <bool COUNTS>
__global__ void __launch_bounds__(256) rows_columns_kernel(const double *__restrict__ X, int64_t ldx,
                                                           const int *__restrict__ rows, int64_t n,
                                                           const int *__restrict__ cols, int n_cols, int cw,
                                                           u64 *kmin, u64 *kmax,                       // !COUNTS
                                                           const int *__restrict__ labels, int C,
                                                           const double *__restrict__ thr,
                                                           unsigned long long *counts) {                // COUNTS
    extern __shared__ int sh[];   // COUNTS: [C][n_cols]
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, rpb = 256 / cw;
    if (COUNTS) {
        for (int i = threadIdx.x; i < n_cols * C; i += 256) sh[i] = 0;
        __syncthreads();
    }
    for (int c0 = 0; c0 < n_cols; c0 += cw) {
        const int ci = c0 + tx;
        const bool has_col = ci < n_cols;
        const int col = has_col ? cols[ci] : 0;
        const double t = (COUNTS && has_col) ? thr[ci] : 0.0;
        double lo = CUDART_INF, hi = -CUDART_INF;
        const int64_t step = (int64_t)gridDim.x * rpb;
        // four rows in flight per thread: the loads of a row do not depend on the previous row
        for (int64_t r = (int64_t)blockIdx.x * rpb + ty; r < n; r += 4 * step) {
            int row[4];
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) row[u] = (r + u * step < n) ? rows[r + u * step] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = (row[u] >= 0 && has_col) ? X[(int64_t)row[u] * ldx + col] : CUDART_NAN;
            if (COUNTS) {
                int lab[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) lab[u] = (row[u] >= 0 && has_col) ? labels[row[u]] : 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)   // strategy.py:151-152: masks = X <= thresholds (NaN compares false)
                    if (x[u] <= t) atomicAdd(&sh[lab[u] * n_cols + ci], 1);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    lo = fmin(lo, x[u]);   // fmin / fmax ignore the NaN of an absent row
                    hi = fmax(hi, x[u]);
                }
            }
        }
        if (!COUNTS && has_col && lo <= hi) {
            atomicMin(kmin + ci, ordered_key(lo));
            atomicMax(kmax + ci, ordered_key(hi));
        }
    }
    if (COUNTS) {
        __syncthreads();
        // shared layout is [class][column] (consecutive threads = consecutive columns: no bank conflicts)
        for (int i = threadIdx.x; i < n_cols * C; i += 256)
            if (sh[i]) atomicAdd(counts + (size_t)(i % n_cols) * C + i / n_cols, (unsigned long long)sh[i]);
    }
}

// regression: per column, sum of y and number of the node's rows with x <= threshold; plus the node's own
// sum / min / max of y (accumulated by the threads of column slot 0).
// The sums are floating point, so their ORDER matters for one thing: two columns whose thresholds cut off the same
// rows must get bit-identical sums (the reference's BLAS product sums every column in the same row order, ties
// exactly and np.argmax takes the first).  So every thread sums its fixed share of the rows in list order into a
// partial slot, and a second kernel adds the slots in slot order: the order depends on the rows only, never on
// the column.
__global__ void __launch_bounds__(256) rows_sums_kernel(const double *__restrict__ X, int64_t ldx,
                                                        const int *__restrict__ rows, int64_t n,
                                                        const int *__restrict__ cols, int n_cols, int cw,
                                                        const double *__restrict__ y, const double *__restrict__ thr,
                                                        double *__restrict__ psum, int *__restrict__ pcnt,
                                                        double *__restrict__ pnode) {
    __shared__ double s_sum[256], s_ns[256], s_lo[256], s_hi[256];
    __shared__ int s_cnt[256];
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, rpb = 256 / cw;
    const int64_t step = (int64_t)gridDim.x * rpb;
    const int64_t slot = blockIdx.x;   // one partial slot per block: its row groups are combined below, in order
    for (int c0 = 0; c0 < n_cols; c0 += cw) {
        const int ci = c0 + tx;
        const bool has_col = ci < n_cols;
        const int col = has_col ? cols[ci] : 0;
        const double t = has_col ? thr[ci] : 0.0;
        const bool node_stats = ci == 0;
        double s = 0.0, ns = 0.0, lo = CUDART_INF, hi = -CUDART_INF;
        int c = 0;
        for (int64_t r = (int64_t)blockIdx.x * rpb + ty; r < n; r += 4 * step) {
            int row[4];
            double x[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) row[u] = (r + u * step < n) ? rows[r + u * step] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[u] = (row[u] >= 0 && has_col) ? X[(int64_t)row[u] * ldx + col] : CUDART_NAN;
                yv[u] = (row[u] >= 0 && has_col) ? y[row[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (x[u] <= t) {   // strategy.py:138: masks = X <= thresholds (NaN compares false)
                    s += yv[u];
                    ++c;
                }
                if (node_stats && row[u] >= 0) {
                    ns += yv[u];
                    lo = fmin(lo, yv[u]);
                    hi = fmax(hi, yv[u]);
                }
            }
        }
        s_sum[threadIdx.x] = s;
        s_cnt[threadIdx.x] = c;
        if (node_stats) {
            s_ns[ty] = ns;
            s_lo[ty] = lo;
            s_hi[ty] = hi;
        }
        __syncthreads();
        if (ty == 0 && has_col) {   // row groups in group order: the order depends on the rows only
            double bs = 0.0;
            int bc = 0;
            for (int g = 0; g < rpb; ++g) {
                bs += s_sum[g * cw + tx];
                bc += s_cnt[g * cw + tx];
            }
            psum[slot * n_cols + ci] = bs;
            pcnt[slot * n_cols + ci] = bc;
            if (node_stats) {
                double bn = 0.0, bl = CUDART_INF, bh = -CUDART_INF;
                for (int g = 0; g < rpb; ++g) {
                    bn += s_ns[g];
                    bl = fmin(bl, s_lo[g]);
                    bh = fmax(bh, s_hi[g]);
                }
                pnode[slot * 3] = bn;
                pnode[slot * 3 + 1] = bl;
                pnode[slot * 3 + 2] = bh;
            }
        }
        __syncthreads();
    }
}

// slots are added in slot order; thread per column, thread 0 of block 0 also reduces the node statistics
__global__ void __launch_bounds__(256) rows_sums_reduce_kernel(const double *__restrict__ psum,
                                                               const int *__restrict__ pcnt,
                                                               const double *__restrict__ pnode, int64_t n_slots,
                                                               int n_cols, double *__restrict__ res) {
    const int ci = blockIdx.x * 256 + threadIdx.x;
    if (ci < n_cols) {
        double s = 0.0;
        long long c = 0;
        for (int64_t p = 0; p < n_slots; ++p) {
            s += psum[p * n_cols + ci];
            c += pcnt[p * n_cols + ci];
        }
        res[ci] = s;
        reinterpret_cast<long long *>(res)[n_cols + ci] = c;
    }
    if (ci == 0) {
        double ns = 0.0, lo = CUDART_INF, hi = -CUDART_INF;
        for (int64_t p = 0; p < n_slots; ++p) {
            ns += pnode[p * 3];
            lo = fmin(lo, pnode[p * 3 + 1]);
            hi = fmax(hi, pnode[p * 3 + 2]);
        }
        res[2 * n_cols] = ns;
        res[2 * n_cols + 1] = lo;
        res[2 * n_cols + 2] = hi;
    }
}

constexpr int SPLIT_BLOCK = 1024;  // rows per block of the split kernels (256 threads x 4)

__global__ void __launch_bounds__(256) rows_split_count_kernel(const double *__restrict__ X, int64_t ldx,
                                                               const int *__restrict__ rows, int64_t n, int col,
                                                               double thr, int *blk_left) {
    __shared__ int sw[8];
    const int64_t base = (int64_t)blockIdx.x * SPLIT_BLOCK;
    int c = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t i = base + u * 256 + threadIdx.x;
        if (i < n && X[(int64_t)rows[i] * ldx + col] <= thr) ++c;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(FULL_MASK, c, m);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sw[w];
        blk_left[blockIdx.x] = t;
    }
}

// exclusive scan of the per-block left counts (one block; n_blocks <= a few 10^4)
__global__ void __launch_bounds__(1024) rows_split_scan_kernel(int *blk_left, int n_blocks, long long *total) {
    __shared__ long long carry;
    __shared__ int sw[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < n_blocks ? blk_left[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, inc, d);
            if ((threadIdx.x & 31) >= d) inc += u;
        }
        if ((threadIdx.x & 31) == 31) sw[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = sw[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, w, d);
                if (threadIdx.x >= d) w += u;
            }
            sw[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? sw[(threadIdx.x >> 5) - 1] : 0;
        if (i < n_blocks) blk_left[i] = (int)(carry + warp_off + inc - v);
        __syncthreads();
        if (threadIdx.x == 1023) carry += warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// stable: left rows keep their order in [0, n_left), right rows theirs in [n_left, n)   (tree_builder.py:31-34)
__global__ void __launch_bounds__(256) rows_split_scatter_kernel(const double *__restrict__ X, int64_t ldx,
                                                                 const int *__restrict__ rows, int64_t n, int col,
                                                                 double thr, const int *__restrict__ blk_left,
                                                                 const long long *__restrict__ total,
                                                                 int *__restrict__ out) {
    __shared__ int sw[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SPLIT_BLOCK;
    const long long n_left = *total;
    int left_before = blk_left[blockIdx.x];
    for (int u = 0; u < 4; ++u) {
        const int64_t i = base + u * 256 + threadIdx.x;
        const bool ok = i < n;
        const int row = ok ? rows[i] : 0;
        const bool goes_left = ok && X[(int64_t)row * ldx + col] <= thr;
        const u32 b = __ballot_sync(FULL_MASK, goes_left);
        if (lane == 0) sw[warp] = __popc(b);
        __syncthreads();
        int warp_off = 0, round_total = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) warp_off += sw[w];
            round_total += sw[w];
        }
        if (ok) {
            const long long lb = left_before + warp_off + __popc(b & ((1u << lane) - 1u));   // left rows before i
            out[goes_left ? lb : n_left + (i - lb)] = row;
        }
        left_before += round_total;
        __syncthreads();
    }
}

int rows_cols_upload(b200dt_ctx *ctx, const int32_t *cols, int n_cols, const double *thr, int **d_cols, double **d_thr) {
    RowSession *s = ctx->row_session;
    if (n_cols < 1) return ctx->fail_msg("rows: no columns");
    for (int i = 0; i < n_cols; ++i)
        if (cols[i] < 0 || cols[i] >= s->F) return ctx->fail_msg("rows: column id out of range");
    char *stage, *dev;
    const size_t bytes = (size_t)n_cols * 12 + 64;
    CKR(pinned_get(ctx, PIN_PLAN, bytes, (void **)&stage));
    CKR(scratch_get(ctx, SB_PLAN, bytes, (void **)&dev));
    memcpy(stage, cols, (size_t)n_cols * 4);
    const size_t thr_off = ((size_t)n_cols * 4 + 15) / 16 * 16;
    if (thr) memcpy(stage + thr_off, thr, (size_t)n_cols * 8);
    CK(cudaMemcpyAsync(dev, stage, thr_off + (size_t)n_cols * 8, cudaMemcpyHostToDevice, ctx->stream));
    *d_cols = (int *)dev;
    *d_thr = (double *)(dev + thr_off);
    return 0;
}

int rows_check_segment(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n) {
    RowSession *s = ctx->row_session;
    if (!s) return ctx->fail_msg("rows: session not begun");
    if ((buffer != 0 && buffer != 1) || start < 0 || n < 0 || start + n > s->N)
        return ctx->fail_msg("rows: segment out of range");
    return 0;
}

int column_width(int n_cols) {   // power of two <= 256 covering the columns of one pass
    int cw = 1;
    while (cw < n_cols && cw < 256) cw <<= 1;
    return cw;
}

}  // namespace

extern "C" {

int b200dt_rows_begin(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F, int64_t ldx,
                      int space, int n_classes, int64_t *class_counts) {
    if (!ctx) return 1;
    if (!X || !labels || !class_counts) return ctx->fail_msg("rows_begin: null argument");
    if (N <= 0 || F <= 0) return ctx->fail_msg("rows_begin: empty input");
    if (N >= (1ll << 31)) return ctx->fail_msg("rows_begin: N must be < 2^31");
    if (n_classes < 1 || n_classes > 4096) return ctx->fail_msg("rows_begin: n_classes must be in [1, 4096]");
    CK(cudaSetDevice(ctx->device));
    end_all_sessions(ctx);
    RowSession *s = new (std::nothrow) RowSession();
    if (!s) return ctx->fail_msg("rows_begin: out of host memory");
    ctx->row_session = s;
    s->N = N;
    s->F = (int)F;
    s->C = n_classes;
    if (space == B200DT_HOST) {
        double *bx;
        int *bl;
        CKR(scratch_get(ctx, SB_X, (size_t)N * F * 8, (void **)&bx));
        CKR(scratch_get(ctx, SB_Y, (size_t)N * 8, (void **)&bl));
        CK(cudaMemcpy2DAsync(bx, (size_t)F * 8, X, (size_t)ldx * 8, (size_t)F * 8, (size_t)N, cudaMemcpyHostToDevice,
                             ctx->stream));
        CK(cudaMemcpyAsync(bl, labels, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
        s->dX = bx;
        s->dlab = bl;
        s->ldx = F;
    } else {
        s->dX = X;
        s->dlab = labels;
        s->ldx = ldx;
    }
    CKR(scratch_get(ctx, SB_ORD_A, (size_t)N * 4 + 64, (void **)&s->rows[0]));
    CKR(scratch_get(ctx, SB_ORD_B, (size_t)N * 4 + 64, (void **)&s->rows[1]));
    unsigned long long *d_hist;
    CKR(scratch_get(ctx, SB_MISC, (size_t)n_classes * 8 + 64, (void **)&d_hist));
    CK(cudaMemsetAsync(d_hist, 0, (size_t)n_classes * 8 + 64, ctx->stream));
    const unsigned blocks = (unsigned)std::min<int64_t>(1024, (N + 255) / 256);
    {
        LaunchScope ls(ctx, KC_MISC, 8.0 * (double)N, 2);
        iota_kernel<<<blocks, 256, 0, ctx->stream>>>(s->rows[0], N);
        label_hist_kernel<<<blocks, 256, (size_t)n_classes * 4, ctx->stream>>>(s->dlab, N, n_classes, d_hist,
                                                                                (int *)(d_hist + n_classes));
    }
    CK(cudaGetLastError());
    std::vector<unsigned long long> h(n_classes + 1);
    CK(cudaMemcpyAsync(h.data(), d_hist, (size_t)(n_classes + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((int)h[n_classes] != 0) return ctx->fail_msg("rows_begin: labels must be integers in [0, n_classes)");
    for (int c = 0; c < n_classes; ++c) class_counts[c] = (int64_t)h[c];
    return 0;
}

int b200dt_rows_minmax(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, const int32_t *cols, int n_cols,
                       double *out_min, double *out_max) {
    if (!ctx) return 1;
    if (!cols || !out_min || !out_max) return ctx->fail_msg("rows_minmax: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(rows_check_segment(ctx, buffer, start, n));
    if (n < 1) return ctx->fail_msg("rows_minmax: zero-size array to reduction operation minimum which has no identity");
    RowSession *s = ctx->row_session;
    int *d_cols;
    double *d_thr;
    CKR(rows_cols_upload(ctx, cols, n_cols, nullptr, &d_cols, &d_thr));
    u64 *d_keys;
    CKR(scratch_get(ctx, SB_RESULT, (size_t)n_cols * 16 + 64, (void **)&d_keys));
    CK(cudaMemsetAsync(d_keys, 0xff, (size_t)n_cols * 8, ctx->stream));
    CK(cudaMemsetAsync(d_keys + n_cols, 0, (size_t)n_cols * 8, ctx->stream));
    const int cw = column_width(n_cols), rpb = 256 / cw;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + rpb - 1) / rpb, (int64_t)ctx->sm_count * 8));
    {
        LaunchScope ls(ctx, KC_ROWS, 8.0 * (double)n * n_cols);
        rows_columns_kernel<false><<<grid, 256, 0, ctx->stream>>>(s->dX, s->ldx, s->rows[buffer] + start, n, d_cols, n_cols,
                                                                  cw, d_keys, d_keys + n_cols, nullptr, 0, nullptr,
                                                                  nullptr);
    }
    CK(cudaGetLastError());
    u64 *h;
    CKR(pinned_get(ctx, PIN_RESULT, (size_t)n_cols * 16, (void **)&h));
    CK(cudaMemcpyAsync(h, d_keys, (size_t)n_cols * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_cols; ++i) {
        out_min[i] = ordered_value(h[i]);
        out_max[i] = ordered_value(h[n_cols + i]);
    }
    return 0;
}

int b200dt_rows_threshold_counts(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, const int32_t *cols, int n_cols,
                                 const double *thresholds, int64_t *out_counts) {
    if (!ctx) return 1;
    if (!cols || !thresholds || !out_counts) return ctx->fail_msg("rows_threshold_counts: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(rows_check_segment(ctx, buffer, start, n));
    RowSession *s = ctx->row_session;
    const int C = s->C;
    int *d_cols;
    double *d_thr;
    CKR(rows_cols_upload(ctx, cols, n_cols, thresholds, &d_cols, &d_thr));
    unsigned long long *d_counts;
    CKR(scratch_get(ctx, SB_RESULT, (size_t)n_cols * C * 8 + 64, (void **)&d_counts));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)n_cols * C * 8, ctx->stream));
    // shared-memory histogram [columns of one launch][C]: at most 40 K counters per launch
    const int cols_per_launch = std::max(1, std::min(n_cols, 40960 / C));
    if ((size_t)cols_per_launch * C * 4 > 48 * 1024)
        CK(cudaFuncSetAttribute(rows_columns_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
    for (int c0 = 0; c0 < n_cols && n > 0; c0 += cols_per_launch) {
        const int nc = std::min(cols_per_launch, n_cols - c0);
        const int cw = column_width(nc), rpb = 256 / cw;
        const unsigned grid =
            (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + rpb - 1) / rpb, (int64_t)ctx->sm_count * 4));
        LaunchScope ls(ctx, KC_ROWS, 8.0 * (double)n * nc);
        rows_columns_kernel<true><<<grid, 256, (size_t)nc * C * 4, ctx->stream>>>(
            s->dX, s->ldx, s->rows[buffer] + start, n, d_cols + c0, nc, cw, nullptr, nullptr, s->dlab, C, d_thr + c0,
            d_counts + (size_t)c0 * C);
    }
    CK(cudaGetLastError());
    unsigned long long *h;
    CKR(pinned_get(ctx, PIN_RESULT, (size_t)n_cols * C * 8, (void **)&h));
    CK(cudaMemcpyAsync(h, d_counts, (size_t)n_cols * C * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < (size_t)n_cols * C; ++i) out_counts[i] = (int64_t)h[i];
    return 0;
}

int b200dt_rows_set_targets(b200dt_ctx *ctx, const double *y, int space) {
    if (!ctx) return 1;
    RowSession *s = ctx->row_session;
    if (!s) return ctx->fail_msg("rows: session not begun");
    if (!y) return ctx->fail_msg("rows_set_targets: null argument");
    CK(cudaSetDevice(ctx->device));
    if (space == B200DT_HOST) {
        double *buf;
        CKR(scratch_get(ctx, SB_YW, (size_t)s->N * 8 + 64, (void **)&buf));
        CK(cudaMemcpyAsync(buf, y, (size_t)s->N * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // the caller's array may be a temporary
        s->dy = buf;
    } else {
        s->dy = y;
    }
    return 0;
}

int b200dt_rows_threshold_sums(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, const int32_t *cols, int n_cols,
                               const double *thresholds, double *out_sums, int64_t *out_counts, double *out_node) {
    if (!ctx) return 1;
    if (!cols || !thresholds || !out_sums || !out_counts || !out_node)
        return ctx->fail_msg("rows_threshold_sums: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(rows_check_segment(ctx, buffer, start, n));
    RowSession *s = ctx->row_session;
    if (!s->dy) return ctx->fail_msg("rows_threshold_sums: no targets registered (b200dt_rows_set_targets)");
    if (n < 1) return ctx->fail_msg("rows_threshold_sums: empty node");
    int *d_cols;
    double *d_thr;
    CKR(rows_cols_upload(ctx, cols, n_cols, thresholds, &d_cols, &d_thr));
    // result block: sums[n_cols] | counts[n_cols] (int64) | node sum | node min | node max
    double *d_res;
    const size_t words = (size_t)2 * n_cols + 3;
    CKR(scratch_get(ctx, SB_RESULT, words * 8 + 64, (void **)&d_res));
    const int cw = column_width(n_cols), rpb = 256 / cw;
    // few slots: they are added serially by one thread per column (and the nodes of deep levels are small)
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 4 * rpb - 1) / (4 * rpb), (int64_t)ctx->sm_count * 4));
    const int64_t n_slots = (int64_t)grid;
    double *d_psum, *d_pnode;
    int *d_pcnt;
    CKR(scratch_get(ctx, SB_L1_A, (size_t)n_slots * n_cols * 8 + 64, (void **)&d_psum));
    CKR(scratch_get(ctx, SB_L1_B, (size_t)n_slots * n_cols * 4 + 64, (void **)&d_pcnt));
    CKR(scratch_get(ctx, SB_L1_C, (size_t)n_slots * 3 * 8 + 64, (void **)&d_pnode));
    {
        LaunchScope ls(ctx, KC_ROWS, 8.0 * (double)n * n_cols, 2);
        rows_sums_kernel<<<grid, 256, 0, ctx->stream>>>(s->dX, s->ldx, s->rows[buffer] + start, n, d_cols, n_cols, cw, s->dy,
                                                        d_thr, d_psum, d_pcnt, d_pnode);
        rows_sums_reduce_kernel<<<(n_cols + 255) / 256, 256, 0, ctx->stream>>>(d_psum, d_pcnt, d_pnode, n_slots, n_cols, d_res);
    }
    CK(cudaGetLastError());
    double *h;
    CKR(pinned_get(ctx, PIN_RESULT, words * 8, (void **)&h));
    CK(cudaMemcpyAsync(h, d_res, words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const long long *hk = reinterpret_cast<const long long *>(h);
    for (int i = 0; i < n_cols; ++i) {
        out_sums[i] = h[i];
        out_counts[i] = (int64_t)hk[n_cols + i];
    }
    out_node[0] = h[2 * n_cols];
    out_node[1] = h[2 * n_cols + 1];
    out_node[2] = h[2 * n_cols + 2];
    return 0;
}

int b200dt_rows_split(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, int32_t col, double threshold,
                      int64_t *n_left) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    CKR(rows_check_segment(ctx, buffer, start, n));
    RowSession *s = ctx->row_session;
    if (col < 0 || col >= s->F) return ctx->fail_msg("rows_split: column id out of range");
    if (n_left) *n_left = 0;
    if (n == 0) return 0;
    const int n_blocks = (int)((n + SPLIT_BLOCK - 1) / SPLIT_BLOCK);
    int *d_blk;
    const size_t total_off = ((size_t)n_blocks * 4 + 15) / 16 * 16;
    CKR(scratch_get(ctx, SB_TILES, total_off + 64, (void **)&d_blk));
    long long *d_total = (long long *)((char *)d_blk + total_off);
    const int *in = s->rows[buffer] + start;
    int *out = s->rows[buffer ^ 1] + start;
    {
        LaunchScope ls(ctx, KC_ROWS, 24.0 * (double)n, 3);
        rows_split_count_kernel<<<n_blocks, 256, 0, ctx->stream>>>(s->dX, s->ldx, in, n, col, threshold, d_blk);
        rows_split_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_blk, n_blocks, d_total);
        rows_split_scatter_kernel<<<n_blocks, 256, 0, ctx->stream>>>(s->dX, s->ldx, in, n, col, threshold, d_blk, d_total,
                                                                      out);
    }
    CK(cudaGetLastError());
    if (!n_left) return 0;   // the caller knows the count (threshold_counts / threshold_sums): no host round trip
    long long h_total = 0;
    CK(cudaMemcpyAsync(&h_total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_left = h_total;
    return 0;
}

int b200dt_rows_end(b200dt_ctx *ctx) {
    if (!ctx) return 1;
    row_session_free(ctx);
    return 0;
}

}  // extern "C"
