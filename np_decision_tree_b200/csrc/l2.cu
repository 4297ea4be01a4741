// K2: L2 (MSE) exhaustive split search for every live node of a level, and the per-node
// arg-best reduction.
//
// replaces greedy_regression_split + best_variance_improvements[_v2] + np_gene
//          (ref decision_tree/one.py:52-63, 81-93, 96-107)
//
// Two kernels share the work:
//  * l2_scan_kernel  -- nodes with more than SC_TILE rows.  One block per (tile, column):
//    coalesced read of 2048 row ids, gather of y, block scan, carry-in from a decoupled
//    look-back over the earlier tiles of the same (node, column), gain at every position,
//    block arg-max.  Single pass: 12 algorithmic bytes per (row, feature) element.  The
//    parallel scan rounds differently from the reference's sequential np.cumsum, so each
//    candidate also carries the runner-up gain; a node whose two best gains are within
//    1e-9 relative is re-searched by the exact kernel.
//  * l2_exact_kernel -- small nodes (and flagged near-ties).  One warp per (node, column)
//    adds y in sorted order ONE ELEMENT AT A TIME in fp64 (shuffle broadcast), with the
//    reference's exact operation order and no FMA contraction, so bit-exact gain ties
//    between features resolve exactly as in the reference (hard part 1).
#include <math_constants.h>

#include "tree_kernels.cuh"

namespace {

struct Cand {
    double g;    // score (maximised)
    double g2;   // best score among the other candidates merged so far
    double s;    // prefix sum at k
    double dev;  // |S - (k+1) mean|
    int k;
};

__device__ __forceinline__ Cand cand_none() {
    Cand c;
    c.g = -CUDART_INF;
    c.g2 = -CUDART_INF;
    c.s = 0.0;
    c.dev = 0.0;
    c.k = 0x7fffffff;
    return c;
}

// same column: higher score wins, then lower position (ref one.py:62)
__device__ __forceinline__ void cand_merge(Cand &a, const Cand &b) {
    const bool b_wins = (b.g > a.g) || (b.g == a.g && b.k < a.k);
    const double loser = b_wins ? a.g : b.g;
    const double g2 = fmax(fmax(a.g2, b.g2), loser);
    if (b_wins) {
        a.g = b.g;
        a.s = b.s;
        a.dev = b.dev;
        a.k = b.k;
    }
    a.g2 = g2;
}

__device__ __forceinline__ Cand cand_shfl_xor(const Cand &c, int m) {
    Cand o;
    o.g = shfl_xor_d(c.g, m);
    o.g2 = shfl_xor_d(c.g2, m);
    o.s = shfl_xor_d(c.s, m);
    o.dev = shfl_xor_d(c.dev, m);
    o.k = __shfl_xor_sync(FULL_MASK, c.k, m);
    return o;
}

__device__ __forceinline__ void cand_warp_reduce(Cand &c) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        Cand o = cand_shfl_xor(c, m);
        cand_merge(c, o);
    }
}

// ---- approximate (parallel-scan) score at position kk of a node with n rows ------------
__device__ __forceinline__ double score_fast(double S, int kk, int n, double mean, int variant, double &dev) {
    if (variant == 1) {
        // ref one.py:57-61: w[k] * (S - (k+1) mean)^2, w = 1 / (float32(k+1) * (n-k-1))
        dev = fabs(S - (double)(kk + 1) * mean);
        const double ab = (double)__int2float_rn(kk + 1) * (double)(n - 1 - kk);
        return dev * dev * rcp_fast(ab);
    }
    // ref one.py:85-90: |sqrt(1/(l r)) S - mean sqrt(l/r)| with float32 ratios
    const float lc = __int2float_rn(kk + 1), rc = __int2float_rn(n - 1 - kk);
    const float ra = __fsqrt_rn(__frcp_rn(__fmul_rn(lc, rc)));
    const float rb = __fsqrt_rn(__fdiv_rn(lc, rc));
    dev = 0.0;
    return fabs((double)ra * S - mean * (double)rb);
}

// ---- reference-order score (no contraction, IEEE division) ------------------------------
__device__ __forceinline__ double score_exact(double S, int kk, int n, double mean, int variant, double &dev) {
    if (variant == 1) {
        dev = fabs(__dsub_rn(S, __dmul_rn((double)(kk + 1), mean)));                       // one.py:57-58
        const double w = __ddiv_rn(1.0, __dmul_rn((double)__int2float_rn(kk + 1), (double)(n - 1 - kk)));  // one.py:52-53
        return __dmul_rn(w, __dmul_rn(dev, dev));                                            // one.py:61
    }
    const float lc = __int2float_rn(kk + 1), rc = __int2float_rn(n - 1 - kk);
    const float ra = __fsqrt_rn(__frcp_rn(__fmul_rn(lc, rc)));                               // one.py:85-86
    const float rb = __fsqrt_rn(__fdiv_rn(lc, rc));                                          // one.py:87-88
    dev = 0.0;
    return fabs(__dsub_rn(__dmul_rn((double)ra, S), __dmul_rn(mean, (double)rb)));          // one.py:89-90
}

constexpr int SY_PITCH = SC_TILE + SC_TILE / 8;  // skew: one pad double per 8 => conflict-free blocked reads

__global__ void __launch_bounds__(SC_THREADS) l2_scan_kernel(LevelArgs a, const TileDesc *__restrict__ tiles,
                                                             int n_tiles, ScanStatus st, u32 epoch, u32 *ticket,
                                                             Partial *__restrict__ partials, int variant) {
    __shared__ double sy[SY_PITCH];
    __shared__ double s_wtot[SC_THREADS / 32];
    __shared__ double s_excl;
    __shared__ u32 s_ticket;
    __shared__ Cand s_c[SC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 t = s_ticket;
    // tile-major order (see sort.cu): in-flight blocks spread over the columns
    const int f = t % a.F_search, ti = t / a.F_search;
    const TileDesc td = tiles[ti];
    const SegDev sg = a.segs[td.seg];
    const int off = td.tin * SC_TILE;
    const int cnt = min(SC_TILE, sg.n - off);
    const int *col = a.orders + (int64_t)f * a.col_stride + sg.start + off;

    // coalesced read of the row ids, gather of y, staged so each thread gets 8 consecutive values
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const int e = i * SC_THREADS + tid;
        double v = 0.0;
        if (e < cnt) v = __ldg(a.y + ld_stream_i32(col + e));
        sy[e + (e >> 3)] = v;
    }
    __syncthreads();
    double p[SC_ITEMS];
    {
        const double *src = sy + tid * (SC_ITEMS + 1);
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; ++i) {
            run += src[i];
            p[i] = run;
        }
    }
    // block exclusive scan of the thread totals
    double incl = p[SC_ITEMS - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = shfl_up_d(incl, d);
        if (lane >= d) incl += o;
    }
    double texcl = shfl_up_d(incl, 1);
    if (lane == 0) texcl = 0.0;
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    double wexcl = 0.0, tile_total = 0.0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w < warp) wexcl += s_wtot[w];
        tile_total += s_wtot[w];
    }
    // carry-in: decoupled look-back over the earlier tiles of this (node, column), by warp 0
    if (warp == 0) {
        const double excl = lookback_f64(st, (int64_t)f * n_tiles + ti, td.tin, epoch, tile_total, lane);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const double offset = s_excl + (wexcl + texcl);
    const double mean = sg.total / (double)sg.n;
    Cand best = cand_none();
    const int e0 = tid * SC_ITEMS;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        const int kk = off + e0 + i;
        if (e0 + i < cnt && kk < sg.n - 1) {
            const double S = offset + p[i];
            double dev;
            const double g = score_fast(S, kk, sg.n, mean, variant, dev);
            if (g > best.g) {
                best.g2 = best.g;
                best.g = g;
                best.s = S;
                best.dev = dev;
                best.k = kk;
            } else if (g > best.g2) {
                best.g2 = g;
            }
        }
    }
    cand_warp_reduce(best);
    if (lane == 0) s_c[warp] = best;
    __syncthreads();
    if (tid == 0) {
        Cand b = s_c[0];
        for (int w = 1; w < SC_THREADS / 32; ++w) cand_merge(b, s_c[w]);
        Partial out;
        out.score = b.g;
        out.aux = b.dev;
        out.lsum = b.s;
        out.score2 = b.g2;
        out.k = b.k == 0x7fffffff ? -1 : b.k;
        out.f = f;
        partials[(int64_t)(sg.pbase + td.tin) * a.F_search + f] = out;
    }
}

// Sum of y over each listed node in the sorted order of column `total_col`, one element at a
// time (the reference's cumsums[-1][-1], one.py:57).  One warp per node.
__global__ void __launch_bounds__(256) exact_total_kernel(LevelArgs a, const int *__restrict__ list, int n_list,
                                                          int total_col, double *__restrict__ exact_total) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_list) return;
    const int lane = threadIdx.x & 31;
    const int seg = list[gw];
    const SegDev sg = a.segs[seg];
    const int *col = a.orders + (int64_t)total_col * a.col_stride + sg.start;
    double S = 0.0;
    double yv = lane < sg.n ? __ldg(a.y + col[lane]) : 0.0;
    for (int base = 0; base < sg.n; base += 32) {
        const double cur = yv;
        const int nx = base + 32 + lane;
        yv = nx < sg.n ? __ldg(a.y + col[nx]) : 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) S = __dadd_rn(S, shfl_d(cur, j));
    }
    if (lane == 0) exact_total[seg] = S;
}

__global__ void __launch_bounds__(256) l2_exact_kernel(LevelArgs a, const int *__restrict__ list, int n_list,
                                                       const double *__restrict__ exact_total,
                                                       Partial *__restrict__ partials, int variant) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_list * a.F_search) return;
    const int lane = threadIdx.x & 31;
    const int f = gw % a.F_search;
    const int seg = list[gw / a.F_search];
    const SegDev sg = a.segs[seg];
    const int n = sg.n;
    const double mean = __ddiv_rn(exact_total[seg], (double)n);  // one.py:57  ys[-1][-1] / n
    const int *col = a.orders + (int64_t)f * a.col_stride + sg.start;
    double S = 0.0;
    Cand best = cand_none();
    double yv = lane < n ? __ldg(a.y + col[lane]) : 0.0;
    for (int base = 0; base < n; base += 32) {
        const double cur = yv;
        const int nx = base + 32 + lane;
        yv = nx < n ? __ldg(a.y + col[nx]) : 0.0;
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            S = __dadd_rn(S, shfl_d(cur, j));  // np.cumsum: strictly sequential fp64 (one.py:97)
            if (lane == j) mine = S;
        }
        const int kk = base + lane;
        if (kk < n - 1) {
            double dev;
            const double g = score_exact(mine, kk, n, mean, variant, dev);
            if (g > best.g) {
                best.g2 = best.g;
                best.g = g;
                best.s = mine;
                best.dev = dev;
                best.k = kk;
            } else if (g > best.g2) {
                best.g2 = g;
            }
        }
    }
    cand_warp_reduce(best);
    if (lane == 0) {
        Partial out;
        out.score = best.g;
        out.aux = best.dev;
        out.lsum = best.s;
        out.score2 = best.g2;
        out.k = best.k == 0x7fffffff ? -1 : best.k;
        out.f = f;
        partials[(int64_t)sg.pbase * a.F_search + f] = out;
    }
}

// ---- per-node reduction over (tiles x columns) with the reference's full tie rule -------
// (score desc, k asc, aux desc, f asc): within one sorted position the reference compares
// |S-(k+1)mean| across features (one.py:59) and across positions the weighted squares
// (one.py:62); aux is 0 for the other criteria, which then tie-break on (k, f) only.
__device__ __forceinline__ bool partial_better(const Partial &b, const Partial &a) {
    if (b.k < 0) return false;
    if (a.k < 0) return true;
    if (b.score != a.score) return b.score > a.score;
    if (b.k != a.k) return b.k < a.k;
    if (b.aux != a.aux) return b.aux > a.aux;
    return b.f < a.f;
}

__device__ __forceinline__ void partial_merge(Partial &a, const Partial &b) {
    const bool bw = partial_better(b, a);
    const double loser = bw ? (a.k < 0 ? -CUDART_INF : a.score) : (b.k < 0 ? -CUDART_INF : b.score);
    const double s2 = fmax(fmax(a.score2, b.score2), loser);
    if (bw) a = b;
    a.score2 = s2;
}

__global__ void __launch_bounds__(128) finalize_kernel(LevelArgs a, const Partial *__restrict__ partials,
                                                       const double *__restrict__ X, int64_t ldx, int col_offset,
                                                       b200dt_candidate *__restrict__ out) {
    __shared__ Partial sp[128];
    const int seg = blockIdx.x;
    const SegDev sg = a.segs[seg];
    const int tid = threadIdx.x;
    const int64_t count = (int64_t)sg.np * a.F_search;
    const Partial *p = partials + (int64_t)sg.pbase * a.F_search;
    Partial b;
    b.score = -CUDART_INF;
    b.aux = 0.0;
    b.lsum = 0.0;
    b.score2 = -CUDART_INF;
    b.k = -1;
    b.f = 0;
    for (int64_t r = tid; r < count; r += 128) partial_merge(b, p[r]);
    sp[tid] = b;
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if (tid < s) {
            Partial mine = sp[tid];
            partial_merge(mine, sp[tid + s]);
            sp[tid] = mine;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const Partial w = sp[0];
        b200dt_candidate c;
        c.gain = w.score;
        c.aux = w.aux;
        c.left_sum = w.lsum;
        c.gain2 = w.score2;
        c.threshold = 0.0;
        c.k = w.k;
        c.f = w.f + col_offset;
        c.exact = sg.exact;
        c.pad = 0;
        if (w.k >= 0) {
            const int row = a.orders[(int64_t)w.f * a.col_stride + sg.start + w.k];
            c.threshold = X[(int64_t)row * ldx + w.f];  // one.py:104-105: a data value, not a midpoint
        }
        out[seg] = c;
    }
}

}  // namespace

int launch_l2_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles, int64_t n_elems,
                   ScanStatus st, u32 epoch, u32 *ticket, Partial *partials, int variant) {
    if (n_tiles <= 0) return 0;
    CK(cudaMemsetAsync(ticket, 0, 4, ctx->stream));
    const int64_t grid = (int64_t)n_tiles * a.F_search;
    if (grid > 0x7fffffffLL) return ctx->fail_msg("l2 scan: grid too large");
    LaunchScope ls(ctx, KC_L2_SCAN, 12.0 * (double)n_elems * a.F_search);
    l2_scan_kernel<<<(unsigned)grid, SC_THREADS, 0, ctx->stream>>>(a, tiles, n_tiles, st, epoch, ticket, partials, variant);
    CK(cudaGetLastError());
    return 0;
}

int launch_l2_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small, int64_t n_elems,
                    int total_col, double *exact_total, Partial *partials, int variant) {
    if (n_small <= 0) return 0;
    {
        LaunchScope ls(ctx, KC_L2_EXACT, 12.0 * (double)n_elems);
        exact_total_kernel<<<(n_small + 7) / 8, 256, 0, ctx->stream>>>(a, small_list, n_small, total_col, exact_total);
    }
    const int64_t warps = (int64_t)n_small * a.F_search;
    LaunchScope ls(ctx, KC_L2_EXACT, 12.0 * (double)n_elems * a.F_search);
    l2_exact_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(a, small_list, n_small, exact_total, partials, variant);
    CK(cudaGetLastError());
    return 0;
}

int launch_finalize(b200dt_ctx *ctx, const LevelArgs &a, const Partial *partials, const double *X, int64_t ldx,
                    int col_offset, b200dt_candidate *out) {
    if (a.n_segs <= 0) return 0;
    LaunchScope ls(ctx, KC_FINALIZE, 0.0);
    finalize_kernel<<<a.n_segs, 128, 0, ctx->stream>>>(a, partials, X, ldx, col_offset, out);
    CK(cudaGetLastError());
    return 0;
}
