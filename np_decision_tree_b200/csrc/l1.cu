// K3: L1 (MAE) split search -- prefix / suffix running medians as an order-statistic scan,
// exact cost scans, arg-min -- and the running median itself.
//
// replaces best_l1_improvements[_v2]                 (ref decision_tree/strategy_l1.py:49-122)
//          np_stream_median.cummedian (native C++)    (ref decision_tree/np_stream_median.pyx:8-39)
//          StreamMedian.__call__                       (ref decision_tree/stream_median.py:28-92)
//
// The reference feeds every column of y[orders] through a two-heap running median, twice
// (forwards for the left child, backwards for the right child), turns the (lower, upper) median
// pairs into per-row cost increments and cumsums them.  The median pair of a prefix of length m
// is a pure selection (order statistics floor((m-1)/2) and floor(m/2)), so any exact
// order-statistic algorithm reproduces it bit for bit.  Here:
//
//  l1_rank_kernel   : the session keeps one extra order column sorted by y; stable partitions keep
//                     it sorted inside every node, so a row's node-local y-rank is its position
//                     there.  Scatter rank[row] and the node-sorted y values.
//  l1_median_walk   : one warp per (node, column, direction).  The ranks of the column, in scan
//                     order, are inserted into a bitmap over ranks held in SHARED memory; the
//                     (lower, upper) median ranks move by at most one set bit per insertion, found
//                     with clz/ffs over bitmap words.  Lane 0 walks; all lanes stage the coalesced
//                     input gather and the output stores through shared memory.
//  l1_cost_scan     : one warp per (node, column, direction): increments from the median pairs
//                     (strategy_l1.py:92-98, 109-114), summed one element at a time in fp64 in
//                     the reference's order (np.cumsum) -- the L1 search is exact by construction.
//  l1_argmin        : cost[k] = R[k+1] + L[k] (strategy_l1.py:119), arg-min per (node, column),
//                     ties to the lowest k; the per-node reduction over columns is the shared
//                     finalize kernel with the (k asc, f asc) rule of the flat argmin (:120).
#include <math_constants.h>

#include <algorithm>

#include "session.cuh"

namespace {

struct L1Seg {
    int start;
    int n;
};

// rank[row] = position of row in the node's y-sorted order; ynode[start+j] = j-th smallest y of the node
__global__ void __launch_bounds__(256) l1_rank_kernel(const int *__restrict__ ycol, const L1Seg *__restrict__ segs,
                                                      const double *__restrict__ y, int *__restrict__ lrank,
                                                      double *__restrict__ ynode) {
    const L1Seg sg = segs[blockIdx.y];
    for (int j = blockIdx.x * 256 + threadIdx.x; j < sg.n; j += gridDim.x * 256) {
        const int row = ycol[sg.start + j];
        lrank[row] = j;
        ynode[sg.start + j] = y[row];
    }
}

// Bitmap over ranks with the two words holding the current medians cached in registers: once the
// set is dense the neighbour search never leaves the cached word, so the per-insertion critical
// path is register arithmetic; the shared-memory read-modify-write of the insertion is off it.
struct Walk {
    u32 *bm;
    int m, lo, hi;
    int wlo, whi;     // word indices of lo / hi
    u32 vlo, vhi;     // cached contents of those words (kept in sync with bm)

    __device__ __forceinline__ void load_lo() { wlo = lo >> 5; vlo = bm[wlo]; }
    __device__ __forceinline__ void load_hi() { whi = hi >> 5; vhi = bm[whi]; }

    // largest set bit strictly below c (c lives in word wc with cached value vc); it exists
    __device__ __forceinline__ int prev_of(int c, int wc, u32 vc) const {
        u32 mk = vc & ((1u << (c & 31)) - 1u);
        int w = wc;
        while (mk == 0u) mk = bm[--w];
        return (w << 5) + 31 - __clz(mk);
    }
    // smallest set bit strictly above c
    __device__ __forceinline__ int next_of(int c, int wc, u32 vc) const {
        u32 mk = (c & 31) == 31 ? 0u : (vc & ~((2u << (c & 31)) - 1u));
        int w = wc;
        while (mk == 0u) mk = bm[++w];
        return (w << 5) + __ffs(mk) - 1;
    }

    __device__ __forceinline__ void insert(int r) {
        const int w = r >> 5;
        const u32 bit = 1u << (r & 31);
        bm[w] |= bit;
        if (m == 0) {
            lo = hi = r;
            wlo = whi = w;
            vlo = vhi = bit;
        } else {
            if (w == wlo) vlo |= bit;
            if (w == whi) vhi |= bit;
            if (m & 1) {
                // odd -> even: the single median c stays one of the pair (pyx:23-33)
                if (r < lo) {
                    lo = prev_of(hi, whi, vhi);
                    if ((lo >> 5) != wlo) load_lo();
                } else {
                    hi = next_of(lo, wlo, vlo);
                    if ((hi >> 5) != whi) load_hi();
                }
            } else {
                // even -> odd: the new median is the old lower, the old upper, or r (pyx:34-38)
                if (r < lo) {
                    hi = lo;
                    whi = wlo;
                    vhi = vlo;
                } else if (r > hi) {
                    lo = hi;
                    wlo = whi;
                    vlo = vhi;
                } else {
                    lo = hi = r;
                    wlo = whi = w;
                    vlo = vhi = bm[w];
                }
            }
        }
        ++m;
    }
};

// One warp per (node, column, direction).  dir 0: positions 0..n-1 (prefix medians), dir 1:
// positions n-1..0 (suffix medians).  out_lo/out_hi[(dir*F+f)*stride + start + pos] = ranks of the
// (lower, upper) median of the elements scanned up to and including position pos.
// All 32 lanes run the walk redundantly on identical state (ranks are broadcast by shuffle, lane j
// keeps the pair of step j), so there is no divergence and no staging through shared memory.
__global__ void __launch_bounds__(32) l1_median_walk_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                            const L1Seg *__restrict__ segs,
                                                            const int *__restrict__ lrank, int *__restrict__ out_lo,
                                                            int *__restrict__ out_hi, u32 *gbitmap,
                                                            int64_t gwords_per_block, int direct_ranks) {
    extern __shared__ u32 smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    const int dir = b & 1;
    const int f = (b >> 1) % F;
    const L1Seg sg = segs[(b >> 1) / F];
    const int n = sg.n;
    const int words = (n + 31) >> 5;
    Walk wk;
    wk.bm = gbitmap ? gbitmap + (int64_t)b * gwords_per_block : smem;
    for (int w = lane; w < words; w += 32) wk.bm[w] = 0u;
    __syncwarp();
    wk.m = wk.lo = wk.hi = wk.wlo = wk.whi = 0;
    wk.vlo = wk.vhi = 0u;
    const int *col = orders + (int64_t)f * stride + sg.start;
    int *olo = out_lo + ((int64_t)dir * F + f) * stride + sg.start;
    int *ohi = out_hi + ((int64_t)dir * F + f) * stride + sg.start;
    // software pipeline: the ranks of the next 32 positions are fetched while this chunk is walked
    auto fetch = [&](int q0) -> int {
        const int q = q0 + lane;
        if (q >= n) return 0;
        const int pos = dir ? n - 1 - q : q;
        return direct_ranks ? col[pos] : lrank[col[pos]];
    };
    int nxt = fetch(0);
    for (int q0 = 0; q0 < n; q0 += 32) {
        const int mine = nxt;
        nxt = fetch(q0 + 32);
        int mylo = 0, myhi = 0;
        if (q0 + 32 <= n) {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                wk.insert(__shfl_sync(FULL_MASK, mine, j));
                if (lane == j) {
                    mylo = wk.lo;
                    myhi = wk.hi;
                }
            }
        } else {
            const int cnt = n - q0;
            for (int j = 0; j < cnt; ++j) {
                wk.insert(__shfl_sync(FULL_MASK, mine, j));
                if (lane == j) {
                    mylo = wk.lo;
                    myhi = wk.hi;
                }
            }
        }
        const int q = q0 + lane;
        if (q < n) {
            const int pos = dir ? n - 1 - q : q;
            olo[pos] = mylo;
            ohi[pos] = myhi;
        }
    }
}

// One warp per (node, column, direction): cost increments and their strictly sequential fp64 sums.
//   dir 0: L[0] = 0, L[p] = L[p-1] + dist(y_p, pair of positions 0..p-1)        strategy_l1.py:89-102
//   dir 1: R[n-1] = inc(n-1), R[p] = R[p+1] + inc(p), pair of positions p..n-1   strategy_l1.py:103-117
__global__ void __launch_bounds__(256) l1_cost_scan_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                           const L1Seg *__restrict__ segs, int n_segs,
                                                           const double *__restrict__ y,
                                                           const double *__restrict__ ynode,
                                                           const int *__restrict__ med_lo,
                                                           const int *__restrict__ med_hi, double *__restrict__ Lc,
                                                           double *__restrict__ Rc, int direct_y) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_segs * F * 2) return;
    const int lane = threadIdx.x & 31;
    const int dir = gw & 1;
    const int f = (gw >> 1) % F;
    const L1Seg sg = segs[(gw >> 1) / F];
    const int n = sg.n;
    const int *col = orders + (int64_t)f * stride + sg.start;
    const int *mlo = med_lo + ((int64_t)dir * F + f) * stride + sg.start;
    const int *mhi = med_hi + ((int64_t)dir * F + f) * stride + sg.start;
    const double *yn = ynode + sg.start;
    double S = 0.0;
    if (dir == 0) {
        double *out = Lc + (int64_t)f * stride + sg.start;
        if (lane == 0) out[0] = 0.0;
        for (int p0 = 1; p0 < n; p0 += 32) {
            const int p = p0 + lane;
            double inc = 0.0;
            if (p < n) {
                const double yv = direct_y ? y[(int64_t)f * stride + sg.start + p] : y[col[p]];
                const double lo = yn[mlo[p - 1]], hi = yn[mhi[p - 1]];
                const double a = yv <= lo ? __dsub_rn(lo, yv) : 0.0;   // strategy_l1.py:93-95
                const double b = yv >= hi ? __dsub_rn(yv, hi) : 0.0;   // strategy_l1.py:96-98
                inc = __dadd_rn(a, b);
            }
            double mine = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                S = __dadd_rn(S, shfl_d(inc, j));
                if (lane == j) mine = S;
            }
            if (p < n) out[p] = mine;
        }
    } else {
        double *out = Rc + (int64_t)f * stride + sg.start;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            const int p = n - 1 - q;
            double inc = 0.0;
            if (q < n) {
                const double yv = direct_y ? y[(int64_t)f * stride + sg.start + p] : y[col[p]];
                const double lo = yn[mlo[p]], hi = yn[mhi[p]];
                const double a = yv <= lo ? __dsub_rn(hi, yv) : 0.0;   // strategy_l1.py:110-111
                const double b = yv >= hi ? __dsub_rn(yv, lo) : 0.0;   // strategy_l1.py:112-114
                inc = __dadd_rn(a, b);
            }
            double mine = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                S = __dadd_rn(S, shfl_d(inc, j));
                if (lane == j) mine = S;
            }
            if (q < n) out[p] = mine;
        }
    }
}

// One warp per (node, column): arg-min over k of R[k+1] + L[k]; ties to the lowest k.
__global__ void __launch_bounds__(256) l1_argmin_kernel(int64_t stride, int F, const SegDev *__restrict__ segs,
                                                        int n_segs, const double *__restrict__ Lc,
                                                        const double *__restrict__ Rc,
                                                        Partial *__restrict__ partials) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_segs * F) return;
    const int lane = threadIdx.x & 31;
    const int f = gw % F;
    const SegDev sg = segs[gw / F];
    const double *L = Lc + (int64_t)f * stride + sg.start;
    const double *R = Rc + (int64_t)f * stride + sg.start;
    double best = CUDART_INF, second = CUDART_INF;
    int bk = 0x7fffffff;
    for (int k = lane; k < sg.n - 1; k += 32) {
        const double c = __dadd_rn(R[k + 1], L[k]);   // strategy_l1.py:119
        if (c < best) {
            second = best;
            best = c;
            bk = k;
        } else if (c < second) {
            second = c;
        }
    }
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1) {
        const double ob = shfl_xor_d(best, mask), os = shfl_xor_d(second, mask);
        const int ok = __shfl_xor_sync(FULL_MASK, bk, mask);
        const bool o_wins = ob < best || (ob == best && ok < bk);
        const double loser = o_wins ? best : ob;
        second = fmin(fmin(second, os), loser);
        if (o_wins) {
            best = ob;
            bk = ok;
        }
    }
    if (lane == 0) {
        Partial out;
        out.score = -best;   // the shared finalize maximises
        out.aux = 0.0;
        out.lsum = 0.0;
        out.score2 = -second;
        out.k = bk == 0x7fffffff ? -1 : bk;
        out.f = f;
        partials[(int64_t)sg.pbase * F + f] = out;
    }
}

// median of each listed node from its y-sorted column (np.median: mean of the two middle values)
__global__ void l1_leaf_median_kernel(const int *__restrict__ ycol, const L1Seg *__restrict__ segs, int n_segs,
                                      const double *__restrict__ y, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_segs) return;
    const L1Seg sg = segs[i];
    const double a = y[ycol[sg.start + ((sg.n - 1) >> 1)]];
    const double b = y[ycol[sg.start + (sg.n >> 1)]];
    out[i] = (sg.n & 1) ? a : __dmul_rn(__dadd_rn(a, b), 0.5);
}

__global__ void gather_pairs_kernel(const int *__restrict__ lo, const int *__restrict__ hi,
                                    const double *__restrict__ sorted, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = sorted[lo[i]];
    out[2 * i + 1] = sorted[hi[i]];
}

__global__ void rank_from_order_kernel(const int *__restrict__ order, const double *__restrict__ a, int64_t n,
                                       int *__restrict__ rank, double *__restrict__ sorted) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i = order[j];
    rank[i] = (int)j;
    sorted[j] = a[i];
}

const int WALK_SMEM_MAX = 220 * 1024;

// launch the median walk for `n_blocks` (node, column, direction) sequences whose longest has max_n elements
int launch_walk(b200dt_ctx *ctx, const int *orders, int64_t stride, int F, const L1Seg *d_segs, int n_segs,
                int max_n, int64_t n_elems, const int *lrank, int *med_lo, int *med_hi, int direct_ranks) {
    const size_t words = (size_t)(max_n + 31) / 32;
    const size_t smem = words * 4 + 16;
    const int blocks = n_segs * F * 2;
    u32 *gbm = nullptr;
    size_t use_smem = smem;
    if (smem > (size_t)WALK_SMEM_MAX) {
        // rank bitmap too large for shared memory (> ~1.8 M rows in one node): keep it in HBM/L2
        CKR(scratch_get(ctx, SB_L1_H, (size_t)blocks * words * 4, (void **)&gbm));
        use_smem = 16;
    }
    static bool attr_done = false;
    if (!attr_done) {
        CK(cudaFuncSetAttribute(l1_median_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK_SMEM_MAX));
        attr_done = true;
    }
    LaunchScope ls(ctx, KC_L1_DESCENT, 16.0 * (double)n_elems * F * 2);
    l1_median_walk_kernel<<<blocks, 32, use_smem, ctx->stream>>>(orders, stride, F, d_segs, lrank, med_lo, med_hi, gbm,
                                                                 (int64_t)words, direct_ranks);
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

// L1 split search of every live node (always in the reference's exact arithmetic).
int l1_level_search(b200dt_ctx *ctx, Session *s, b200dt_candidate *d_out, bool exact_only) {
    (void)exact_only;
    const int n_live = (int)s->live.size();
    if (s->ycol < 0) return ctx->fail_msg("L1 search: session has no y-order column");
    const int F = s->F_local;
    const int64_t stride = s->col_stride;
    std::vector<L1Seg> segs(n_live);
    s->h_segs.resize(n_live);
    int max_n = 0;
    int64_t n_elems = 0;
    for (int i = 0; i < n_live; ++i) {
        const HNode &nd = s->nodes[s->live[i]];
        segs[i] = L1Seg{nd.start, nd.n};
        SegDev sg;
        sg.total = 0.0;
        sg.total_w = 0.0;
        sg.start = nd.start;
        sg.n = nd.n;
        sg.pbase = i;
        sg.np = 1;
        sg.nleft = 0;
        sg.exact = 1;
        sg.pstep = 1;
        sg.pside = 0;
        s->h_segs[i] = sg;
        max_n = std::max(max_n, nd.n);
        n_elems += nd.n;
    }
    L1Seg *d_l1segs;
    int *lrank, *med_lo, *med_hi;
    double *ynode, *Lc, *Rc;
    CKR(scratch_get(ctx, SB_SEG, (size_t)n_live * sizeof(SegDev) + 64, (void **)&s->d_segs));
    CKR(scratch_get(ctx, SB_TILES, (size_t)n_live * sizeof(L1Seg) + 64, (void **)&d_l1segs));
    CKR(scratch_get(ctx, SB_PARTIAL, (size_t)(n_live + 1) * F * sizeof(Partial) + 64, (void **)&s->d_partials));
    CKR(scratch_get(ctx, SB_L1_A, (size_t)s->N * 4 + 64, (void **)&lrank));
    CKR(scratch_get(ctx, SB_L1_B, (size_t)stride * 8 + 64, (void **)&ynode));
    CKR(scratch_get(ctx, SB_L1_C, (size_t)2 * F * stride * 4 + 64, (void **)&med_lo));
    CKR(scratch_get(ctx, SB_L1_D, (size_t)2 * F * stride * 4 + 64, (void **)&med_hi));
    CKR(scratch_get(ctx, SB_L1_E, (size_t)F * stride * 8 + 64, (void **)&Lc));
    CKR(scratch_get(ctx, SB_L1_F, (size_t)F * stride * 8 + 64, (void **)&Rc));
    void *stage;
    const size_t seg_bytes = (size_t)n_live * sizeof(SegDev);
    CKR(pinned_get(ctx, PIN_SEARCH, seg_bytes + 64 + (size_t)n_live * sizeof(L1Seg), &stage));
    memcpy(stage, s->h_segs.data(), seg_bytes);
    memcpy((char *)stage + seg_bytes + 64, segs.data(), (size_t)n_live * sizeof(L1Seg));
    CK(cudaMemcpyAsync(s->d_segs, stage, seg_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_l1segs, (char *)stage + seg_bytes + 64, (size_t)n_live * sizeof(L1Seg),
                       cudaMemcpyHostToDevice, ctx->stream));
    const int *orders = s->ord[s->cur];
    const int *ycol = orders + (int64_t)s->ycol * stride;
    {
        int gx = std::min(1024, (max_n + 255) / 256);
        for (int i0 = 0; i0 < n_live; i0 += 65535) {
            const int ni = std::min(65535, n_live - i0);
            LaunchScope ls(ctx, KC_L1_RANK, 24.0 * (double)n_elems);
            l1_rank_kernel<<<dim3(gx, ni), 256, 0, ctx->stream>>>(ycol, d_l1segs + i0, s->dy, lrank, ynode);
        }
    }
    CKR(launch_walk(ctx, orders, stride, F, d_l1segs, n_live, max_n, n_elems, lrank, med_lo, med_hi, 0));
    {
        const int64_t warps = (int64_t)n_live * F * 2;
        LaunchScope ls(ctx, KC_L1_COST, 44.0 * (double)n_elems * F);
        l1_cost_scan_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(orders, stride, F, d_l1segs, n_live,
                                                                                  s->dy, ynode, med_lo, med_hi, Lc, Rc, 0);
    }
    {
        const int64_t warps = (int64_t)n_live * F;
        LaunchScope ls(ctx, KC_L1_COST, 16.0 * (double)n_elems * F);
        l1_argmin_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(stride, F, s->d_segs, n_live, Lc, Rc,
                                                                               s->d_partials);
    }
    CK(cudaGetLastError());
    LevelArgs a;
    a.orders = orders;
    a.col_stride = stride;
    a.F_search = F;
    a.y = s->dy;
    a.ypay = nullptr;
    a.wpay = nullptr;
    a.segs = s->d_segs;
    a.n_segs = n_live;
    CKR(launch_finalize(ctx, a, s->d_partials, F, nullptr, s->dX, s->ldx, s->col_offset, d_out));
    return 0;
}

// median of the nodes in `pending` (read from the CURRENT order buffer), then clear the list
int l1_leaf_medians(b200dt_ctx *ctx, Session *s, std::vector<int> &pending) {
    const int n = (int)pending.size();
    if (n == 0) return 0;
    std::vector<L1Seg> segs(n);
    for (int i = 0; i < n; ++i) segs[i] = L1Seg{s->nodes[pending[i]].start, s->nodes[pending[i]].n};
    L1Seg *d_segs;
    double *d_out;
    CKR(scratch_get(ctx, SB_TMP_I64, (size_t)n * (sizeof(L1Seg) + 8) + 64, (void **)&d_segs));
    d_out = (double *)(d_segs + n);
    void *stage;
    CKR(pinned_get(ctx, PIN_L1, (size_t)n * (sizeof(L1Seg) + 8) + 64, &stage));
    memcpy(stage, segs.data(), (size_t)n * sizeof(L1Seg));
    CK(cudaMemcpyAsync(d_segs, stage, (size_t)n * sizeof(L1Seg), cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, KC_MISC, 0.0);
        l1_leaf_median_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(
            s->ord[s->cur] + (int64_t)s->ycol * s->col_stride, d_segs, n, s->dy, d_out);
    }
    CK(cudaGetLastError());
    double *h_out = (double *)((char *)stage + (size_t)n * sizeof(L1Seg));
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) s->nodes[pending[i]].value = h_out[i];
    pending.clear();
    return 0;
}

// Running (lower, upper) median of a 1-D device array: out[2*i], out[2*i+1] for prefix 0..i.
int l1_cummedian_device(b200dt_ctx *ctx, const double *d_a, int64_t n, double *d_out) {
    if (n >= (1ll << 30)) return ctx->fail_msg("cummedian: n must be < 2^30");
    const int64_t stride = (n + 31) / 32 * 32;
    int *order, *alt, *rank, *med_lo, *med_hi;
    double *sorted;
    L1Seg *d_seg;
    CKR(scratch_get(ctx, SB_ORD_A, (size_t)stride * 4, (void **)&order));
    CKR(scratch_get(ctx, SB_ORD_B, (size_t)stride * 4, (void **)&alt));
    CKR(scratch_get(ctx, SB_L1_A, (size_t)stride * 4 + 64, (void **)&rank));
    CKR(scratch_get(ctx, SB_L1_B, (size_t)stride * 8 + 64, (void **)&sorted));
    CKR(scratch_get(ctx, SB_L1_C, (size_t)2 * stride * 4 + 64, (void **)&med_lo));
    CKR(scratch_get(ctx, SB_L1_D, (size_t)2 * stride * 4 + 64, (void **)&med_hi));
    CKR(scratch_get(ctx, SB_TILES, sizeof(L1Seg) + 64, (void **)&d_seg));
    // ranks by a stable sort of the values (equal values are interchangeable for a selection)
    CKR(sort_columns_device(ctx, d_a, n, 1, 1, order, alt, stride));
    {
        LaunchScope ls(ctx, KC_L1_RANK, 24.0 * (double)n);
        rank_from_order_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(order, d_a, n, rank, sorted);
    }
    L1Seg seg{0, (int)n};
    void *stage;
    CKR(pinned_get(ctx, PIN_L1, sizeof(L1Seg), &stage));
    memcpy(stage, &seg, sizeof seg);
    CK(cudaMemcpyAsync(d_seg, stage, sizeof seg, cudaMemcpyHostToDevice, ctx->stream));
    // the "column" is the rank sequence itself (direct_ranks = 1); only direction 0 is used, but the
    // walk kernel always runs both directions, so the outputs have two planes
    CKR(launch_walk(ctx, rank, stride, 1, d_seg, 1, (int)n, n, nullptr, med_lo, med_hi, 1));
    {
        LaunchScope ls(ctx, KC_MISC, 40.0 * (double)n);
        gather_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(med_lo, med_hi, sorted, n, d_out);
    }
    CK(cudaGetLastError());
    return 0;
}
