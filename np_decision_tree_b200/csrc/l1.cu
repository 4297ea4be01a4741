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
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

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
// L2ONLY: the bitmap lives in global memory and is shared with nobody: insertions are fire-and-forget
// reductions (no load on the insertion's path) and reads go to L2, where those reductions land.
template <bool L2ONLY>
struct Walk {
    u32 *bm;
    int m, lo, hi;
    int wlo, whi;     // word indices of lo / hi
    u32 vlo, vhi;     // cached contents of those words (kept in sync with bm)
    bool writer;      // all lanes walk the same state redundantly; one of them updates a global bitmap

    __device__ __forceinline__ u32 rd(int w) const {
        // L2ONLY: only lane 0 wrote the bitmap, so only lane 0's read is ordered after those writes
        if (L2ONLY) return __shfl_sync(FULL_MASK, ld_relaxed_u32(bm + w), 0);
        return bm[w];
    }
    __device__ __forceinline__ void set(int w, u32 bit) {
        if (L2ONLY) {
            if (writer) asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(bm + w), "r"(bit) : "memory");
        } else {
            bm[w] |= bit;
        }
    }
    __device__ __forceinline__ void load_lo() { wlo = lo >> 5; vlo = rd(wlo); }
    __device__ __forceinline__ void load_hi() { whi = hi >> 5; vhi = rd(whi); }

    // largest set bit strictly below c (c lives in word wc with cached value vc); it exists.
    // wv = the full content of the word it was found in (no second read when the caller caches it)
    __device__ __forceinline__ int prev_of(int c, int wc, u32 vc, u32 &wv) const {
        u32 mk = vc & ((1u << (c & 31)) - 1u);
        int w = wc;
        wv = vc;
        while (mk == 0u) mk = wv = rd(--w);
        return (w << 5) + 31 - __clz(mk);
    }
    // smallest set bit strictly above c
    __device__ __forceinline__ int next_of(int c, int wc, u32 vc, u32 &wv) const {
        u32 mk = (c & 31) == 31 ? 0u : (vc & ~((2u << (c & 31)) - 1u));
        int w = wc;
        wv = vc;
        while (mk == 0u) mk = wv = rd(++w);
        return (w << 5) + __ffs(mk) - 1;
    }

    __device__ __forceinline__ void insert(int r) {
        const int w = r >> 5;
        const u32 bit = 1u << (r & 31);
        set(w, bit);
        if (m == 0) {
            lo = hi = r;
            wlo = whi = w;
            vlo = vhi = bit;
        } else {
            if (w == wlo) vlo |= bit;
            if (w == whi) vhi |= bit;
            if (m & 1) {
                // odd -> even: the single median c stays one of the pair (pyx:23-33)
                // (here lo == hi; the word of the moved end was just read by the neighbour search)
                u32 wv;
                if (r < lo) {
                    lo = prev_of(hi, whi, vhi, wv);
                    wlo = lo >> 5;
                    vlo = wv;
                } else {
                    hi = next_of(lo, wlo, vlo, wv);
                    whi = hi >> 5;
                    vhi = wv;
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
                    vlo = vhi = rd(w) | bit;   // (| bit: the word as of this insertion, whatever the read saw)
                }
            }
        }
        ++m;
    }
};

// ---- chunked walks (nodes of >= 2 * WALK_CHUNK_MIN rows) ------------------------------------------
// A walk is a serial chain of n insertions.  It is cut into P chunks that run concurrently: the state at
// the start of chunk c -- the set of the first c*S ranks and its median pair -- is an order statistic,
// not a recurrence: every chunk sets the bits of its OWN ranks (l1_chunk_bits_kernel), an exclusive
// prefix-OR over the chunks turns them into start bitmaps (l1_chunk_prefix_kernel), and the chunk finds
// its start pair by selecting the k-th set bit.  Bitmaps live in global memory (one per (walk, chunk)).
constexpr int WALK_CHUNK_MIN = 16384;
constexpr int WALK_CHUNKS_MAX = 32;

__host__ __device__ inline int walk_chunk_len(int n, int P) { return (((n + P - 1) / P) + 31) & ~31; }

// Walk chunks are NOT of equal length: an insertion at set size m costs about t0 + t_ld * n / (32 m) -- the
// chance of leaving the cached word when the set is dense, the number of empty words scanned when it is
// sparse, one L2 round trip each -- so early chunks are short.  bounds[c] = first insertion of chunk c as a
// fraction of n (bounds[0] = 0, bounds[P] = 1); a chunk covers [align32(n * bounds[c]), align32(n * bounds[c+1])).
__host__ __device__ inline int walk_chunk_begin(int n, const float *bounds, int c, int P) {
    if (c <= 0) return 0;
    if (c >= P) return n;
    const int q = ((int)((double)n * (double)bounds[c]) + 31) & ~31;
    return q < n ? q : n;
}

static void walk_chunk_bounds(int n, int P, std::vector<float> &bounds) {
    // us per insertion / per word fetched: register path, L2 round trip, chunk 0's cached path.  Calibrated on
    // B200 at 1 M rows x 100 columns (the walk phase is issue-bound -- all 32 lanes run the same scalar walk --
    // so the result is flat around these values: 20.6 ms here, 28.7 ms at t0 = 0.4, t_ld = 0.6)
    double t0 = 0.10, t_ld = 0.15, t_l1 = 0.10;
    if (const char *e = getenv("B200DT_L1_T0")) t0 = atof(e);
    if (const char *e = getenv("B200DT_L1_TLD")) t_ld = atof(e);
    if (const char *e = getenv("B200DT_L1_TL1")) t_l1 = atof(e);
    const double A = t_ld * (double)n / 32.0, A0 = t_l1 * (double)n / 32.0;
    // cost of a chunk [a, b): t0 (b - a) + A ln(b / a)  (chunk 0, from the empty set, with A0 and a = 1)
    auto end_of = [&](double a, double T, bool first) {   // largest b with cost(a, b) <= T
        double lo = a, hi = (double)n;
        const double AA = first ? A0 : A, aa = first ? 1.0 : a;
        if (t0 * (hi - a) + AA * std::log(hi / aa) <= T) return hi;
        for (int it = 0; it < 60; ++it) {
            const double mid = 0.5 * (lo + hi);
            (t0 * (mid - a) + AA * std::log(std::max(mid, 1.0) / aa) <= T ? lo : hi) = mid;
        }
        return lo;
    };
    auto chunks_needed = [&](double T, std::vector<double> *out) {
        double a = 0.0;
        int c = 0;
        if (out) out->assign(1, 0.0);
        while (a < (double)n && c < 4 * P) {
            a = std::max(end_of(a, T, c == 0), a + 32.0);
            if (out) out->push_back(std::min(a, (double)n));
            ++c;
        }
        return c;
    };
    double Tlo = 0.0, Thi = t0 * n + A * 20.0 + A0 * 20.0;
    for (int it = 0; it < 60; ++it) {
        const double T = 0.5 * (Tlo + Thi);
        (chunks_needed(T, nullptr) <= P ? Thi : Tlo) = T;
    }
    std::vector<double> b;
    chunks_needed(Thi, &b);
    bounds.assign(P + 1, 1.0f);
    for (int c = 0; c <= P; ++c) bounds[c] = c < (int)b.size() ? (float)(b[c] / (double)n) : 1.0f;
    bounds[0] = 0.0f;
    bounds[P] = 1.0f;
}

__global__ void __launch_bounds__(256) l1_chunk_bits_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                            const L1Seg *__restrict__ segs,
                                                            const int *__restrict__ lrank, u32 *gb, int64_t gwords,
                                                            int P, const float *__restrict__ bounds,
                                                            int direct_ranks) {
    const int c = blockIdx.x, b = blockIdx.y;
    const int dir = b & 1;
    const int f = (b >> 1) % F;
    const L1Seg sg = segs[(b >> 1) / F];
    const int n = sg.n, words = (n + 31) >> 5;
    u32 *bm = gb + ((int64_t)b * P + c) * gwords;
    for (int w = threadIdx.x; w < words; w += 256) bm[w] = 0u;
    __syncthreads();
    const int *col = orders + (int64_t)f * stride + sg.start;
    const int q_end = walk_chunk_begin(n, bounds, c + 1, P);
    for (int q = walk_chunk_begin(n, bounds, c, P) + threadIdx.x; q < q_end; q += 256) {
        const int pos = dir ? n - 1 - q : q;
        const int r = direct_ranks ? col[pos] : lrank[col[pos]];
        atomicOr(bm + (r >> 5), 1u << (r & 31));
    }
}

// bitmap of chunk c <- OR of the bitmaps of chunks 0 .. c-1
__global__ void __launch_bounds__(256) l1_chunk_prefix_kernel(const L1Seg *__restrict__ segs, int F, u32 *gb,
                                                              int64_t gwords, int P) {
    const int b = blockIdx.y;
    const L1Seg sg = segs[(b >> 1) / F];
    const int words = (sg.n + 31) >> 5;
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= words) return;
    u32 *p = gb + (int64_t)b * P * gwords + w;
    u32 run = 0u;
    for (int c = 0; c < P; ++c) {
        const u32 t = p[(int64_t)c * gwords];
        p[(int64_t)c * gwords] = run;
        run |= t;
    }
}

// One warp per (node, column, direction).  dir 0: positions 0..n-1 (prefix medians), dir 1:
// positions n-1..0 (suffix medians).  out_lo/out_hi[(dir*F+f)*stride + start + pos] = ranks of the
// (lower, upper) median of the elements scanned up to and including position pos.
// All 32 lanes run the walk redundantly on identical state (ranks are broadcast by shuffle, lane j
// keeps the pair of step j), so there is no divergence and no staging through shared memory.
template <bool L2ONLY>
__device__ __forceinline__ void walk_chunk(Walk<L2ONLY> &wk, const int *__restrict__ col, const int *__restrict__ lrank,
                                           int direct_ranks, int dir, int n, int q_begin, int q_end, int *olo,
                                           int *ohi, int lane) {
    // software pipeline: the ranks of the next 32 positions are fetched while this group is walked
    auto fetch = [&](int q0) -> int {
        const int q = q0 + lane;
        if (q >= q_end) return 0;
        const int pos = dir ? n - 1 - q : q;
        return direct_ranks ? col[pos] : lrank[col[pos]];
    };
    int nxt = fetch(q_begin);
    for (int q0 = q_begin; q0 < q_end; q0 += 32) {
        const int mine = nxt;
        nxt = fetch(q0 + 32);
        int mylo = 0, myhi = 0;
        if (q0 + 32 <= q_end) {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                wk.insert(__shfl_sync(FULL_MASK, mine, j));
                if (lane == j) {
                    mylo = wk.lo;
                    myhi = wk.hi;
                }
            }
        } else {
            const int cnt = q_end - q0;
            for (int j = 0; j < cnt; ++j) {
                wk.insert(__shfl_sync(FULL_MASK, mine, j));
                if (lane == j) {
                    mylo = wk.lo;
                    myhi = wk.hi;
                }
            }
        }
        const int q = q0 + lane;
        if (q < q_end) {
            const int pos = dir ? n - 1 - q : q;
            olo[pos] = mylo;
            ohi[pos] = myhi;
        }
    }
}

// positions (rank values) of the k1-th and k2-th set bits (0-based, k1 <= k2 <= k1 + 1) of a bitmap
__device__ __forceinline__ void select_two(const u32 *bm, int words, int k1, int k2, int lane, int &p1, int &p2) {
    p1 = p2 = -1;
    int run = 0;
    for (int w0 = 0; w0 < words && p2 < 0; w0 += 32) {
        const u32 v = (w0 + lane < words) ? ld_relaxed_u32(bm + w0 + lane) : 0u;
        const int pc = __popc(v);
        int inc = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += u;
        }
        const int before = run + inc - pc;
        if (p1 < 0) {
            const u32 has = __ballot_sync(FULL_MASK, before <= k1 && k1 < before + pc);
            if (has) {
                const int src = __ffs(has) - 1;
                const int pos = ((w0 + lane) << 5) + (int)__fns(v, 0, k1 - before + 1);
                p1 = __shfl_sync(FULL_MASK, pos, src);
            }
        }
        const u32 has2 = __ballot_sync(FULL_MASK, before <= k2 && k2 < before + pc);
        if (has2) {
            const int src = __ffs(has2) - 1;
            const int pos = ((w0 + lane) << 5) + (int)__fns(v, 0, k2 - before + 1);
            p2 = __shfl_sync(FULL_MASK, pos, src);
        }
        run += __shfl_sync(FULL_MASK, inc, 31);
    }
}

// P == 1: one warp per walk, bitmap in shared memory (or, for huge nodes, in global memory).
// P  > 1: one warp per (chunk, walk), chunk-major block order; bitmaps prepared by the two kernels above.
__global__ void __launch_bounds__(32) l1_median_walk_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                            const L1Seg *__restrict__ segs,
                                                            const int *__restrict__ lrank, int *__restrict__ out_lo,
                                                            int *__restrict__ out_hi, u32 *gbitmap,
                                                            int64_t gwords_per_block, int direct_ranks, int P,
                                                            int n_walks, const float *__restrict__ bounds) {
    extern __shared__ u32 smem[];
    const int lane = threadIdx.x;
    const int c = blockIdx.x / n_walks, b = blockIdx.x % n_walks;
    const int dir = b & 1;
    const int f = (b >> 1) % F;
    const L1Seg sg = segs[(b >> 1) / F];
    const int n = sg.n;
    const int words = (n + 31) >> 5;
    const int *col = orders + (int64_t)f * stride + sg.start;
    int *olo = out_lo + ((int64_t)dir * F + f) * stride + sg.start;
    int *ohi = out_hi + ((int64_t)dir * F + f) * stride + sg.start;
    if (P == 1) {
        Walk<false> wk;
        wk.writer = true;
        wk.bm = gbitmap ? gbitmap + (int64_t)b * gwords_per_block : smem;
        for (int w = lane; w < words; w += 32) wk.bm[w] = 0u;
        __syncwarp();
        wk.m = wk.lo = wk.hi = wk.wlo = wk.whi = 0;
        wk.vlo = wk.vhi = 0u;
        walk_chunk(wk, col, lrank, direct_ranks, dir, n, 0, n, olo, ohi, lane);
        return;
    }
    const int q_begin = walk_chunk_begin(n, bounds, c, P), q_end = walk_chunk_begin(n, bounds, c + 1, P);
    if (q_begin >= q_end) return;
    u32 *bm = gbitmap + ((int64_t)b * P + c) * gwords_per_block;
    if (q_begin == 0) {
        // the first chunk starts from the empty set and is sparse throughout (long scans over empty words):
        // plain cached accesses keep its private bitmap in L1; these blocks come first, one or two per SM
        Walk<false> wk;
        wk.writer = true;
        wk.bm = bm;
        wk.m = wk.lo = wk.hi = wk.wlo = wk.whi = 0;
        wk.vlo = wk.vhi = 0u;
        walk_chunk(wk, col, lrank, direct_ranks, dir, n, q_begin, q_end, olo, ohi, lane);
    } else {
        Walk<true> wk;
        wk.writer = lane == 0;
        wk.bm = bm;
        wk.m = q_begin;   // ranks already in the set; lower / upper median = order statistics (m-1)/2 and m/2
        select_two(bm, words, (q_begin - 1) >> 1, q_begin >> 1, lane, wk.lo, wk.hi);
        wk.load_lo();
        wk.load_hi();
        walk_chunk(wk, col, lrank, direct_ranks, dir, n, q_begin, q_end, olo, ohi, lane);
    }
}

// Cost increments and their running sums, one sequence per (node, column, direction):
//   dir 0: L[0] = 0, L[p] = L[p-1] + dist(y_p, pair of positions 0..p-1)        strategy_l1.py:89-102
//   dir 1: R[n-1] = inc(n-1), R[p] = R[p+1] + inc(p), pair of positions p..n-1   strategy_l1.py:103-117
// Sequence index q: dir 0 -> position p = q (the increment of q = 0 is 0), dir 1 -> p = n-1-q.
struct CostSeq {
    const int *col, *mlo, *mhi;
    const double *yn, *y;
    int64_t ybase;
    int direct_y, n, dir;

    template <int DIR>
    __device__ __forceinline__ double inc_d(int q) const {
        const int p = DIR ? n - 1 - q : q;
        if (DIR == 0 && p == 0) return 0.0;
        const double yv = direct_y ? y[ybase + p] : y[col[p]];
        const int pm = DIR ? p : p - 1;
        const double lo = yn[mlo[pm]], hi = yn[mhi[pm]];
        if (DIR == 0) {
            const double a = yv <= lo ? __dsub_rn(lo, yv) : 0.0;   // strategy_l1.py:93-95
            const double b = yv >= hi ? __dsub_rn(yv, hi) : 0.0;   // strategy_l1.py:96-98
            return __dadd_rn(a, b);
        }
        const double a = yv <= lo ? __dsub_rn(hi, yv) : 0.0;       // strategy_l1.py:110-111
        const double b = yv >= hi ? __dsub_rn(yv, lo) : 0.0;       // strategy_l1.py:112-114
        return __dadd_rn(a, b);
    }
    __device__ __forceinline__ double inc(int q) const { return dir ? inc_d<1>(q) : inc_d<0>(q); }
    // running sums of q_begin .. q_end-1 added ONE ELEMENT AT A TIME to S (np.cumsum's order)
    template <int DIR>
    __device__ __forceinline__ void scan_d(int q_begin, int q_end, double S, double *out, int lane) const {
        for (int q0 = q_begin; q0 < q_end; q0 += 32) {
            const int q = q0 + lane;
            const double v = q < q_end ? inc_d<DIR>(q) : 0.0;
            double mine = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                S = __dadd_rn(S, shfl_d(v, j));
                if (lane == j) mine = S;
            }
            if (q < q_end) out[DIR ? n - 1 - q : q] = mine;
        }
    }
    __device__ __forceinline__ void scan(int q_begin, int q_end, double S, double *out, int lane) const {
        if (dir)
            scan_d<1>(q_begin, q_end, S, out, lane);
        else
            scan_d<0>(q_begin, q_end, S, out, lane);
    }
};

__device__ __forceinline__ CostSeq cost_seq(int gw, const int *orders, int64_t stride, int F, const L1Seg *segs,
                                            const double *y, const double *ynode, const int *med_lo,
                                            const int *med_hi, int direct_y, double *Lc, double *Rc, double **out,
                                            int *seg_index) {
    CostSeq c;
    c.dir = gw & 1;
    const int f = (gw >> 1) % F;
    *seg_index = (gw >> 1) / F;
    const L1Seg sg = segs[*seg_index];
    c.n = sg.n;
    c.col = orders + (int64_t)f * stride + sg.start;
    c.mlo = med_lo + ((int64_t)c.dir * F + f) * stride + sg.start;
    c.mhi = med_hi + ((int64_t)c.dir * F + f) * stride + sg.start;
    c.yn = ynode + sg.start;
    c.y = y;
    c.ybase = (int64_t)f * stride + sg.start;
    c.direct_y = direct_y;
    *out = (c.dir ? Rc : Lc) + (int64_t)f * stride + sg.start;
    return c;
}

// exact: one warp per sequence, strictly sequential fp64 sums -- nodes flagged exact in dsegs
__global__ void __launch_bounds__(256) l1_cost_scan_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                           const L1Seg *__restrict__ segs, int n_segs,
                                                           const SegDev *__restrict__ dsegs,
                                                           const double *__restrict__ y,
                                                           const double *__restrict__ ynode,
                                                           const int *__restrict__ med_lo,
                                                           const int *__restrict__ med_hi, double *__restrict__ Lc,
                                                           double *__restrict__ Rc, int direct_y) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_segs * F * 2) return;
    double *out;
    int seg;
    const CostSeq c = cost_seq(gw, orders, stride, F, segs, y, ynode, med_lo, med_hi, direct_y, Lc, Rc, &out, &seg);
    if (!dsegs[seg].exact) return;
    c.scan(0, c.n, 0.0, out, threadIdx.x & 31);
}

// approximate (large nodes): the sequence is cut into P chunks.  PASS 0: sum of each chunk (any order);
// PASS 1: each chunk scans its range starting from the sum of the chunks before it.  The values differ
// from the sequential ones only in rounding; the near-tie rule (fit.cu) sends close calls to the exact kernel.
template <int PASS>
__global__ void __launch_bounds__(256) l1_cost_chunk_kernel(const int *__restrict__ orders, int64_t stride, int F,
                                                            const L1Seg *__restrict__ segs, int n_segs,
                                                            const SegDev *__restrict__ dsegs,
                                                            const double *__restrict__ y,
                                                            const double *__restrict__ ynode,
                                                            const int *__restrict__ med_lo,
                                                            const int *__restrict__ med_hi, double *__restrict__ Lc,
                                                            double *__restrict__ Rc, int direct_y, int P,
                                                            double *__restrict__ chunk_tot) {
    const int64_t gwc = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gwc >= (int64_t)n_segs * F * 2 * P) return;
    const int lane = threadIdx.x & 31;
    const int gw = (int)(gwc / P), ch = (int)(gwc % P);
    double *out;
    int seg;
    const CostSeq c = cost_seq(gw, orders, stride, F, segs, y, ynode, med_lo, med_hi, direct_y, Lc, Rc, &out, &seg);
    if (dsegs[seg].exact) return;
    const int S = walk_chunk_len(c.n, P);
    const int q_begin = ch * S, q_end = min(c.n, q_begin + S);
    if (PASS == 0) {
        double t = 0.0;
        for (int q = q_begin + lane; q < q_end; q += 32) t += c.inc(q);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) t += shfl_xor_d(t, m);
        if (lane == 0) chunk_tot[gwc] = t;
    } else {
        if (q_begin >= c.n) return;
        double base = 0.0;
        for (int k = 0; k < ch; ++k) base += chunk_tot[(int64_t)gw * P + k];
        c.scan(q_begin, q_end, base, out, lane);
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
        out.ppos = -1;
        out.pad_ = 0;
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
// B200DT_L1_EXACT=1: always the sequential cost sums (no chunked approximation)
bool l1_exact_env() {
    const char *e = getenv("B200DT_L1_EXACT");
    return e && e[0] == '1';
}

int launch_walk(b200dt_ctx *ctx, const int *orders, int64_t stride, int F, const L1Seg *d_segs, int n_segs,
                int max_n, int64_t n_elems, const int *lrank, int *med_lo, int *med_hi, int direct_ranks) {
    const size_t words = (size_t)(max_n + 31) / 32;
    const size_t smem = words * 4 + 16;
    const int n_walks = n_segs * F * 2;
    // chunks per walk: enough rows per chunk, and at most 8 GB of bitmaps
    int P = 1;
    const char *env = getenv("B200DT_L1_CHUNKS");
    if (max_n >= 2 * WALK_CHUNK_MIN) {
        // as many chunks as keep every (walk, chunk) resident at once: a second wave would double the time
        // (with more walks than half the slots the GPU is full anyway: no chunking; this also keeps the 2-D
        //  grids of the preparation kernels below the 65535 limit of gridDim.y)
        const int slots = ctx->sm_count * 32;
        P = std::min(std::min(WALK_CHUNKS_MAX, max_n / WALK_CHUNK_MIN), slots / std::max(1, n_walks));
        if (P < 2) P = 1;
    }
    if (env) P = std::max(1, std::min(WALK_CHUNKS_MAX, atoi(env)));
    if (n_walks > 65535) P = 1;
    while (P > 1 && (size_t)n_walks * P * words * 4 > ((size_t)8 << 30)) --P;
    u32 *gbm = nullptr;
    float *d_bounds = nullptr;
    size_t use_smem = smem;
    if (P > 1) {
        CKR(scratch_get(ctx, SB_L1_H, (size_t)n_walks * P * words * 4 + 256, (void **)&gbm));
        d_bounds = (float *)(gbm + (size_t)n_walks * P * words);
        std::vector<float> bounds;
        walk_chunk_bounds(max_n, P, bounds);
        void *stage;
        CKR(pinned_get(ctx, PIN_WALK, (size_t)(P + 1) * 4, &stage));
        memcpy(stage, bounds.data(), (size_t)(P + 1) * 4);
        CK(cudaMemcpyAsync(d_bounds, stage, (size_t)(P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        use_smem = 16;
        LaunchScope ls(ctx, KC_L1_DESCENT, 0.0, 2);
        l1_chunk_bits_kernel<<<dim3(P, n_walks), 256, 0, ctx->stream>>>(orders, stride, F, d_segs, lrank, gbm,
                                                                        (int64_t)words, P, d_bounds, direct_ranks);
        l1_chunk_prefix_kernel<<<dim3((unsigned)((words + 255) / 256), n_walks), 256, 0, ctx->stream>>>(
            d_segs, F, gbm, (int64_t)words, P);
        CK(cudaGetLastError());
    } else if (smem > (size_t)WALK_SMEM_MAX) {
        // rank bitmap too large for shared memory (> ~1.8 M rows in one node): keep it in HBM/L2
        CKR(scratch_get(ctx, SB_L1_H, (size_t)n_walks * words * 4, (void **)&gbm));
        use_smem = 16;
    }
    // (per context = per device: a process may drive several GPUs)
    if (!(ctx->attr_mask & 2u)) {
        CK(cudaFuncSetAttribute(l1_median_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK_SMEM_MAX));
        ctx->attr_mask |= 2u;
    }
    LaunchScope ls(ctx, KC_L1_DESCENT, 16.0 * (double)n_elems * F * 2);
    l1_median_walk_kernel<<<(unsigned)((int64_t)n_walks * P), 32, use_smem, ctx->stream>>>(
        orders, stride, F, d_segs, lrank, med_lo, med_hi, gbm, (int64_t)words, direct_ranks, P, n_walks, d_bounds);
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

// L1 split search of every live node (always in the reference's exact arithmetic).
int l1_level_search(b200dt_ctx *ctx, Session *s, b200dt_candidate *d_out, bool exact_only) {
    // exact_only: a re-search after a near tie.  The ranks and median pairs of this level are still in their
    // scratch buffers; only the cost sums of the flagged nodes are redone, in the reference's sequential order.
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
        // large nodes: chunked (parallel-order) cost sums, exact re-search on a near tie
        sg.exact = (nd.n < 2 * WALK_CHUNK_MIN || s->force_exact[i] || l1_exact_env()) ? 1 : 0;
        sg.pstep = 1;
        sg.pside = 0;
        sg.vstart = -1;
        sg.cbase = 0;
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
    if (!exact_only) {
        int gx = std::min(1024, (max_n + 255) / 256);
        for (int i0 = 0; i0 < n_live; i0 += 65535) {
            const int ni = std::min(65535, n_live - i0);
            LaunchScope ls(ctx, KC_L1_RANK, 24.0 * (double)n_elems);
            l1_rank_kernel<<<dim3(gx, ni), 256, 0, ctx->stream>>>(ycol, d_l1segs + i0, s->dy, lrank, ynode);
        }
        CKR(launch_walk(ctx, orders, stride, F, d_l1segs, n_live, max_n, n_elems, lrank, med_lo, med_hi, 0));
    }
    {
        const int64_t warps = (int64_t)n_live * F * 2;
        bool any_exact = false, any_chunked = false;
        for (const SegDev &sg : s->h_segs) (sg.exact ? any_exact : any_chunked) = true;
        if (any_exact) {
            LaunchScope ls(ctx, KC_L1_COST, 44.0 * (double)n_elems * F);
            l1_cost_scan_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(
                orders, stride, F, d_l1segs, n_live, s->d_segs, s->dy, ynode, med_lo, med_hi, Lc, Rc, 0);
        }
        if (any_chunked && !exact_only) {
            const int P = std::max(2, std::min(WALK_CHUNKS_MAX, max_n / WALK_CHUNK_MIN));
            double *tot;
            CKR(scratch_get(ctx, SB_L1_G, (size_t)warps * P * 8 + 64, (void **)&tot));
            const unsigned grid = (unsigned)((warps * P + 7) / 8);
            LaunchScope ls(ctx, KC_L1_COST, 88.0 * (double)n_elems * F, 2);
            l1_cost_chunk_kernel<0><<<grid, 256, 0, ctx->stream>>>(orders, stride, F, d_l1segs, n_live, s->d_segs, s->dy,
                                                                   ynode, med_lo, med_hi, Lc, Rc, 0, P, tot);
            l1_cost_chunk_kernel<1><<<grid, 256, 0, ctx->stream>>>(orders, stride, F, d_l1segs, n_live, s->d_segs, s->dy,
                                                                   ynode, med_lo, med_hi, Lc, Rc, 0, P, tot);
        }
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
    a.row_map = nullptr;
    a.node_best = nullptr;
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
