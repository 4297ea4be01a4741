// K6: exhaustive multi-class split search (gini, precision-at-1) and the classification fit.
//
// replaces greedy_classification / greedy_classification_p_at_k   (ref decision_tree/strategy.py:272-283, 313-323)
//          surrogate_gini_improvements / surrogate_p_at_k_improvements        (ref strategy.py:32-55, 79-100)
//          build_classification_tree -> build_tree                          (ref tree_builder.py:56-96, 149-159)
//
// The reference re-sorts every node (`np.argsort(X[rows])`), gathers the one-hot targets into an
// (n, F, C) array and cumsums it.  Here the columns are sorted ONCE; the class id of a row rides in
// the bits above its row id inside the 32-bit order entry, so one 4-byte stream per (row, feature)
// carries everything a level needs, and the stable partition (partition.cu, PAY == 3) keeps every
// node's columns sorted -- for distinct feature values exactly the order the reference's per-node
// argsort produces.
//
// All scanned quantities are whole numbers, so the search is EXACT (no summation-order issue):
// with r_j = number of earlier same-class rows of sorted position j,
//     sum_c L_c(k)^2        = sum_{j<=k} (2 r_j + 1)                              =: A(k)
//     sum_c R_c(k)^2        = sum_c T_c^2 - 2 sum_{j<=k} T_{c_j} + A(k)           (R_c = T_c - L_c)
//     max_c L_c(k)          = max_{j<=k} (r_j + 1)
//     max_c R_c(k)          = max_{j>k} (T_{c_j} - r_j)
// i.e. plain integer prefix sums / prefix maxima of per-row terms once r_j is known.  r_j needs the
// per-class counts at the start of the tile: pass 1 writes per-tile class histograms, pass 2 turns
// them into exclusive prefixes per (node, column), pass 3 re-reads the tile and scores it.
// Algorithmic bytes: 2 x 4 per live (row, feature) per level (+ 4 + 1 + 4 for the partition).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include <math_constants.h>

#include "class_div.cuh"
#include "session.cuh"

constexpr int GT_ROWS = 32;
constexpr int GT_TILE = 32 * GT_ROWS;  // sorted positions per warp tile

struct CSeg {
    long long st2;   // sum_c T_c^2 of the node
    int start, n;
    int tile_base;   // first tile of this node in the level's tile list
    int ntiles;
    int tmax;        // max_c T_c
    int pad;
};

struct CPart {
    double score;
    int k;  // position inside the node; INT_MAX = no candidate in this tile
    int pad;
};

struct CCand {
    double score;      // best surrogate value of the node
    double threshold;  // feature value at (k, f)
    int k, f;
    int k_ext;         // last position whose feature value equals the threshold (value-based split)
    int pad;
};

constexpr int MAX_CLASS_CHUNKS_HOST = 8;  // n_classes <= 256

namespace {

__device__ __forceinline__ u32 ld_stream_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// orders[f][p] |= label[row] << row_bits
__global__ void __launch_bounds__(256) pack_labels_kernel(u32 *orders, int64_t col_stride, int F, int64_t N,
                                                          const int *__restrict__ labels, int row_bits) {
    const int f = blockIdx.y;
    u32 *col = orders + (int64_t)f * col_stride;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < N; p += (int64_t)gridDim.x * 256) {
        const u32 row = col[p];
        col[p] = row | ((u32)__ldg(labels + row) << row_bits);
    }
}

__global__ void __launch_bounds__(256) label_check_hist_kernel(const int *__restrict__ labels, int64_t N, int C,
                                                               unsigned long long *hist, int *bad) {
    __shared__ int sh[32 * MAX_CLASS_CHUNKS_HOST];
    const u32 lt = lanemask_lt();
    for (int c = threadIdx.x; c < C; c += 256) sh[c] = 0;
    __syncthreads();
    const int64_t n_round = (N + 255) / 256 * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        int c = i < N ? labels[i] : -1;
        if (i < N && (c < 0 || c >= C)) {
            *bad = 1;
            c = -1;
        }
        const u32 peers = __match_any_sync(FULL_MASK, c);
        if (c >= 0 && (peers & lt) == 0) atomicAdd(&sh[c], __popc(peers));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256)
        if (sh[c]) atomicAdd(hist + c, (unsigned long long)sh[c]);
}

// ---- pass 1: class histogram of every (tile, column) ------------------------------------
__global__ void __launch_bounds__(256) class_hist_kernel(const u32 *__restrict__ orders, int64_t col_stride,
                                                         const CSeg *__restrict__ segs,
                                                         const TileDesc *__restrict__ tiles, int n_tiles, int F,
                                                         int C, int row_bits, int *__restrict__ hist) {
    extern __shared__ int sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    int *h = sm + warp * (C + 1);
    const int64_t total = (int64_t)n_tiles * F;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total; t += (int64_t)gridDim.x * 8) {
        const int f = (int)(t % F), ti = (int)(t / F);
        const TileDesc td = tiles[ti];
        const CSeg sg = segs[td.seg];
        const int off = td.tin * GT_TILE;
        const int cnt = min(GT_TILE, sg.n - off);
        const u32 *col = orders + (int64_t)f * col_stride + sg.start + off;
        u32 v[GT_ROWS];
#pragma unroll
        for (int i = 0; i < GT_ROWS; ++i) v[i] = (i * 32 + lane < cnt) ? ld_stream_u32(col + i * 32 + lane) : 0u;
        for (int c = lane; c <= C; c += 32) h[c] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < GT_ROWS; ++i) {
            const int c = (i * 32 + lane < cnt) ? (int)(v[i] >> row_bits) : C;  // slot C absorbs the padding
            const u32 peers = __match_any_sync(FULL_MASK, c);
            if ((peers & lt) == 0) h[c] += __popc(peers);  // one writer per class
            __syncwarp();
        }
        int *out = hist + ((int64_t)f * n_tiles + ti) * C;
        for (int c = lane; c < C; c += 32) out[c] = h[c];
        __syncwarp();
    }
}

// ---- pass 2: exclusive prefix of the histograms along the tiles of each (node, column) -----
// one warp per (node, column), lanes = classes.  Also the largest right-hand class count at
// the end of every tile (precision-at-1 needs it as the seed of its backward pass).
constexpr int MAX_CLASS_CHUNKS = MAX_CLASS_CHUNKS_HOST;

__global__ void __launch_bounds__(256) class_scan_kernel(const CSeg *__restrict__ segs, int n_segs, int F, int C,
                                                         int n_tiles, const int *__restrict__ tot,
                                                         const int *__restrict__ hist, int *__restrict__ excl,
                                                         int *__restrict__ rmax_end) {
    const int lane = threadIdx.x & 31;
    const int64_t n_work = (int64_t)n_segs * F;
    const int nch = (C + 31) / 32;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_work;
         w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int seg = (int)(w / F), f = (int)(w % F);
        const CSeg sg = segs[seg];
        int run[MAX_CLASS_CHUNKS], T[MAX_CLASS_CHUNKS];
#pragma unroll
        for (int ch = 0; ch < MAX_CLASS_CHUNKS; ++ch) {
            run[ch] = 0;
            const int c = ch * 32 + lane;
            T[ch] = (ch < nch && c < C) ? tot[(int64_t)seg * C + c] : 0;
        }
        const int64_t base = ((int64_t)f * n_tiles + sg.tile_base) * C;
#pragma unroll 4
        for (int t = 0; t < sg.ntiles; ++t) {
            int rm = 0;
#pragma unroll
            for (int ch = 0; ch < MAX_CLASS_CHUNKS; ++ch) {
                const int c = ch * 32 + lane;
                if (ch < nch && c < C) {
                    const int64_t idx = base + (int64_t)t * C + c;
                    const int hv = hist[idx];
                    excl[idx] = run[ch];
                    run[ch] += hv;
                    rm = max(rm, T[ch] - run[ch]);
                }
            }
            if (rmax_end) {
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) rm = max(rm, __shfl_xor_sync(FULL_MASK, rm, m));
                if (lane == 0) rmax_end[(int64_t)f * n_tiles + sg.tile_base + t] = rm;
            }
        }
    }
}

// ---- pass 3: score every split position of the tile, keep the best ------------------------
// CRIT 0: gini surrogate (strategy.py:32-55); CRIT 1: precision-at-1 surrogate (strategy.py:79-100)
template <int CRIT>
__global__ void __launch_bounds__(256) class_score_kernel(const u32 *__restrict__ orders, int64_t col_stride,
                                                          const CSeg *__restrict__ segs,
                                                          const TileDesc *__restrict__ tiles, int n_tiles, int F,
                                                          int C, int row_bits, const int *__restrict__ excl,
                                                          const int *__restrict__ tot,
                                                          const int *__restrict__ rmax_end,
                                                          CPart *__restrict__ partials) {
    extern __shared__ int sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    const int per_warp = 2 * (C + 1) + (CRIT == 1 ? 2 * GT_TILE : 0);
    int *cnt = sm + warp * per_warp;  // running class counts (slot C: padding lanes)
    int *T = cnt + (C + 1);           // class totals of the node
    int *s_ml = T + (C + 1);          // CRIT 1: max_c L_c at each position
    int *s_sr = s_ml + GT_TILE;       // CRIT 1: T_c - r of each position
    const int64_t total = (int64_t)n_tiles * F;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total; t += (int64_t)gridDim.x * 8) {
        const int f = (int)(t % F), ti = (int)(t / F);
        const TileDesc td = tiles[ti];
        const CSeg sg = segs[td.seg];
        const int off = td.tin * GT_TILE;
        const int n = sg.n;
        const int cnt_here = min(GT_TILE, n - off);
        const u32 *col = orders + (int64_t)f * col_stride + sg.start + off;
        u32 v[GT_ROWS];
#pragma unroll
        for (int i = 0; i < GT_ROWS; ++i) v[i] = (i * 32 + lane < cnt_here) ? ld_stream_u32(col + i * 32 + lane) : 0u;
        const int *ex = excl + ((int64_t)f * n_tiles + ti) * C;
        long long a0 = 0, b0 = 0;
        int m0 = 0;
        for (int c = lane; c < C; c += 32) {
            const int e = ex[c], tc = tot[(int64_t)td.seg * C + c];
            cnt[c] = e;
            T[c] = tc;
            a0 += (long long)e * e;
            b0 += (long long)tc * e;
            m0 = max(m0, e);
        }
        if (lane == 0) {
            cnt[C] = 0;
            T[C] = 0;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            a0 += __shfl_xor_sync(FULL_MASK, a0, m);
            b0 += __shfl_xor_sync(FULL_MASK, b0, m);
            m0 = max(m0, __shfl_xor_sync(FULL_MASK, m0, m));
        }
        __syncwarp();
        double best = -CUDART_INF;
        int best_k = 0x7fffffff;
        long long a_run = a0, b_run = b0;
        int ml_run = m0;
#pragma unroll
        for (int i = 0; i < GT_ROWS; ++i) {
            const int e = i * 32 + lane;
            const bool ok = e < cnt_here;
            const int c = ok ? (int)(v[i] >> row_bits) : C;
            const u32 peers = __match_any_sync(FULL_MASK, c);
            const int r = cnt[c] + __popc(peers & lt);  // earlier rows of the same class in this node
            __syncwarp();
            if ((peers & lt) == 0) cnt[c] += __popc(peers);
            __syncwarp();
            const int kk = off + e;
            if (CRIT == 0) {
                u32 a = ok ? 2u * (u32)r + 1u : 0u, b = ok ? (u32)T[c] : 0u;  // row sums stay below 2^32 (N <= 2^26)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 ua = __shfl_up_sync(FULL_MASK, a, d), ub = __shfl_up_sync(FULL_MASK, b, d);
                    if (lane >= d) {
                        a += ua;
                        b += ub;
                    }
                }
                if (ok && kk < n - 1) {
                    const long long SL = a_run + (long long)a;
                    const long long SR = sg.st2 - 2 * (b_run + (long long)b) + SL;
                    // strategy.py:52-55: left_sq / left_count + right_sq / right_count, IEEE fp64
                    const double sc = __dadd_rn(div_counts((double)SL, (double)(kk + 1)),
                                                div_counts((double)SR, (double)(n - kk - 1)));
                    if (sc > best) {
                        best = sc;
                        best_k = kk;
                    }
                }
                a_run += (long long)__shfl_sync(FULL_MASK, a, 31);
                b_run += (long long)__shfl_sync(FULL_MASK, b, 31);
            } else {
                int ml = ok ? r + 1 : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, ml, d);
                    if (lane >= d) ml = max(ml, u);
                }
                ml = max(ml, ml_run);
                s_ml[e] = ml;
                s_sr[e] = ok ? T[c] - r : 0;
                ml_run = __shfl_sync(FULL_MASK, ml, 31);
            }
        }
        if (CRIT == 1) {
            __syncwarp();
            int mr_run = rmax_end[(int64_t)f * n_tiles + ti];  // largest class count right of this tile
#pragma unroll 4
            for (int i = GT_ROWS - 1; i >= 0; --i) {
                const int e = i * 32 + lane;
                int sfx = s_sr[e];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_down_sync(FULL_MASK, sfx, d);
                    if (lane + d < 32) sfx = max(sfx, u);
                }
                int after = __shfl_down_sync(FULL_MASK, sfx, 1);  // max over the later lanes of this row
                if (lane == 31) after = 0;
                const int mr = max(mr_run, after);
                const int kk = off + e;
                if (e < cnt_here && kk < n - 1) {
                    const double sc = (double)(s_ml[e] + mr);  // strategy.py:97-100
                    if (sc >= best) {  // positions are visited in decreasing order: ties go to the lower one
                        best = sc;
                        best_k = kk;
                    }
                }
                mr_run = max(mr_run, __shfl_sync(FULL_MASK, sfx, 0));
            }
            __syncwarp();
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double ob = shfl_xor_d(best, m);
            const int ok2 = __shfl_xor_sync(FULL_MASK, best_k, m);
            if (ob > best || (ob == best && ok2 < best_k)) {
                best = ob;
                best_k = ok2;
            }
        }
        if (lane == 0) {
            CPart p;
            p.score = best;
            p.k = best_k;
            p.pad = 0;
            partials[(int64_t)f * n_tiles + ti] = p;
        }
        __syncwarp();
    }
}

// ---- pass 3, gini with at most 32 classes: lane-contiguous scoring ---------------------------
// Each lane owns 32 CONSECUTIVE positions of the tile, so the running sums are plain per-lane
// registers and no per-row warp scan is needed.  A lane's class counters {L_c, R_c = T_c - L_c}
// live in shared memory as one 64-bit word per (class, lane) (bank = lane: conflict-free); their
// start values come from one warp scan per class over the lanes' chunk histograms.  Adding a
// row of class c moves sum_c L_c^2 by 2 L_c + 1 and sum_c R_c^2 by -(2 R_c - 1).
constexpr int LANE_MAX_CLASSES = 32;
constexpr int STAGE_PITCH = 33;

//
// LB = true: single pass.  The class counts at the tile start come from a decoupled look-back along
// the tiles of the (node, column) -- lane c carries class c, one tagged 64-bit status word per
// (tile, column, class) as in lookback_count -- instead of the histogram + scan passes.  The grid
// must be co-resident and tiles are taken in tile-major order (a tile's predecessor has a smaller
// work index), exactly like the L2 scan (l2.cu).
//
// CRIT 1 (precision-at-1): max_c L_c only grows along the chunk (forward pass, stored per position) and
// max_c R_c only grows when walking the chunk backwards from its end state, so two sweeps over the
// same counters give max_c L_c(k) + max_c R_c(k) for every position -- no division, integer scores.
template <bool LB, int CRIT>
__global__ void __launch_bounds__(256) class_score_lane_kernel(const u32 *__restrict__ orders, int64_t col_stride,
                                                               const CSeg *__restrict__ segs,
                                                               const TileDesc *__restrict__ tiles, int n_tiles, int F,
                                                               int C, int row_bits, const int *__restrict__ excl,
                                                               const int *__restrict__ tot, u64 *status, u32 epoch,
                                                               unsigned long long *node_best,
                                                               CPart *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = (size_t)(CRIT == 1 ? 2 : 1) * 32 * STAGE_PITCH * 4 + (size_t)C * 32 * 8;
    u64 *pk = reinterpret_cast<u64 *>(smraw + warp * per_warp);             // [C][32] {L (low), R (high)}
    u32 *stage = reinterpret_cast<u32 *>(smraw + warp * per_warp + (size_t)C * 32 * 8);  // [32][33]
    int *mlbuf = reinterpret_cast<int *>(stage + 32 * STAGE_PITCH);         // CRIT 1: [32][33] max_c L_c per position
    const int64_t total = (int64_t)n_tiles * F;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total; t += (int64_t)gridDim.x * 8) {
        const int f = (int)(t % F), ti = (int)(t / F);
        const TileDesc td = tiles[ti];
        const CSeg sg = segs[td.seg];
        const int off = td.tin * GT_TILE;
        const int n = sg.n;
        const int cnt_here = min(GT_TILE, n - off);
        const u32 *col = orders + (int64_t)f * col_stride + sg.start + off;
        // coalesced load, transposed through shared memory: element e -> stage[e / 32][e % 32]
#pragma unroll
        for (int i = 0; i < GT_ROWS; ++i)
            stage[i * STAGE_PITCH + lane] = (i * 32 + lane < cnt_here) ? ld_stream_u32(col + i * 32 + lane) : 0xffffffffu;
        for (int c = 0; c < C; ++c) pk[c * 32 + lane] = 0ull;
        __syncwarp();
        const int my0 = lane * 32;                       // first position of this lane inside the tile
        const int mine = max(0, min(32, cnt_here - my0));  // valid positions of this lane
        const u32 *my = stage + lane * STAGE_PITCH;
        u32 *pk32 = reinterpret_cast<u32 *>(pk);
        for (int j = 0; j < mine; ++j) {
            const int c = (int)(my[j] >> row_bits);
            pk32[(c * 32 + lane) * 2] += 1u;
        }
        // start counts of every class for this lane: counts at the tile start + the earlier lanes' rows
        const int *ex = LB ? nullptr : excl + ((int64_t)f * n_tiles + ti) * C;
        const int *tt = tot + (int64_t)td.seg * C;
        int base = 0;  // LB: lane c holds the count of class c at the tile start
        if (LB) {
            int my_agg = 0;
            for (int c = 0; c < C; ++c) {
                const int h = (int)pk32[(c * 32 + lane) * 2];
                int inc = h;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, inc, d);
                    if (lane >= d) inc += u;
                }
                pk32[(c * 32 + lane) * 2 + 1] = (u32)(inc - h);   // rows of class c in the earlier lanes
                const int agg = __shfl_sync(FULL_MASK, inc, 31);
                if (lane == c) my_agg = agg;
            }
            const u64 tag = (u64)(epoch & 0x3fffffffu) << 32;
            u64 *st = status + ((int64_t)f * n_tiles + ti) * C + lane;
            const bool has_class = lane < C;
            if (td.tin == 0) {
                if (has_class) st_relaxed_u64(st, (2ull << 62) | tag | (u32)my_agg);
            } else {
                if (has_class) st_relaxed_u64(st, (1ull << 62) | tag | (u32)my_agg);
                bool done = !has_class;
                const u64 *look = st - C;
                while (__any_sync(FULL_MASK, !done)) {
                    if (!done) {
                        const u64 w = ld_relaxed_u64(look);
                        if ((w & 0x3fffffff00000000ull) == tag && (w >> 62) != 0) {
                            base += (int)(u32)w;
                            if ((w >> 62) == 2)
                                done = true;
                            else
                                look -= C;
                        }
                    }
                }
                if (has_class) st_relaxed_u64(st, (2ull << 62) | tag | (u32)(base + my_agg));
            }
        }
        long long SL = 0, SR = 0;
        int ml = 0;   // CRIT 1: largest left-hand class count so far
        for (int c = 0; c < C; ++c) {
            int L0;
            if (LB) {
                L0 = __shfl_sync(FULL_MASK, base, c) + (int)pk32[(c * 32 + lane) * 2 + 1];
            } else {
                const int h = (int)pk32[(c * 32 + lane) * 2];
                int inc = h;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, inc, d);
                    if (lane >= d) inc += u;
                }
                L0 = __ldg(ex + c) + inc - h;
            }
            const int R0 = __ldg(tt + c) - L0;
            pk[c * 32 + lane] = (u64)(u32)L0 | ((u64)(u32)R0 << 32);
            SL += (long long)L0 * L0;
            SR += (long long)R0 * R0;
            ml = max(ml, L0);
        }
        double best = -CUDART_INF;
        int best_k = 0x7fffffff;
        if (CRIT == 0) {
            // node_best[seg] = the largest EXACT score any tile of this node (any column) has produced so far.
            // A candidate whose cheap fp32 upper bound is below it cannot be the node's argmax (strictly smaller
            // exact score), so its two exact divisions are skipped; ties are never skipped.
            const bool filter = node_best != nullptr;
            const double bound = filter ? __longlong_as_double((long long)ld_relaxed_u64(node_best + td.seg)) : -1.0;
            // float shadows of the two sums of squares that are >= the true values BY CONSTRUCTION: converted once per
            // chunk rounding up, then advanced with round-up additions / subtractions of rounded-up / rounded-down
            // terms; reciprocals of rounded-down counts, inflated by the approximation's 1-ulp error; round-up
            // product and sum.  So `ub` is a rigorous upper bound of the fp64 score (just not a tight one)
            float SLf = __ll2float_ru(SL), SRf = __ll2float_ru(SR);
            for (int j = 0; j < mine; ++j) {
                const int c = (int)(my[j] >> row_bits);
                const u64 w = pk[c * 32 + lane];
                const int L = (int)(u32)w, R = (int)(u32)(w >> 32);
                SL += 2ll * L + 1;
                SR -= 2ll * R - 1;
                pk[c * 32 + lane] = (u64)(u32)(L + 1) | ((u64)(u32)(R - 1) << 32);
                const int kk = off + my0 + j;
                bool exact = kk < n - 1;
                if (filter) {
                    SLf = __fadd_ru(SLf, __int2float_ru(2 * L + 1));
                    SRf = __fsub_ru(SRf, __int2float_rd(2 * R - 1));
                    float rl, rr;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(__int2float_rd(kk + 1)));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(__int2float_rd(n - kk - 1)));
                    rl = __fmul_ru(rl, 1.0000003f);   // rcp.approx: max relative error 2^-23
                    rr = __fmul_ru(rr, 1.0000003f);
                    const float ub = __fmaf_ru(SLf, rl, __fmul_ru(SRf, rr));
                    exact = exact && (double)ub >= fmax(bound, best);
                }
                if (exact) {
                    // strategy.py:52-55: left_sq / left_count + right_sq / right_count, IEEE fp64
                    const double sc = __dadd_rn(div_counts((double)SL, (double)(kk + 1)),
                                                div_counts((double)SR, (double)(n - kk - 1)));
                    if (sc > best) {   // positions increase: strict ">" keeps the lowest one
                        best = sc;
                        best_k = kk;
                    }
                }
            }
        } else {
            int *myml = mlbuf + lane * STAGE_PITCH;
            for (int j = 0; j < mine; ++j) {   // forward: the chunk's rows join the left side one by one
                const int c = (int)(my[j] >> row_bits);
                const u64 w = pk[c * 32 + lane];
                const int L = (int)(u32)w + 1, R = (int)(u32)(w >> 32) - 1;
                ml = max(ml, L);
                myml[j] = ml;
                pk[c * 32 + lane] = (u64)(u32)L | ((u64)(u32)R << 32);
            }
            int mr = 0;   // largest right-hand class count behind this lane's chunk
            for (int c = 0; c < C; ++c) mr = max(mr, (int)pk32[(c * 32 + lane) * 2 + 1]);
            int best_i = -1;
            for (int j = mine - 1; j >= 0; --j) {   // backward: rows return to the right side
                const int kk = off + my0 + j;
                if (kk < n - 1) {
                    const int sc = myml[j] + mr;   // strategy.py:97-100
                    if (sc >= best_i) {            // positions decrease: ">=" keeps the lowest one
                        best_i = sc;
                        best_k = kk;
                    }
                }
                const int c = (int)(my[j] >> row_bits);
                const int R = (int)pk32[(c * 32 + lane) * 2 + 1] + 1;
                pk32[(c * 32 + lane) * 2 + 1] = (u32)R;
                mr = max(mr, R);
            }
            if (best_i >= 0) best = (double)best_i;
        }
        __syncwarp();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double ob = shfl_xor_d(best, m);
            const int ok2 = __shfl_xor_sync(FULL_MASK, best_k, m);
            if (ob > best || (ob == best && ok2 < best_k)) {
                best = ob;
                best_k = ok2;
            }
        }
        if (lane == 0) {
            CPart p;
            p.score = best;
            p.k = best_k;
            p.pad = 0;
            partials[(int64_t)f * n_tiles + ti] = p;
            // scores are positive, so their bit patterns order like the values
            if (CRIT == 0 && node_best && best_k != 0x7fffffff)
                atomicMax(node_best + td.seg, (unsigned long long)__double_as_longlong(best));
        }
        __syncwarp();
    }
}

// ---- per-node arg-best over (tile, column): flat row-major rule (position, then column) -----
// strategy.py:277-278: np.unravel_index(np.argmax(improvements), shape) on an (n-1, F) array
constexpr int CF_SLICES = 64;

struct CBest {
    double score;
    int k, f;
};

__device__ __forceinline__ bool cbest_better(double os, int ok, int of, double ms, int mk, int mf) {
    if (ok == 0x7fffffff) return false;
    if (mk == 0x7fffffff) return true;
    return os > ms || (os == ms && (ok < mk || (ok == mk && of < mf)));
}

// block-wide reduction of one candidate per thread; result valid in thread 0
__device__ __forceinline__ CBest cbest_block_reduce(CBest mine) {
    __shared__ double s_score[256];
    __shared__ int s_k[256], s_f[256];
    s_score[threadIdx.x] = mine.score;
    s_k[threadIdx.x] = mine.k;
    s_f[threadIdx.x] = mine.f;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if ((int)threadIdx.x < m &&
            cbest_better(s_score[threadIdx.x + m], s_k[threadIdx.x + m], s_f[threadIdx.x + m], s_score[threadIdx.x],
                         s_k[threadIdx.x], s_f[threadIdx.x])) {
            s_score[threadIdx.x] = s_score[threadIdx.x + m];
            s_k[threadIdx.x] = s_k[threadIdx.x + m];
            s_f[threadIdx.x] = s_f[threadIdx.x + m];
        }
        __syncthreads();
    }
    CBest r;
    r.score = s_score[0];
    r.k = s_k[0];
    r.f = s_f[0];
    return r;
}

// stage 1: grid (n_live, CF_SLICES): every block reduces one slice of the node's records
__global__ void __launch_bounds__(256) class_finalize_slices_kernel(const CSeg *__restrict__ segs, int F, int n_tiles,
                                                                    const CPart *__restrict__ partials,
                                                                    CBest *__restrict__ sliced) {
    const CSeg sg = segs[blockIdx.x];
    const int64_t n_rec = (int64_t)sg.ntiles * F;
    const int64_t per = (n_rec + CF_SLICES - 1) / CF_SLICES;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(n_rec, r0 + per);
    CBest b;
    b.score = -CUDART_INF;
    b.k = 0x7fffffff;
    b.f = 0x7fffffff;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
        const int f = (int)(r / sg.ntiles), tt = (int)(r % sg.ntiles);
        const CPart p = partials[(int64_t)f * n_tiles + sg.tile_base + tt];
        if (cbest_better(p.score, p.k, f, b.score, b.k, b.f)) {
            b.score = p.score;
            b.k = p.k;
            b.f = f;
        }
    }
    b = cbest_block_reduce(b);
    if (threadIdx.x == 0) sliced[(int64_t)blockIdx.x * CF_SLICES + blockIdx.y] = b;
}

// stage 2: one block per node: best slice, then threshold and the end of its run of equal values
__global__ void __launch_bounds__(256) class_finalize_kernel(const CSeg *__restrict__ segs,
                                                             const CBest *__restrict__ sliced,
                                                             const u32 *__restrict__ orders, int64_t col_stride,
                                                             int row_mask, const double *__restrict__ X, int64_t ldx,
                                                             int col_offset, CCand *__restrict__ out) {
    const CSeg sg = segs[blockIdx.x];
    CBest b;
    b.score = -CUDART_INF;
    b.k = 0x7fffffff;
    b.f = 0x7fffffff;
    if (threadIdx.x < CF_SLICES) b = sliced[(int64_t)blockIdx.x * CF_SLICES + threadIdx.x];
    b = cbest_block_reduce(b);
    if (threadIdx.x == 0) {
        CCand c;
        c.score = b.score;
        c.k = b.k;
        c.f = b.f;
        c.k_ext = c.k;
        c.threshold = 0.0;
        c.pad = 0;
        if (c.k != 0x7fffffff) {
            const u32 *col = orders + (int64_t)c.f * col_stride + sg.start;
            const double thr = X[(int64_t)(col[c.k] & (u32)row_mask) * ldx + c.f];  // strategy.py:279-280
            // tree_builder.py:31-34 splits by value (<= threshold): equal values right of k go left too
            int lo = c.k, hi = sg.n - 1;
            while (lo < hi) {
                const int mid = lo + (hi - lo + 1) / 2;
                if (X[(int64_t)(col[mid] & (u32)row_mask) * ldx + c.f] <= thr)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            c.threshold = thr;
            c.k_ext = lo;
            c.f += col_offset;   // candidates carry GLOBAL column ids
        } else {
            c.k = -1;
        }
        out[blockIdx.x] = c;
    }
}

// left rows of every split node: side-mask bits + their class counts
__global__ void __launch_bounds__(256) class_mark_kernel(const u32 *__restrict__ orders, int64_t col_stride,
                                                         const int *__restrict__ marks, int C, int row_bits,
                                                         u32 *mask, int *lcount) {
    extern __shared__ int sh[];
    const int m = blockIdx.y;   // marks: [f_local, start, k, slot of the split node in lcount]
    const int f = marks[4 * m], start = marks[4 * m + 1], k = marks[4 * m + 2], slot = marks[4 * m + 3];
    for (int c = threadIdx.x; c < C; c += 256) sh[c] = 0;
    __syncthreads();
    const u32 *col = orders + (int64_t)f * col_stride + start;
    const u32 rmask = (1u << row_bits) - 1u;
    for (int j = blockIdx.x * 256 + threadIdx.x; j <= k; j += gridDim.x * 256) {
        const u32 v = col[j];
        const u32 row = v & rmask;
        atomicOr(&mask[row >> 5], 1u << (row & 31));
        atomicAdd(&sh[v >> row_bits], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256)
        if (sh[c]) atomicAdd(&lcount[(int64_t)slot * C + c], sh[c]);
}

// numpy's pairwise float64 summation (np.sum of a contiguous 1-D array), for the host-side
// improvement test: surrogate_to_gini sums C squared class fractions (strategy.py:107-108)
double np_pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double r = -0.0;
        for (int64_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

template <typename T>
int upload_vec(b200dt_ctx *ctx, int pin_slot, size_t pin_off, const std::vector<T> &src, T *dst) {
    if (src.empty()) return 0;
    void *stage;
    CKR(pinned_get(ctx, pin_slot, pin_off + src.size() * sizeof(T), &stage));
    memcpy((char *)stage + pin_off, src.data(), src.size() * sizeof(T));
    CK(cudaMemcpyAsync(dst, (char *)stage + pin_off, src.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// advance the look-back epoch (shared with fit.cu's status buffers); on wrap-around clear them first
int class_next_epoch(b200dt_ctx *ctx) {
    if (ctx->epoch >= (1u << 30) - 2) {
        for (int id : {SB_STATUS_F, SB_STATUS_VAL, SB_STATUS_VAL2, SB_STATUS_P, SB_L1_H})
            if (ctx->scratch[id].p) CK(cudaMemsetAsync(ctx->scratch[id].p, 0, ctx->scratch[id].cap, ctx->stream));
        ctx->epoch = 0;
    }
    ctx->epoch += 1;
    return 0;
}

// B200DT_CLASS_NOFILTER=1: evaluate every gini candidate exactly (no node-wide running bound)
bool class_nofilter_env() {
    const char *e = getenv("B200DT_CLASS_NOFILTER");
    return e && e[0] == '1';
}

int resident_grid(b200dt_ctx *ctx, const void *kern, size_t smem, int64_t work_warps) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * ctx->sm_count;
    const int64_t need = (work_warps + 7) / 8;
    if (grid > need) grid = need;
    return (int)std::max<int64_t>(grid, 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side: the level-synchronous classification session
// ---------------------------------------------------------------------------------------
struct CNode {
    int parent = -1, is_left = 1, depth = 0, start = 0, n = 0;
    int feature = -2, left = -1, right = -1;
    double thr = -2.0, value = 0.0;
    bool leaf = false;
};


struct ClassSession {
    int64_t N = 0, stride = 0, ldx = 0, min_sample_leaf = 1;
    int F = 0, F_global = 0, col_offset = 0, C = 0, criterion = B200DT_GINI, max_depth = 0;
    int row_bits = 1, row_mask = 1, cur = 0;
    double min_improvement = 0.0;
    bool three_pass = false;     // B200DT_CLASS_3PASS=1: histogram + scan + score passes even where one pass would do
    bool single_node = false;    // b200dt_class_split_node: the root is searched whatever its size
    const double *dX = nullptr;
    int *ord[2] = {nullptr, nullptr};
    u32 *d_mask = nullptr;       // single-rank fits: the session's own side mask
    std::vector<CNode> nodes;
    std::vector<long long> counts;  // counts[node * C + c]
    std::vector<int> live;
    // state of the current level
    std::vector<CSeg> segs;
    std::vector<TileDesc> tiles;
    std::vector<int> tot, marks;
    std::vector<CCand> win;         // winners of the live nodes (after decide)
    std::vector<int> split_idx;     // indices into live of the nodes that split
    size_t pin_off = 0;
    CSeg *d_segs = nullptr;

    bool leaf_by_size(const CNode &nd) const {
        // base.py:13-21 (rows <= 2 * min_sample_leaf; a one-hot block with a single distinct value: no rows,
        // or a single class column) and tree_builder.py:74 (depth limit)
        return (int64_t)nd.n <= 2 * min_sample_leaf || nd.n == 0 || C == 1 || nd.depth >= max_depth;
    }
    double first_argmax(int node) const {  // tree_builder.py:20-21: np.argmax of the class counts
        int best = 0;
        for (int c = 1; c < C; ++c)
            if (counts[(size_t)node * C + c] > counts[(size_t)node * C + best]) best = c;
        return (double)best;
    }
};

void class_session_free(b200dt_ctx *ctx) {
    delete ctx->class_session;
    ctx->class_session = nullptr;
}

static int class_begin(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F, int64_t ldx,
                       int space, int C, int criterion, int max_depth, double min_improvement,
                       int64_t min_sample_leaf, int64_t col_offset, int64_t F_global, bool single_node) {
    if (N <= 0 || F <= 0) return ctx->fail_msg("fit_classes: empty input");
    if (criterion != B200DT_GINI && criterion != B200DT_P_AT_1) return ctx->fail_msg("fit_classes: unknown criterion");
    if (C < 1 || C > 32 * MAX_CLASS_CHUNKS) return ctx->fail_msg("fit_classes: n_classes must be in [1, 256]");
    if (N > (1ll << 26)) return ctx->fail_msg("fit_classes: N must be <= 2^26 (exact integer class statistics)");
    if (F > 65535) return ctx->fail_msg("fit_classes: at most 65535 columns per GPU (shard the columns over more ranks)");
    if (max_depth < 0 || max_depth > 30) return ctx->fail_msg("fit_classes: max_depth out of range [0, 30]");
    if (col_offset < 0 || col_offset + F > F_global) return ctx->fail_msg("fit_classes: column slice out of range");
    int row_bits = 1;
    while ((1ll << row_bits) < N) ++row_bits;
    if ((int64_t)C > (1ll << (32 - row_bits)))
        return ctx->fail_msg("fit_classes: n_classes * N does not fit the packed 32-bit (class, row) entries");
    end_all_sessions(ctx);
    ClassSession *s = new (std::nothrow) ClassSession();
    if (!s) return ctx->fail_msg("fit_classes: out of host memory");
    ctx->class_session = s;
    s->N = N;
    s->F = (int)F;
    s->F_global = (int)F_global;
    s->col_offset = (int)col_offset;
    s->C = C;
    s->criterion = criterion;
    s->max_depth = max_depth;
    s->min_improvement = min_improvement;
    s->min_sample_leaf = min_sample_leaf;
    s->row_bits = row_bits;
    s->row_mask = (int)((1u << row_bits) - 1u);
    s->stride = (N + 31) / 32 * 32;
    s->single_node = single_node;
    const char *env3 = getenv("B200DT_CLASS_3PASS");
    s->three_pass = env3 && env3[0] == '1';

    // inputs on the device
    const double *dX = X;
    const int *dlab = labels;
    int64_t dld = ldx;
    if (space == B200DT_HOST) {
        double *bx;
        int *bl;
        CKR(scratch_get(ctx, SB_X, (size_t)N * F * 8, (void **)&bx));
        CKR(scratch_get(ctx, SB_Y, (size_t)N * 8, (void **)&bl));
        CK(cudaMemcpy2DAsync(bx, (size_t)F * 8, X, (size_t)ldx * 8, (size_t)F * 8, (size_t)N, cudaMemcpyHostToDevice,
                             ctx->stream));
        CK(cudaMemcpyAsync(bl, labels, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
        dX = bx;
        dlab = bl;
        dld = F;
    }
    s->dX = dX;
    s->ldx = dld;
    // root class totals (+ label range check)
    unsigned long long *d_roothist;
    CKR(scratch_get(ctx, SB_MISC, (size_t)C * 8 + 64, (void **)&d_roothist));
    CK(cudaMemsetAsync(d_roothist, 0, (size_t)C * 8 + 64, ctx->stream));
    {
        LaunchScope ls(ctx, KC_MISC, 4.0 * (double)N);
        label_check_hist_kernel<<<(unsigned)std::min<int64_t>(1024, (N + 255) / 256), 256, 0, ctx->stream>>>(
            dlab, N, C, d_roothist, (int *)(d_roothist + C));
    }
    CK(cudaGetLastError());
    std::vector<unsigned long long> h_root(C + 1);
    CK(cudaMemcpyAsync(h_root.data(), d_roothist, (size_t)(C + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((int)h_root[C] != 0) return ctx->fail_msg("fit_classes: labels must be integers in [0, n_classes)");

    // one sort per fit; the class id then joins the row id
    CKR(scratch_get(ctx, SB_ORD_A, (size_t)F * s->stride * 4, (void **)&s->ord[0]));
    CKR(scratch_get(ctx, SB_ORD_B, (size_t)F * s->stride * 4, (void **)&s->ord[1]));
    CKR(sort_columns_device(ctx, dX, N, F, dld, s->ord[0], s->ord[1], s->stride));
    {
        LaunchScope ls(ctx, KC_PAYLOAD, 12.0 * (double)N * F);
        pack_labels_kernel<<<dim3((unsigned)std::min<int64_t>(2048, (N + 255) / 256), (unsigned)F), 256, 0, ctx->stream>>>(
            (u32 *)s->ord[0], s->stride, (int)F, N, dlab, row_bits);
    }
    CK(cudaGetLastError());
    s->cur = 0;
    CKR(scratch_get(ctx, SB_MASK, (size_t)((N + 31) / 32) * 4 + 64, (void **)&s->d_mask));

    s->nodes.assign(1, CNode());
    s->counts.assign(C, 0);
    for (int c = 0; c < C; ++c) s->counts[c] = (long long)h_root[c];
    s->nodes[0].n = (int)N;
    s->live.clear();
    if (!single_node && s->leaf_by_size(s->nodes[0])) {
        s->nodes[0].leaf = true;
        s->nodes[0].value = s->first_argmax(0);
    } else {
        s->live.push_back(0);
    }
    return 0;
}

// Search every live node over this rank's columns; one CCand per live node into d_out (device).
static int class_search(b200dt_ctx *ctx, CCand *d_out) {
    ClassSession *s = ctx->class_session;
    if (!s) return ctx->fail_msg("class session: not begun");
    const int n_live = (int)s->live.size();
    if (n_live == 0) return 0;
    const int C = s->C, F = s->F, row_bits = s->row_bits;
    const int criterion = s->criterion;
    s->segs.resize(n_live);
    s->tiles.clear();
    s->tot.resize((size_t)n_live * C);
    int64_t elems = 0;
    for (int i = 0; i < n_live; ++i) {
        const CNode &nd = s->nodes[s->live[i]];
        if (nd.n < 2)
            return ctx->fail_msg("fit_classes: attempt to split a node with a single row (min_sample_leaf < 1); "
                                 "the reference raises ValueError here (argmax of an empty sequence)");
        CSeg sg;
        sg.start = nd.start;
        sg.n = nd.n;
        sg.tile_base = (int)s->tiles.size();
        sg.ntiles = (nd.n + GT_TILE - 1) / GT_TILE;
        sg.st2 = 0;
        sg.tmax = 0;
        sg.pad = 0;
        for (int c = 0; c < C; ++c) {
            const long long tc = s->counts[(size_t)s->live[i] * C + c];
            s->tot[(size_t)i * C + c] = (int)tc;
            sg.st2 += tc * tc;
            sg.tmax = std::max(sg.tmax, (int)tc);
        }
        s->segs[i] = sg;
        for (int t = 0; t < sg.ntiles; ++t) s->tiles.push_back(TileDesc{i, t});
        elems += nd.n;
    }
    int n_tiles = (int)s->tiles.size();
    CSeg *d_segs;
    TileDesc *d_tiles;
    int *d_tot, *d_hist = nullptr, *d_excl = nullptr, *d_rmax = nullptr;
    CPart *d_part;
    const bool lane_path = C <= LANE_MAX_CLASSES;
    const bool single_pass = lane_path && !s->three_pass;
    const size_t hist_elems = (size_t)n_tiles * F * C;
    CKR(scratch_get(ctx, SB_SEG, (size_t)n_live * sizeof(CSeg) + 64, (void **)&d_segs));
    CKR(scratch_get(ctx, SB_TILES, (size_t)n_tiles * sizeof(TileDesc) + 64, (void **)&d_tiles));
    CKR(scratch_get(ctx, SB_TMP2, (size_t)n_live * C * 4 + 64, (void **)&d_tot));
    if (!single_pass) {
        CKR(scratch_get(ctx, SB_L1_A, hist_elems * 4 + 64, (void **)&d_hist));
        CKR(scratch_get(ctx, SB_L1_B, hist_elems * 4 + 64, (void **)&d_excl));
        CKR(scratch_get(ctx, SB_L1_C, (size_t)n_tiles * F * 4 + 64, (void **)&d_rmax));
    }
    CKR(scratch_get(ctx, SB_PARTIAL, (size_t)n_tiles * F * sizeof(CPart) + 64, (void **)&d_part));
    s->d_segs = d_segs;
    size_t po = 0;
    CKR(upload_vec(ctx, PIN_CLASS, po, s->segs, d_segs));
    po += s->segs.size() * sizeof(CSeg) + 64;
    CKR(upload_vec(ctx, PIN_CLASS, po, s->tiles, d_tiles));
    po += s->tiles.size() * sizeof(TileDesc) + 64;
    CKR(upload_vec(ctx, PIN_CLASS, po, s->tot, d_tot));
    po += s->tot.size() * 4 + 64;
    s->pin_off = po;

    const u32 *d_ord = (const u32 *)s->ord[s->cur];
    int64_t stride = s->stride;
    const int64_t work = (int64_t)n_tiles * F;
    if (!single_pass) {
        {
            const size_t smem = (size_t)8 * (C + 1) * 4;
            const int grid = resident_grid(ctx, (const void *)class_hist_kernel, smem, work);
            LaunchScope ls(ctx, KC_CLASS_HIST, 4.0 * (double)elems * F);
            class_hist_kernel<<<grid, 256, smem, ctx->stream>>>(d_ord, stride, d_segs, d_tiles, n_tiles, F, C, row_bits,
                                                                d_hist);
        }
        CK(cudaGetLastError());
        {
            const int64_t warps = (int64_t)n_live * F;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 7) / 8, (int64_t)ctx->sm_count * 16));
            LaunchScope ls(ctx, KC_MISC, 8.0 * (double)hist_elems);
            class_scan_kernel<<<grid, 256, 0, ctx->stream>>>(d_segs, n_live, F, C, n_tiles, d_tot, d_hist, d_excl,
                                                             criterion == B200DT_P_AT_1 ? d_rmax : nullptr);
        }
        CK(cudaGetLastError());
    }
    if (lane_path) {
        const bool patk = criterion == B200DT_P_AT_1;
        const size_t smem = (size_t)8 * ((size_t)(patk ? 2 : 1) * 32 * STAGE_PITCH * 4 + (size_t)C * 32 * 8);
        const void *kern = patk ? (single_pass ? (const void *)class_score_lane_kernel<true, 1>
                                               : (const void *)class_score_lane_kernel<false, 1>)
                                : (single_pass ? (const void *)class_score_lane_kernel<true, 0>
                                               : (const void *)class_score_lane_kernel<false, 0>);
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = resident_grid(ctx, kern, smem, work);
        u64 *d_status = nullptr;
        u32 epoch = 0;
        if (single_pass) {
            CKR(scratch_get(ctx, SB_STATUS_F, ((size_t)work * C + 1) * 8, (void **)&d_status, true));
            CKR(class_next_epoch(ctx));
            epoch = ctx->epoch;
        }
        const CSeg *a_segs = d_segs;
        const TileDesc *a_tiles = d_tiles;
        const int *a_excl = d_excl, *a_tot = d_tot;
        int a_F = F, a_C = C, a_bits = row_bits;
        unsigned long long *d_node_best = nullptr;
        if (!patk && !class_nofilter_env()) {
            CKR(scratch_get(ctx, SB_L1_E, (size_t)n_live * 8 + 64, (void **)&d_node_best));
            CK(cudaMemsetAsync(d_node_best, 0, (size_t)n_live * 8, ctx->stream));   // +0.0: below every score
        }
        void *args[] = {&d_ord, &stride, &a_segs, &a_tiles, &n_tiles, &a_F, &a_C, &a_bits, &a_excl, &a_tot, &d_status,
                        &epoch, &d_node_best, &d_part};
        LaunchScope ls(ctx, KC_CLASS_SCORE, 4.0 * (double)elems * F);
        if (single_pass)
            CK(launch_resident(kern, (unsigned)grid, 256, args, smem, ctx->stream));   // tiles wait on earlier tiles
        else
            CK(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(256), args, smem, ctx->stream));
    } else {
        const bool patk = criterion == B200DT_P_AT_1;
        const size_t smem = (size_t)8 * (2 * (C + 1) + (patk ? 2 * GT_TILE : 0)) * 4;
        const void *kern = patk ? (const void *)class_score_kernel<1> : (const void *)class_score_kernel<0>;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = resident_grid(ctx, kern, smem, work);
        LaunchScope ls(ctx, KC_CLASS_SCORE, 4.0 * (double)elems * F);
        if (patk)
            class_score_kernel<1><<<grid, 256, smem, ctx->stream>>>(d_ord, stride, d_segs, d_tiles, n_tiles, F, C, row_bits,
                                                                    d_excl, d_tot, d_rmax, d_part);
        else
            class_score_kernel<0><<<grid, 256, smem, ctx->stream>>>(d_ord, stride, d_segs, d_tiles, n_tiles, F, C, row_bits,
                                                                    d_excl, d_tot, d_rmax, d_part);
    }
    CK(cudaGetLastError());
    {
        CBest *d_sliced;
        CKR(scratch_get(ctx, SB_L1_G, (size_t)n_live * CF_SLICES * sizeof(CBest) + 64, (void **)&d_sliced));
        LaunchScope ls(ctx, KC_FINALIZE, (double)work * sizeof(CPart), 2);
        class_finalize_slices_kernel<<<dim3(n_live, CF_SLICES), 256, 0, ctx->stream>>>(d_segs, F, n_tiles, d_part, d_sliced);
        class_finalize_kernel<<<n_live, 256, 0, ctx->stream>>>(d_segs, d_sliced, d_ord, stride, s->row_mask, s->dX, s->ldx,
                                                               s->col_offset, d_out);
    }
    CK(cudaGetLastError());
    return 0;
}

// Pick the winner of every live node among the ranks' candidates ([rank][node] records), apply the leaf
// test, and -- for the winning columns this rank owns -- set the left rows' bits in d_mask and count
// their classes into d_lcount ([split node][class]; zero where another rank owns the column).
static int class_decide(b200dt_ctx *ctx, const CCand *d_gathered, int n_ranks, u32 *d_mask, int *d_lcount,
                        int64_t *n_split_out) {
    ClassSession *s = ctx->class_session;
    if (!s) return ctx->fail_msg("class session: not begun");
    *n_split_out = 0;
    const int n_live = (int)s->live.size();
    if (n_live == 0) return 0;
    const int C = s->C;
    CCand *h;
    CKR(pinned_get(ctx, PIN_RESULT, (size_t)n_ranks * n_live * sizeof(CCand), (void **)&h));
    CK(cudaMemcpyAsync(h, d_gathered, (size_t)n_ranks * n_live * sizeof(CCand), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s->win.resize(n_live);
    s->split_idx.clear();
    s->marks.clear();
    std::vector<double> frac(C);
    int max_left = 0, n_owned = 0;
    for (int i = 0; i < n_live; ++i) {
        // flat argmax over (position, GLOBAL column): strategy.py:277-278
        CCand w = h[i];
        for (int r = 1; r < n_ranks; ++r) {
            const CCand &c = h[(size_t)r * n_live + i];
            if (c.k < 0) continue;
            if (w.k < 0 || c.score > w.score || (c.score == w.score && (c.k < w.k || (c.k == w.k && c.f < w.f)))) w = c;
        }
        if (w.k < 0) return ctx->fail_msg("fit_classes: a node produced no split candidate");
        s->win[i] = w;
        const int nid = s->live[i];
        const double nn = (double)s->nodes[nid].n;
        double improvement;
        if (s->criterion == B200DT_GINI) {
            for (int c = 0; c < C; ++c) {
                const double fr = (double)s->counts[(size_t)nid * C + c] / nn;
                frac[c] = fr * fr;
            }
            improvement = w.score / nn - np_pairwise_sum(frac.data(), C);  // strategy.py:107-108
        } else {
            improvement = (w.score - (double)s->segs[i].tmax) / nn;        // strategy.py:103-104
        }
        if (improvement < s->min_improvement) {   // tree_builder.py:83-84
            s->nodes[nid].leaf = true;
            s->nodes[nid].value = s->first_argmax(nid);
            continue;
        }
        s->nodes[nid].feature = w.f;
        s->nodes[nid].thr = w.threshold;
        const int slot = (int)s->split_idx.size();
        s->split_idx.push_back(i);
        const int fl = w.f - s->col_offset;
        if (fl >= 0 && fl < s->F) {   // this rank owns the winning column
            s->marks.push_back(fl);
            s->marks.push_back(s->nodes[nid].start);
            s->marks.push_back(w.k_ext);
            s->marks.push_back(slot);
            max_left = std::max(max_left, w.k_ext + 1);
            ++n_owned;
        }
    }
    const int n_split = (int)s->split_idx.size();
    *n_split_out = n_split;
    if (n_split == 0) {
        s->live.clear();
        return 0;
    }
    CK(cudaMemsetAsync(d_mask, 0, (size_t)((s->N + 31) / 32) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_lcount, 0, (size_t)n_split * C * 4, ctx->stream));
    if (n_owned > 0) {
        int *d_marks;
        CKR(scratch_get(ctx, SB_TMP_I64, (size_t)n_owned * 16 + 64, (void **)&d_marks));
        CKR(upload_vec(ctx, PIN_CLASS, s->pin_off, s->marks, d_marks));
        s->pin_off += s->marks.size() * 4 + 64;
        const int gx = std::max(1, std::min(2048, (max_left + 255) / 256));
        for (int m0 = 0; m0 < n_owned; m0 += 65535) {
            const int nm = std::min(65535, n_owned - m0);
            LaunchScope ls(ctx, KC_MARK, 0.0);
            class_mark_kernel<<<dim3(gx, nm), 256, (size_t)C * 4, ctx->stream>>>(
                (const u32 *)s->ord[s->cur], s->stride, d_marks + 4 * m0, C, s->row_bits, d_mask, d_lcount);
        }
        CK(cudaGetLastError());
    }
    return 0;
}

// Children of the split nodes from the (all-reduced) left class counts, then the stable partition.
static int class_partition(b200dt_ctx *ctx, const u32 *d_mask, const int *d_lcount) {
    ClassSession *s = ctx->class_session;
    if (!s) return ctx->fail_msg("class session: not begun");
    const int n_split = (int)s->split_idx.size();
    if (n_split == 0) return 0;
    const int C = s->C;
    int *h_lcount;
    CKR(pinned_get(ctx, PIN_MISC, (size_t)n_split * C * 4, (void **)&h_lcount));
    CK(cudaMemcpyAsync(h_lcount, d_lcount, (size_t)n_split * C * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int> next_live;
    std::vector<SegDev> psegs;
    std::vector<TileDesc> ptiles;
    int64_t pelems = 0;
    for (int j = 0; j < n_split; ++j) {
        const int i = s->split_idx[j], nid = s->live[i];
        const int start = s->nodes[nid].start, n = s->nodes[nid].n, nl = s->win[i].k_ext + 1;
        const int depth = s->nodes[nid].depth;
        long long lsum = 0;
        for (int c = 0; c < C; ++c) lsum += h_lcount[(size_t)j * C + c];
        if (lsum != nl) return ctx->fail_msg("fit_classes: left class counts do not add up (mask / count exchange)");
        int child[2];
        for (int side = 0; side < 2; ++side) {
            CNode ch;
            ch.parent = nid;
            ch.is_left = side == 0;
            ch.depth = depth + 1;
            ch.start = side == 0 ? start : start + nl;
            ch.n = side == 0 ? nl : n - nl;
            s->nodes.push_back(ch);
            child[side] = (int)s->nodes.size() - 1;
            s->counts.resize(s->nodes.size() * (size_t)C);
            for (int c = 0; c < C; ++c) {
                const long long lc = h_lcount[(size_t)j * C + c];
                s->counts[(size_t)child[side] * C + c] = side == 0 ? lc : s->counts[(size_t)nid * C + c] - lc;
            }
        }
        s->nodes[nid].left = child[0];
        s->nodes[nid].right = child[1];
        bool any_search = false;
        for (int side = 0; side < 2; ++side) {
            CNode &ch = s->nodes[child[side]];
            if (s->leaf_by_size(ch)) {
                ch.leaf = true;
                ch.value = s->first_argmax(child[side]);
            } else {
                next_live.push_back(child[side]);
                any_search = true;
            }
        }
        if (any_search) {
            SegDev sg;
            sg.total = sg.total_w = 0.0;
            sg.start = start;
            sg.n = n;
            sg.pbase = 0;
            sg.np = (n + PT_TILE - 1) / PT_TILE;
            sg.nleft = nl;
            sg.exact = 0;
            sg.pstep = 1;
            sg.pside = 0;
            sg.vstart = -1;
            sg.cbase = 0;
            const int si = (int)psegs.size();
            psegs.push_back(sg);
            for (int t = 0; t < sg.np; ++t) ptiles.push_back(TileDesc{si, t});
            pelems += n;
        }
    }
    if (!psegs.empty()) {
        SegDev *d_psegs;
        TileDesc *d_ptiles;
        u64 *pstatus;
        CKR(scratch_get(ctx, SB_PLAN, psegs.size() * sizeof(SegDev) + 64, (void **)&d_psegs));
        CKR(scratch_get(ctx, SB_TILES, ptiles.size() * sizeof(TileDesc) + 64, (void **)&d_ptiles));
        CKR(upload_vec(ctx, PIN_CLASS, s->pin_off, psegs, d_psegs));
        s->pin_off += psegs.size() * sizeof(SegDev) + 64;
        CKR(upload_vec(ctx, PIN_CLASS, s->pin_off, ptiles, d_ptiles));
        CKR(scratch_get(ctx, SB_STATUS_P, (ptiles.size() * (size_t)s->F + 1) * 8, (void **)&pstatus, true));
        CKR(class_next_epoch(ctx));
        CKR(launch_partition(ctx, s->ord[s->cur], s->ord[s->cur ^ 1], s->stride, s->F, d_psegs, d_ptiles,
                             (int)ptiles.size(), pelems, d_mask, pstatus, ctx->epoch, nullptr, nullptr, nullptr, nullptr,
                             nullptr, s->row_mask));
        s->cur ^= 1;
        // the pinned staging area is rewritten by the next level: let the copies finish first
        CK(cudaStreamSynchronize(ctx->stream));
    }
    s->live.swap(next_live);
    s->split_idx.clear();
    return 0;
}

static int class_finish(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right, int64_t *feature,
                        double *threshold, double *value, int32_t *n_node_samples, int64_t capacity,
                        int64_t *node_count) {
    ClassSession *s = ctx->class_session;
    if (!s) return ctx->fail_msg("class session: not begun");
    if (!s->live.empty()) return ctx->fail_msg("class session: finish called before the tree is complete");
    // pre-order ids, left first (base.py:47-52; tree_builder.py:91-94 pushes right, then left)
    const int64_t count = (int64_t)s->nodes.size();
    if (count > capacity) return ctx->fail_msg("fit_classes: tree arrays too small");
    std::vector<int> new_id(count, -1), stack(1, 0);
    int next = 0;
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        new_id[v] = next++;
        if (!s->nodes[v].leaf) {
            stack.push_back(s->nodes[v].right);
            stack.push_back(s->nodes[v].left);
        }
    }
    for (int64_t v = 0; v < count; ++v) {
        const CNode &nd = s->nodes[v];
        const int id = new_id[v];
        if (nd.leaf) {
            children_left[id] = -1;
            children_right[id] = -1;
            feature[id] = -2;
            threshold[id] = -2.0;
            value[id] = nd.value;
        } else {
            children_left[id] = new_id[nd.left];
            children_right[id] = new_id[nd.right];
            feature[id] = nd.feature;
            threshold[id] = nd.thr;
            value[id] = 0.0;   // the reference leaves internal values uninitialised (base.py:36)
        }
        if (n_node_samples) n_node_samples[id] = nd.n;
    }
    *node_count = count;
    class_session_free(ctx);
    return 0;
}

static_assert(sizeof(CCand) == sizeof(b200dt_class_candidate), "candidate layout");

extern "C" {

int b200dt_class_session_begin(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F_local,
                               int64_t ldx, int space, int n_classes, int criterion, int max_depth,
                               double min_improvement, int64_t min_sample_leaf, int64_t col_offset, int64_t F_global) {
    if (!ctx) return 1;
    if (!X || !labels) return ctx->fail_msg("class_session_begin: null argument");
    CK(cudaSetDevice(ctx->device));
    return class_begin(ctx, X, labels, N, F_local, ldx, space, n_classes, criterion, max_depth, min_improvement,
                       min_sample_leaf, col_offset, F_global, false);
}

int b200dt_class_session_live(b200dt_ctx *ctx, int64_t *n_live) {
    if (!ctx) return 1;
    if (!ctx->class_session) return ctx->fail_msg("class session: not begun");
    *n_live = (int64_t)ctx->class_session->live.size();
    return 0;
}

int b200dt_class_session_search(b200dt_ctx *ctx, b200dt_class_candidate *d_out) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    return class_search(ctx, reinterpret_cast<CCand *>(d_out));
}

int b200dt_class_session_decide(b200dt_ctx *ctx, const b200dt_class_candidate *d_gathered, int n_ranks,
                                uint32_t *d_side_mask, int32_t *d_left_counts, int64_t *n_split) {
    if (!ctx) return 1;
    if (!d_gathered || !d_side_mask || !d_left_counts || !n_split || n_ranks < 1)
        return ctx->fail_msg("class_session_decide: bad argument");
    CK(cudaSetDevice(ctx->device));
    return class_decide(ctx, reinterpret_cast<const CCand *>(d_gathered), n_ranks, d_side_mask, d_left_counts, n_split);
}

int b200dt_class_session_partition(b200dt_ctx *ctx, const uint32_t *d_side_mask, const int32_t *d_left_counts) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    return class_partition(ctx, d_side_mask, d_left_counts);
}

int b200dt_class_session_finish(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right, int64_t *feature,
                                double *threshold, double *value, int32_t *n_node_samples, int64_t capacity,
                                int64_t *node_count) {
    if (!ctx) return 1;
    if (!children_left || !children_right || !feature || !threshold || !value || !node_count)
        return ctx->fail_msg("class_session_finish: null argument");
    return class_finish(ctx, children_left, children_right, feature, threshold, value, n_node_samples, capacity,
                        node_count);
}

int b200dt_fit_classes(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F, int64_t ldx,
                       int space, int n_classes, int criterion, int max_depth, double min_improvement,
                       int64_t min_sample_leaf, int64_t *children_left, int64_t *children_right, int64_t *feature,
                       double *threshold, double *value, int32_t *n_node_samples, int64_t capacity,
                       int64_t *node_count) {
    if (!ctx) return 1;
    if (!X || !labels || !children_left || !children_right || !feature || !threshold || !value || !node_count)
        return ctx->fail_msg("fit_classes: null argument");
    CK(cudaSetDevice(ctx->device));
    int r = class_begin(ctx, X, labels, N, F, ldx, space, n_classes, criterion, max_depth, min_improvement,
                        min_sample_leaf, 0, F, false);
    while (r == 0 && !ctx->class_session->live.empty()) {
        ClassSession *s = ctx->class_session;
        const int n_live = (int)s->live.size();
        CCand *d_cand;
        int *d_lcount;
        int64_t n_split = 0;
        r = scratch_get(ctx, SB_RESULT, (size_t)n_live * sizeof(CCand) + 64, (void **)&d_cand);
        if (r == 0) r = scratch_get(ctx, SB_L1_D, (size_t)n_live * n_classes * 4 + 64, (void **)&d_lcount);
        if (r == 0) r = class_search(ctx, d_cand);
        if (r == 0) r = class_decide(ctx, d_cand, 1, s->d_mask, d_lcount, &n_split);
        if (r == 0) r = class_partition(ctx, s->d_mask, d_lcount);
    }
    if (r == 0)
        r = class_finish(ctx, children_left, children_right, feature, threshold, value, n_node_samples, capacity,
                         node_count);
    class_session_free(ctx);
    return r;
}

int b200dt_class_split_node(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t n, int64_t F, int64_t ldx,
                            int space, int n_classes, int criterion, int64_t *out_f, int64_t *out_k,
                            double *out_threshold, double *out_score) {
    if (!ctx) return 1;
    if (!X || !labels || !out_f || !out_k || !out_threshold || !out_score)
        return ctx->fail_msg("class_split_node: null argument");
    if (n < 2) return ctx->fail_msg("class_split_node: attempt to get argmax of an empty sequence (n < 2)");
    CK(cudaSetDevice(ctx->device));
    int r = class_begin(ctx, X, labels, n, F, ldx, space, n_classes, criterion, 1, 0.0, 0, 0, F, true);
    CCand c;
    if (r == 0) {
        CCand *d_cand;
        r = scratch_get(ctx, SB_RESULT, sizeof(CCand) + 64, (void **)&d_cand);
        if (r == 0) r = class_search(ctx, d_cand);
        if (r == 0) {
            CK(cudaMemcpyAsync(&c, d_cand, sizeof(CCand), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    class_session_free(ctx);
    if (r) return r;
    if (c.k < 0) return ctx->fail_msg("class_split_node: the node cannot be searched");
    *out_f = c.f;
    *out_k = c.k;
    *out_threshold = c.threshold;
    *out_score = c.score;
    return 0;
}

}  // extern "C"
