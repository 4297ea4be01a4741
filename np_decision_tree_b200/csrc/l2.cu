// K2: L2 (MSE) exhaustive split search for every live node of a level, and the per-node
// arg-best reduction.
//
// replaces greedy_regression_split + best_variance_improvements[_v2] + np_gene
//          (ref decision_tree/one.py:52-63, 81-93, 96-107)
//
// Two kernels share the work:
//  * l2_scan_kernel  -- nodes with more than EXACT_ROWS rows.  One WARP per (tile, column):
//    coalesced read of 512 row ids, gather of y, warp scan, carry-in from a decoupled
//    look-back over the earlier tiles of the same (node, column), gain at every position,
//    warp arg-max.  No block barriers; persistent grid, static tile order.  Single pass: 12 algorithmic bytes per (row, feature) element.  The
//    parallel scan rounds differently from the reference's sequential np.cumsum, so each
//    candidate also carries the runner-up gain; a node whose two best gains are within
//    1e-9 relative is re-searched by the exact kernel.
//  * l2_exact_kernel -- small nodes (and flagged near-ties).  One warp per (node, column)
//    adds y in sorted order ONE ELEMENT AT A TIME in fp64 (shuffle broadcast), with the
//    reference's exact operation order and no FMA contraction, so bit-exact gain ties
//    between features resolve exactly as in the reference (hard part 1).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "l2_score.cuh"

namespace {


template <int VARIANT>
__global__ void __launch_bounds__(256, 3) l2_scan_kernel(LevelArgs a, const TileDesc *__restrict__ tiles,
                                                         int n_tiles, ScanStatus st, u32 epoch,
                                                         Partial *__restrict__ partials) {
    constexpr int variant = VARIANT;
    __shared__ double sy_all[8][SY_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sy = sy_all[warp];
    const int64_t total_tiles = (int64_t)n_tiles * a.F_search;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % a.F_search), ti = (int)(t / a.F_search);
        const TileDesc td = tiles[ti];
        const SegDev sg = a.segs[td.seg];
        const int off = td.tin * WT_TILE;
        const int cnt = min(WT_TILE, sg.n - off);
        const int *col = a.orders + (int64_t)f * a.col_stride + sg.start + off;
        if (a.ypay) {
            // streaming path: y travels with the row ids, no gather
            const double *yp = a.ypay + (int64_t)f * a.col_stride + sg.start + off;
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const int e = i * 32 + lane;
                sy[i * 34 + lane + (lane >> 4)] = e < cnt ? __ldcs(yp + e) : 0.0;
            }
        } else if (cnt == WT_TILE) {
            // coalesced read of the row ids, gather of y, staged so each lane gets 16 consecutive values
            int r[WT_ITEMS];
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) r[i] = ld_stream_i32(col + i * 32 + lane);
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) sy[i * 34 + lane + (lane >> 4)] = __ldg(a.y + r[i]);
        } else {
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const int e = i * 32 + lane;
                sy[i * 34 + lane + (lane >> 4)] = e < cnt ? __ldg(a.y + ld_stream_i32(col + e)) : 0.0;
            }
        }
        __syncwarp();
        double p[WT_ITEMS];
        {
            const double *src = sy + lane * 17;  // element lane*16+i sits at lane*16+i + lane
            double run = 0.0;
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                run += src[i];
                p[i] = run;
            }
        }
        __syncwarp();  // sy may be overwritten by the next tile
        double incl = p[WT_ITEMS - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = shfl_up_d(incl, d);
            if (lane >= d) incl += o;
        }
        double texcl = shfl_up_d(incl, 1);
        if (lane == 0) texcl = 0.0;
        const double tile_total = shfl_d(incl, 31);
        const double excl = lookback_f64(st, (int64_t)f * n_tiles + sg.cbase + td.tin, td.tin, epoch, tile_total, lane);
        const double offset = excl + texcl;

        // score of every position of this lane; best and runner-up tracked branch-free
        const int n = sg.n;
        const double mean = sg.total * rcp_fast1((double)n);   // approximate search: 1e-12 is plenty
        const int k0 = off + lane * WT_ITEMS;
        double m1 = -CUDART_INF, m2 = -CUDART_INF, sbest = 0.0, dbest = 0.0;
        int ibest = -1;
        if (variant == 1 && n <= (1 << 24)) {
            // gain = dev^2 / ((k+1)(n-k-1)); the denominator is advanced exactly in fp64:
            // den(k+1) - den(k) = n - 2k - 3.  float32(k+1) is exact for n <= 2^24 (one.py:52-53)
            double km = (double)(k0 + 1) * mean;
            double den = (double)(k0 + 1) * (double)(n - 1 - k0);
            double dd = (double)(n - 2 * k0 - 3);
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const double S = offset + p[i];
                const double dev = S - km;
                double x;  // 1/den to ~1e-12: hardware seed + one Newton step (approximate search only)
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(den));
                x = fma(x, fma(-den, x, 1.0), x);
                double g = dev * dev * x;
                if (!(lane * WT_ITEMS + i < cnt && k0 + i < n - 1)) g = -CUDART_INF;
                m2 = fmax(m2, fmin(m1, g));
                if (g > m1) {
                    m1 = g;
                    sbest = S;
                    dbest = dev;
                    ibest = i;
                }
                km += mean;
                den += dd;
                dd -= 2.0;
            }
            dbest = fabs(dbest);
        } else {
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const int kk = k0 + i;
                const double S = offset + p[i];
                double dev = 0.0;
                double g = -CUDART_INF;
                if (lane * WT_ITEMS + i < cnt && kk < n - 1) g = score_fast(S, kk, n, mean, variant, dev);
                m2 = fmax(m2, fmin(m1, g));
                if (g > m1) {
                    m1 = g;
                    sbest = S;
                    dbest = dev;
                    ibest = i;
                }
            }
        }
        Cand best;
        best.g = m1;
        best.g2 = m2;
        best.s = sbest;
        best.dev = dbest;
        best.k = ibest < 0 ? 0x7fffffff : k0 + ibest;
        cand_warp_reduce(best);
        if (lane == 0) {
            Partial out;
            out.score = best.g;
            out.aux = best.dev;
            out.lsum = best.s;
            out.score2 = best.g2;
            out.k = best.k == 0x7fffffff ? -1 : best.k;
            out.f = f;
            out.ppos = -1;
            out.pad_ = 0;
            partials[(int64_t)(sg.pbase + td.tin * sg.pstep + sg.pside) * a.F_search + f] = out;
        }
    }
}

// warp arg-best of per-lane (best gain, position) pairs: (gain desc, position asc) decides the winner lane, whose
// prefix sum / deviation are broadcast; the runner-up is the best of the lanes' runner-ups and of the losing lanes'
// bests.  29 shuffles instead of the 55 of a full-record butterfly.
__device__ __forceinline__ void warp_argbest(double m1, double m2, double &sbest, double &dbest, int kbest, double &wg,
                                             int &wk, double &r2) {
    wg = m1;
    wk = kbest;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const double og = shfl_xor_d(wg, m);
        const int ok = __shfl_xor_sync(FULL_MASK, wk, m);
        if (og > wg || (og == wg && ok < wk)) {
            wg = og;
            wk = ok;
        }
    }
    const bool winner = kbest == wk && wk != 0x7fffffff;
    const int wl = __ffs(__ballot_sync(FULL_MASK, winner)) - 1;
    r2 = winner ? m2 : fmax(m1, m2);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) r2 = fmax(r2, shfl_xor_d(r2, m));
    if (wl >= 0) {
        sbest = shfl_d(sbest, wl);
        dbest = shfl_d(dbest, wl);
    }
}

// Streaming variant (sessions that carry y as a payload next to the row ids): the tile is 1024
// positions of ypay, staged once in shared memory and read twice by the lane that owns 32
// consecutive positions -- first to sum them (tile aggregate -> look-back), then to rebuild the
// running prefix while scoring.  No per-lane prefix array, rolled loops, and the per-tile costs
// (look-back, warp arg-max, segment decode) are amortised over twice as many elements.
constexpr int ST_PITCH = ST_TILE + ST_TILE / 32;  // 1 pad double per 32: conflict-free blocked reads

template <int VARIANT>
__global__ void __launch_bounds__(256, 3) l2_scan_stream_kernel(LevelArgs a, const TileDesc *__restrict__ tiles,
                                                                int n_tiles, ScanStatus st, u32 epoch,
                                                                Partial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sy = reinterpret_cast<double *>(scan_smem) + warp * ST_PITCH;
    const double *mine = sy + lane * (ST_ITEMS + 1);   // position lane*32 + j sits at lane*33 + j
    const int64_t total_tiles = (int64_t)n_tiles * a.F_search;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    // the (tile, node) descriptors of a warp's NEXT tile are fetched while the current tile's y values travel: two
    // dependent look-ups per tile leave the critical path
    struct SegLite {
        double total;
        int start, n, pbase, pstep, pside, cbase;
    };
    auto seg_lite = [&](int seg) {
        const SegDev *p = a.segs + seg;
        SegLite l;
        l.total = p->total;
        l.start = p->start;
        l.n = p->n;
        l.pbase = p->pbase;
        l.pstep = p->pstep;
        l.pside = p->pside;
        l.cbase = p->cbase;
        return l;
    };
    TileDesc td_next{0, 0};
    SegLite sg_next{};
    {
        const int64_t t0 = (int64_t)blockIdx.x * 8 + warp;
        if (t0 < total_tiles) {
            td_next = tiles[(int)(t0 / a.F_search)];
            sg_next = seg_lite(td_next.seg);
        }
    }
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % a.F_search);
        const TileDesc td = td_next;
        const SegLite sg = sg_next;
        const int off = td.tin * ST_TILE;
        const int cnt = min(ST_TILE, sg.n - off);
        const double *yp = a.ypay + (int64_t)f * a.col_stride + sg.start + off;
        // best gain any warp has found in this node so far (bits of a non-negative double; 0 at level start)
        long long nb_seen = 0;
        if (VARIANT == 1) nb_seen = (long long)ld_relaxed_u64((const u64 *)(a.node_best + td.seg));
        const bool more = t + stride_tiles < total_tiles;
        if (more) td_next = tiles[(int)((t + stride_tiles) / a.F_search)];
        if (cnt == ST_TILE) {
#pragma unroll
            for (int i = 0; i < ST_ITEMS; ++i) sy[i * 33 + lane] = __ldcs(yp + i * 32 + lane);
        } else {
#pragma unroll 4
            for (int i = 0; i < ST_ITEMS; ++i) {
                const int e = i * 32 + lane;
                sy[i * 33 + lane] = e < cnt ? __ldcs(yp + e) : 0.0;
            }
        }
        if (more) sg_next = seg_lite(td_next.seg);
        __syncwarp();
        double ts0 = 0.0, ts1 = 0.0, ts2 = 0.0, ts3 = 0.0;   // four independent chains (fp64 add latency)
#pragma unroll
        for (int j = 0; j < ST_ITEMS; j += 4) {
            ts0 += mine[j];
            ts1 += mine[j + 1];
            ts2 += mine[j + 2];
            ts3 += mine[j + 3];
        }
        const double tsum = (ts0 + ts1) + (ts2 + ts3);
        double incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = shfl_up_d(incl, d);
            if (lane >= d) incl += o;
        }
        const double tile_total = shfl_d(incl, 31);
        const double excl = lookback_f64(st, (int64_t)f * n_tiles + sg.cbase + td.tin, td.tin, epoch, tile_total, lane);
        double S = excl + (incl - tsum);   // sum of everything before this lane's first position

        const int n = sg.n;
        const double mean = sg.total * rcp_fast1((double)n);
        const int k0 = off + lane * ST_ITEMS;
        // positions this lane may score: inside the tile and not the node's last row
        const int nvalid = max(0, min(ST_ITEMS, min(cnt - lane * ST_ITEMS, n - 1 - k0)));
        double m1 = -CUDART_INF, m2 = -CUDART_INF, sbest = 0.0, dbest = 0.0;
        int kbest = 0x7fffffff;
        if (VARIANT == 1 && n <= (1 << 24)) {
            // gain(k) = dev^2 / den, den = (k+1)(n-1-k) advanced exactly.  Almost no position can beat -- or come
            // within the near-tie margin of -- the best gain the NODE has produced so far (`node_best`, a running
            // maximum shared by all warps working on the node), so positions are first tested without the division:
            //      dev^2 > 0.9999999 * bound * den
            // and only the survivors (a vanishing fraction once the bound has risen) are scored in full.  The
            // position of the node's maximum and everything within 1e-7 of it always survive, whatever the timing,
            // so the winner, the tie rule and the near-tie flag (1e-9) do not depend on the bound.
            double km = (double)(k0 + 1) * mean;
            double den = (double)(k0 + 1) * (double)(n - 1 - k0);
            double dd = (double)(n - 2 * k0 - 3);
            double gate = 0.9999999 * __longlong_as_double(nb_seen);
#pragma unroll 4
            for (int j = 0; j < nvalid; ++j) {
                S += mine[j];
                const double dev = S - km;
                if (dev * dev > gate * den) {
                    const double g = dev * dev * rcp_fast1(den);
                    m2 = fmax(m2, fmin(m1, g));
                    if (g > m1) {
                        m1 = g;
                        sbest = S;
                        dbest = dev;
                        kbest = k0 + j;
                        gate = fmax(gate, 0.9999999 * g);
                    }
                }
                km += mean;
                den += dd;
                dd -= 2.0;
            }
            dbest = fabs(dbest);
        } else {
            for (int j = 0; j < nvalid; ++j) {
                S += mine[j];
                double dev = 0.0;
                const double g = score_fast(S, k0 + j, n, mean, VARIANT, dev);
                m2 = fmax(m2, fmin(m1, g));
                if (g > m1) {
                    m1 = g;
                    sbest = S;
                    dbest = dev;
                    kbest = k0 + j;
                }
            }
        }
        __syncwarp();  // sy is overwritten by the next tile
        double wg, r2;
        int wk;
        warp_argbest(m1, m2, sbest, dbest, kbest, wg, wk, r2);
        if (lane == 0) {
            Partial out;
            out.score = wg;
            out.aux = dbest;
            out.lsum = sbest;
            out.score2 = r2;
            out.k = wk == 0x7fffffff ? -1 : wk;
            DCHECK(out.k < sg.n - 1 && out.k >= -1);
            out.f = f;
            out.ppos = -1;
            out.pad_ = 0;
            partials[(int64_t)(sg.pbase + td.tin * sg.pstep + sg.pside) * a.F_search + f] = out;
            // raise the node's bound (non-negative doubles order like their bit patterns; NaN never gets here)
            if (VARIANT == 1 && wg > __longlong_as_double(nb_seen))
                atomicMax(a.node_best + td.seg, __double_as_longlong(wg));
        }
    }
}

// The same search with the tile brought into shared memory by BULK ASYNCHRONOUS COPIES (TMA, cp.async.bulk + mbarrier),
// double-buffered: while a warp works on tile i, the 8 KB of its tile i+1 are already in flight into its other buffer,
// and the copy engine writes shared memory directly (no LDG + STS pair per element; the register-staged kernel keeps the
// LSU pipe 65 % busy and stalls a quarter of its time on its loads, ncu r02).  Selected with B200DT_SCAN_BULK=1.
// MEASURED (10 M x 256, profiles/README.md): correct, LSU 65 -> 28 %, but 6.4 ms per level against 5.1 ms: two 8.7 KB
// buffers per warp leave 12 warps per SM, too few to cover the dependent fp64 prefix chain, and the per-lane copies
// (needed for a conflict-free padded layout) are serialised through the uniform datapath (ELECT / R2UR loop, 32 x
// UBLKCP per tile).  Kept as the measured alternative; the default is the register-staged kernel above.
// Every lane issues ONE 256-byte copy: its own 32 positions go to its own row of the buffer.  Rows are 34 doubles
// (272 B) apart: 16-byte aligned as the copy engine requires, and a quarter-warp's 128-bit reads of its rows hit 32
// different banks (lane l covers banks 4l .. 4l+3).  Global sources must be 16-byte aligned too, so the tile grid of
// a node whose first position is odd is shifted down by one element (`o`); that one foreign element is never read.
constexpr int SB_ROW = ST_ITEMS + 2;              // doubles per lane row
constexpr int SB_BUF = 32 * SB_ROW;               // doubles per buffer (8704 B)
constexpr int SCAN_WARPS = 4;                     // warps per block: 4 x 2 x 8704 B = 68 KB, 3 blocks per SM
constexpr size_t SCAN_SMEM = (size_t)SCAN_WARPS * 2 * SB_BUF * sizeof(double) + SCAN_WARPS * 2 * sizeof(u64);

template <int VARIANT>
__global__ void __launch_bounds__(SCAN_WARPS * 32, 3) l2_scan_bulk_kernel(LevelArgs a, const TileDesc *__restrict__ tiles,
                                                                          int n_tiles, ScanStatus st, u32 epoch,
                                                                          Partial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wbuf = reinterpret_cast<double *>(scan_smem) + (size_t)warp * 2 * SB_BUF;
    u64 *bars = reinterpret_cast<u64 *>(reinterpret_cast<double *>(scan_smem) + (size_t)SCAN_WARPS * 2 * SB_BUF) + warp * 2;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int64_t total_tiles = (int64_t)n_tiles * a.F_search;
    const int64_t stride_tiles = (int64_t)gridDim.x * SCAN_WARPS;
    const int64_t t0 = (int64_t)blockIdx.x * SCAN_WARPS + warp;

    // start the copies of tile t into buffer b: lane l copies the valid part of its own 32 positions
    auto issue = [&](int64_t t, int b) {
        const int f = (int)(t % a.F_search);
        const TileDesc td = tiles[t / a.F_search];
        const int start = a.segs[td.seg].start, n = a.segs[td.seg].n;
        const int o = start & 1;
        const int q0 = td.tin * ST_TILE - o + lane * ST_ITEMS;          // node position of this lane's element 0
        const int jhi = max(0, min(ST_ITEMS, n - q0));
        const u32 bytes = (u32)((jhi * 8 + 15) & ~15);
        const u32 total = __reduce_add_sync(FULL_MASK, bytes);
        if (lane == 0) mbar_arrive_expect_tx(&bars[b], total);
        __syncwarp();
        if (bytes)
            bulk_g2s(wbuf + (size_t)b * SB_BUF + lane * SB_ROW, a.ypay + (int64_t)f * a.col_stride + start + q0, bytes,
                     &bars[b]);
    };

    if (t0 < total_tiles) issue(t0, 0);
    int it = 0;
    for (int64_t t = t0; t < total_tiles; t += stride_tiles, ++it) {
        const int b = it & 1;
        if (t + stride_tiles < total_tiles) {
            fence_proxy_async_smem();      // buffer b^1 was read (generic proxy) in the previous iteration
            issue(t + stride_tiles, b ^ 1);
        }
        const int f = (int)(t % a.F_search), ti = (int)(t / a.F_search);
        const TileDesc td = tiles[ti];
        const SegDev sg = a.segs[td.seg];
        const int n = sg.n;
        const int o = sg.start & 1;
        const int q0 = td.tin * ST_TILE - o + lane * ST_ITEMS;
        const int jlo = max(0, -q0);                                  // 1 only for the shifted first element of a node
        const int jhi = max(jlo, min(ST_ITEMS, n - q0));              // positions of the node: [jlo, jhi)
        const int jsc = max(jlo, min(ST_ITEMS, n - 1 - q0));          // positions that can be split points: [jlo, jsc)
        // warp-uniform: every lane owns 32 scoreable positions (all tiles but a node's first / last)
        const bool full = td.tin * ST_TILE - o >= 0 && td.tin * ST_TILE - o + ST_TILE <= n - 1;
        // best gain any warp has found in this node so far (bits of a non-negative double; 0 at level start)
        long long nb_seen = 0;
        if (VARIANT == 1) nb_seen = (long long)ld_relaxed_u64((const u64 *)(a.node_best + td.seg));
        const double *mine = wbuf + (size_t)b * SB_BUF + lane * SB_ROW;
        const double2 *mine2 = reinterpret_cast<const double2 *>(mine);
        mbar_wait(&bars[b], (u32)(it >> 1) & 1u);

        double tsum = 0.0;
        if (full) {
#pragma unroll
            for (int i = 0; i < ST_ITEMS / 2; ++i) {
                const double2 v = mine2[i];
                tsum += v.x;
                tsum += v.y;
            }
        } else {
            for (int j = jlo; j < jhi; ++j) tsum += mine[j];
        }
        double incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double x = shfl_up_d(incl, d);
            if (lane >= d) incl += x;
        }
        const double tile_total = shfl_d(incl, 31);
        const double excl = lookback_f64(st, (int64_t)f * n_tiles + sg.cbase + td.tin, td.tin, epoch, tile_total, lane);
        double S = excl + (incl - tsum);   // sum of everything before this lane's first position

        const double mean = sg.total * rcp_fast1((double)n);
        double m1 = -CUDART_INF, m2 = -CUDART_INF, sbest = 0.0, dbest = 0.0;
        int kbest = 0x7fffffff;
        if (VARIANT == 1 && n <= (1 << 24)) {
            // gain(k) = dev^2 / den, den = (k+1)(n-1-k) advanced exactly.  Almost no position can beat -- or come
            // within the near-tie margin of -- the best gain the NODE has produced so far (`node_best`, a running
            // maximum shared by all warps working on the node), so positions are first tested without the division:
            //      dev^2 > 0.9999999 * bound * den
            // and only the survivors (a vanishing fraction once the bound has risen) are scored in full.  The
            // position of the node's maximum and everything within 1e-7 of it always survive, whatever the timing,
            // so the winner, the tie rule and the near-tie flag (1e-9) do not depend on the bound.
            const int k0 = q0 + jlo;
            double km = (double)(k0 + 1) * mean;
            double den = (double)(k0 + 1) * (double)(n - 1 - k0);
            double dd = (double)(n - 2 * k0 - 3);
            double gate = 0.9999999 * __longlong_as_double(nb_seen);
#define L2_SCORE_ONE(YV, KPOS)                                          \
    {                                                                   \
        S += (YV);                                                      \
        const double dev = S - km;                                      \
        if (dev * dev > gate * den) {                                   \
            const double g = dev * dev * rcp_fast1(den);                \
            m2 = fmax(m2, fmin(m1, g));                                 \
            if (g > m1) {                                               \
                m1 = g;                                                 \
                sbest = S;                                              \
                dbest = dev;                                            \
                kbest = (KPOS);                                         \
                gate = fmax(gate, 0.9999999 * g);                       \
            }                                                           \
        }                                                               \
        km += mean;                                                     \
        den += dd;                                                      \
        dd -= 2.0;                                                      \
    }
            if (full) {
#pragma unroll 4
                for (int i = 0; i < ST_ITEMS / 2; ++i) {
                    const double2 v = mine2[i];
                    L2_SCORE_ONE(v.x, q0 + 2 * i)
                    L2_SCORE_ONE(v.y, q0 + 2 * i + 1)
                }
            } else {
                for (int j = jlo; j < jsc; ++j) L2_SCORE_ONE(mine[j], q0 + j)
            }
#undef L2_SCORE_ONE
            dbest = fabs(dbest);
        } else {
            for (int j = jlo; j < jsc; ++j) {
                S += mine[j];
                double dev = 0.0;
                const double g = score_fast(S, q0 + j, n, mean, VARIANT, dev);
                m2 = fmax(m2, fmin(m1, g));
                if (g > m1) {
                    m1 = g;
                    sbest = S;
                    dbest = dev;
                    kbest = q0 + j;
                }
            }
        }
        __syncwarp();  // every lane is done with buffer b: the next iteration refills it
        double wg, r2;
        int wk;
        warp_argbest(m1, m2, sbest, dbest, kbest, wg, wk, r2);
        if (lane == 0) {
            Partial out;
            out.score = wg;
            out.aux = dbest;
            out.lsum = sbest;
            out.score2 = r2;
            out.k = wk == 0x7fffffff ? -1 : wk;
            DCHECK(out.k < sg.n - 1 && out.k >= -1);
            out.f = f;
            out.ppos = -1;
            out.pad_ = 0;
            partials[(int64_t)(sg.pbase + td.tin * sg.pstep + sg.pside) * a.F_search + f] = out;
            // raise the node's bound (non-negative doubles order like their bit patterns; NaN never gets here)
            if (VARIANT == 1 && wg > __longlong_as_double(nb_seen))
                atomicMax(a.node_best + td.seg, __double_as_longlong(wg));
        }
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
    const double *yp = a.ypay ? a.ypay + (int64_t)total_col * a.col_stride + sg.start : nullptr;
    double S = 0.0;
    double yv = lane < sg.n ? (yp ? yp[lane] : __ldg(a.y + col[lane])) : 0.0;
    for (int base = 0; base < sg.n; base += 32) {
        const double cur = yv;
        const int nx = base + 32 + lane;
        yv = nx < sg.n ? (yp ? yp[nx] : __ldg(a.y + col[nx])) : 0.0;
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
    const double *yp = a.ypay ? a.ypay + (int64_t)f * a.col_stride + sg.start : nullptr;
    double S = 0.0;
    Cand best = cand_none();
    double yv = lane < n ? (yp ? yp[lane] : __ldg(a.y + col[lane])) : 0.0;
    for (int base = 0; base < n; base += 32) {
        const double cur = yv;
        const int nx = base + 32 + lane;
        yv = nx < n ? (yp ? yp[nx] : __ldg(a.y + col[nx])) : 0.0;
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
        out.ppos = -1;
        out.pad_ = 0;
        partials[(int64_t)(sg.pbase + sg.pside) * a.F_search + f] = out;
    }
}

// ---- per-node reduction over (tiles x columns) with the reference's full tie rule -------
// (score desc, k asc, aux desc, f asc): within one sorted position the reference compares
// |S-(k+1)mean| across features (one.py:59) and across positions the weighted squares
// (one.py:62); aux is 0 for the other criteria, which then tie-break on (k, f) only.
// FLAT: the criteria that take a flat row-major argmax (four.py:28, one.py:91, strategy_l1.py:120)
// break ties on (k, f) only; there the aux slot carries data (the left weight sum), not a tie key.
template <bool FLAT>
__device__ __forceinline__ bool partial_better(const Partial &b, const Partial &a) {
    if (b.k < 0) return false;
    if (a.k < 0) return true;
    if (b.score != a.score) return b.score > a.score;
    if (b.k != a.k) return b.k < a.k;
    if (!FLAT && b.aux != a.aux) return b.aux > a.aux;
    return b.f < a.f;
}

template <bool FLAT>
__device__ __forceinline__ void partial_merge(Partial &a, const Partial &b) {
    const bool bw = partial_better<FLAT>(b, a);
    const double loser = bw ? (a.k < 0 ? -CUDART_INF : a.score) : (b.k < 0 ? -CUDART_INF : b.score);
    const double s2 = fmax(fmax(a.score2, b.score2), loser);
    if (bw) a = b;
    a.score2 = s2;
}

__device__ __forceinline__ Partial partial_none() {
    Partial b;
    b.score = -CUDART_INF;
    b.aux = 0.0;
    b.lsum = 0.0;
    b.score2 = -CUDART_INF;
    b.k = -1;
    b.f = 0;
    b.ppos = -1;
    b.pad_ = 0;
    return b;
}

template <bool FLAT>
__device__ __forceinline__ Partial partial_block_reduce(Partial b, Partial *sp) {
    const int tid = threadIdx.x;
    sp[tid] = b;
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if (tid < s) {
            Partial mine = sp[tid];
            partial_merge<FLAT>(mine, sp[tid + s]);
            sp[tid] = mine;
        }
        __syncthreads();
    }
    return sp[0];
}

constexpr int FIN_SLICES = 64;  // stage-1 blocks per node with many candidates

// stage 1: blockIdx.y = node, blockIdx.x = slice of that node's (tiles x columns) candidates
template <bool FLAT>
__global__ void __launch_bounds__(128) finalize_slices_kernel(LevelArgs a, const Partial *__restrict__ partials,
                                                              Partial *__restrict__ sliced) {
    __shared__ Partial sp[128];
    const int seg = blockIdx.y;
    const SegDev sg = a.segs[seg];
    const int64_t count = (int64_t)sg.np * a.F_search;
    Partial b = partial_none();
    for (int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x; r < count; r += (int64_t)FIN_SLICES * 128) {
        const int64_t j = r / a.F_search, f = r - j * a.F_search;
        partial_merge<FLAT>(b, partials[(sg.pbase + j * sg.pstep + sg.pside) * a.F_search + f]);
    }
    b = partial_block_reduce<FLAT>(b, sp);
    if (threadIdx.x == 0) sliced[(int64_t)seg * FIN_SLICES + blockIdx.x] = b;
}

// stage 2 (or the only stage when `sliced` is null): one block per node
template <bool FLAT>
__global__ void __launch_bounds__(128) finalize_kernel(LevelArgs a, const Partial *__restrict__ partials,
                                                       const Partial *__restrict__ sliced,
                                                       const double *__restrict__ X, int64_t ldx, int col_offset,
                                                       b200dt_candidate *__restrict__ out) {
    __shared__ Partial sp[128];
    const int seg = blockIdx.x;
    const SegDev sg = a.segs[seg];
    const int tid = threadIdx.x;
    Partial b = partial_none();
    if (sliced) {
        for (int r = tid; r < FIN_SLICES; r += 128) partial_merge<FLAT>(b, sliced[(int64_t)seg * FIN_SLICES + r]);
    } else {
        const int64_t count = (int64_t)sg.np * a.F_search;
        for (int64_t r = tid; r < count; r += 128) {
            const int64_t j = r / a.F_search, f = r - j * a.F_search;
            partial_merge<FLAT>(b, partials[(sg.pbase + j * sg.pstep + sg.pside) * a.F_search + f]);
        }
    }
    const Partial w = partial_block_reduce<FLAT>(b, sp);
    if (tid == 0) {
        b200dt_candidate c;
        c.gain = w.score;
        c.aux = w.aux;
        c.left_sum = w.lsum;
        c.gain2 = w.score2;
        c.threshold = 0.0;
        c.k = w.k;
        c.f = w.f + col_offset;
        c.exact = sg.exact;
        c.pad = w.ppos;   // position inside the parent's segment while the node's rows have not been moved
        if (w.k >= 0 && X) {
            int row = sg.vstart >= 0 ? a.orders[(int64_t)w.f * a.col_stride + sg.vstart + w.ppos]
                                     : a.orders[(int64_t)w.f * a.col_stride + sg.start + w.k];
            DCHECK(row >= 0);
            if (a.row_map) row = a.row_map[row];
            DCHECK(row >= 0 && w.k < sg.n);
            c.threshold = X[(int64_t)row * ldx + w.f];  // one.py:104-105: a data value, not a midpoint
        }
        out[seg] = c;
    }
}

// ---- stand-alone scoring of caller-supplied column cumsums (one.py:56-93 as plain functions) -------------
// cums: (n, F) row-major.  One thread per sorted position k < n-1 replays the reference's arithmetic over the
// F columns with correctly rounded, uncontracted operations; blocks reduce to one record each.
//   variant 1: best_variance_improvements     (one.py:56-63): row-wise argmax of |S-(k+1)mean| over features (lowest
//              feature on ties), then argmax over rows of np_gene * value^2 (lowest row on ties)
//   variant 0: best_variance_improvements_v0  (one.py:66-78): flat row-major argmax of square(ra*S - mean*rb)
//   variant 2: best_variance_improvements_v2  (one.py:81-93): flat argmax of |ra*S - mean*rb|, returns its square
struct CumBest {
    double score;
    int k, f;
};

__global__ void __launch_bounds__(256) score_cumsums_kernel(const double *__restrict__ cums, int n, int F, int variant,
                                                            CumBest *__restrict__ out) {
    __shared__ CumBest sb[256];
    const int k = blockIdx.x * 256 + threadIdx.x;
    CumBest best;
    best.score = -CUDART_INF;
    best.k = 0x7fffffff;
    best.f = 0x7fffffff;
    if (k < n - 1) {
        const double mean = __ddiv_rn(cums[(int64_t)(n - 1) * F + (F - 1)], (double)n);   // ys[-1][-1] / n
        const double *row = cums + (int64_t)k * F;
        if (variant == 1) {
            const double resid = __dmul_rn((double)(k + 1), mean);                          // one.py:57
            double v = -1.0;
            int bf = 0;
            for (int f = 0; f < F; ++f) {
                const double t = fabs(__dsub_rn(row[f], resid));                             // one.py:58
                if (t > v || (t != t && v == v)) {   // np.argmax: first maximum, NaN wins
                    v = t;
                    bf = f;
                }
            }
            const double w = __ddiv_rn(1.0, __dmul_rn((double)__int2float_rn(k + 1), (double)(n - 1 - k)));
            best.score = __dmul_rn(w, __dmul_rn(v, v));                                      // one.py:61
            best.k = k;
            best.f = bf;
        } else {
            const float lc = __int2float_rn(k + 1), rc = __int2float_rn(n - 1 - k);
            const float ra = __fsqrt_rn(__frcp_rn(__fmul_rn(lc, rc)));
            const float rb = __fsqrt_rn(__fdiv_rn(lc, rc));
            const double mb = __dmul_rn(mean, (double)rb);
            for (int f = 0; f < F; ++f) {
                const double d = __dsub_rn(__dmul_rn((double)ra, row[f]), mb);
                const double sc = variant == 0 ? __dmul_rn(d, d) : fabs(d);
                if (sc > best.score || (sc != sc && best.score == best.score)) {
                    best.score = sc;
                    best.k = k;
                    best.f = f;
                }
            }
        }
    }
    sb[threadIdx.x] = best;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if (threadIdx.x < s) {
            const CumBest a = sb[threadIdx.x], b = sb[threadIdx.x + s];
            // first maximum in row-major order; a NaN score wins over numbers like np.argmax
            const bool a_nan = a.score != a.score, b_nan = b.score != b.score;
            bool take_b;
            if (a_nan != b_nan) take_b = b_nan;
            else if (!a_nan && a.score != b.score) take_b = b.score > a.score;
            else take_b = b.k < a.k || (b.k == a.k && b.f < a.f);
            if (take_b) sb[threadIdx.x] = b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sb[0];
}

}  // namespace

static int persistent_grid(b200dt_ctx *ctx, const void *kernel, int threads, int64_t max_blocks) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int64_t g = (int64_t)per_sm * ctx->sm_count;
    if (g > max_blocks) g = max_blocks;
    return (int)(g < 1 ? 1 : g);
}

int launch_l2_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles, int64_t n_elems,
                   ScanStatus st, u32 epoch, u32 *ticket, Partial *partials, int variant, bool bulk) {
    (void)ticket;
    if (n_tiles <= 0) return 0;
    const int64_t total = (int64_t)n_tiles * a.F_search;
    LaunchScope ls(ctx, KC_L2_SCAN, 12.0 * (double)n_elems * a.F_search);
    if (a.ypay) {
        // streaming kernel: tiles of ST_TILE positions (the host builds the tile list accordingly)
        const size_t smem = bulk ? SCAN_SMEM : 8 * ST_PITCH * sizeof(double);
        const int threads = bulk ? SCAN_WARPS * 32 : 256;
        // (per context = per device: a process may drive several GPUs)
        if (!(ctx->attr_mask & 4u)) {
            CK(cudaFuncSetAttribute(l2_scan_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * ST_PITCH * sizeof(double))));
            CK(cudaFuncSetAttribute(l2_scan_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * ST_PITCH * sizeof(double))));
            CK(cudaFuncSetAttribute(l2_scan_bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM));
            CK(cudaFuncSetAttribute(l2_scan_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM));
            ctx->attr_mask |= 4u;
        }
        const void *kern = bulk ? (variant == 1 ? (const void *)l2_scan_bulk_kernel<1> : (const void *)l2_scan_bulk_kernel<2>)
                                : (variant == 1 ? (const void *)l2_scan_stream_kernel<1> : (const void *)l2_scan_stream_kernel<2>);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const int wpb = threads / 32;
        int64_t grid = std::min<int64_t>((int64_t)per_sm * ctx->sm_count, (total + wpb - 1) / wpb);
        LevelArgs a_ = a;
        void *args[] = {&a_, &tiles, &n_tiles, &st, &epoch, &partials};
        CK(launch_resident(kern, (unsigned)grid, threads, args, smem, ctx->stream));
        return 0;
    }
    // every block must be resident at once (tiles wait on earlier tiles): grid <= occupancy limit
    const void *kern = variant == 1 ? (const void *)l2_scan_kernel<1> : (const void *)l2_scan_kernel<2>;
    const int grid = persistent_grid(ctx, kern, 256, (total + 7) / 8);
    LevelArgs a_ = a;
    void *args[] = {&a_, &tiles, &n_tiles, &st, &epoch, &partials};
    CK(launch_resident(kern, (unsigned)grid, 256, args, 0, ctx->stream));
    return 0;
}

// node totals in the order of column `total_col` (one.py:57), for the listed nodes
int launch_l2_exact_totals(b200dt_ctx *ctx, const LevelArgs &a, const int *list, int n_list, int64_t n_elems,
                           int total_col, double *exact_total) {
    if (n_list <= 0) return 0;
    LaunchScope ls(ctx, KC_L2_EXACT, 12.0 * (double)n_elems);
    exact_total_kernel<<<(n_list + 7) / 8, 256, 0, ctx->stream>>>(a, list, n_list, total_col, exact_total);
    CK(cudaGetLastError());
    return 0;
}

// totals_given: exact_total[] already holds the totals of the listed nodes (feature-sharded fits: they come from
// the rank that owns the last global column)
int launch_l2_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small, int64_t n_elems,
                    int total_col, double *exact_total, Partial *partials, int variant, bool totals_given) {
    if (n_small <= 0) return 0;
    if (!totals_given) CKR(launch_l2_exact_totals(ctx, a, small_list, n_small, n_elems, total_col, exact_total));
    const int64_t warps = (int64_t)n_small * a.F_search;
    LaunchScope ls(ctx, KC_L2_EXACT, 12.0 * (double)n_elems * a.F_search);
    l2_exact_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(a, small_list, n_small, exact_total, partials, variant);
    CK(cudaGetLastError());
    return 0;
}

// feature-sharded fits: the candidate records of nodes whose total is still to be exchanged carry the total
// (owner of the last global column: f = that column) or nothing (other ranks: f = -1); exact = 2 marks them
__global__ void emit_totals_kernel(const int *__restrict__ list, int n_list, const double *__restrict__ exact_total,
                                   int owner_col, b200dt_candidate *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    const int seg = list[i];
    b200dt_candidate c;
    c.gain = 0.0;
    c.aux = 0.0;
    c.left_sum = owner_col >= 0 ? exact_total[seg] : 0.0;
    c.gain2 = 0.0;
    c.threshold = 0.0;
    c.k = -1;
    c.f = owner_col;
    c.exact = 2;
    c.pad = 0;
    out[seg] = c;
}

// nodes that are NOT searched in this round must not leave stale candidate slots for the finalize pass
__global__ void clear_partials_kernel(const int *__restrict__ list, int n_list, const SegDev *__restrict__ segs, int F,
                                      Partial *__restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list * F) return;
    const SegDev sg = segs[list[i / F]];
    Partial b;
    b.score = -CUDART_INF;
    b.aux = 0.0;
    b.lsum = 0.0;
    b.score2 = -CUDART_INF;
    b.k = -1;
    b.f = i % F;
    b.ppos = -1;
    b.pad_ = 0;
    partials[(int64_t)(sg.pbase + sg.pside) * F + (i % F)] = b;
}

int launch_clear_partials(b200dt_ctx *ctx, const LevelArgs &a, const int *list, int n_list, Partial *partials) {
    if (n_list <= 0) return 0;
    LaunchScope ls(ctx, KC_MISC, 0.0);
    const int64_t n = (int64_t)n_list * a.F_search;
    clear_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(list, n_list, a.segs, a.F_search, partials);
    CK(cudaGetLastError());
    return 0;
}

int launch_emit_totals(b200dt_ctx *ctx, const int *list, int n_list, const double *exact_total, int owner_col,
                       b200dt_candidate *out) {
    if (n_list <= 0) return 0;
    LaunchScope ls(ctx, KC_MISC, 0.0);
    emit_totals_kernel<<<(n_list + 127) / 128, 128, 0, ctx->stream>>>(list, n_list, exact_total, owner_col, out);
    CK(cudaGetLastError());
    return 0;
}

int launch_finalize(b200dt_ctx *ctx, const LevelArgs &a, const Partial *partials, int64_t max_records,
                    Partial *sliced, const double *X, int64_t ldx, int col_offset, b200dt_candidate *out,
                    bool flat_rule) {
    if (a.n_segs <= 0) return 0;
    const Partial *stage1 = nullptr;
    if (max_records > 64 * 1024 && sliced && a.n_segs <= 65535) {   // (gridDim.y limit; deeper levels have small nodes)
        LaunchScope ls(ctx, KC_FINALIZE, 0.0);
        if (flat_rule)
            finalize_slices_kernel<true><<<dim3(FIN_SLICES, a.n_segs), 128, 0, ctx->stream>>>(a, partials, sliced);
        else
            finalize_slices_kernel<false><<<dim3(FIN_SLICES, a.n_segs), 128, 0, ctx->stream>>>(a, partials, sliced);
        stage1 = sliced;
    }
    LaunchScope ls(ctx, KC_FINALIZE, 0.0);
    if (flat_rule)
        finalize_kernel<true><<<a.n_segs, 128, 0, ctx->stream>>>(a, partials, stage1, X, ldx, col_offset, out);
    else
        finalize_kernel<false><<<a.n_segs, 128, 0, ctx->stream>>>(a, partials, stage1, X, ldx, col_offset, out);
    CK(cudaGetLastError());
    return 0;
}

// replaces one.best_variance_improvements / _v0 / _v2 as stand-alone functions   (ref one.py:56-93)
extern "C" int b200dt_l2_score_cumsums(b200dt_ctx *ctx, const double *cumsums, int64_t n, int64_t F, int variant,
                                       int64_t *out_k, int64_t *out_f, double *out_score) {
    if (!ctx) return 1;
    if (!cumsums || !out_k || !out_f || !out_score) return ctx->fail_msg("l2_score_cumsums: null argument");
    if (n < 2 || F < 1) return ctx->fail_msg("l2_score_cumsums: attempt to get argmax of an empty sequence (n < 2)");
    if (n >= (1ll << 31) || F >= (1ll << 31)) return ctx->fail_msg("l2_score_cumsums: shape out of range");
    if (variant < 0 || variant > 2) return ctx->fail_msg("l2_score_cumsums: variant must be 0, 1 or 2");
    CK(cudaSetDevice(ctx->device));
    if (ctx->session || ctx->class_session || ctx->row_session)
        return ctx->fail_msg("l2_score_cumsums: a fit session is in progress on this context");
    double *d_c;
    CumBest *d_out;
    const int blocks = (int)((n - 1 + 255) / 256);
    CKR(scratch_get(ctx, SB_X, (size_t)n * F * 8, (void **)&d_c));
    CKR(scratch_get(ctx, SB_OUT, (size_t)blocks * sizeof(CumBest), (void **)&d_out));
    CK(cudaMemcpyAsync(d_c, cumsums, (size_t)n * F * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, KC_L2_EXACT, 8.0 * (double)n * F);
        score_cumsums_kernel<<<blocks, 256, 0, ctx->stream>>>(d_c, (int)n, (int)F, variant, d_out);
    }
    CK(cudaGetLastError());
    std::vector<CumBest> h(blocks);
    CK(cudaMemcpyAsync(h.data(), d_out, (size_t)blocks * sizeof(CumBest), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CumBest best = h[0];
    for (int b = 1; b < blocks; ++b) {   // blocks are in row order: a strictly greater score (or the first NaN) wins
        const bool a_nan = best.score != best.score, b_nan = h[b].score != h[b].score;
        if ((b_nan && !a_nan) || (!a_nan && !b_nan && h[b].score > best.score)) best = h[b];
    }
    *out_k = best.k;
    *out_f = best.f;
    *out_score = variant == 2 ? best.score * best.score : best.score;   // one.py:93 returns the square
    return 0;
}
