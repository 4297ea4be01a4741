// Two levels per data movement (L2 fits with the y payload).
//
// replaces, for a PAIR of levels:  2 x split_orders                       (ref decision_tree/one.py:110-118)
//                                  + greedy_regression_split of the level in between   (ref one.py:96-107, 52-63)
//
// The stable partition is bound by one scattered side-bit look-up per element plus 24 B of streaming per
// element (partition.cu).  The children of a split can be SEARCHED without moving anything: along the
// parent's sorted order the running sums of the left rows / of the right rows ARE the children's prefix sums
// along their own order, and a child's position is the count of its rows seen so far.  So after the decisions of
// level d the data stays where it is; `l2_children_scan_kernel` streams the parent's (row, y) order once, looks up
// each row's side bit and scores every split position of BOTH children; the decisions of level d+1 mark their
// left rows through positions in the parent's order; and ONE 4-way partition (2-bit code per row) then moves the
// data for both levels.  Per pair of levels that is one streaming search + one side-bit search + one 4-way
// partition instead of two searches + two partitions.
//
// One warp owns one tile of ST_TILE (1024) parent positions of one (node, column); persistent co-resident grid,
// static tile-major schedule, carry between tiles = (left count, left sum, right sum) through `lookback3`.
#include <algorithm>

#include "l2_score.cuh"

namespace {

constexpr int CS_PITCH = ST_TILE + ST_TILE / 32;  // 1 pad double per 32: conflict-free blocked reads

template <int VARIANT>
__global__ void __launch_bounds__(256, 3) l2_children_scan_kernel(
    const int *__restrict__ ord, const double *__restrict__ ypay, int64_t col_stride, int F_search,
    const PlanSeg *__restrict__ plan, const TileDesc *__restrict__ tiles, int n_tiles,
    const u32 *__restrict__ mask, u64 *status, u32 epoch, Partial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sy = reinterpret_cast<double *>(cs_smem) + warp * CS_PITCH;
    const double *mine = sy + lane * (ST_ITEMS + 1);   // position lane*32 + j sits at lane*33 + j
    const int64_t total_tiles = (int64_t)n_tiles * F_search;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % F_search), ti = (int)(t / F_search);
        const TileDesc td = tiles[ti];
        const PlanSeg sg = plan[td.seg];
        const int off = td.tin * ST_TILE;
        const int cnt = min(ST_TILE, sg.n - off);
        const int64_t cbase = (int64_t)f * col_stride + sg.start + off;
        const double *yp = ypay + cbase;
        const int *col = ord + cbase;
        // ---- striped loads: y into shared memory ---------------------------------------------------------------
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
        // ---- side bits: the ballot of striped row i holds exactly the 32 positions lane i owns in blocked order.
        //      Two rounds of 16 rows: all row ids of a round, then all their mask words, are in flight together
        u32 mybits = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = (h * 16 + u) * 32 + lane;
                r[u] = e < cnt ? ld_stream_i32(col + e) : -1;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) r[u] = r[u] >= 0 ? (int)((__ldg(mask + (r[u] >> 5)) >> (r[u] & 31)) & 1u) : 0;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const u32 b = __ballot_sync(FULL_MASK, r[u] != 0);
                if (lane == h * 16 + u) mybits = b;
            }
        }
        __syncwarp();
        const int nmine = max(0, min(ST_ITEMS, cnt - lane * ST_ITEMS));   // positions of this lane inside the node
        const u32 vmask = nmine >= 32 ? 0xffffffffu : ((1u << nmine) - 1u);
        // ---- this lane's sums per side, then the exclusive prefixes across lanes and tiles -----------------
        double sL = 0.0, sA = 0.0;
#pragma unroll 8
        for (int j = 0; j < ST_ITEMS; ++j) {
            const double y = mine[j];   // (0 beyond the node)
            sA += y;
            sL += ((mybits >> j) & 1u) ? y : 0.0;
        }
        const double sR = sA - sL;   // (rounding differs from a separate sum: approximate search only)
        const int cl = __popc(mybits & vmask);
        double iL = sL, iR = sR;
        int ic = cl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double oL = shfl_up_d(iL, d), oR = shfl_up_d(iR, d);
            const int oc = __shfl_up_sync(FULL_MASK, ic, d);
            if (lane >= d) {
                iL += oL;
                iR += oR;
                ic += oc;
            }
        }
        Carry3 tile;
        tile.c = __shfl_sync(FULL_MASK, ic, 31);
        tile.l = shfl_d(iL, 31);
        tile.r = shfl_d(iR, 31);
        const Carry3 pre = lookback3(status, (int64_t)f * n_tiles + ti, td.tin, epoch, tile, lane);

        // ---- score every split position of both children: one COMPACTED pass per side over the set bits of this
        //      lane's side mask, so the inner body is the single-sided one of the streaming search (l2.cu) --------
        const int nL = sg.nleft, nR = sg.n - sg.nleft;
        const int p0 = off + lane * ST_ITEMS;                    // parent position of this lane's first element
        const int kL0 = pre.c + (ic - cl);                       // left rows before it = next left child position
        double m1L = -CUDART_INF, m2L = -CUDART_INF, sbL = 0.0, dbL = 0.0;
        double m1R = -CUDART_INF, m2R = -CUDART_INF, sbR = 0.0, dbR = 0.0;
        int kbL = 0x7fffffff, kbR = 0x7fffffff, pbL = -1, pbR = -1;
        const bool scoreL = sg.score_l != 0, scoreR = sg.score_r != 0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !scoreL : !scoreR) continue;   // (warp-uniform)
            const int nc = side == 0 ? nL : nR;
            const double mean = (side == 0 ? sg.total_l : sg.total_r) * rcp_fast1((double)max(nc, 1));
            double S = side == 0 ? pre.l + (iL - sL) : pre.r + (iR - sR);   // child sum before this lane
            int k = side == 0 ? kL0 : p0 - kL0;                             // child position of its next row
            u32 bits = (side == 0 ? mybits : ~mybits) & vmask;
            double m1 = -CUDART_INF, m2 = -CUDART_INF, sb = 0.0, db = 0.0;
            int kb = 0x7fffffff, pb = -1;
            if (VARIANT == 1 && sg.n <= (1 << 24)) {
                double km = (double)(k + 1) * mean, den = (double)(k + 1) * (double)(nc - 1 - k);
                double dd = (double)(nc - 2 * k - 3);
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    S += mine[j];
                    const double dev = S - km;
                    const double g = k < nc - 1 ? dev * dev * rcp_fast1(den) : -CUDART_INF;   // not the child's last row
                    m2 = fmax(m2, fmin(m1, g));
                    if (g > m1) {
                        m1 = g;
                        sb = S;
                        db = dev;
                        kb = k;
                        pb = p0 + j;
                    }
                    km += mean;
                    den += dd;
                    dd -= 2.0;
                    ++k;
                }
                db = fabs(db);
            } else {
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    S += mine[j];
                    double dev = 0.0;
                    const double g = k < nc - 1 ? score_fast(S, k, nc, mean, VARIANT, dev) : -CUDART_INF;
                    m2 = fmax(m2, fmin(m1, g));
                    if (g > m1) {
                        m1 = g;
                        sb = S;
                        db = dev;
                        kb = k;
                        pb = p0 + j;
                    }
                    ++k;
                }
            }
            if (side == 0) {
                m1L = m1; m2L = m2; sbL = sb; dbL = db; kbL = kb; pbL = pb;
            } else {
                m1R = m1; m2R = m2; sbR = sb; dbR = db; kbR = kb; pbR = pb;
            }
        }
        __syncwarp();  // sy is overwritten by the next tile
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !scoreL : !scoreR) continue;   // (warp-uniform)
            Cand best;
            best.g = side == 0 ? m1L : m1R;
            best.g2 = side == 0 ? m2L : m2R;
            best.s = side == 0 ? sbL : sbR;
            best.dev = side == 0 ? dbL : dbR;
            best.k = side == 0 ? kbL : kbR;
            int ppos = side == 0 ? pbL : pbR;
            // (score desc, child position asc); the parent position travels with the winner
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const Cand o = cand_shfl_xor(best, m);
                const int op = __shfl_xor_sync(FULL_MASK, ppos, m);
                const bool o_wins = (o.g > best.g) || (o.g == best.g && o.k < best.k);
                cand_merge(best, o);
                if (o_wins) ppos = op;
            }
            if (lane == 0) {
                Partial out;
                out.score = best.g;
                out.aux = best.dev;
                out.lsum = best.s;
                out.score2 = best.g2;
                out.k = best.k == 0x7fffffff ? -1 : best.k;
                out.f = f;
                out.ppos = ppos;
                out.pad_ = 0;
                partials[(int64_t)(sg.fbase + 2 * td.tin + side) * F_search + f] = out;
            }
        }
    }
}

// Left rows of split nodes whose rows still lie inside their parent's segment: the rows of the node's side at
// parent positions <= ppos (the winning position) of the winning column.   marks: [f, parent start, ppos, side]
__global__ void __launch_bounds__(256) mark_left_virtual_kernel(const int *__restrict__ orders, int64_t col_stride,
                                                                const int *__restrict__ marks,
                                                                const u32 *__restrict__ parent_mask, u32 *mask) {
    const int m = blockIdx.y;
    const int f = marks[4 * m], start = marks[4 * m + 1], ppos = marks[4 * m + 2], side = marks[4 * m + 3];
    const int *col = orders + (int64_t)f * col_stride + start;
    for (int j = blockIdx.x * 256 + threadIdx.x; j <= ppos; j += gridDim.x * 256) {
        const int row = col[j];
        const u32 in_left_child = (parent_mask[row >> 5] >> (row & 31)) & 1u;
        if ((int)in_left_child == (side == 0 ? 1 : 0)) atomicOr(&mask[row >> 5], 1u << (row & 31));
    }
}

// codes[row] (2 bits) = A | B << 1: A = row is in the left child (level d), B = row is in the left grandchild of
// its child (level d+1).  16 rows per 32-bit word.
__global__ void __launch_bounds__(256) merge_codes_kernel(const u32 *__restrict__ mask_a, const u32 *__restrict__ mask_b,
                                                          int64_t n_words, u32 *__restrict__ codes) {
    for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * 256) {
        const u32 a = mask_a[w], b = mask_b[w];
        u32 lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            lo |= (((a >> i) & 1u) | (((b >> i) & 1u) << 1)) << (2 * i);
            hi |= (((a >> (16 + i)) & 1u) | (((b >> (16 + i)) & 1u) << 1)) << (2 * i);
        }
        codes[2 * w] = lo;
        codes[2 * w + 1] = hi;
    }
}

// 4-way stable partition of (row, y): classes LL | LR | RL | RR of every planned parent (Plan4), in that order.
// Carry between tiles: the counts of LL, LR, RL (RR follows from the position), through lookback3.
__global__ void __launch_bounds__(256, 3) partition4_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                            const double *__restrict__ y_in, double *__restrict__ y_out,
                                                            int64_t col_stride, const Plan4 *__restrict__ plan,
                                                            const TileDesc *__restrict__ tiles, int n_tiles,
                                                            const u32 *__restrict__ codes, u64 *status, u32 epoch,
                                                            int n_cols) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    const int64_t total_tiles = (int64_t)n_tiles * n_cols;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % n_cols), ti = (int)(t / n_cols);
        const TileDesc td = tiles[ti];
        const Plan4 sg = plan[td.seg];
        const int off = td.tin * P4_TILE;
        const int cnt = min(P4_TILE, sg.n - off);
        const int64_t cbase = (int64_t)f * col_stride + sg.start;
        const int *col = in + cbase + off;
        int rows[P4_ITEMS];
#pragma unroll
        for (int i = 0; i < P4_ITEMS; ++i) rows[i] = (i * 32 + lane < cnt) ? ld_stream_i32(col + i * 32 + lane) : -1;
        u32 abits = 0, bbits = 0;   // bit i: item i of this lane
#pragma unroll
        for (int i = 0; i < P4_ITEMS; ++i) {
            const int row = rows[i];
            const u32 code = row >= 0 ? ((__ldg(codes + (row >> 4)) >> ((row & 15) * 2)) & 3u) : 0u;
            abits |= (code & 1u) << i;
            bbits |= (code >> 1) << i;
        }
        int c_ll = 0, c_lr = 0, c_rl = 0;
#pragma unroll
        for (int i = 0; i < P4_ITEMS; ++i) {
            const bool ok = i * 32 + lane < cnt;
            const bool A = (abits >> i) & 1u, B = (bbits >> i) & 1u;
            c_ll += __popc(__ballot_sync(FULL_MASK, ok && A && B));
            c_lr += __popc(__ballot_sync(FULL_MASK, ok && A && !B));
            c_rl += __popc(__ballot_sync(FULL_MASK, ok && !A && B));
        }
        Carry3 tile;
        tile.c = c_ll;
        tile.l = (double)c_lr;   // whole numbers: exact in fp64
        tile.r = (double)c_rl;
        const Carry3 pre = lookback3(status, (int64_t)f * n_tiles + ti, td.tin, epoch, tile, lane);
        int b_ll = pre.c, b_lr = (int)pre.l, b_rl = (int)pre.r;
        int b_rr = off - b_ll - b_lr - b_rl;   // RR rows before this tile
        int *ocol = out + cbase;
        const double *ycol = y_in + cbase + off;
        double *oy = y_out + cbase;
#pragma unroll
        for (int i0 = 0; i0 < P4_ITEMS; i0 += 4) {
            double yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = (i0 + u) * 32 + lane;
                yv[u] = e < cnt ? __ldcs(ycol + e) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                const int e = i * 32 + lane;
                const bool ok = e < cnt;
                const bool A = (abits >> i) & 1u, B = (bbits >> i) & 1u;
                const u32 m_ll = __ballot_sync(FULL_MASK, ok && A && B);
                const u32 m_lr = __ballot_sync(FULL_MASK, ok && A && !B);
                const u32 m_rl = __ballot_sync(FULL_MASK, ok && !A && B);
                const u32 m_rr = __ballot_sync(FULL_MASK, ok && !A && !B);
                if (ok) {
                    int dest;
                    if (A)
                        dest = B ? b_ll + __popc(m_ll & lt) : sg.n_ll + b_lr + __popc(m_lr & lt);
                    else
                        dest = B ? sg.n_l + b_rl + __popc(m_rl & lt) : sg.n_l + sg.n_rl + b_rr + __popc(m_rr & lt);
                    __stcs(ocol + dest, rows[i]);
                    __stcs(oy + dest, yv[u]);
                }
                b_ll += __popc(m_ll);
                b_lr += __popc(m_lr);
                b_rl += __popc(m_rl);
                b_rr += __popc(m_rr);
            }
        }
    }
}

int resident_blocks(b200dt_ctx *ctx, const void *kern, size_t smem, int64_t work_warps) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * ctx->sm_count;
    if (grid > (work_warps + 7) / 8) grid = (work_warps + 7) / 8;
    return (int)std::max<int64_t>(1, grid);
}

}  // namespace

int launch_children_scan(b200dt_ctx *ctx, const int *ord, const double *ypay, int64_t col_stride, int F_search,
                         const PlanSeg *plan, const TileDesc *tiles, int n_tiles, int64_t n_elems, const u32 *mask,
                         u64 *status, u32 epoch, Partial *partials, int variant) {
    if (n_tiles <= 0) return 0;
    const size_t smem = 8 * CS_PITCH * sizeof(double);
    const void *kern = variant == 1 ? (const void *)l2_children_scan_kernel<1> : (const void *)l2_children_scan_kernel<2>;
    if (!(ctx->attr_mask & 128u)) {
        CK(cudaFuncSetAttribute(l2_children_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(l2_children_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_mask |= 128u;
    }
    const int grid = resident_blocks(ctx, kern, smem, (int64_t)n_tiles * F_search);
    LaunchScope ls(ctx, KC_CHILD_SCAN, 12.0 * (double)n_elems * F_search);
    void *args[] = {&ord, &ypay, &col_stride, &F_search, &plan, &tiles, &n_tiles, &mask, &status, &epoch, &partials};
    CK(launch_resident(kern, (unsigned)grid, 256, args, smem, ctx->stream));
    return 0;
}

int launch_mark_left_virtual(b200dt_ctx *ctx, const int *orders, int64_t col_stride, const int *marks, int n_marks,
                             int max_span, const u32 *parent_mask, u32 *mask) {
    if (n_marks <= 0) return 0;
    const int gx = std::max(1, std::min(2048, (max_span + 255) / 256));
    for (int m0 = 0; m0 < n_marks; m0 += 65535) {
        const int nm = std::min(65535, n_marks - m0);
        LaunchScope ls(ctx, KC_MARK, 0.0);
        mark_left_virtual_kernel<<<dim3(gx, nm), 256, 0, ctx->stream>>>(orders, col_stride, marks + 4 * m0, parent_mask,
                                                                         mask);
    }
    CK(cudaGetLastError());
    return 0;
}

// plan4: host array of n_plan records {start, n, n_l, n_ll, n_rl, 0}; tiles: P4_TILE tiles of those segments
int launch_partition4(b200dt_ctx *ctx, const int *orders_in, int *orders_out, const double *y_in, double *y_out,
                      int64_t col_stride, int F_store, const void *d_plan4, const TileDesc *tiles, int n_tiles,
                      int64_t n_elems, const u32 *mask_a, const u32 *mask_b, int64_t N, u32 *codes, u64 *status,
                      u32 epoch) {
    if (n_tiles <= 0) return 0;
    const int64_t n_words = (N + 31) / 32;
    {
        LaunchScope ls(ctx, KC_MISC, 16.0 * (double)n_words);
        merge_codes_kernel<<<(unsigned)std::min<int64_t>(1024, (n_words + 255) / 256), 256, 0, ctx->stream>>>(
            mask_a, mask_b, n_words, codes);
    }
    CK(cudaGetLastError());
    const void *kern = (const void *)partition4_kernel;
    const int grid = resident_blocks(ctx, kern, 0, (int64_t)n_tiles * F_store);
    const Plan4 *plan = (const Plan4 *)d_plan4;
    LaunchScope ls(ctx, KC_PARTITION, 2 * 9.0 * (double)n_elems * F_store);
    const u32 *codes_c = codes;
    void *args[] = {&orders_in, &orders_out, &y_in, &y_out, &col_stride, &plan, &tiles, &n_tiles, &codes_c, &status,
                    &epoch, &F_store};
    CK(launch_resident(kern, (unsigned)grid, 256, args, 0, ctx->stream));
    return 0;
}
