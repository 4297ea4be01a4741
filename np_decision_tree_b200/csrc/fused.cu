// K2+K4 fused: one streaming pass per tree level that PARTITIONS every split node and, on the
// way, SEARCHES its two children.
//
// replaces, per level:  split_orders                      (ref decision_tree/one.py:110-118)
//                       + greedy_regression_split of both children   (ref one.py:96-107, 52-63)
//
// The stand-alone search gathers y through the row ids: one 32-byte L2 sector per 8 useful bytes,
// which ncu showed to be the bound (L1 wavefronts / L2, not HBM).  Here y travels WITH the row id
// (a payload column per feature), so the level is pure streaming: read (row, y) 12 B, look up the
// row's side bit, write (row, y) 12 B to the child's range.  Because the left / right prefix
// sums of y along the parent's order ARE the children's prefix sums along their own order, the
// gain of every split position of both children is evaluated in the same pass (the children's
// sizes and totals are known from the parent's winning split).  Candidates go to two slots per
// (parent tile, column): one per side.
//
// One warp owns one tile of 512 parent positions of one (node, column); persistent grid, static
// tile-major schedule, no block barrier (see l2.cu).  The carry between tiles is the triple
// (left count, left sum, right sum), exchanged through self-validating 48-byte slots.
// Algorithmic bytes: 21 per element (the 12 of the search + the 9 of the partition it replaces);
// actual traffic 24 (+ the mask look-up).
#include <algorithm>

#include "l2_score.cuh"

namespace {

template <int VARIANT>
__global__ void __launch_bounds__(256, 3) level_fused_kernel(
    const int *__restrict__ ord_in, const double *__restrict__ y_in, int *__restrict__ ord_out,
    double *__restrict__ y_out, int64_t col_stride, int n_cols, int F_search, const PlanSeg *__restrict__ plan,
    const TileDesc *__restrict__ tiles, int n_tiles, const u32 *__restrict__ mask, u64 *status, u32 epoch,
    Partial *__restrict__ partials) {
    // per warp: y of the tile (striped write, blocked read), the per-side prefix sums in blocked
    // order, and the 16 side ballots
    extern __shared__ __align__(16) unsigned char fused_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sy = reinterpret_cast<double *>(fused_smem) + warp * SY_PITCH;
    double *sp = reinterpret_cast<double *>(fused_smem) + (8 + warp) * SY_PITCH + lane * 17;  // this lane's 16 blocked slots
    u32 *sbal = reinterpret_cast<u32 *>(fused_smem + 16 * SY_PITCH * sizeof(double)) + warp * 16;
    const u32 lt = lanemask_lt();
    const int64_t total_tiles = (int64_t)n_tiles * n_cols;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % n_cols), ti = (int)(t / n_cols);
        const TileDesc td = tiles[ti];
        const PlanSeg sg = plan[td.seg];
        const int off = td.tin * WT_TILE;
        const int cnt = min(WT_TILE, sg.n - off);
        const int64_t cbase = (int64_t)f * col_stride + sg.start;
        const int *col = ord_in + cbase + off;
        const double *ycol = y_in + cbase + off;

        // ---- load (row, y), striped; y is staged in shared memory for the blocked scan --------
        int r[WT_ITEMS];
        if (cnt == WT_TILE) {
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) r[i] = ld_stream_i32(col + i * 32 + lane);
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) sy[i * 34 + lane + (lane >> 4)] = __ldcs(ycol + i * 32 + lane);
        } else {
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const int e = i * 32 + lane;
                r[i] = e < cnt ? ld_stream_i32(col + e) : -1;
                sy[i * 34 + lane + (lane >> 4)] = e < cnt ? __ldcs(ycol + e) : 0.0;
            }
        }
        // ---- side of every row (one.py:111-113: bool_aux[orders]) ------------------------------
        u32 leftbits = 0;
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            const int row = r[i];
            const bool is_left = row >= 0 && ((__ldg(mask + (row >> 5)) >> (row & 31)) & 1u);
            leftbits |= (is_left ? 1u : 0u) << i;
        }
        int tile_left = 0;
        u32 mybal = 0;
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            const u32 b = __ballot_sync(FULL_MASK, (leftbits >> i) & 1u);
            if (lane == i) mybal = b;
            tile_left += __popc(b);
        }
        if (lane < WT_ITEMS) sbal[lane] = mybal;
        __syncwarp();

        // ---- blocked order: lane owns positions lane*16 .. lane*16+15 ------------------------------
        const u32 fl16 = (sbal[lane >> 1] >> ((lane & 1) * 16)) & 0xffffu;
        // inclusive prefix sum of each element's OWN side inside this lane (kept in shared memory:
        // a rolled loop keeps the code small -- the unrolled version stalled on instruction fetch)
        double sL = 0.0, sR = 0.0;
        {
            const double *src = sy + lane * 17;
#pragma unroll 4
            for (int j = 0; j < WT_ITEMS; ++j) {
                const double yb = src[j];
                const bool L = (fl16 >> j) & 1u;
                sL += L ? yb : 0.0;
                sR += L ? 0.0 : yb;
                sp[j] = L ? sL : sR;
            }
        }
        const int cl = __popc(fl16);
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
        tile.c = tile_left;
        tile.l = shfl_d(iL, 31);
        tile.r = shfl_d(iR, 31);
        const Carry3 pre = lookback3(status, (int64_t)f * n_tiles + ti, td.tin, epoch, tile, lane);

        // ---- search of the two children --------------------------------------------------------------
        if (f < F_search && (sg.score_l | sg.score_r)) {
            const int nL = sg.nleft, nR = sg.n - sg.nleft;
            const double meanL = sg.total_l / (double)nL, meanR = sg.total_r / (double)nR;
            const double baseL = pre.l + (iL - sL), baseR = pre.r + (iR - sR);   // sums before this lane
            int lefts = pre.c + (ic - cl);                                      // left rows before this lane
            const int p0 = off + lane * WT_ITEMS;                               // parent position of item 0
            // one running best per side, in registers (no dynamic indexing)
            double m1L = -CUDART_INF, m2L = -CUDART_INF, sbL = 0.0, dbL = 0.0;
            double m1R = -CUDART_INF, m2R = -CUDART_INF, sbR = 0.0, dbR = 0.0;
            int kbL = 0x7fffffff, kbR = 0x7fffffff;
#pragma unroll 2
            for (int j = 0; j < WT_ITEMS; ++j) {
                const bool L = (fl16 >> j) & 1u;
                const int kk = L ? lefts : (p0 + j - lefts);   // position inside the child
                const int nc = L ? nL : nR;
                const double S = (L ? baseL : baseR) + sp[j];
                double dev = 0.0, g = -CUDART_INF;
                const bool scored = L ? (sg.score_l != 0) : (sg.score_r != 0);
                if (scored && lane * WT_ITEMS + j < cnt && kk < nc - 1)
                    g = score_fast(S, kk, nc, L ? meanL : meanR, VARIANT, dev);
                // children positions only grow along the parent order, so "first maximum" = lowest k
                const double gl = L ? g : -CUDART_INF, gr = L ? -CUDART_INF : g;
                m2L = fmax(m2L, fmin(m1L, gl));
                m2R = fmax(m2R, fmin(m1R, gr));
                if (gl > m1L) {
                    m1L = gl;
                    sbL = S;
                    dbL = dev;
                    kbL = kk;
                }
                if (gr > m1R) {
                    m1R = gr;
                    sbR = S;
                    dbR = dev;
                    kbR = kk;
                }
                lefts += L ? 1 : 0;
            }
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (side == 0 ? !sg.score_l : !sg.score_r) continue;
                Cand best;
                best.g = side == 0 ? m1L : m1R;
                best.g2 = side == 0 ? m2L : m2R;
                best.s = side == 0 ? sbL : sbR;
                best.dev = side == 0 ? dbL : dbR;
                best.k = side == 0 ? kbL : kbR;
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
                    partials[(int64_t)(sg.fbase + 2 * td.tin + side) * F_search + f] = out;
                }
            }
        }

        // ---- stable partition: (row, y) to the children's ranges (one.py:114-116) ---------------------
        int lb_row = pre.c;
        int *ocol = ord_out + cbase;
        double *oy = y_out + cbase;
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            const u32 b = sbal[i];
            const int e = i * 32 + lane;
            if (e < cnt) {
                const int lb = lb_row + __popc(b & lt);  // left rows before this element
                const int dest = ((leftbits >> i) & 1u) ? lb : sg.nleft + (off + e - lb);
                ocol[dest] = r[i];
                oy[dest] = sy[i * 34 + lane + (lane >> 4)];
            }
            lb_row += __popc(b);
        }
        __syncwarp();  // sy / sbal are reused by the next tile
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

int launch_level_fused(b200dt_ctx *ctx, const int *ord_in, const double *y_in, int *ord_out, double *y_out,
                       int64_t col_stride, int F_store, int F_search, const PlanSeg *plan, const TileDesc *tiles,
                       int n_tiles, int64_t n_elems, const u32 *mask, u64 *status, u32 epoch, Partial *partials,
                       int variant) {
    if (n_tiles <= 0) return 0;
    const int64_t total = (int64_t)n_tiles * F_store;
    const void *kern = variant == 1 ? (const void *)level_fused_kernel<1> : (const void *)level_fused_kernel<2>;
    const size_t smem = 16 * SY_PITCH * sizeof(double) + 8 * 16 * sizeof(u32);
    // (per context = per device: a process may drive several GPUs)
    if (!(ctx->attr_mask & 1u)) {
        CK(cudaFuncSetAttribute(level_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(level_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_mask |= 1u;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * ctx->sm_count;  // all blocks resident: tiles wait on earlier tiles
    if (grid > (total + 7) / 8) grid = (total + 7) / 8;
    LaunchScope ls(ctx, KC_LEVEL_FUSED, 21.0 * (double)n_elems * F_store);
    void *args[] = {&ord_in, &y_in, &ord_out, &y_out, &col_stride, &F_store, &F_search, &plan, &tiles, &n_tiles,
                    &mask, &status, &epoch, &partials};
    CK(launch_resident(kern, (unsigned)grid, 256, args, smem, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int launch_gather_payload(b200dt_ctx *ctx, const int *orders, const double *y, double *ypay, int64_t col_stride,
                          int F_store, int64_t N) {
    if (N <= 0 || F_store <= 0) return 0;
    int gx = (int)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->sm_count * 16 / std::max(1, std::min(F_store, 16)) + 1);
    LaunchScope ls(ctx, KC_PAYLOAD, 20.0 * (double)N * F_store);
    gather_payload_kernel<<<dim3((unsigned)gx, (unsigned)F_store), 256, 0, ctx->stream>>>(orders, y, ypay, col_stride, N);
    CK(cudaGetLastError());
    return 0;
}
