// K2 with sample weights: the split search of the reference's level-synchronous builder.
//
// replaces best_variance_improvements(ys, start, end, weights)      (ref decision_tree/four.py:21-30)
//
//   sum_y, sum_w : column prefix sums of y*w and of w in sorted order        four.py:22-23
//   expected     = sum_w * total_y / total_w   (totals from the LAST column)  four.py:24
//   gain[k, f]   = (sum_y - expected)^2 / (sum_w (total_w - sum_w)), k < n-1  four.py:25-27
//   flat row-major argmax => (gain desc, k asc, f asc)                       four.py:28-29
//
// Two payload columns travel with every row id (y*w and w), so both kernels stream.
//  * l2w_exact_kernel   nodes <= EXACT_ROWS rows and near-ties: one warp per (node, column), both
//    sums advanced one element at a time in fp64 with the reference's operation order.
//  * l2w_scan_stream_kernel   larger nodes: one warp per 512-position tile, both channels staged in
//    shared memory, (sum_y, sum_w) carried through one fence-free look-back.
#include <algorithm>

#include "l2_score.cuh"

namespace {

__global__ void __launch_bounds__(256) mul_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                  int64_t n, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __dmul_rn(a[i], b[i]);   // four.py:117  y = y * weight
}

// totals of y*w and w of each listed node in the sorted order of column total_col (four.py:24)
__global__ void __launch_bounds__(256) exactw_total_kernel(LevelArgs a, const int *__restrict__ list, int n_list,
                                                           int total_col, double *__restrict__ ty,
                                                           double *__restrict__ tw) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_list) return;
    const int lane = threadIdx.x & 31;
    const int seg = list[gw];
    const SegDev sg = a.segs[seg];
    const double *yp = a.ypay + (int64_t)total_col * a.col_stride + sg.start;
    const double *wp = a.wpay + (int64_t)total_col * a.col_stride + sg.start;
    double Sy = 0.0, Sw = 0.0;
    for (int base = 0; base < sg.n; base += 32) {
        const int i = base + lane;
        const double yv = i < sg.n ? yp[i] : 0.0, wv = i < sg.n ? wp[i] : 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            Sy = __dadd_rn(Sy, shfl_d(yv, j));
            Sw = __dadd_rn(Sw, shfl_d(wv, j));
        }
    }
    if (lane == 0) {
        ty[seg] = Sy;
        tw[seg] = Sw;
    }
}

__global__ void __launch_bounds__(256) l2w_exact_kernel(LevelArgs a, const int *__restrict__ list, int n_list,
                                                        const double *__restrict__ ty, const double *__restrict__ tw,
                                                        Partial *__restrict__ partials) {
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gw >= n_list * a.F_search) return;
    const int lane = threadIdx.x & 31;
    const int f = gw % a.F_search;
    const int seg = list[gw / a.F_search];
    const SegDev sg = a.segs[seg];
    const int n = sg.n;
    const double Ty = ty[seg], Tw = tw[seg];
    const double *yp = a.ypay + (int64_t)f * a.col_stride + sg.start;
    const double *wp = a.wpay + (int64_t)f * a.col_stride + sg.start;
    double Sy = 0.0, Sw = 0.0;
    Cand best = cand_none();
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const double yv = i < n ? yp[i] : 0.0, wv = i < n ? wp[i] : 0.0;
        double my_y = 0.0, my_w = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            Sy = __dadd_rn(Sy, shfl_d(yv, j));   // np.cumsum, strictly sequential (four.py:22-23)
            Sw = __dadd_rn(Sw, shfl_d(wv, j));
            if (lane == j) {
                my_y = Sy;
                my_w = Sw;
            }
        }
        if (i < n - 1) {
            const double expected = __ddiv_rn(__dmul_rn(my_w, Ty), Tw);                     // four.py:24
            const double d = __dsub_rn(my_y, expected);
            const double sq = __dmul_rn(d, d);                                              // four.py:25
            const double ratio = __ddiv_rn(1.0, __dmul_rn(my_w, __dsub_rn(Tw, my_w)));      // four.py:26
            const double g = __dmul_rn(ratio, sq);                                          // four.py:27
            if (g > best.g) {
                best.g2 = best.g;
                best.g = g;
                best.s = my_y;
                best.dev = my_w;   // the left weight sum rides in the aux slot
                best.k = i;
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

__global__ void __launch_bounds__(256, 3) l2w_scan_stream_kernel(LevelArgs a, const TileDesc *__restrict__ tiles,
                                                                 int n_tiles, u64 *status, u32 epoch,
                                                                 Partial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char w_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sy = reinterpret_cast<double *>(w_smem) + warp * SY_PITCH;
    double *sw = reinterpret_cast<double *>(w_smem) + (8 + warp) * SY_PITCH;
    const double *my = sy + lane * 17, *mw = sw + lane * 17;   // position lane*16 + j sits at lane*17 + j
    const int64_t total_tiles = (int64_t)n_tiles * a.F_search;
    const int64_t stride_tiles = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < total_tiles; t += stride_tiles) {
        const int f = (int)(t % a.F_search), ti = (int)(t / a.F_search);
        const TileDesc td = tiles[ti];
        const SegDev sg = a.segs[td.seg];
        const int off = td.tin * WT_TILE;
        const int cnt = min(WT_TILE, sg.n - off);
        const int64_t cb = (int64_t)f * a.col_stride + sg.start + off;
#pragma unroll 4
        for (int i = 0; i < WT_ITEMS; ++i) {
            const int e = i * 32 + lane;
            sy[i * 34 + lane + (lane >> 4)] = e < cnt ? __ldcs(a.ypay + cb + e) : 0.0;
            sw[i * 34 + lane + (lane >> 4)] = e < cnt ? __ldcs(a.wpay + cb + e) : 0.0;
        }
        __syncwarp();
        double ts_y = 0.0, ts_w = 0.0;
#pragma unroll 8
        for (int j = 0; j < WT_ITEMS; ++j) {
            ts_y += my[j];
            ts_w += mw[j];
        }
        double iy = ts_y, iw = ts_w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double oy = shfl_up_d(iy, d), ow = shfl_up_d(iw, d);
            if (lane >= d) {
                iy += oy;
                iw += ow;
            }
        }
        Carry3 tile;
        tile.c = 0;
        tile.l = shfl_d(iy, 31);
        tile.r = shfl_d(iw, 31);
        const Carry3 pre = lookback3(status, (int64_t)f * n_tiles + sg.cbase + td.tin, td.tin, epoch, tile, lane);
        double Sy = pre.l + (iy - ts_y), Sw = pre.r + (iw - ts_w);
        const int n = sg.n;
        const double Ty = sg.total, Tw = sg.total_w;
        const double slope = Ty * rcp_fast1(Tw);
        const int k0 = off + lane * WT_ITEMS;
        const int nvalid = max(0, min(WT_ITEMS, min(cnt - lane * WT_ITEMS, n - 1 - k0)));
        double m1 = -CUDART_INF, m2 = -CUDART_INF, sbest = 0.0, wbest = 0.0;
        int kbest = 0x7fffffff;
#pragma unroll 4
        for (int j = 0; j < nvalid; ++j) {
            Sy += my[j];
            Sw += mw[j];
            const double dev = Sy - Sw * slope;
            const double g = dev * dev * rcp_fast1(Sw * (Tw - Sw));
            m2 = fmax(m2, fmin(m1, g));
            if (g > m1) {
                m1 = g;
                sbest = Sy;
                wbest = Sw;
                kbest = k0 + j;
            }
        }
        __syncwarp();
        Cand best;
        best.g = m1;
        best.g2 = m2;
        best.s = sbest;
        best.dev = wbest;
        best.k = kbest;
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

}  // namespace

int launch_mul(b200dt_ctx *ctx, const double *a, const double *b, int64_t n, double *out) {
    LaunchScope ls(ctx, KC_MISC, 24.0 * (double)n);
    mul_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(a, b, n, out);
    CK(cudaGetLastError());
    return 0;
}

int launch_l2w_scan(b200dt_ctx *ctx, const LevelArgs &a, const TileDesc *tiles, int n_tiles, int64_t n_elems,
                    u64 *status, u32 epoch, Partial *partials) {
    if (n_tiles <= 0) return 0;
    const int64_t total = (int64_t)n_tiles * a.F_search;
    const size_t smem = 16 * SY_PITCH * sizeof(double);
    // (per context = per device: a process may drive several GPUs)
    if (!(ctx->attr_mask & 32u)) {
        CK(cudaFuncSetAttribute(l2w_scan_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_mask |= 32u;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l2w_scan_stream_kernel, 256, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    unsigned grid = (unsigned)std::min<int64_t>((int64_t)per_sm * ctx->sm_count, (total + 7) / 8);
    LaunchScope ls(ctx, KC_L2_SCAN, 20.0 * (double)n_elems * a.F_search);
    LevelArgs a_ = a;
    void *args[] = {&a_, &tiles, &n_tiles, &status, &epoch, &partials};
    CK(launch_resident((const void *)l2w_scan_stream_kernel, grid, 256, args, smem, ctx->stream));
    return 0;
}

int launch_l2w_exact(b200dt_ctx *ctx, const LevelArgs &a, const int *small_list, int n_small, int64_t n_elems,
                     int total_col, double *ty, double *tw, Partial *partials) {
    if (n_small <= 0) return 0;
    {
        LaunchScope ls(ctx, KC_L2_EXACT, 16.0 * (double)n_elems);
        exactw_total_kernel<<<(n_small + 7) / 8, 256, 0, ctx->stream>>>(a, small_list, n_small, total_col, ty, tw);
    }
    const int64_t warps = (int64_t)n_small * a.F_search;
    LaunchScope ls(ctx, KC_L2_EXACT, 16.0 * (double)n_elems * a.F_search);
    l2w_exact_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(a, small_list, n_small, ty, tw, partials);
    CK(cudaGetLastError());
    return 0;
}
