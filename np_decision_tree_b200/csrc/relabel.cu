// Row relabelling: make the row ids of every node of one level CONTIGUOUS, once per fit.
//
// Why (measured, profiles/r02_partition_lookup_ubench.md): K4 looks up one side bit per element in an
// N-bit mask.  With the original row ids a node's rows are scattered over the whole mask, every warp
// look-up touches 32 different lines, and the kernel is bound by L1 miss handling (15-16 ms per level at
// 10 M x 256).  Once the rows of a node have consecutive ids, the bits a tile needs lie in a window of
// n/8 bytes that stays L1-resident, and the same kernel runs at the HBM roofline of the bytes it moves
// (10.4 ms).  The relabelling itself replaces the side-bit look-up of ONE level by an id look-up in a
// table (same cost), so it is free: from the next level on every look-up is local.
//
// New ids: the child that will occupy positions [start, start + n) of the order arrays gets the ids
// [start, start + n), its rows numbered in increasing OLD id.  That definition depends only on the side
// mask and the segment table, which every rank of a feature-sharded fit holds identically, so all
// ranks agree on the ids (the mask all-reduce stays meaningful) without any extra exchange.
//
// The reference has no counterpart: ids are internal (the tree stores thresholds, not rows; the one
// place that needs an original row, the threshold fetch X[row, f] of one.py:104-105, goes through
// row_map).
#include <algorithm>

#include "tree_kernels.cuh"

namespace {

constexpr int RL_CHUNK = 1024;       // rows per warp chunk
constexpr u32 RL_NONE = 0xffffu;     // row takes no part in this level's partition

// (plan segment, side) of every row of the split nodes, read off ONE column (all columns hold the
// same row sets per segment).  One warp per PT_TILE positions.
__global__ void __launch_bounds__(256) relabel_bins_kernel(const int *__restrict__ col, const SegDev *__restrict__ segs,
                                                           const TileDesc *__restrict__ tiles, int n_tiles,
                                                           const u32 *__restrict__ mask,
                                                           unsigned short *__restrict__ bin_of_row) {
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tiles; t += gridDim.x * 8) {
        const TileDesc td = tiles[t];
        const SegDev sg = segs[td.seg];
        const int off = td.tin * PT_TILE;
#pragma unroll 4
        for (int i = 0; i < PT_TILE / 32; ++i) {
            const int e = off + i * 32 + lane;
            if (e < sg.n) {
                const int row = col[sg.start + e];
                const u32 left = (mask[row >> 5] >> (row & 31)) & 1u;
                bin_of_row[row] = (unsigned short)(2 * td.seg + (left ? 0 : 1));
            }
        }
    }
}

// counts[bin][chunk]: rows of each bin inside each chunk of RL_CHUNK consecutive ids.  One warp per chunk,
// its bin counters in shared memory (one leader lane per distinct bin per step: plain read-modify-write).
__global__ void __launch_bounds__(256) relabel_count_kernel(const unsigned short *__restrict__ bin_of_row, int64_t N,
                                                            int K, int n_chunks, int *__restrict__ counts) {
    extern __shared__ int rl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *bins = rl_smem + warp * K;
    const u32 lt = lanemask_lt();
    for (int chunk = blockIdx.x * 8 + warp; chunk < n_chunks; chunk += gridDim.x * 8) {
        for (int b = lane; b < K; b += 32) bins[b] = 0;
        __syncwarp();
        const int64_t r0 = (int64_t)chunk * RL_CHUNK;
        for (int it = 0; it < RL_CHUNK / 32; ++it) {
            const int64_t r = r0 + it * 32 + lane;
            const u32 b = r < N ? (u32)bin_of_row[r] : RL_NONE;
            const u32 peers = __match_any_sync(FULL_MASK, b);
            if (b != RL_NONE && (peers & lt) == 0u) bins[b] += __popc(peers);
            __syncwarp();
        }
        for (int b = lane; b < K; b += 32) counts[(int64_t)b * n_chunks + chunk] = bins[b];
        __syncwarp();
    }
}

// per bin: exclusive scan of its chunk counts, offset by the first id of the bin.  One block per bin.
__global__ void __launch_bounds__(256) relabel_scan_kernel(int *__restrict__ counts, int n_chunks,
                                                           const int *__restrict__ bin_start) {
    __shared__ int wsum[8];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *row = counts + (int64_t)blockIdx.x * n_chunks;
    if (tid == 0) carry_s = bin_start[blockIdx.x];
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 256) {
        const int i = base + tid;
        const int c = i < n_chunks ? row[i] : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int before = carry_s;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (i < n_chunks) row[i] = before + incl - c;
        __syncthreads();
        if (tid == 255) carry_s = before + incl;
        __syncthreads();
    }
}

// new id of every participating row (stable inside its bin), and the map new id -> ORIGINAL row
__global__ void __launch_bounds__(256) relabel_assign_kernel(const unsigned short *__restrict__ bin_of_row, int64_t N,
                                                             int K, int n_chunks, const int *__restrict__ offsets,
                                                             int *__restrict__ newid, const int *__restrict__ map_old,
                                                             int *__restrict__ map_new) {
    extern __shared__ int rl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *bins = rl_smem + warp * K;
    const u32 lt = lanemask_lt();
    for (int chunk = blockIdx.x * 8 + warp; chunk < n_chunks; chunk += gridDim.x * 8) {
        for (int b = lane; b < K; b += 32) bins[b] = offsets[(int64_t)b * n_chunks + chunk];
        __syncwarp();
        const int64_t r0 = (int64_t)chunk * RL_CHUNK;
        for (int it = 0; it < RL_CHUNK / 32; ++it) {
            const int64_t r = r0 + it * 32 + lane;
            const u32 b = r < N ? (u32)bin_of_row[r] : RL_NONE;
            const u32 peers = __match_any_sync(FULL_MASK, b);
            int base = 0;
            if (b != RL_NONE) base = bins[b];
            __syncwarp();
            if (b != RL_NONE) {
                if ((peers & lt) == 0u) bins[b] = base + __popc(peers);
                const int nid = base + __popc(peers & lt);
                DCHECK(nid >= 0 && (int64_t)nid < N);
                newid[r] = nid;
                map_new[nid] = map_old ? map_old[r] : (int)r;
            }
            __syncwarp();
        }
    }
}

}  // namespace

// Computes d_newid[old id] for every row of the plan's split nodes and d_map_new[new id] = original row.
// d_bin_start: device [2 * n_segs] first id of (segment, left) / (segment, right); scratch: d_bins (N u16),
// d_counts (2 * n_segs * n_chunks ints).
int launch_relabel(b200dt_ctx *ctx, const int *col, const SegDev *d_segs, int n_segs, const TileDesc *d_tiles,
                   int n_tiles, const u32 *mask, int64_t N, const int *d_bin_start, unsigned short *d_bins,
                   int *d_counts, int *d_newid, const int *d_map_old, int *d_map_new) {
    const int K = 2 * n_segs;
    const int n_chunks = (int)((N + RL_CHUNK - 1) / RL_CHUNK);
    const size_t smem = (size_t)8 * K * sizeof(int);
    if (smem > 48 * 1024) return ctx->fail_msg("relabel: too many nodes");
    LaunchScope ls(ctx, KC_RELABEL, 0.0, 4);
    CK(cudaMemsetAsync(d_bins, 0xff, (size_t)N * 2, ctx->stream));
    const int g = std::min(ctx->sm_count * 8, (n_tiles + 7) / 8);
    relabel_bins_kernel<<<g, 256, 0, ctx->stream>>>(col, d_segs, d_tiles, n_tiles, mask, d_bins);
    const int gc = std::min(ctx->sm_count * 8, (n_chunks + 7) / 8);
    relabel_count_kernel<<<gc, 256, smem, ctx->stream>>>(d_bins, N, K, n_chunks, d_counts);
    relabel_scan_kernel<<<K, 256, 0, ctx->stream>>>(d_counts, n_chunks, d_bin_start);
    relabel_assign_kernel<<<gc, 256, smem, ctx->stream>>>(d_bins, N, K, n_chunks, d_counts, d_newid, d_map_old, d_map_new);
    CK(cudaGetLastError());
    return 0;
}

int relabel_chunks(int64_t N) { return (int)((N + RL_CHUNK - 1) / RL_CHUNK); }
