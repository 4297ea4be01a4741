// K1: per-column argsort of a row-major (N, F) float64 matrix.
//
// replaces np.argsort(X, axis=0)   (ref decision_tree/one.py:139, strategy_l1.py:147,177)
//
// Stable LSD radix sort on the order-preserving uint64 image of each key, payload = int32 row id,
// every column of a batch in the same launch.  Two paths, same result:
//
//  * 64-bit path (always correct): 8 onesweep passes of 8-bit digits over the full key.
//      prepare 16 B + passes 20 + 6*24 + 16 = 196 algorithmic bytes per element.
//  * 32-bit path (default, falls back per column): a radix pass costs the same per key whatever
//    the key width (a latency chain per tile; cutting instructions changed nothing), so the lever
//    is fewer bytes and passes.  Sort a 32-bit order-preserving IMAGE of the key (linear quantisation of the
//    value between the column's min and max; 4 passes of 4-byte keys), then one fix-up pass streams
//    the sorted images, gathers the full key only of elements whose image is shared (~1 % for smooth
//    data) and stably sorts those short runs.  min/max 8 + prepare 12 + passes 12+16+16+16 + fix-up
//    12 = 92 bytes per element.  A run longer than FIX_MAX_RUN (heavy duplicates, integer-valued,
//    constant or outlier-dominated columns) flags its column, which the 64-bit path re-sorts.
//
//   prepare : X[n][f] (coalesced 64-byte row pieces) -> keys[f][n] through a shared-memory
//             transpose, plus the digit histograms of every pass (shared-memory atomics)
//   pass p  : "onesweep": one read of (key, row id), one scattered write.  A tile's global
//             digit offsets come from a decoupled look-back over the tiles before it in the
//             same column; tiles are handed out by an atomic ticket in tile-major order.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;           // radix pass, 8-byte keys (4-byte keys: RsCfg below)
#ifndef RS_ITEMS_CFG
#define RS_ITEMS_CFG 16
#endif
#ifndef RS_THREADS4_CFG
#define RS_THREADS4_CFG 512   // threads per block of the 4-byte-key pass on wide column batches
#endif
#ifndef RS_BATCH_COLS
#define RS_BATCH_COLS 32      // columns per sort batch (B200DT_SORT_BATCH overrides): see sort_columns_device
#endif
#ifndef RS_BIG_TILE_MIN_COLS
#define RS_BIG_TILE_MIN_COLS 128
#endif
constexpr int RS_ITEMS = RS_ITEMS_CFG;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per tile
// Block shape of the radix pass.  A tile leaves as 256 digit runs, so the tile size sets the run length and divides
// the per-tile look-back and barriers: for 4-byte keys 512 threads x 16 keys (8192-key tiles, 2 blocks per SM in 64
// registers) take 52.6 ms for the four passes of 256 columns against 55.9 at 4096 (4 blocks per SM) and 66.1 at
// 16384 (one block per SM).  With few columns (a rank of a feature-sharded fit: 64 or 32) the smaller tile wins
// (12.25 vs 12.77 ms at 64 columns, 6.19 vs 6.42 at 32: fewer, longer tiles in flight per look-back chain), so the
// host picks the shape per column batch.  8-byte keys always run 256 threads, 3 blocks per SM in 80 registers.
template <typename KeyT, int THREADS_P>
struct RsCfg {
    static constexpr int THREADS = THREADS_P;
    static constexpr int MINB = sizeof(KeyT) == 4 ? 1024 / THREADS : 3;
    static constexpr int TILE = THREADS * RS_ITEMS;
    static constexpr int WARPS = THREADS / 32;
};
constexpr u32 RS_FLAG_AGG = 1u << 30;
constexpr u32 RS_FLAG_INCL = 2u << 30;
constexpr u32 RS_VAL_MASK = (1u << 30) - 1;

constexpr int PREP_COLS = 8;     // columns per prepare block: 64-byte row pieces
constexpr int PREP_ROWS = 256;   // rows per chunk
constexpr int PREP_PITCH = PREP_ROWS + 4;   // a warp stores 8 columns x 4 rows: word (col * PITCH + row) -> bank 4 col + row
// Shared-memory slot of histogram bin (column c, digit slot d, value v): [d][v][c].  A warp's 32 atomics are 8
// columns x 4 rows; with a [c][d][256] layout the bank is v % 32 whatever the column (3.5 wavefronts per atomic on
// random digits, 69 % of the kernel's shared-memory wavefronts were bank-conflict replays, ncu r02).  Interleaving the
// columns gives column c the banks c, c + 8, c + 16, c + 24, so only the 4 lanes of one column can collide, and the
// slot is one scaled add away from the digit.
__device__ __forceinline__ int prep_hist_slot(int c, int d, u32 v) {
    return (d * 256 + (int)v) * PREP_COLS + c;
}

#ifndef FIX_TILE_CFG
#define FIX_TILE_CFG 1024
#endif
constexpr int FIX_TILE = FIX_TILE_CFG;   // fix-up: positions per block
constexpr int FIX_MAX_RUN = 32;  // longest run (equal top-32 bits) repaired in place

// float64 -> uint64 whose unsigned order equals the float order.
// NaN sorts last (like numpy); -0.0 is folded onto +0.0 so the sort equals a stable argsort.
__device__ __forceinline__ u64 sortable_key(double x) {
    if (x != x) return ~0ull;
    u64 b = (u64)__double_as_longlong(x);
    if (x == 0.0) b = 0ull;
    u64 flip = (u64)((long long)b >> 63) | 0x8000000000000000ull;
    return b ^ flip;
}

// 32-bit image of a key for the fast path: LINEAR quantisation of the value between the column's
// finite minimum and maximum.  x -> (x - mn) * scale is monotone non-decreasing in fp64, so the
// image is order-preserving (ties only merge neighbours); unlike the top 32 bits of the fp64 bit
// pattern (20 mantissa bits per binade) it spreads a smooth distribution over all 2^32 slots, so
// for 10 M standard-normal values ~1 % of the elements share a slot instead of ~70 %.
struct ColRange {
    double mn;      // smallest finite value of the column
    double scale;   // (2^32 - 2) / (max - min), or 0 when the column is constant / not finite
};

__device__ __forceinline__ u32 quantise(double x, const ColRange r) {
    if (x != x) return 0xffffffffu;                       // NaN last, like the full key
    double q = (x - r.mn) * r.scale;                      // -inf -> -inf, +inf -> +inf (scale > 0)
    if (!(q > 0.0)) q = 0.0;                              // also catches 0 * inf = NaN when scale == 0
    if (q > 4294967294.0) q = 4294967294.0;
    return (u32)__double2ull_rz(q);
}

// per column: min / max over the finite values, as order-preserving u64 keys (atomicMin/Max)
// `sample` > 1: only every sample-th group of 32 rows is read (the range is then a sampled one, see col_range_kernel)
__global__ void __launch_bounds__(256) col_minmax_kernel(const double *__restrict__ X, int64_t N, int64_t ldx,
                                                         int fb, u64 *__restrict__ kmin, u64 *__restrict__ kmax,
                                                         int sample) {
    __shared__ u64 smin[256], smax[256];
    const int tid = threadIdx.x;
    const int cg = blockIdx.y;
    const int ncols = min(PREP_COLS, fb - cg * PREP_COLS);
    const int tcol = tid & (PREP_COLS - 1), trow = tid >> 3;
    u64 lo = ~0ull, hi = 0ull;
    // four rows in flight per thread (one load per iteration left the memory system idle)
    const int64_t step = (int64_t)gridDim.x * 32 * sample;
    if (tcol < ncols) {
        const double *base = X + cg * PREP_COLS + tcol;
        for (int64_t r = (int64_t)blockIdx.x * 32 * sample + trow; r < N; r += 4 * step) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = (r + u * step < N) ? base[(r + u * step) * ldx] : CUDART_NAN;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (x[u] == x[u] && fabs(x[u]) != CUDART_INF) {
                    const u64 k = sortable_key(x[u]);
                    lo = k < lo ? k : lo;
                    hi = k > hi ? k : hi;
                }
            }
        }
    }
    smin[tid] = lo;
    smax[tid] = hi;
    __syncthreads();
    for (int s = 128; s >= PREP_COLS; s >>= 1) {   // threads with equal (tid & 7) hold the same column
        if (tid < s) {
            smin[tid] = smin[tid + s] < smin[tid] ? smin[tid + s] : smin[tid];
            smax[tid] = smax[tid + s] > smax[tid] ? smax[tid + s] : smax[tid];
        }
        __syncthreads();
    }
    if (tid < ncols) {
        atomicMin((unsigned long long *)&kmin[cg * PREP_COLS + tid], (unsigned long long)smin[tid]);
        atomicMax((unsigned long long *)&kmax[cg * PREP_COLS + tid], (unsigned long long)smax[tid]);
    }
}

__device__ __forceinline__ double key_to_double(u64 k) {
    const u64 b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double((long long)b);
}

// widen > 0: the min / max come from a row SAMPLE, so the range is stretched by `widen` of its span on both
// sides.  Values beyond it clamp to the first / last slot (still monotone); they tie there and the fix-up
// pass orders them by the full key, or sends the column to the 64-bit path if there are more than it repairs.
__global__ void col_range_kernel(const u64 *__restrict__ kmin, const u64 *__restrict__ kmax, int fb,
                                 ColRange *__restrict__ out, double widen) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= fb) return;
    ColRange r;
    r.mn = 0.0;
    r.scale = 0.0;
    if (kmin[f] <= kmax[f]) {
        double mn = key_to_double(kmin[f]), mx = key_to_double(kmax[f]);
        const double pad = widen * (mx - mn);
        if (pad > 0.0 && pad < CUDART_INF) {
            mn -= pad;
            mx += pad;
        }
        const double span = mx - mn;
        r.mn = mn;
        if (span > 0.0 && span < CUDART_INF) r.scale = 4294967294.0 / span;
    }
    out[f] = r;
}

template <typename KeyT>
__device__ __forceinline__ KeyT narrow_key(u64 k);
template <>
__device__ __forceinline__ u64 narrow_key<u64>(u64 k) { return k; }
template <>
__device__ __forceinline__ u32 narrow_key<u32>(u64 k) { return (u32)(k >> 32); }

// histograms: ghist[col][8][256] (only the first sizeof(KeyT) digit slots are used)
template <typename KeyT>
__global__ void __launch_bounds__(256) sort_prepare_kernel(const double *__restrict__ X, int64_t N,
                                                           int64_t ldx, int fb, KeyT *__restrict__ keys,
                                                           int64_t kstride, u32 *__restrict__ ghist,
                                                           const ColRange *__restrict__ ranges) {
    constexpr int NPASS = (int)sizeof(KeyT);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *hist = reinterpret_cast<u32 *>(smem_raw);                                  // [8 cols][NPASS][256]
    KeyT *tile = reinterpret_cast<KeyT *>(smem_raw + PREP_COLS * NPASS * 256 * 4);  // [8][PREP_PITCH]
    const int tid = threadIdx.x;
    const int cg = blockIdx.y;
    const int ncols = min(PREP_COLS, fb - cg * PREP_COLS);
    for (int i = tid; i < PREP_COLS * NPASS * 256; i += 256) hist[i] = 0;
    __syncthreads();
    const int tcol = tid & (PREP_COLS - 1);
    const int trow = tid >> 3;  // 32 rows per sweep
    ColRange rng;
    rng.mn = 0.0;
    rng.scale = 0.0;
    if (sizeof(KeyT) == 4 && tcol < ncols) rng = ranges[cg * PREP_COLS + tcol];
    // software pipeline: the loads of the block's NEXT chunk are in flight while this one is quantised, counted and
    // written (ncu r02: 22 % of the stall samples sat on the first use of the loaded values, 10 % on the barriers).
    // Two register buffers take turns (no copies); once pipelined the kernel is issue-bound (78 %), so the
    // per-key work is kept short: one shift + one scaled add per histogram slot, no bounds tests in full chunks.
    const bool active = tcol < ncols;
    const double *xcol = X + cg * PREP_COLS + tcol;
    u32 *hmine = hist + tcol;
    KeyT *tmine = tile + tcol * PREP_PITCH + trow;
    auto load_chunk = [&](int64_t chunk, double (&dst)[PREP_ROWS / 32]) {
        const int64_t r0 = chunk * PREP_ROWS;
        const double *px = xcol + (r0 + trow) * ldx;
        if (active && r0 + PREP_ROWS <= N) {
#pragma unroll
            for (int it = 0; it < PREP_ROWS / 32; ++it) dst[it] = px[(int64_t)it * 32 * ldx];
        } else {
#pragma unroll
            for (int it = 0; it < PREP_ROWS / 32; ++it)
                dst[it] = (active && r0 + it * 32 + trow < N) ? px[(int64_t)it * 32 * ldx] : 0.0;
        }
    };
    auto process_chunk = [&](int64_t chunk, const double (&src)[PREP_ROWS / 32]) {
        const int64_t r0 = chunk * PREP_ROWS;
        const int rows_here = (int)min((int64_t)PREP_ROWS, N - r0);
        if (active) {
#pragma unroll
            for (int it = 0; it < PREP_ROWS / 32; ++it) {
                if (rows_here == PREP_ROWS || it * 32 + trow < rows_here) {
                    const KeyT k = sizeof(KeyT) == 4 ? (KeyT)quantise(src[it], rng) : narrow_key<KeyT>(sortable_key(src[it]));
                    tmine[it * 32] = k;
#pragma unroll
                    for (int d = 0; d < NPASS; ++d)
                        atomicAdd(hmine + d * (PREP_COLS * 256) + (int)((u32)(k >> (8 * d)) & 255u) * PREP_COLS, 1u);
                }
            }
        }
        __syncthreads();
        if (tid < rows_here) {
            KeyT *ko = keys + (int64_t)cg * PREP_COLS * kstride + r0 + tid;
            const KeyT *ti = tile + tid;
            if (ncols == PREP_COLS) {
#pragma unroll
                for (int j = 0; j < PREP_COLS; ++j) ko[(int64_t)j * kstride] = ti[j * PREP_PITCH];
            } else {
                for (int j = 0; j < ncols; ++j) ko[(int64_t)j * kstride] = ti[j * PREP_PITCH];
            }
        }
        __syncthreads();
    };
    static_assert(PREP_ROWS == 256, "one output row per thread");
    double xa[PREP_ROWS / 32], xb[PREP_ROWS / 32];
    const int64_t n_chunks = (N + PREP_ROWS - 1) / PREP_ROWS;
    int64_t chunk = blockIdx.x;
    if (chunk < n_chunks) load_chunk(chunk, xa);
    while (chunk < n_chunks) {
        if (chunk + gridDim.x < n_chunks) load_chunk(chunk + gridDim.x, xb);
        process_chunk(chunk, xa);
        chunk += gridDim.x;
        if (chunk >= n_chunks) break;
        if (chunk + gridDim.x < n_chunks) load_chunk(chunk + gridDim.x, xa);
        process_chunk(chunk, xb);
        chunk += gridDim.x;
    }
    for (int i = tid; i < ncols * NPASS * 256; i += 256) {
        const int col = i / (NPASS * 256), rest = i % (NPASS * 256);
        const u32 c = hist[prep_hist_slot(col, rest >> 8, (u32)(rest & 255))];
        if (c) atomicAdd(&ghist[((int64_t)cg * PREP_COLS + col) * 8 * 256 + rest], c);
    }
}

// exclusive scan of each 256-bin histogram: one block per (column, digit)
__global__ void __launch_bounds__(256) sort_scan_hist_kernel(const u32 *__restrict__ ghist,
                                                             u32 *__restrict__ gbase) {
    __shared__ u32 wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 c = ghist[(int64_t)blockIdx.x * 256 + tid];
    u32 incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    gbase[(int64_t)blockIdx.x * 256 + tid] = base + incl - c;
}

template <typename KeyT, int THREADS_P>
struct RsSmem {
    u32 warp_hist[RsCfg<KeyT, THREADS_P>::WARPS][256];
    u32 gofs[256];
    u32 warp_sums[8];
    u32 ticket;
    u32 pad[3];
    KeyT keys[RsCfg<KeyT, THREADS_P>::TILE];
    int vals[RsCfg<KeyT, THREADS_P>::TILE];
};

template <typename KeyT>
__device__ __forceinline__ KeyT ld_stream_key(const KeyT *p);
template <>
__device__ __forceinline__ u64 ld_stream_key<u64>(const u64 *p) { return ld_stream_u64(p); }
template <>
__device__ __forceinline__ u32 ld_stream_key<u32>(const u32 *p) { return (u32)ld_stream_i32((const int *)p); }

#define RS_LOAD_VALS()                                                                                              \
    do {                                                                                                            \
        if (full) {                                                                                                 \
            _Pragma("unroll") for (int i = 0; i < RS_ITEMS; ++i)                                                    \
                val[i] = FIRST ? (int)base + o0 + i * 32 : ld_stream_i32(vin + o0 + i * 32);                        \
        } else {                                                                                                    \
            _Pragma("unroll") for (int i = 0; i < RS_ITEMS; ++i)                                                    \
                val[i] = FIRST ? (int)base + o0 + i * 32 : (o0 + i * 32 < count ? ld_stream_i32(vin + o0 + i * 32) : 0); \
        }                                                                                                           \
    } while (0)

template <typename KeyT, bool FIRST, bool LAST, int THREADS_P>
__global__ void __launch_bounds__(RsCfg<KeyT, THREADS_P>::THREADS, RsCfg<KeyT, THREADS_P>::MINB) radix_pass_kernel(
    const KeyT *__restrict__ keys_in, const int *__restrict__ vals_in, KeyT *__restrict__ keys_out,
    int *__restrict__ vals_out, const u32 *__restrict__ gbase, u32 *status, u32 *ticket_ctr, int64_t N,
    int64_t kstride, int64_t vstride, int tiles_per_col, int n_cols, int pass) {
    using Cfg = RsCfg<KeyT, THREADS_P>;
    constexpr int THREADS_ = Cfg::THREADS, TILE_ = Cfg::TILE, WARPS_ = Cfg::WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem<KeyT, THREADS_P> &sm = *reinterpret_cast<RsSmem<KeyT, THREADS_P> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sm.ticket = atomicAdd(ticket_ctr, 1u);
    for (int i = tid; i < WARPS_ * 256; i += THREADS_) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const u32 t = sm.ticket;
    // tile-major order: consecutive tickets are the same tile of different columns, so the
    // blocks in flight spread over all columns and every look-back chain stays short
    const int col = t % n_cols, tile = t / n_cols;
    const int64_t base = (int64_t)tile * TILE_;
    const int count = (int)min((int64_t)TILE_, N - base);
    const bool full = count == TILE_;
    const int shift = pass * 8;
    // everything below indexes with 32-bit tile-local offsets from these bases
    const KeyT *kin = keys_in + (int64_t)col * kstride + base;
    const int *vin = FIRST ? nullptr : vals_in + (int64_t)col * vstride + base;
    const int o0 = warp * (32 * RS_ITEMS) + lane;  // this lane's item 0; item i is o0 + 32 i

    KeyT key[RS_ITEMS];
    int val[RS_ITEMS];
    if (full) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) key[i] = ld_stream_key<KeyT>(kin + o0 + i * 32);
    } else {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i)
            key[i] = o0 + i * 32 < count ? ld_stream_key<KeyT>(kin + o0 + i * 32) : (KeyT)~(KeyT)0;
    }
    // stable rank of each key among the keys of its warp with the same digit.  All peer masks
    // first (independent ballots pipeline), then, item by item, one shared atomic per distinct
    // digit by its lowest lane, whose old count is broadcast to the peers; items are issued in
    // order, so the per-digit counts accumulate in the stable order.
    const u32 lt = lanemask_lt();
    u32 rank[RS_ITEMS];
#define RS_DIGIT(i) ((u32)(key[i] >> shift) & 255u)
    // Peer masks from 8 ballots per item, NOT match.any: measured on B200 (scripts/ubench/rank_ubench.cu) one
    // warp-wide MATCH.ANY costs ~61 SM cycles -- 150 G keys/s chip-wide, i.e. 17 ms of the 22 ms this pass took --
    // while the ballot form ranks 330 G keys/s.
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const u32 d = RS_DIGIT(i);
        u32 peers = full ? FULL_MASK : __ballot_sync(FULL_MASK, o0 + i * 32 < count);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            // peers &= bit ? v : ~v in four instructions (LOP3 -> P, VOTE, @!P LOP3, LOP3); the C++ form of the same
            // line compiled to seven (the predicate was derived twice)
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                ".reg .b32 t;\n\t"
                "and.b32 t, %1, %2;\n\t"
                "setp.ne.u32 p, t, 0;\n\t"
                "vote.sync.ballot.b32 t, p, 0xffffffff;\n\t"
                "@!p not.b32 t, t;\n\t"
                "and.b32 %0, %0, t;\n\t"
                "}" : "+r"(peers) : "r"(d), "r"(1u << b));
        }
        rank[i] = peers;
    }
    u32 *my_hist = sm.warp_hist[warp];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const u32 peers = rank[i];
        const u32 below = peers & lt;
        u32 old = 0;
        if (below == 0u && ((peers >> lane) & 1u)) old = atomicAdd(&my_hist[RS_DIGIT(i)], (u32)__popc(peers));
        old = __shfl_sync(FULL_MASK, old, __ffs(peers) - 1);
        rank[i] = old + __popc(below);
    }
    // the row ids are not needed before the reorder: requested here they travel during the look-back, and the
    // ranking above runs in 64 registers = 4 blocks per SM (56.0 instead of 62.2 ms for the four 4-byte passes;
    // an L2 prefetch of these lines at the top of the kernel changes nothing)
    RS_LOAD_VALS();
    __syncthreads();
    if (tid < 256) {
        // thread `tid` owns digit `tid`: exclusive offsets of the warps, tile total
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS_; ++w) {
            const u32 c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += c;
        }
        const u32 total = run;
        // decoupled look-back along the tiles of this column, one chain per digit
        u32 *st = status + ((int64_t)col * tiles_per_col) * 256 + tid;
        u32 excl = 0;
        if (tile == 0) {
            st_relaxed_u32(st, RS_FLAG_INCL | total);
        } else {
            st_relaxed_u32(st + tile * 256, RS_FLAG_AGG | total);
            int look = tile - 1;
            while (true) {
                const u32 sv = ld_relaxed_u32(st + look * 256);
                const u32 fl = sv & ~RS_VAL_MASK;
                if (fl == 0) continue;
                excl += sv & RS_VAL_MASK;
                if (fl == RS_FLAG_INCL) break;
                --look;
            }
            st_relaxed_u32(st + tile * 256, RS_FLAG_INCL | (excl + total));
        }
        // tile-local start of each digit's run (exclusive scan over the 256 digits by 8 warps)
        u32 incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        asm volatile("bar.sync 1, 256;");  // only the 8 digit warps
        u32 wofs = 0;
        for (int w = 0; w < warp; ++w) wofs += sm.warp_sums[w];
        const u32 bstart = wofs + incl - total;
        // fold the digit's start into the warp offsets: a key's slot is warp_hist[w][d] + rank
#pragma unroll
        for (int w = 0; w < WARPS_; ++w) sm.warp_hist[w][tid] += bstart;
        sm.gofs[tid] = gbase[((int64_t)col * 8 + pass) * 256 + tid] + excl - bstart;  // mod 2^32
    }
    __syncthreads();
    // reorder inside the tile so that each digit's keys leave as one contiguous run
    if (full) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const u32 pos = my_hist[RS_DIGIT(i)] + rank[i];
            DCHECK(pos < (u32)TILE_);
            sm.keys[pos] = key[i];
            sm.vals[pos] = val[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            if (o0 + i * 32 < count) {
                const u32 pos = my_hist[RS_DIGIT(i)] + rank[i];
                sm.keys[pos] = key[i];
                sm.vals[pos] = val[i];
            }
        }
    }
    __syncthreads();
    KeyT *kout = LAST ? nullptr : keys_out + (int64_t)col * kstride;
    int *vout = vals_out + (int64_t)col * vstride;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int pos = j * THREADS_ + tid;
        if (full || pos < count) {
            const KeyT k = sm.keys[pos];
            const u32 dest = sm.gofs[(u32)(k >> shift) & 255u] + (u32)pos;
            DCHECK((int64_t)dest < N);
            if (!LAST) kout[dest] = k;
            vout[dest] = sm.vals[pos];
        }
    }
}

// Fix-up of the 32-bit path.  in[f][j] is sorted (stably) by the 32-bit image q[f][j].  A block
// takes FIX_TILE positions plus a halo; elements whose image is shared with a neighbour gather
// their full key (one 32-byte sector; ~1 % of the elements for smooth data), and the thread at the
// start of each run insertion-sorts it (stable) by the full key.  A position is written by the tile
// that owns the START of its run; runs longer than FIX_MAX_RUN flag the column (its output is then
// discarded and the column re-sorted by the 64-bit path).
// y / ypay (optional): the fit's y payload, ypay[f][j] = y[row at position j], leaves in the same pass (a separate
// gather pass re-reads every row id).
// Elements whose image is unique (~99 % for smooth data) are written as soon as their neighbours have been seen;
// only the tied ones take the key gather / insertion sort / run-ownership path.
__global__ void __launch_bounds__(256) sort_fixup_kernel(const double *__restrict__ X, int64_t ldx, int64_t N,
                                                         const u32 *__restrict__ q, int64_t qstride,
                                                         const int *__restrict__ in, int *__restrict__ out,
                                                         int64_t col_stride, int *__restrict__ col_flags,
                                                         const double *__restrict__ y, double *__restrict__ ypay) {
    __shared__ u64 skey[FIX_TILE + FIX_MAX_RUN + 1];
    __shared__ u32 sq[FIX_TILE + FIX_MAX_RUN + 2];
    __shared__ int srow[FIX_TILE + FIX_MAX_RUN + 1];
    __shared__ int any_tied;
    const int f = blockIdx.y;
    const int64_t a = (int64_t)blockIdx.x * FIX_TILE;   // first position this tile owns
    const int *col = in + (int64_t)f * col_stride;
    const u32 *qc = q + (int64_t)f * qstride;
    int *ocol = out + (int64_t)f * col_stride;
    double *ycol = ypay ? ypay + (int64_t)f * col_stride : nullptr;
    // slot s <-> position a - 1 + s  (slot 0 = left neighbour, only used to recognise a run that
    // started in an earlier tile)
    const int nslots = (int)min((int64_t)(FIX_TILE + FIX_MAX_RUN + 1), N - a + 1);
    const int first = a == 0 ? 1 : 0;   // lowest slot holding a real element
    const int owned = (int)min((int64_t)FIX_TILE, N - a);
    if (threadIdx.x == 0) any_tied = 0;
    for (int s = first + threadIdx.x; s < nslots; s += 256) {
        srow[s] = ld_stream_i32(col + a - 1 + s);
        sq[s] = (u32)ld_stream_i32((const int *)qc + a - 1 + s);
    }
    __syncthreads();
    bool mine_tied = false;
    for (int s = first + threadIdx.x; s < nslots; s += 256) {
        const u32 v = sq[s];
        const bool tied = (s > first && sq[s - 1] == v) || (s + 1 < nslots && sq[s + 1] == v);
        if (tied) {
            skey[s] = sortable_key(__ldg(X + (int64_t)srow[s] * ldx + f));
            mine_tied = true;
        } else if (s >= 1 && s <= owned) {
            // unique image: the element is in its final place
            const int row = srow[s];
            __stcs(ocol + a - 1 + s, row);
            if (ycol) __stcs(ycol + a - 1 + s, __ldg(y + row));
        }
    }
    if (mine_tied) any_tied = 1;
    __syncthreads();
    if (!any_tied) return;
    for (int s = 1 + threadIdx.x; s <= owned; s += 256) {
        const u32 v = sq[s];
        if (s > first && sq[s - 1] == v) continue;   // not the start of a run
        int len = 1;
        while (s + len < nslots && len <= FIX_MAX_RUN && sq[s + len] == v) ++len;
        if (len == 1) continue;
        if (len > FIX_MAX_RUN) {
            col_flags[f] = 1;   // too long to repair here: this column takes the 64-bit path
            continue;
        }
        // stable insertion sort of slots [s, s+len) by the full key (equal keys are already in row
        // order, the radix passes being stable)
        for (int p = 1; p < len; ++p) {
            const u64 k = skey[s + p];
            const int r = srow[s + p];
            int t = p;
            while (t > 0 && skey[s + t - 1] > k) {
                skey[s + t] = skey[s + t - 1];
                srow[s + t] = srow[s + t - 1];
                --t;
            }
            skey[s + t] = k;
            srow[s + t] = r;
        }
    }
    __syncthreads();
    for (int s = 1 + threadIdx.x; s < nslots; s += 256) {
        const u32 v = sq[s];
        const bool tied = (s > first && sq[s - 1] == v) || (s + 1 < nslots && sq[s + 1] == v);
        if (!tied) continue;
        int b = s, steps = 0;
        while (b > first && sq[b - 1] == v && steps <= FIX_MAX_RUN) {
            --b;
            ++steps;
        }
        // b == 0: the run began in an earlier tile; b > owned: it begins in the next one
        if (steps <= FIX_MAX_RUN && b >= 1 && b <= owned) {
            const int row = srow[s];
            ocol[a - 1 + s] = row;
            if (ycol) ycol[a - 1 + s] = __ldg(y + row);
        }
    }
}

template <typename KeyT, int THREADS_P>
int set_pass_attributes(b200dt_ctx *ctx) {
    const int pass_smem = (int)sizeof(RsSmem<KeyT, THREADS_P>);
    CK(cudaFuncSetAttribute(radix_pass_kernel<KeyT, true, false, THREADS_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, pass_smem));
    CK(cudaFuncSetAttribute(radix_pass_kernel<KeyT, false, false, THREADS_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, pass_smem));
    CK(cudaFuncSetAttribute(radix_pass_kernel<KeyT, false, true, THREADS_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, pass_smem));
    return 0;
}

// the passes alternate v0, v1, ...; the last one writes v1 (and, with keep_keys, the sorted keys into keys0).
// `status` is sized for 4096-key tiles, the smallest shape.
template <typename KeyT, int THREADS_P>
int run_radix_passes(b200dt_ctx *ctx, int64_t N, int fb, KeyT *keys0, KeyT *keys1, int *v0, int *v1, int64_t col_stride,
                     const u32 *gbase, u32 *status, bool keep_keys) {
    using Cfg = RsCfg<KeyT, THREADS_P>;
    constexpr int NPASS = (int)sizeof(KeyT);
    cudaStream_t s = ctx->stream;
    const size_t pass_smem = sizeof(RsSmem<KeyT, THREADS_P>);
    int *v[2] = {v0, v1};
    KeyT *k[2] = {keys0, keys1};
    const int tiles = (int)((N + Cfg::TILE - 1) / Cfg::TILE);
    u32 *ticket = status + (size_t)fb * tiles * 256;
    for (int p = 0; p < NPASS; ++p) {
        CK(cudaMemsetAsync(status, 0, (size_t)fb * tiles * 256 * 4 + 16, s));
        const unsigned grid = (unsigned)((int64_t)fb * tiles);
        const KeyT *kin = k[p & 1];
        KeyT *kout = k[(p + 1) & 1];
        const int *vin = v[(p + 1) & 1];
        int *vout = v[p & 1];
        const double K = (double)sizeof(KeyT);
        const bool last_plain = p == NPASS - 1 && !keep_keys;
        const double bytes = (p == 0 ? 2 * K + 4 : last_plain ? K + 8 : 2 * K + 8) * (double)N * fb;
        LaunchScope ls(ctx, KC_SORT_PASS, bytes);
        if (p == 0)
            radix_pass_kernel<KeyT, true, false, THREADS_P><<<grid, Cfg::THREADS, pass_smem, s>>>(
                kin, nullptr, kout, vout, gbase, status, ticket, N, N, col_stride, tiles, fb, p);
        else if (last_plain)
            radix_pass_kernel<KeyT, false, true, THREADS_P><<<grid, Cfg::THREADS, pass_smem, s>>>(
                kin, vin, nullptr, vout, gbase, status, ticket, N, N, col_stride, tiles, fb, p);
        else
            radix_pass_kernel<KeyT, false, false, THREADS_P><<<grid, Cfg::THREADS, pass_smem, s>>>(
                kin, vin, kout, vout, gbase, status, ticket, N, N, col_stride, tiles, fb, p);
    }
    CK(cudaGetLastError());
    return 0;
}

template <typename KeyT>
int run_radix_sort(b200dt_ctx *ctx, const double *dX, int64_t N, int fb, int64_t ldx, KeyT *keys0, KeyT *keys1,
                   int *v0, int *v1, int64_t col_stride, u32 *ghist, u32 *gbase, u32 *status, int tiles,
                   const ColRange *ranges, bool keep_keys) {
    // the passes alternate v0, v1, ...; the last one writes v1 (and, with keep_keys, the sorted
    // keys into keys0)
    constexpr int NPASS = (int)sizeof(KeyT);
    cudaStream_t s = ctx->stream;
    const int groups = (fb + PREP_COLS - 1) / PREP_COLS;
    const size_t prep_smem = PREP_COLS * NPASS * 256 * 4 + PREP_COLS * PREP_PITCH * sizeof(KeyT);
    // (per context = per device: a process may drive several GPUs)
    const u32 attr_bit = sizeof(KeyT) == 4 ? 16u : 64u;   // this function is instantiated per key type
    if (!(ctx->attr_mask & attr_bit)) {
        CK(cudaFuncSetAttribute(sort_prepare_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
        CKR((set_pass_attributes<KeyT, 256>(ctx)));
        if constexpr (sizeof(KeyT) == 4) CKR((set_pass_attributes<KeyT, RS_THREADS4_CFG>(ctx)));
        ctx->attr_mask |= attr_bit;
    }
    CK(cudaMemsetAsync(ghist, 0, (size_t)groups * PREP_COLS * 8 * 256 * 4, s));
    {
        int64_t chunks = (N + PREP_ROWS - 1) / PREP_ROWS;
        int64_t rb = (3ll * ctx->sm_count * 2 + groups - 1) / groups;
        if (rb > chunks) rb = chunks;
        if (rb < 1) rb = 1;
        LaunchScope ls(ctx, KC_SORT_PREPARE, (8.0 + sizeof(KeyT)) * (double)N * fb);
        sort_prepare_kernel<KeyT><<<dim3((unsigned)rb, groups), 256, prep_smem, s>>>(dX, N, ldx, fb, keys0, N, ghist, ranges);
    }
    {
        LaunchScope ls(ctx, KC_MISC, 0.0);
        sort_scan_hist_kernel<<<fb * 8, 256, 0, s>>>(ghist, gbase);
    }
    CK(cudaGetLastError());
    // 4-byte keys: big tiles when the batch has enough columns to keep every look-back chain short
    if constexpr (sizeof(KeyT) == 4) {
        if (fb >= RS_BIG_TILE_MIN_COLS)
            return run_radix_passes<KeyT, RS_THREADS4_CFG>(ctx, N, fb, keys0, keys1, v0, v1, col_stride, gbase, status,
                                                           keep_keys);
    }
    return run_radix_passes<KeyT, 256>(ctx, N, fb, keys0, keys1, v0, v1, col_stride, gbase, status, keep_keys);
}

}  // namespace

// Sort every column of dX (device, row-major, row stride ldx).  Result: d_out[f*col_stride + j]
// = row id of the j-th smallest value of column f.  d_alt is scratch of the same shape.
// ext_keys0/1 (optional): caller-provided key double-buffer of ext_key_capacity uint64 each (the
// fit lends its y-payload buffers, which are idle during the sort), instead of own scratch.
int launch_gather_payload(b200dt_ctx *ctx, const int *orders, const double *y, double *ypay, int64_t col_stride,
                          int F_store, int64_t N);

// y / ypay_out (optional): also produce the y payload ypay_out[f][j] = y[d_out[f][j]] (fused into the fix-up pass).
int sort_columns_device(b200dt_ctx *ctx, const double *dX, int64_t N, int64_t F, int64_t ldx,
                        int *d_out, int *d_alt, int64_t col_stride, u64 *ext_keys0, u64 *ext_keys1,
                        int64_t ext_key_capacity, const double *y, double *ypay_out) {
    if (N <= 0 || F <= 0) return 0;
    if (N > (int64_t)RS_VAL_MASK) return ctx->fail_msg("argsort: N must be < 2^30");
    const int tiles = (int)((N + RS_TILE - 1) / RS_TILE);
    int64_t fb_max;
    const bool ext = ext_keys0 && ext_keys1 && ext_key_capacity >= (int64_t)PREP_COLS * N;
    if (ext) {
        fb_max = ext_key_capacity / N;
    } else {
        // column batch: bound the key double-buffer to a fraction of the free memory
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = free_b / 3;
        const size_t cap = (size_t)24 << 30;
        if (budget > cap) budget = cap;
        fb_max = (int64_t)(budget / (16ull * (size_t)N));
    }
    {
        // columns per sort batch: a pass over fewer columns keeps fewer scatter streams open (256 per column) and
        // is faster per column -- four passes over 256 columns take 52.5 ms in batches of 128 (8192-key tiles),
        // 52.7 in batches of 64, 49.0 in batches of 32, 48.9 / 49.6 in batches of 24 / 16 (where prepare loses more
        // than the passes gain).  Ranks of a feature-sharded fit and host-resident inputs already came in 32s.
        int64_t cap = RS_BATCH_COLS;
        if (const char *e = getenv("B200DT_SORT_BATCH")) cap = std::max<int64_t>(PREP_COLS, atoll(e));
        if (fb_max > cap) fb_max = cap;
    }
    fb_max = (fb_max / PREP_COLS) * PREP_COLS;
    if (fb_max < PREP_COLS) fb_max = PREP_COLS;
    int64_t Fb = F < fb_max ? ((F + PREP_COLS - 1) / PREP_COLS) * PREP_COLS : fb_max;
    u64 *keys0 = ext_keys0, *keys1 = ext_keys1;
    u32 *hist, *status;
    if (!ext || Fb * N > ext_key_capacity) {
        CKR(scratch_get(ctx, SB_KEYS0, (size_t)Fb * N * 8, (void **)&keys0));
        CKR(scratch_get(ctx, SB_KEYS1, (size_t)Fb * N * 8, (void **)&keys1));
    }
    CKR(scratch_get(ctx, SB_SORT_HIST, (size_t)Fb * 8 * 256 * 4 * 2 + (size_t)F * 4 + 64, (void **)&hist));
    const size_t status_bytes = (size_t)Fb * tiles * 256 * 4 + 64;
    CKR(scratch_get(ctx, SB_SORT_STATUS, status_bytes, (void **)&status));
    u32 *ghist = hist, *gbase = hist + (size_t)Fb * 8 * 256;
    int *flags = (int *)(gbase + (size_t)Fb * 8 * 256);

    const char *e64 = getenv("B200DT_SORT64");   // read per call, like every other switch
    const bool use32 = !(e64 && e64[0] == '1');
    if (use32) {
        // ---- 32-bit path: radix sort of the 32-bit images into d_alt, fix-up into d_out ------------
        CK(cudaMemsetAsync(flags, 0, (size_t)F * 4, ctx->stream));
        u64 *kmin, *kmax;
        ColRange *ranges;
        CKR(scratch_get(ctx, SB_SORT_RANGE, (size_t)Fb * (8 + 8 + sizeof(ColRange)) + 64, (void **)&kmin));
        kmax = kmin + Fb;
        ranges = (ColRange *)(kmax + Fb);
        for (int64_t c0 = 0; c0 < F; c0 += Fb) {
            const int fb = (int)(F - c0 < Fb ? F - c0 : Fb);
            const int groups = (fb + PREP_COLS - 1) / PREP_COLS;
            CK(cudaMemsetAsync(kmin, 0xff, (size_t)Fb * 8, ctx->stream));
            CK(cudaMemsetAsync(kmax, 0x00, (size_t)Fb * 8, ctx->stream));
            {
                // big columns: quantise between the min / max of a 1/32 row sample, stretched by 25 % on both
                // sides (a full pass over X just for the range cost 7 ms of the C4 fit); the result is the exact
                // stable order either way (fix-up / 64-bit fall-back)
                const int sample = N >= (1 << 20) ? 32 : 1;
                const int64_t groups32 = (N + 32 * (int64_t)sample - 1) / (32 * (int64_t)sample);
                int64_t rb = std::min<int64_t>(groups32, std::max<int64_t>(1, 6ll * ctx->sm_count / groups));
                LaunchScope ls(ctx, KC_SORT_PREPARE, 8.0 * (double)N * fb / sample, 2);
                col_minmax_kernel<<<dim3((unsigned)rb, groups), 256, 0, ctx->stream>>>(dX + c0, N, ldx, fb, kmin, kmax,
                                                                                       sample);
                col_range_kernel<<<(fb + 127) / 128, 128, 0, ctx->stream>>>(kmin, kmax, fb, ranges,
                                                                             sample > 1 ? 0.25 : 0.0);
            }
            // 4 passes: keys0 -> keys1 -> keys0 -> keys1 -> keys0 (sorted images), rows end in d_alt
            CKR(run_radix_sort<u32>(ctx, dX + c0, N, fb, ldx, (u32 *)keys0, (u32 *)keys1, d_out + c0 * col_stride,
                                    d_alt + c0 * col_stride, col_stride, ghist, gbase, status, tiles, ranges, true));
            LaunchScope ls(ctx, KC_SORT_FIXUP, (ypay_out ? 28.0 : 12.0) * (double)N * fb);
            sort_fixup_kernel<<<dim3((unsigned)((N + FIX_TILE - 1) / FIX_TILE), (unsigned)fb), 256, 0, ctx->stream>>>(
                dX + c0, ldx, N, (const u32 *)keys0, N, d_alt + c0 * col_stride, d_out + c0 * col_stride, col_stride,
                flags + c0, y, ypay_out ? ypay_out + c0 * col_stride : nullptr);
        }
        CK(cudaGetLastError());
        // columns with a long run of equal 32-bit images take the 64-bit path
        int *h_flags;
        CKR(pinned_get(ctx, 5, (size_t)F * 4, (void **)&h_flags));
        CK(cudaMemcpyAsync(h_flags, flags, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<int> redo;
        for (int64_t f = 0; f < F; ++f)
            if (h_flags[f]) redo.push_back((int)f);
        size_t i = 0;
        while (i < redo.size()) {  // contiguous ranges of flagged columns are re-sorted together
            size_t j = i;
            while (j + 1 < redo.size() && redo[j + 1] == redo[j] + 1 && (int64_t)(j + 1 - i) < Fb) ++j;
            const int c0 = redo[i], fb = (int)(j - i + 1);
            CKR(run_radix_sort<u64>(ctx, dX + c0, N, fb, ldx, keys0, keys1, d_alt + (int64_t)c0 * col_stride,
                                    d_out + (int64_t)c0 * col_stride, col_stride, ghist, gbase, status, tiles, nullptr,
                                    false));
            if (ypay_out)   // the fix-up's payload of a re-sorted column is stale
                CKR(launch_gather_payload(ctx, d_out + (int64_t)c0 * col_stride, y, ypay_out + (int64_t)c0 * col_stride,
                                          col_stride, fb, N));
            i = j + 1;
        }
        return 0;
    }
    // ---- 64-bit path -----------------------------------------------------------------------------
    for (int64_t c0 = 0; c0 < F; c0 += Fb) {
        const int fb = (int)(F - c0 < Fb ? F - c0 : Fb);
        CKR(run_radix_sort<u64>(ctx, dX + c0, N, fb, ldx, keys0, keys1, d_alt + c0 * col_stride,
                                d_out + c0 * col_stride, col_stride, ghist, gbase, status, tiles, nullptr, false));
    }
    if (ypay_out) CKR(launch_gather_payload(ctx, d_out, y, ypay_out, col_stride, (int)F, N));
    return 0;
}
