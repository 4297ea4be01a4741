// Host side of libb200tree: the level-synchronous tree-growing session, the one-GPU fit,
// the single-node ops used by the parity tests, and the C ABI (include/b200dt.h).
//
// The reference grows the tree depth-first with one Python iteration per node
// (ref decision_tree/one.py:141-160).  Here all nodes of a depth are contiguous segments of
// one feature-major order array and are searched / partitioned by single launches; node ids
// are renumbered to the reference's pre-order at the end (one.py:26-31,157-160).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#include "session.cuh"

int predict_device(b200dt_ctx *ctx, const double *dX, int64_t N, int64_t F, int64_t ldx, const int *d_feature,
                   const int *d_left, const int *d_right, const double *d_threshold, const double *d_value,
                   int n_nodes, double *d_out_value, int *d_out_leaf, double avg_path_hint);
int l1_level_search(b200dt_ctx *ctx, Session *s, b200dt_candidate *d_out, bool exact_only);
int l1_cummedian_device(b200dt_ctx *ctx, const double *d_a, int64_t n, double *d_out);
int l1_leaf_medians(b200dt_ctx *ctx, Session *s, std::vector<int> &pending);

// nodes with at most this many rows are searched in the reference's sequential fp64 order
static const int EXACT_MAX_ROWS = EXACT_ROWS;
// relative gap under which two candidate scores are treated as a tie to be re-searched exactly
static const double NEAR_TIE_EPS = 1e-9;

// ---------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) block_sum_kernel(const double *__restrict__ y, int64_t n, double *partial) {
    __shared__ double sw[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += y[i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += shfl_xor_d(acc, m);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sw[w];
        partial[blockIdx.x] = t;
    }
}

// feature-major int32 [F][stride] -> row-major int64 [N][F]
__global__ void orders_to_rowmajor_kernel(const int *__restrict__ in, int64_t stride, int64_t N, int64_t F,
                                          int64_t *__restrict__ out) {
    __shared__ int tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32, f0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t f = f0 + j, n = n0 + threadIdx.x;
        if (f < F && n < N) tile[j][threadIdx.x] = in[f * stride + n];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t n = n0 + j, f = f0 + threadIdx.x;
        if (f < F && n < N) out[n * F + f] = (int64_t)tile[threadIdx.x][j];
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// uploads through pinned staging
// ---------------------------------------------------------------------------------------

template <typename T>
static int upload(b200dt_ctx *ctx, int pin_slot, size_t pin_off, const T *src, size_t count, T *dst) {
    if (count == 0) return 0;
    void *stage;
    CKR(pinned_get(ctx, pin_slot, pin_off + count * sizeof(T), &stage));
    memcpy((char *)stage + pin_off, src, count * sizeof(T));
    CK(cudaMemcpyAsync(dst, (char *)stage + pin_off, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// advance the look-back epoch; on wrap-around clear every status buffer first
static int next_epoch(b200dt_ctx *ctx) {
    if (ctx->epoch >= (1u << 30) - 2) {
        for (int id : {SB_STATUS_F, SB_STATUS_VAL, SB_STATUS_VAL2, SB_STATUS_P, SB_L1_H})
            if (ctx->scratch[id].p) CK(cudaMemsetAsync(ctx->scratch[id].p, 0, ctx->scratch[id].cap, ctx->stream));
        ctx->epoch = 0;
    }
    ctx->epoch += 1;
    return 0;
}

static int device_sum(b200dt_ctx *ctx, const double *d, int64_t n, double *out) {
    double *part;
    const int blocks = (int)std::min<int64_t>(1024, (n + 255) / 256);
    CKR(scratch_get(ctx, SB_MISC, 1024 * 8 + 64, (void **)&part));
    {
        LaunchScope ls(ctx, KC_MISC, 8.0 * (double)n);
        block_sum_kernel<<<blocks, 256, 0, ctx->stream>>>(d, n, part);
    }
    CK(cudaGetLastError());
    void *stage;
    CKR(pinned_get(ctx, PIN_MISC, 1024 * 8, &stage));
    CK(cudaMemcpyAsync(stage, part, (size_t)blocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    for (int i = 0; i < blocks; ++i) t += ((double *)stage)[i];
    *out = t;
    return 0;
}

// ---------------------------------------------------------------------------------------
// host -> device copy of a PAGEABLE matrix, column batch by column batch
// ---------------------------------------------------------------------------------------
// A user of the reference passes ordinary numpy arrays.  From pageable memory cudaMemcpy2DAsync is synchronous
// and staged by the driver in strided 512-byte pieces: 3.4 s for the 20.5 GB of C4 (6 GB/s), and the column
// batches no longer overlap the sort.  Instead a few host threads pack (rows chunk x column batch) pieces into
// pinned slots and a feeder thread issues the asynchronous copies in order on the copy stream and records one
// event per column batch -- the same events the pinned path records -- so the sort of batch b still runs while
// batch b+1 is on PCIe.
struct BouncePipe {
    std::mutex m;
    std::condition_variable cv;
    std::vector<char> ready;          // item packed
    int issued = 0;                   // items whose copy has been enqueued (in order)
    int batches_recorded = 0;         // column batches whose event has been recorded
    std::atomic<int> next{0};
    bool failed = false;
    std::string err;
    std::vector<std::thread> threads;
    ~BouncePipe() {
        for (auto &t : threads)
            if (t.joinable()) t.join();
    }
};

static bool host_pointer_is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// starts the pipeline; the caller waits for batch b with bounce_wait_batch before it enqueues a stream wait
// on ctx->copy_events[b]
static int bounce_start(b200dt_ctx *ctx, std::unique_ptr<BouncePipe> &pipe, const double *X, int64_t N, int64_t F_local,
                        int64_t ldx, double *dX, int64_t h2d_cols, int n_batches) {
    const int64_t chunk_rows = 32768;
    const int chunks = (int)((N + chunk_rows - 1) / chunk_rows);
    const int n_items = chunks * n_batches;
    unsigned hw = std::thread::hardware_concurrency();
    const int workers = (int)std::max(2u, std::min(12u, hw ? hw / 2 : 4u));
    const int slots = 2 * workers + 2;
    const size_t slot_bytes = (size_t)chunk_rows * h2d_cols * 8;
    char *stage;
    CKR(pinned_get(ctx, PIN_BOUNCE, (size_t)slots * slot_bytes, (void **)&stage));
    std::vector<cudaEvent_t> slot_ev(slots);
    for (auto &e : slot_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    pipe.reset(new BouncePipe());
    BouncePipe *bp = pipe.get();
    bp->ready.assign(n_items, 0);
    const int device = ctx->device;
    cudaStream_t cs = ctx->copy_stream;
    std::vector<cudaEvent_t> batch_ev(ctx->copy_events.begin(), ctx->copy_events.begin() + n_batches);
    auto item_geom = [=](int i, int64_t &r0, int64_t &rows, int64_t &c0, int64_t &fb) {
        const int b = i / chunks, c = i % chunks;
        r0 = (int64_t)c * chunk_rows;
        rows = std::min<int64_t>(chunk_rows, N - r0);
        c0 = (int64_t)b * h2d_cols;
        fb = std::min<int64_t>(h2d_cols, F_local - c0);
    };
    auto packer = [=]() {
            cudaSetDevice(device);
            while (true) {
                const int i = bp->next.fetch_add(1);
                if (i >= n_items) return;
                const int slot = i % slots;
                if (i >= slots) {   // the slot's previous copy must have left it
                    std::unique_lock<std::mutex> lk(bp->m);
                    bp->cv.wait(lk, [&] { return bp->issued > i - slots || bp->failed; });
                    if (bp->failed) return;
                    lk.unlock();
                    cudaEventSynchronize(slot_ev[slot]);
                }
                int64_t r0, rows, c0, fb;
                item_geom(i, r0, rows, c0, fb);
                char *dst = stage + (size_t)slot * slot_bytes;
                const size_t width = (size_t)fb * 8;
                const double *src = X + r0 * ldx + c0;
                for (int64_t r = 0; r < rows; ++r) memcpy(dst + (size_t)r * width, src + r * ldx, width);
                {
                    std::lock_guard<std::mutex> lk(bp->m);
                    bp->ready[i] = 1;
                }
                bp->cv.notify_all();
            }
        };
    auto feeder = [=]() {   // feeder: copies leave in item order, one event per column batch
        cudaSetDevice(device);
        for (int i = 0; i < n_items; ++i) {
            {
                std::unique_lock<std::mutex> lk(bp->m);
                bp->cv.wait(lk, [&] { return bp->ready[i] != 0 || bp->failed; });
                if (bp->failed) return;
            }
            int64_t r0, rows, c0, fb;
            item_geom(i, r0, rows, c0, fb);
            const int slot = i % slots;
            cudaError_t e = cudaMemcpy2DAsync(dX + r0 * F_local + c0, (size_t)F_local * 8, stage + (size_t)slot * slot_bytes,
                                              (size_t)fb * 8, (size_t)fb * 8, (size_t)rows, cudaMemcpyHostToDevice, cs);
            if (e == cudaSuccess) e = cudaEventRecord(slot_ev[slot], cs);
            const bool last_of_batch = (i % chunks) == chunks - 1;
            if (e == cudaSuccess && last_of_batch) e = cudaEventRecord(batch_ev[i / chunks], cs);
            {
                std::lock_guard<std::mutex> lk(bp->m);
                if (e != cudaSuccess) {
                    bp->failed = true;
                    bp->err = std::string("pageable H2D pipeline: ") + cudaGetErrorString(e);
                } else {
                    bp->issued = i + 1;
                    if (last_of_batch) bp->batches_recorded = i / chunks + 1;
                }
            }
            bp->cv.notify_all();
            if (e != cudaSuccess) return;
        }
        cudaStreamSynchronize(cs);   // the slots are reused by the next fit
        for (auto ev : slot_ev) cudaEventDestroy(ev);
    };
    // thread creation can fail (resource limits): the feeder is indispensable, the packers degrade gracefully
    int started = 0;
    try {
        bp->threads.emplace_back(feeder);
        for (int w = 0; w < workers; ++w) {
            bp->threads.emplace_back(packer);
            ++started;
        }
    } catch (const std::exception &) {
    }
    if (started == 0) {
        {
            std::lock_guard<std::mutex> lk(bp->m);
            bp->failed = true;
            bp->err = "pageable H2D pipeline: could not start host threads";
        }
        bp->cv.notify_all();
        pipe.reset();   // joins whatever was started
        return ctx->fail_msg("pageable H2D pipeline: could not start host threads");
    }
    return 0;
}

static int bounce_wait_batch(b200dt_ctx *ctx, BouncePipe *bp, int b) {
    std::unique_lock<std::mutex> lk(bp->m);
    bp->cv.wait(lk, [&] { return bp->batches_recorded > b || bp->failed; });
    if (bp->failed) return ctx->fail_msg(bp->err);
    return 0;
}

// ---------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------
static void session_free(b200dt_ctx *ctx) {
    delete ctx->session;
    ctx->session = nullptr;
}
void tree_session_free(b200dt_ctx *ctx) { session_free(ctx); }

static int make_child(Session *s, int parent, int is_left, int start, int n, double total) {
    HNode c;
    c.parent = parent;
    c.is_left = is_left;
    c.depth = s->nodes[parent].depth + 1;
    c.start = start;
    c.n = n;
    c.total = total;
    s->nodes.push_back(c);
    return (int)s->nodes.size() - 1;
}

static bool node_is_leaf_by_size(const Session *s, const HNode &nd) {
    // ref one.py:40-46,145: rows <= min_sample_leaf, or depth limit
    return nd.n <= s->min_sample_leaf || nd.depth >= s->max_depth;
}

static int session_begin_impl(b200dt_ctx *ctx, const double *X, const double *y, int64_t N, int64_t F_local,
                              int64_t ldx, int space, int criterion, int max_depth, double min_improvement,
                              int64_t min_sample_leaf, int variant, int64_t col_offset, int64_t F_global,
                              const double *last_col, int64_t last_col_ld, const double *weight = nullptr) {
    NvtxRange nvtx("b200dt: begin (H2D + sort)");
    if (N <= 0 || F_local <= 0) return ctx->fail_msg("fit: empty input");
    if (N >= (1ll << 30)) return ctx->fail_msg("fit: N must be < 2^30");
    if (F_local > 65534) return ctx->fail_msg("fit: at most 65534 columns per GPU (shard the columns over more ranks)");
    if (max_depth < 0 || max_depth > 30) return ctx->fail_msg("fit: max_depth out of range [0, 30]");
    if (criterion != B200DT_L2 && criterion != B200DT_L1) return ctx->fail_msg("fit: unknown criterion");
    end_all_sessions(ctx);
    Session *s = new (std::nothrow) Session();
    if (!s) return ctx->fail_msg("fit: out of host memory");
    ctx->session = s;
    s->N = N;
    s->F_local = (int)F_local;
    s->F_global = (int)F_global;
    s->col_offset = (int)col_offset;
    s->criterion = criterion;
    s->max_depth = max_depth;
    s->variant = variant;
    s->min_improvement = min_improvement;
    s->min_sample_leaf = min_sample_leaf;
    const bool owns_last = (col_offset + F_local == F_global);
    // (last_col is accepted for compatibility and no longer used: see Session::sharded)
    (void)last_col;
    (void)last_col_ld;
    s->sharded = F_local != F_global;
    int F_store = (int)F_local;
    s->total_col = owns_last ? (int)F_local - 1 : -1;
    if (criterion == B200DT_L1) {
        s->ycol = F_store;
        F_store += 1;
    }
    s->F_store = F_store;
    s->col_stride = (N + 31) / 32 * 32;

    // inputs on the device
    // host inputs: X is copied in column batches on a second stream so that the sort of batch b
    // overlaps the PCIe transfer of batch b+1 (pinned host memory makes the copies asynchronous)
    const int64_t h2d_cols = 32;   // (64: the sort of the LAST batch, which nothing overlaps, took 27 ms instead of 13)
    int n_h2d = 0;
    std::unique_ptr<BouncePipe> bounce;   // pageable host input: joined when this function returns
    if (space == B200DT_HOST) {
        double *dX, *dy;
        CKR(scratch_get(ctx, SB_X, (size_t)N * F_local * 8, (void **)&dX));
        CKR(scratch_get(ctx, SB_Y, (size_t)N * 8, (void **)&dy));
        if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        n_h2d = (int)((F_local + h2d_cols - 1) / h2d_cols);
        while ((int)ctx->copy_events.size() < n_h2d + 1) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->copy_events.push_back(e);
        }
        // the copy stream must not overtake earlier work of the main stream on these buffers
        CK(cudaEventRecord(ctx->copy_events[n_h2d], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[n_h2d], 0));
        CK(cudaMemcpyAsync(dy, y, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
        if ((size_t)N * F_local * 8 >= ((size_t)64 << 20) && host_pointer_is_pageable(X)) {
            CKR(bounce_start(ctx, bounce, X, N, F_local, ldx, dX, h2d_cols, n_h2d));
        } else {
            for (int b = 0; b < n_h2d; ++b) {
                const int64_t c0 = b * h2d_cols, fb = std::min<int64_t>(h2d_cols, F_local - c0);
                CK(cudaMemcpy2DAsync(dX + c0, (size_t)F_local * 8, X + c0, (size_t)ldx * 8, (size_t)fb * 8, (size_t)N,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
                CK(cudaEventRecord(ctx->copy_events[b], ctx->copy_stream));
            }
        }
        s->dX = dX;
        s->ldx = F_local;
        s->dy = dy;
    } else {
        s->dX = X;
        s->ldx = ldx;
        s->dy = y;
    }
    if (weight) {
        if (criterion != B200DT_L2 || variant != 1)
            return ctx->fail_msg("fit: sample weights are defined for the L2 criterion (v == 1) only");
        // four.py:116-117: y <- y * weight once; the search then works on (y*w, w)
        const double *dw = weight;
        if (space == B200DT_HOST) {
            double *buf;
            CKR(scratch_get(ctx, SB_W, (size_t)N * 8, (void **)&buf));
            CK(cudaMemcpyAsync(buf, weight, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
            // y travels on the copy stream, ahead of the first X batch: wait for that batch's event
            if (bounce) CKR(bounce_wait_batch(ctx, bounce.get(), 0));
            if (n_h2d > 0) CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[0], 0));
            dw = buf;
        }
        double *yw;
        CKR(scratch_get(ctx, SB_YW, (size_t)N * 8, (void **)&yw));
        CKR(launch_mul(ctx, s->dy, dw, N, yw));
        s->dy = yw;
        s->dw = dw;
    }
    CKR(scratch_get(ctx, SB_ORD_A, (size_t)F_store * s->col_stride * 4, (void **)&s->ord[0]));
    CKR(scratch_get(ctx, SB_ORD_B, (size_t)F_store * s->col_stride * 4, (void **)&s->ord[1]));
    s->cur = 0;
    u64 *k0 = nullptr, *k1 = nullptr;
    int64_t kcap = 0;
    int payload_cols_done = 0;
    {
        const char *t = getenv("B200DT_TWO_LEVEL");
        s->two_level = t && t[0] == '1';
        // rows whose side bits stay L1-resident (1.5 M rows = 183 KB of mask); B200DT_RELABEL_ROWS overrides
        // (0 = off; small values make small test problems take the relabelled path)
        s->relabel_window = 1500000;
        const char *sb = getenv("B200DT_SCAN_BULK");
        s->scan_bulk = sb && sb[0] == '1';
        const char *w = getenv("B200DT_RELABEL_ROWS");
        if (w && w[0]) s->relabel_window = atoll(w);
    }
    if (criterion == B200DT_L2) {
        // y-payload double buffer; idle during the sort, so it doubles as the radix key scratch
        kcap = (int64_t)F_store * s->col_stride;
        CKR(scratch_get(ctx, SB_YPAY0, (size_t)kcap * 8, (void **)&s->ypay[0]));
        CKR(scratch_get(ctx, SB_YPAY1, (size_t)kcap * 8, (void **)&s->ypay[1]));
        // the sort borrows only the ALTERNATE payload buffer (two halves), so the primary one can be
        // filled column batch by column batch while later batches are still in flight
        kcap /= 2;
        k0 = (u64 *)s->ypay[1];
        k1 = (u64 *)s->ypay[1] + kcap;
    }
    // one sort per fit (one.py:139)
    if (n_h2d > 0) {
        for (int b = 0; b < n_h2d; ++b) {
            const int64_t c0 = b * h2d_cols, fb = std::min<int64_t>(h2d_cols, F_local - c0);
            if (bounce) CKR(bounce_wait_batch(ctx, bounce.get(), b));   // the event must be RECORDED before the wait
            CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[b], 0));
            // (y is on the device before the first batch: its payload comes out of the sort's fix-up pass)
            CKR(sort_columns_device(ctx, s->dX + c0, N, fb, s->ldx, s->ord[0] + c0 * s->col_stride,
                                    s->ord[1] + c0 * s->col_stride, s->col_stride, k0, k1, kcap, s->dy,
                                    s->ypay[0] ? s->ypay[0] + c0 * s->col_stride : nullptr));
        }
        payload_cols_done = (int)F_local;
    } else {
        CKR(sort_columns_device(ctx, s->dX, N, F_local, s->ldx, s->ord[0], s->ord[1], s->col_stride, k0, k1, kcap,
                                s->dy, s->ypay[0]));
        if (s->ypay[0]) payload_cols_done = (int)F_local;
    }
    if (s->ycol >= 0) {
        CKR(sort_columns_device(ctx, s->dy, N, 1, 1, s->ord[0] + (int64_t)s->ycol * s->col_stride,
                                s->ord[1] + (int64_t)s->ycol * s->col_stride, s->col_stride));
    }
    if (s->ypay[0] && payload_cols_done < F_store)
        CKR(launch_gather_payload(ctx, s->ord[0] + (int64_t)payload_cols_done * s->col_stride, s->dy,
                                  s->ypay[0] + (int64_t)payload_cols_done * s->col_stride, s->col_stride,
                                  F_store - payload_cols_done, N));
    if (s->dw) {
        CKR(scratch_get(ctx, SB_WPAY0, (size_t)F_store * s->col_stride * 8, (void **)&s->wpay[0]));
        CKR(scratch_get(ctx, SB_WPAY1, (size_t)F_store * s->col_stride * 8, (void **)&s->wpay[1]));
        CKR(launch_gather_payload(ctx, s->ord[0], s->dw, s->wpay[0], s->col_stride, F_store, N));
    }
    // (the look-back status arrays are sized per launch from the actual tile count)
    CKR(scratch_get(ctx, SB_MASK, (size_t)((N + 31) / 32) * 4 + 64, (void **)&s->d_own_mask));

    HNode root;
    root.start = 0;
    root.n = (int)N;
    CKR(device_sum(ctx, s->dy, N, &root.total));
    if (s->dw) CKR(device_sum(ctx, s->dw, N, &root.total_w));
    s->nodes.clear();
    s->nodes.push_back(root);
    s->live.clear();
    s->depth = 0;
    if (node_is_leaf_by_size(s, s->nodes[0])) {
        s->nodes[0].leaf = true;
        s->nodes[0].value = s->dw ? root.total / root.total_w : root.total / (double)N;
        if (criterion == B200DT_L1) s->median_pending_cur.push_back(0);
    } else {
        s->live.push_back(0);
    }
    s->force_exact.assign(s->live.size(), 0);
    s->known_total.assign(s->live.size(), 0.0);
    s->has_total.assign(s->live.size(), 0);
    return 0;
}

// Build and upload the segment table of the live nodes; returns lists of tiled / sequential nodes.
static int upload_level(b200dt_ctx *ctx, Session *s, std::vector<TileDesc> &tiles, std::vector<int> &small,
                        int64_t &tile_elems, int64_t &small_elems, bool only_forced) {
    const int n_live = (int)s->live.size();
    s->h_segs.resize(n_live);
    tiles.clear();
    small.clear();
    tile_elems = small_elems = 0;
    std::vector<int> scanned;   // live nodes searched by the tiled scan
    // candidate slots: [0, fused_slots) were written by the fused level kernel of the parents; behind
    // them every live node reserves slots of its own (its tiles when it must be scanned here -- the
    // root, or any node of a gather-path session -- else one, for the sequential search)
    int pbase = s->fused_slots;
    const int scan_tile = s->dw ? WT_TILE : l2_scan_tile(s->ypay[0] != nullptr);
    for (int i = 0; i < n_live; ++i) {
        const HNode &nd = s->nodes[s->live[i]];
        // the streaming search copies its tiles with 16-byte aligned bulk copies: a node that starts at an odd
        // position has its tile grid shifted down by one element (l2.cu)
        const int shift = (s->scan_bulk && s->ypay[0] != nullptr && !s->dw) ? (nd.start & 1) : 0;
        const int ntiles = (nd.n + shift + scan_tile - 1) / scan_tile;
        SegDev sg;
        sg.total = nd.total;
        sg.total_w = nd.total_w;
        sg.start = nd.start;
        sg.n = nd.n;
        sg.nleft = 0;
        sg.exact = (nd.n <= EXACT_MAX_ROWS || s->force_exact[i]) ? 1 : 0;
        sg.pstep = 1;
        sg.pside = 0;
        sg.vstart = -1;
        sg.cbase = 0;
        sg.pbase = pbase;
        sg.np = sg.exact ? 1 : ntiles;
        const bool from_parent = nd.has_fused && !sg.exact;
        if (from_parent) {
            sg.pbase = nd.fp_base;
            sg.np = nd.fp_np;
            sg.pstep = 2;
            sg.pside = nd.fp_side;
            sg.vstart = nd.vstart;   // >= 0: searched from its parent's order, rows not moved yet (children.cu)
        }
        pbase += nd.has_fused ? 1 : ntiles;
        s->h_segs[i] = sg;
        if (sg.exact) {
            if (!only_forced || s->force_exact[i]) {
                small.push_back(i);
                small_elems += nd.n;
            }
        } else if (!only_forced && !from_parent) {
            scanned.push_back(i);
            tile_elems += nd.n;
        }
    }
    // Tile order of the scan: ROUND-ROBIN over the nodes (tile 0 of every node, then tile 1 of every node, ...).  A
    // persistent grid has a fixed number of tiles in flight; in node-major order they would all be consecutive tiles
    // of ONE (node, column) chain when a rank owns few columns (~110 at 32 columns), and every look-back then waits
    // for the slowest of ~110 predecessors.  Interleaving the nodes divides that by the number of live nodes.  A tile
    // still only depends on lower-numbered tiles of its own chain.
    {
        int max_np = 0, cbase = 0;
        for (int i : scanned) {
            max_np = std::max(max_np, s->h_segs[i].np);
            s->h_segs[i].cbase = cbase;    // chain slots stay node-major, only the processing order is interleaved
            cbase += s->h_segs[i].np;
        }
        for (int t = 0; t < max_np; ++t)
            for (int i : scanned)
                if (t < s->h_segs[i].np) tiles.push_back(TileDesc{i, t});
    }
    CKR(scratch_get(ctx, SB_SEG, (size_t)n_live * sizeof(SegDev) + 64, (void **)&s->d_segs));
    CKR(scratch_get(ctx, SB_PARTIAL, (size_t)(pbase + 1) * s->F_local * sizeof(Partial) + 64, (void **)&s->d_partials));
    s->next_partial_slots = 0;
    CKR(scratch_get(ctx, SB_TMP2, (size_t)n_live * 8 + 64, (void **)&s->d_exact_total));
    CKR(upload(ctx, PIN_SEARCH, 0, s->h_segs.data(), (size_t)n_live, s->d_segs));
    return 0;
}

static int materialise_pending(b200dt_ctx *ctx);

static int session_search_impl(b200dt_ctx *ctx, b200dt_candidate *d_out, bool exact_only) {
    NvtxRange nvtx(exact_only ? "b200dt: level search (exact)" : "b200dt: level search");
    Session *s = ctx->session;
    if (!s) return ctx->fail_msg("session: not begun");
    const int n_live = (int)s->live.size();
    if (n_live == 0) return 0;
    for (int i = 0; i < n_live; ++i)
        if (s->nodes[s->live[i]].n < 2)
            return ctx->fail_msg("fit: attempt to split a node with a single row (min_sample_leaf < 1); "
                                 "the reference raises ValueError here (argmax of an empty sequence)");
    if (s->criterion == B200DT_L1) return l1_level_search(ctx, s, d_out, exact_only);
    if (s->pending && exact_only) {
        // near tie on the second level of a pair: move the rows (plain partition), then search this level again
        // in the ordinary way (sequential order for the flagged nodes, parallel scan for the others)
        CKR(materialise_pending(ctx));
        exact_only = false;
    }
    std::vector<TileDesc> tiles;
    std::vector<int> small;
    int64_t tile_elems, small_elems;
    CKR(upload_level(ctx, s, tiles, small, tile_elems, small_elems, exact_only));
    TileDesc *d_tiles;
    int *d_list;
    const size_t seg_bytes = (size_t)n_live * sizeof(SegDev);
    CKR(scratch_get(ctx, SB_TILES, (tiles.size() + 1) * sizeof(TileDesc) + (small.size() + 1) * 4 + 64, (void **)&d_tiles));
    d_list = (int *)(d_tiles + tiles.size() + 1);
    CKR(upload(ctx, PIN_SEARCH, seg_bytes + 64, tiles.data(), tiles.size(), d_tiles));
    CKR(upload(ctx, PIN_SEARCH, seg_bytes + 64 + (tiles.size() + 1) * sizeof(TileDesc), small.data(), small.size(), d_list));
    LevelArgs a;
    a.orders = s->ord[s->cur];
    a.col_stride = s->col_stride;
    a.F_search = s->F_local;
    a.y = s->dy;
    a.ypay = s->ypay[s->cur];
    a.segs = s->d_segs;
    a.n_segs = n_live;
    a.wpay = s->wpay[s->cur];
    a.row_map = s->relabeled ? s->d_rowmap : nullptr;
    CKR(scratch_get(ctx, SB_NODE_BEST, (size_t)n_live * 8 + 64, (void **)&a.node_best));
    CK(cudaMemsetAsync(a.node_best, 0, (size_t)n_live * 8, ctx->stream));
    int64_t max_records = 0;
    for (const SegDev &sg : s->h_segs) max_records = std::max<int64_t>(max_records, (int64_t)sg.np * s->F_local);
    Partial *d_sliced;
    CKR(scratch_get(ctx, SB_L1_G, (size_t)n_live * FIN_SLICES_HOST * sizeof(Partial) + 64, (void **)&d_sliced));
    // look-back status: one entry per (tile, column); a freshly (re)allocated buffer is zeroed,
    // which no epoch ever matches
    const size_t n_status = tiles.size() * (size_t)s->F_local + 1;
    u32 *ticket = (u32 *)((char *)ctx->scratch[SB_MASK].p + (size_t)((s->N + 31) / 32) * 4);
    CKR(next_epoch(ctx));
    if (s->dw) {
        // weighted search (four.py:21-30): two payload channels, flat (k, f) tie rule
        u64 *st3;
        double *d_tw;
        CKR(scratch_get(ctx, SB_STATUS_F, n_status * 48, (void **)&st3, true));
        CKR(scratch_get(ctx, SB_TW, (size_t)n_live * 8 + 64, (void **)&d_tw));
        CKR(launch_l2w_scan(ctx, a, d_tiles, (int)tiles.size(), tile_elems, st3, ctx->epoch, s->d_partials));
        CKR(launch_l2w_exact(ctx, a, d_list, (int)small.size(), small_elems, s->total_col, s->d_exact_total, d_tw,
                             s->d_partials));
        CKR(launch_finalize(ctx, a, s->d_partials, max_records, d_sliced, s->dX, s->ldx, s->col_offset, d_out, true));
        return 0;
    }
    ScanStatus st;
    CKR(scratch_get(ctx, SB_STATUS_F, n_status * 16, (void **)&st.slot, true));
    CKR(launch_l2_scan(ctx, a, d_tiles, (int)tiles.size(), tile_elems, st, ctx->epoch, ticket, s->d_partials, s->variant,
                       s->scan_bulk));
    if (!s->sharded) {
        CKR(launch_l2_exact(ctx, a, d_list, (int)small.size(), small_elems, s->total_col, s->d_exact_total,
                            s->d_partials, s->variant));
        CKR(launch_finalize(ctx, a, s->d_partials, max_records, d_sliced, s->dX, s->ldx, s->col_offset, d_out));
        return 0;
    }
    // feature-sharded: nodes whose total is known (it came back with the previous gather) are searched now; for the
    // others this round only publishes the total -- the owner of the last global column computes it, every rank
    // emits a marked record, and `decide` asks for another exact round
    std::vector<int> known, unknown;
    int64_t known_elems = 0, unknown_elems = 0;
    for (int i : small) {
        if (s->has_total[i]) {
            known.push_back(i);
            known_elems += s->nodes[s->live[i]].n;
        } else {
            unknown.push_back(i);
            unknown_elems += s->nodes[s->live[i]].n;
        }
    }
    int *d_known, *d_unknown;
    CKR(scratch_get(ctx, SB_TMP_I64, (known.size() + unknown.size() + 2) * 4 + 128, (void **)&d_known));
    d_unknown = d_known + known.size() + 1;
    const size_t off_lists = seg_bytes + 64 + (tiles.size() + 1) * sizeof(TileDesc) + (small.size() + 1) * 4 + 64;
    CKR(upload(ctx, PIN_SEARCH, off_lists, known.data(), known.size(), d_known));
    CKR(upload(ctx, PIN_SEARCH, off_lists + (known.size() + 1) * 4, unknown.data(), unknown.size(), d_unknown));
    if (!known.empty()) {
        CKR(upload(ctx, PIN_SEARCH, off_lists + (known.size() + unknown.size() + 2) * 4 + 64, s->known_total.data(),
                   (size_t)n_live, s->d_exact_total));
        CKR(launch_l2_exact(ctx, a, d_known, (int)known.size(), known_elems, 0, s->d_exact_total, s->d_partials,
                            s->variant, true));
    }
    CKR(launch_clear_partials(ctx, a, d_unknown, (int)unknown.size(), s->d_partials));
    const bool owner = s->total_col >= 0;
    if (owner && !unknown.empty())
        CKR(launch_l2_exact_totals(ctx, a, d_unknown, (int)unknown.size(), unknown_elems, s->total_col, s->d_exact_total));
    CKR(launch_finalize(ctx, a, s->d_partials, max_records, d_sliced, s->dX, s->ldx, s->col_offset, d_out));
    CKR(launch_emit_totals(ctx, d_unknown, (int)unknown.size(), s->d_exact_total, owner ? s->F_global - 1 : -1, d_out));
    return 0;
}

// winner of one node among the ranks' candidates: (score desc, k asc, aux desc, f asc)
static bool cand_better(const b200dt_candidate &b, const b200dt_candidate &a, bool flat) {
    if (b.k < 0) return false;
    if (a.k < 0) return true;
    if (b.gain != a.gain) return b.gain > a.gain;
    if (b.k != a.k) return b.k < a.k;
    if (!flat && b.aux != a.aux) return b.aux > a.aux;   // flat argmax criteria carry data in aux
    return b.f < a.f;
}

static int session_decide_impl(b200dt_ctx *ctx, const b200dt_candidate *d_gathered, int n_ranks, u32 *d_mask,
                               int *need_partition, int *need_exact) {
    NvtxRange nvtx("b200dt: level decide");
    Session *s = ctx->session;
    if (!s) return ctx->fail_msg("session: not begun");
    *need_partition = 0;
    *need_exact = 0;
    const int n_live = (int)s->live.size();
    if (n_live == 0) return 0;
    b200dt_candidate *h;
    CKR(pinned_get(ctx, PIN_RESULT, (size_t)n_ranks * n_live * sizeof(b200dt_candidate), (void **)&h));
    CK(cudaMemcpyAsync(h, d_gathered, (size_t)n_ranks * n_live * sizeof(b200dt_candidate), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<b200dt_candidate> win(n_live);
    bool any_flag = false;
    // records marked exact == 2 carry a node total (feature-sharded fits), not a candidate: remember the total and
    // have the node searched sequentially in the next exact round
    std::vector<char> totals_only(n_live, 0);
    for (int i = 0; i < n_live; ++i)
        for (int r = 0; r < n_ranks; ++r) {
            const b200dt_candidate &c = h[(size_t)r * n_live + i];
            if (c.exact != 2) continue;
            totals_only[i] = 1;
            if (c.f == s->F_global - 1) {
                s->known_total[i] = c.left_sum;
                s->has_total[i] = 1;
            }
        }
    for (int i = 0; i < n_live; ++i) {
        if (totals_only[i]) {
            if (!s->has_total[i]) return ctx->fail_msg("fit: no rank supplied the total of a node (last global column)");
            s->force_exact[i] = 1;
            any_flag = true;
            continue;
        }
        b200dt_candidate best = h[i];
        double second = best.gain2;
        for (int r = 1; r < n_ranks; ++r) {
            const b200dt_candidate &c = h[(size_t)r * n_live + i];
            second = std::max(second, c.gain2);
            if (cand_better(c, best, s->dw != nullptr)) {
                if (best.k >= 0) second = std::max(second, best.gain);
                best = c;
            } else if (c.k >= 0) {
                second = std::max(second, c.gain);
            }
        }
        win[i] = best;
        if (best.k < 0) return ctx->fail_msg("fit: a node produced no split candidate");
        if (!std::isfinite(best.gain)) continue;  // NaN/inf targets: nothing to disambiguate
        // a runner-up within NEAR_TIE_EPS of the winner of an approximate (parallel-scan) search
        if (!best.exact && second >= best.gain - NEAR_TIE_EPS * std::fabs(best.gain)) {
            s->force_exact[i] = 1;
            any_flag = true;
        }
    }
    if (any_flag) {
        *need_exact = 1;
        return 0;
    }
    // apply: leaf or split (ref one.py:145-160)
    std::vector<int> marks;
    int max_left = 0;
    std::vector<int> next_live;
    s->plan_segs.clear();
    s->plan_fused.clear();
    s->plan_tiles.clear();
    s->plan_elems = 0;
    s->next_fused_slots = 0;
    s->plan_parents.clear();
    s->plan_two_level = false;
    // two levels per data movement (children.cu): this level's children may be searched from the parents' order
    // (first level of a pair), or the live nodes already are such children (second level: `was_pending`)
    const bool was_pending = s->pending;
    const bool pair_candidate = s->two_level && !was_pending && s->ypay[0] != nullptr && !s->dw &&
                                s->criterion == B200DT_L2;
    std::vector<PlanSeg> pair_plan;        // one per plan seg, filled while the plan is built
    std::vector<int> pair_children;        // (left child, right child) node ids per plan seg
    std::vector<int> vmarks;               // [f, parent start, parent position, side] of virtual split nodes
    int max_vspan = 0;
    for (int i = 0; i < n_live; ++i) {
        const int nid = s->live[i];
        const b200dt_candidate &w = win[i];
        double score_for_test;
        if (s->criterion == B200DT_L1)
            score_for_test = -w.gain;  // children cost (strategy_l1.py:158: same "<" test)
        else if (s->variant == 1)
            score_for_test = w.gain;
        else
            score_for_test = w.gain * w.gain;  // one.py:93 returns the square of the best ratio score
        const bool stop = (s->flags & B200DT_LEAF_IF_LE) ? score_for_test <= s->min_improvement   // four.py:143
                                                         : score_for_test < s->min_improvement;   // one.py:150
        if (stop) {
            HNode &nd = s->nodes[nid];
            nd.leaf = true;
            nd.value = s->dw ? nd.total / nd.total_w : nd.total / (double)nd.n;
            if (s->criterion == B200DT_L1) s->median_pending_cur.push_back(nid);
            continue;
        }
        const int start = s->nodes[nid].start, n = s->nodes[nid].n;
        const double total = s->nodes[nid].total, total_w = s->nodes[nid].total_w;
        const int nl = w.k + 1;
        s->nodes[nid].feature = w.f;
        s->nodes[nid].thr = w.threshold;
        s->nodes[nid].value = s->dw ? total / total_w : total / (double)n;
        const int lc = make_child(s, nid, 1, start, nl, w.left_sum);
        const int rc = make_child(s, nid, 0, start + nl, n - nl, total - w.left_sum);
        if (s->dw) {   // weighted: the candidate's aux slot is the left weight sum
            s->nodes[lc].total_w = w.aux;
            s->nodes[rc].total_w = total_w - w.aux;
        }
        s->nodes[nid].left = lc;
        s->nodes[nid].right = rc;
        bool any_search = false;
        for (int c : {lc, rc}) {
            HNode &ch = s->nodes[c];
            if (node_is_leaf_by_size(s, ch)) {
                ch.leaf = true;
                ch.value = s->dw ? ch.total / ch.total_w : ch.total / (double)ch.n;
                if (s->criterion == B200DT_L1) s->median_pending_next.push_back(c);
            } else {
                next_live.push_back(c);
                any_search = true;
            }
        }
        // L1 leaves read their median from the partitioned y-order column, so always partition
        if (any_search || s->criterion == B200DT_L1) {
            SegDev sg;
            sg.total = 0.0;
            sg.total_w = 0.0;
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
            const int si = (int)s->plan_segs.size();
            s->plan_segs.push_back(sg);
            for (int t = 0; t < sg.np; ++t) s->plan_tiles.push_back(TileDesc{si, t});
            s->plan_elems += n;
            s->plan_parents.push_back(PendParent{nid, start, n, nl});
            if (pair_candidate) {
                PlanSeg ps;
                ps.total_l = w.left_sum;
                ps.total_r = total - w.left_sum;
                ps.start = start;
                ps.n = n;
                ps.nleft = nl;
                ps.fbase = 0;
                ps.score_l = ps.score_r = 0;
                pair_plan.push_back(ps);
                pair_children.push_back(lc);
                pair_children.push_back(rc);
            }
            const int fl = w.f - s->col_offset;
            if (fl >= 0 && fl < s->F_local) {  // this rank owns the winning column: it marks the left rows
                const HNode &me = s->nodes[nid];
                if (me.vstart >= 0) {
                    // the node's rows have not been moved: its left rows are the rows of its side up to the
                    // winner's position in the PARENT's order (candidate.pad)
                    vmarks.push_back(fl);
                    vmarks.push_back(me.vstart);
                    vmarks.push_back(w.pad);
                    vmarks.push_back(me.fp_side);
                    max_vspan = std::max(max_vspan, w.pad + 1);
                } else {
                    marks.push_back(fl);
                    marks.push_back(start);
                    marks.push_back(w.k);
                    max_left = std::max(max_left, nl);
                }
            }
        }
    }
    // relabel the rows with this partition?  Once, at the first level whose children all fit the window;
    // payload sessions only (no kernel gathers y or weights by row id there).  Every rank of a sharded fit
    // takes the same decision from the same plan.
    s->plan_relabel = false;
    if (!s->relabeled && !s->two_level && s->ypay[0] != nullptr && s->relabel_window > 0 &&
        s->N > s->relabel_window && !s->plan_segs.empty() && s->plan_segs.size() <= 512) {
        int64_t largest = 0;
        for (int c : next_live) largest = std::max<int64_t>(largest, s->nodes[c].n);
        s->plan_relabel = largest <= s->relabel_window;
    }
    if (!s->plan_segs.empty()) {
        *need_partition = 1;
        CK(cudaMemsetAsync(d_mask, 0, (size_t)((s->N + 31) / 32) * 4, ctx->stream));
        int *d_marks;
        CKR(scratch_get(ctx, SB_TMP_I64, (marks.size() + vmarks.size()) * 4 + 128, (void **)&d_marks));
        if (!marks.empty()) {
            CKR(upload(ctx, PIN_PLAN, 0, marks.data(), marks.size(), d_marks));
            CKR(launch_mark_left(ctx, s->ord[s->cur], s->col_stride, d_marks, (int)marks.size() / 3, max_left, d_mask));
        }
        if (!vmarks.empty()) {
            int *d_vmarks = d_marks + (marks.size() + 15) / 16 * 16;
            CKR(upload(ctx, PIN_PLAN, (marks.size() + 15) / 16 * 64, vmarks.data(), vmarks.size(), d_vmarks));
            CKR(launch_mark_left_virtual(ctx, s->ord[s->cur], s->col_stride, d_vmarks, (int)vmarks.size() / 4, max_vspan,
                                         s->d_mask_a, d_mask));
        }
    }
    // first level of a pair?  Only if every node the next level searches is large enough for the parallel scan
    // (the sequential exact kernel needs the node's rows in place)
    if (pair_candidate && !s->plan_segs.empty() && !next_live.empty()) {
        bool ok = true;
        for (int c : next_live) ok = ok && s->nodes[c].n > EXACT_MAX_ROWS;
        if (ok) {
            int slots = 0;
            for (size_t i = 0; i < pair_plan.size(); ++i) {
                PlanSeg &ps = pair_plan[i];
                const int np = (ps.n + ST_TILE - 1) / ST_TILE;
                ps.fbase = slots;
                for (int side = 0; side < 2; ++side) {
                    HNode &ch = s->nodes[pair_children[2 * i + side]];
                    if (ch.leaf) continue;
                    ch.has_fused = true;
                    ch.fp_base = ps.fbase;
                    ch.fp_np = np;
                    ch.fp_side = side;
                    ch.vstart = ps.start;
                    (side == 0 ? ps.score_l : ps.score_r) = 1;
                }
                slots += 2 * np;
            }
            s->next_fused_slots = slots;
            s->plan_fused = pair_plan;
            s->plan_two_level = true;
        }
    }
    // candidate slots the next level needs: the fused ones + each live node's own reserve (this must
    // be allocated BEFORE the fused kernel writes, and not grow afterwards)
    size_t slots = (size_t)s->next_fused_slots + 1;
    for (int c : next_live) {
        const HNode &ch = s->nodes[c];
        slots += ch.has_fused ? 1 : (size_t)(ch.n + WT_TILE - 1) / WT_TILE;   // upper bound for either tile size
    }
    s->next_partial_slots = slots;
    s->fused_slots = s->next_fused_slots;
    s->live.swap(next_live);
    s->force_exact.assign(s->live.size(), 0);
    s->known_total.assign(s->live.size(), 0.0);
    s->has_total.assign(s->live.size(), 0);
    s->depth += 1;
    return 0;
}

// The rows of the pending level's nodes still lie in their parents' segments: move them now with the plain
// binary partition (mask of the first level of the pair).  Used when the second level needs its rows in place.
static int materialise_pending(b200dt_ctx *ctx) {
    Session *s = ctx->session;
    if (!s->pending) return 0;
    SegDev *d_segs;
    TileDesc *d_tiles;
    u64 *pstatus;
    const size_t seg_bytes = s->pend_segs.size() * sizeof(SegDev);
    CKR(scratch_get(ctx, SB_SEG, seg_bytes + 64, (void **)&d_segs));
    CKR(scratch_get(ctx, SB_TILES, s->pend_tiles.size() * sizeof(TileDesc) + 64, (void **)&d_tiles));
    CKR(upload(ctx, PIN_PLAN, 0, s->pend_segs.data(), s->pend_segs.size(), d_segs));
    CKR(upload(ctx, PIN_PLAN, seg_bytes + 64, s->pend_tiles.data(), s->pend_tiles.size(), d_tiles));
    CKR(scratch_get(ctx, SB_STATUS_P, (s->pend_tiles.size() * (size_t)s->F_store + 1) * 8, (void **)&pstatus, true));
    CKR(next_epoch(ctx));
    CKR(launch_partition(ctx, s->ord[s->cur], s->ord[s->cur ^ 1], s->col_stride, s->F_store, d_segs, d_tiles,
                         (int)s->pend_tiles.size(), s->pend_elems, s->d_mask_a, pstatus, ctx->epoch, nullptr,
                         s->ypay[s->cur], s->ypay[s->cur ^ 1], nullptr, nullptr));
    s->cur ^= 1;
    for (int c : s->live) {
        s->nodes[c].has_fused = false;
        s->nodes[c].vstart = -1;
    }
    s->fused_slots = 0;
    s->pending = false;
    s->pend.clear();
    s->pend_segs.clear();
    s->pend_tiles.clear();
    return 0;
}

static int session_partition_impl(b200dt_ctx *ctx, const u32 *d_mask) {
    NvtxRange nvtx("b200dt: level partition");
    Session *s = ctx->session;
    if (!s) return ctx->fail_msg("session: not begun");
    if (s->plan_segs.empty()) return 0;
    const size_t mask_bytes = (size_t)((s->N + 31) / 32) * 4;
    if (s->pending) {
        // second level of a pair: ONE 4-way partition moves the rows for both levels (children.cu)
        std::vector<Plan4> plan4;
        std::vector<TileDesc> tiles4;
        int64_t elems = 0;
        std::vector<char> is_live(s->nodes.size(), 0);
        for (int c : s->live) is_live[c] = 1;
        for (const PendParent &pp : s->pend) {
            const HNode &par = s->nodes[pp.node];
            Plan4 p4;
            p4.start = pp.start;
            p4.n = pp.n;
            p4.n_l = pp.nleft;
            p4.n_ll = p4.n_rl = 0;
            p4.pad = 0;
            bool needed = false;
            for (int side = 0; side < 2; ++side) {
                const HNode &ch = s->nodes[side == 0 ? par.left : par.right];
                // a child that is a leaf, or whose split has no grandchild to search (decide marks no rows for
                // it): all its rows keep code B = 0 and stay together, in order
                if (ch.left < 0 || !(is_live[ch.left] || is_live[ch.right])) continue;
                (side == 0 ? p4.n_ll : p4.n_rl) = s->nodes[ch.left].n;
                needed = true;
            }
            if (!needed) continue;
            const int si = (int)plan4.size();
            plan4.push_back(p4);
            for (int t = 0; t < (pp.n + P4_TILE - 1) / P4_TILE; ++t) tiles4.push_back(TileDesc{si, t});
            elems += pp.n;
        }
        if (!plan4.empty()) {
            Plan4 *d_plan4;
            TileDesc *d_tiles;
            u32 *d_codes;
            u64 *pstatus;
            CKR(scratch_get(ctx, SB_PLAN4, plan4.size() * sizeof(Plan4) + 64, (void **)&d_plan4));
            CKR(scratch_get(ctx, SB_TILES, tiles4.size() * sizeof(TileDesc) + 64, (void **)&d_tiles));
            CKR(scratch_get(ctx, SB_CODES, 2 * mask_bytes + 64, (void **)&d_codes));
            const size_t mark_bytes = (size_t)s->plan_segs.size() * 32 + 256;   // marks staged first in PIN_PLAN
            CKR(upload(ctx, PIN_PLAN, mark_bytes, plan4.data(), plan4.size(), d_plan4));
            CKR(upload(ctx, PIN_PLAN, mark_bytes + plan4.size() * sizeof(Plan4) + 64, tiles4.data(), tiles4.size(), d_tiles));
            CKR(scratch_get(ctx, SB_STATUS_P, (tiles4.size() * (size_t)s->F_store + 1) * 48, (void **)&pstatus, true));
            CKR(next_epoch(ctx));
            CKR(launch_partition4(ctx, s->ord[s->cur], s->ord[s->cur ^ 1], s->ypay[s->cur], s->ypay[s->cur ^ 1],
                                  s->col_stride, s->F_store, d_plan4, d_tiles, (int)tiles4.size(), elems, s->d_mask_a,
                                  d_mask, s->N, d_codes, pstatus, ctx->epoch));
            s->cur ^= 1;
        }
        s->pending = false;
        s->pend.clear();
        s->pend_segs.clear();
        s->pend_tiles.clear();
        s->plan_segs.clear();
        s->plan_fused.clear();
        s->plan_tiles.clear();
        return 0;
    }
    if (s->plan_two_level) {
        // first level of a pair: nothing moves; the children are searched from the parents' order
        CKR(scratch_get(ctx, SB_MASK_A, mask_bytes + 64, (void **)&s->d_mask_a));
        CK(cudaMemcpyAsync(s->d_mask_a, d_mask, mask_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        std::vector<TileDesc> stiles;
        for (size_t i = 0; i < s->plan_fused.size(); ++i)
            for (int t = 0; t < (s->plan_fused[i].n + ST_TILE - 1) / ST_TILE; ++t) stiles.push_back(TileDesc{(int)i, t});
        PlanSeg *d_plan;
        TileDesc *d_tiles;
        u64 *pstatus;
        const size_t plan_bytes = s->plan_fused.size() * sizeof(PlanSeg);
        const size_t mark_bytes = (size_t)s->plan_segs.size() * 32 + 256;
        CKR(scratch_get(ctx, SB_PLAN, plan_bytes + 64, (void **)&d_plan));
        CKR(scratch_get(ctx, SB_TILES, stiles.size() * sizeof(TileDesc) + 64, (void **)&d_tiles));
        CKR(upload(ctx, PIN_PLAN, mark_bytes, s->plan_fused.data(), s->plan_fused.size(), d_plan));
        CKR(upload(ctx, PIN_PLAN, mark_bytes + plan_bytes + 64, stiles.data(), stiles.size(), d_tiles));
        CKR(scratch_get(ctx, SB_STATUS_P, (stiles.size() * (size_t)s->F_local + 1) * 48, (void **)&pstatus, true));
        CKR(scratch_get(ctx, SB_PARTIAL, (s->next_partial_slots + 1) * s->F_local * sizeof(Partial) + 64,
                        (void **)&s->d_partials));
        CKR(next_epoch(ctx));
        CKR(launch_children_scan(ctx, s->ord[s->cur], s->ypay[s->cur], s->col_stride, s->F_local, d_plan, d_tiles,
                                 (int)stiles.size(), s->plan_elems, s->d_mask_a, pstatus, ctx->epoch, s->d_partials,
                                 s->variant));
        s->pending = true;
        s->pend = s->plan_parents;
        s->pend_segs = s->plan_segs;
        s->pend_tiles = s->plan_tiles;
        s->pend_elems = s->plan_elems;
        s->plan_two_level = false;
        s->plan_segs.clear();
        s->plan_fused.clear();
        s->plan_tiles.clear();
        return 0;
    }
    // L1: unsplit leaves of this level are only valid in the current buffer -- read them first
    if (s->criterion == B200DT_L1) CKR(l1_leaf_medians(ctx, s, s->median_pending_cur));
    SegDev *d_segs;
    TileDesc *d_tiles;
    const size_t seg_bytes = s->plan_segs.size() * sizeof(SegDev);
    const size_t mark_bytes = (size_t)s->plan_segs.size() * 12 + 64;  // marks staged first in PIN_PLAN
    CKR(scratch_get(ctx, SB_SEG, seg_bytes + 64, (void **)&d_segs));
    CKR(scratch_get(ctx, SB_TILES, s->plan_tiles.size() * sizeof(TileDesc) + 64, (void **)&d_tiles));
    CKR(upload(ctx, PIN_PLAN, mark_bytes, s->plan_segs.data(), s->plan_segs.size(), d_segs));
    CKR(upload(ctx, PIN_PLAN, mark_bytes + seg_bytes + 64, s->plan_tiles.data(), s->plan_tiles.size(), d_tiles));
    u32 *ticket = (u32 *)((char *)ctx->scratch[SB_MASK].p + (size_t)((s->N + 31) / 32) * 4);
    u64 *pstatus;
    CKR(scratch_get(ctx, SB_STATUS_P, (s->plan_tiles.size() * (size_t)s->F_store + 1) * 8, (void **)&pstatus, true));
    int *d_newid = nullptr;
    if (s->plan_relabel) {
        // new ids = the children's position ranges (relabel.cu); computed from the (reduced) side mask
        const int n_segs = (int)s->plan_segs.size();
        std::vector<int> bin_start(2 * (size_t)n_segs);
        for (int i = 0; i < n_segs; ++i) {
            bin_start[2 * i] = s->plan_segs[i].start;
            bin_start[2 * i + 1] = s->plan_segs[i].start + s->plan_segs[i].nleft;
        }
        int *d_bin_start, *d_counts, *d_map;
        unsigned short *d_bins;
        const size_t off_bins = mark_bytes + seg_bytes + 64 + s->plan_tiles.size() * sizeof(TileDesc) + 64;
        CKR(scratch_get(ctx, SB_NEWID, (size_t)s->N * 4 + 64, (void **)&d_newid));
        CKR(scratch_get(ctx, SB_ROWMAP, (size_t)s->N * 4 + 64, (void **)&d_map));
        CKR(scratch_get(ctx, SB_RL_BINS, (size_t)s->N * 2 + 64, (void **)&d_bins));
        CKR(scratch_get(ctx, SB_RL_COUNTS, ((size_t)2 * n_segs * relabel_chunks(s->N) + 2 * (size_t)n_segs) * 4 + 64,
                        (void **)&d_counts));
        d_bin_start = d_counts + (size_t)2 * n_segs * relabel_chunks(s->N);
        CKR(upload(ctx, PIN_PLAN, off_bins, bin_start.data(), bin_start.size(), d_bin_start));
        CKR(launch_relabel(ctx, s->ord[s->cur], d_segs, n_segs, d_tiles, (int)s->plan_tiles.size(), d_mask, s->N,
                           d_bin_start, d_bins, d_counts, d_newid, nullptr, d_map));
        s->d_rowmap = d_map;
    }
    CKR(next_epoch(ctx));
    CKR(launch_partition(ctx, s->ord[s->cur], s->ord[s->cur ^ 1], s->col_stride, s->F_store, d_segs, d_tiles,
                         (int)s->plan_tiles.size(), s->plan_elems, d_mask, pstatus, ctx->epoch, ticket,
                         s->ypay[s->cur], s->ypay[s->cur ^ 1], s->wpay[s->cur], s->wpay[s->cur ^ 1], 0, d_newid));
    if (s->plan_relabel) {
        s->relabeled = true;
        s->plan_relabel = false;
    }
    s->cur ^= 1;
    s->plan_segs.clear();
    s->plan_fused.clear();
    s->plan_tiles.clear();
    if (s->criterion == B200DT_L1) CKR(l1_leaf_medians(ctx, s, s->median_pending_next));
    return 0;
}

static int session_finish_impl(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right, int64_t *feature,
                               double *threshold, double *value, int32_t *n_node_samples, int64_t capacity,
                               int64_t *node_count) {
    Session *s = ctx->session;
    if (!s) return ctx->fail_msg("session: not begun");
    if (!s->live.empty()) return ctx->fail_msg("session: finish called before the tree is complete");
    if (s->criterion == B200DT_L1) {
        // no partition followed the last decisions: everything pending lives in the current buffer
        CKR(l1_leaf_medians(ctx, s, s->median_pending_cur));
        CKR(l1_leaf_medians(ctx, s, s->median_pending_next));
    }
    const int64_t count = (int64_t)s->nodes.size();
    if (count > capacity) return ctx->fail_msg("fit: tree arrays too small");
    std::vector<int> new_id(count, -1);
    if (s->flags & B200DT_NUMBER_BY_LEVEL) {
        // four.py:113-170: a level is processed as [left children of the previous level's split nodes,
        // in their order] + [their right children]; ids are handed out in processing order
        std::vector<int> level(1, 0), next_level;
        int next = 0;
        while (!level.empty()) {
            for (int v : level) new_id[v] = next++;
            next_level.clear();
            for (int v : level)
                if (!s->nodes[v].leaf) next_level.push_back(s->nodes[v].left);
            for (int v : level)
                if (!s->nodes[v].leaf) next_level.push_back(s->nodes[v].right);
            level.swap(next_level);
        }
    } else {
        // pre-order ids, left first (one.py:26-31: ids are handed out as tasks are popped;
        // one.py:157-160: right is pushed first so left is popped first)
        std::vector<int> stack;
        stack.push_back(0);
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
    }
    for (int64_t v = 0; v < count; ++v) {
        const HNode &nd = s->nodes[v];
        const int id = new_id[v];
        if (nd.leaf) {
            children_left[id] = -1;
            children_right[id] = -1;
            feature[id] = -2;
            threshold[id] = -2.0;
        } else {
            children_left[id] = new_id[nd.left];
            children_right[id] = new_id[nd.right];
            feature[id] = nd.feature;
            threshold[id] = nd.thr;
        }
        value[id] = nd.value;
        if (n_node_samples) n_node_samples[id] = nd.n;
    }
    *node_count = count;
    return 0;
}

// accessors used by l1.cu
int session_cur_orders(Session *s, const int **orders, int64_t *col_stride) {
    *orders = s->ord[s->cur];
    *col_stride = s->col_stride;
    return 0;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static std::string g_create_error;

extern "C" {

int b200dt_abi_version(void) { return B200DT_ABI_VERSION; }

int b200dt_create(b200dt_ctx **out, int device_id) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("b200dt_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); this library has no CPU fallback";
        return 1;
    }
    if (device_id < 0 || device_id >= n) {
        g_create_error = "b200dt_create: device id out of range";
        return 1;
    }
    e = cudaSetDevice(device_id);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    b200dt_ctx *ctx = new (std::nothrow) b200dt_ctx();
    if (!ctx) {
        g_create_error = "out of host memory";
        return 1;
    }
    ctx->device = device_id;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_id);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        delete ctx;
        return 1;
    }
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete ctx;
        return 1;
    }
    ctx->stream = ctx->own_stream;
    ctx->scratch.resize(SB_COUNT);
    *out = ctx;
    return 0;
}

void b200dt_destroy(b200dt_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    prof_drain(ctx);
    session_free(ctx);
    class_session_free(ctx);
    row_session_free(ctx);
    for (auto &b : ctx->scratch)
        if (b.p) cudaFree(b.p);
    for (auto p : ctx->pinned)
        if (p) cudaFreeHost(p);
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    for (auto e : ctx->copy_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *b200dt_last_error(const b200dt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b200dt_set_stream(b200dt_ctx *ctx, void *cuda_stream) {
    // NULL is CUDA's legacy default stream (what torch.cuda.current_stream() is unless the caller
    // switched streams); (void*)-1 selects the context's own non-blocking stream again
    ctx->stream = cuda_stream == (void *)-1 ? ctx->own_stream : (cudaStream_t)cuda_stream;
    return 0;
}

int b200dt_synchronize(b200dt_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dt_session_begin(b200dt_ctx *ctx, const double *X, const double *y, int64_t N, int64_t F_local, int64_t ldx,
                         int space, int criterion, int max_depth, double min_improvement, int64_t min_sample_leaf,
                         int variant, int64_t col_offset, int64_t F_global, const double *last_col,
                         int64_t last_col_ld) {
    CK(cudaSetDevice(ctx->device));
    return session_begin_impl(ctx, X, y, N, F_local, ldx, space, criterion, max_depth, min_improvement,
                              min_sample_leaf, variant, col_offset, F_global, last_col, last_col_ld);
}

int b200dt_session_live(b200dt_ctx *ctx, int64_t *n_live) {
    if (!ctx->session) return ctx->fail_msg("session: not begun");
    *n_live = (int64_t)ctx->session->live.size();
    return 0;
}

int b200dt_session_search(b200dt_ctx *ctx, b200dt_candidate *d_out) {
    CK(cudaSetDevice(ctx->device));
    return session_search_impl(ctx, d_out, false);
}

int b200dt_session_search_exact(b200dt_ctx *ctx, b200dt_candidate *d_out) {
    CK(cudaSetDevice(ctx->device));
    return session_search_impl(ctx, d_out, true);
}

int b200dt_session_decide(b200dt_ctx *ctx, const b200dt_candidate *d_gathered, int n_ranks, uint32_t *d_side_mask,
                          int *need_partition, int *need_exact) {
    CK(cudaSetDevice(ctx->device));
    return session_decide_impl(ctx, d_gathered, n_ranks, d_side_mask, need_partition, need_exact);
}

int b200dt_session_partition(b200dt_ctx *ctx, const uint32_t *d_side_mask) {
    CK(cudaSetDevice(ctx->device));
    return session_partition_impl(ctx, d_side_mask);
}

int b200dt_session_finish(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right, int64_t *feature,
                          double *threshold, double *value, int64_t capacity, int64_t *node_count) {
    CK(cudaSetDevice(ctx->device));
    int r = session_finish_impl(ctx, children_left, children_right, feature, threshold, value, nullptr, capacity,
                                node_count);
    session_free(ctx);
    return r;
}

int b200dt_fit_ex(b200dt_ctx *ctx, const double *X, const double *y, const double *weight, int64_t N, int64_t F,
                  int64_t ldx, int space, int criterion, int max_depth, double min_improvement,
                  int64_t min_sample_leaf, int variant, int flags,
                  int64_t *children_left, int64_t *children_right, int64_t *feature, double *threshold, double *value,
                  int32_t *n_node_samples, int64_t capacity, int64_t *node_count) {
    CK(cudaSetDevice(ctx->device));
    CKR(session_begin_impl(ctx, X, y, N, F, ldx, space, criterion, max_depth, min_improvement, min_sample_leaf,
                           variant, 0, F, nullptr, 0, weight));
    Session *s = ctx->session;
    s->flags = flags;
    while (!s->live.empty()) {
        const int n_live = (int)s->live.size();
        CKR(scratch_get(ctx, SB_RESULT, (size_t)n_live * sizeof(b200dt_candidate) + 64, (void **)&s->d_cand));
        CKR(session_search_impl(ctx, s->d_cand, false));
        int need_part = 0, need_exact = 0;
        CKR(session_decide_impl(ctx, s->d_cand, 1, s->d_own_mask, &need_part, &need_exact));
        if (need_exact) {
            CKR(session_search_impl(ctx, s->d_cand, true));
            CKR(session_decide_impl(ctx, s->d_cand, 1, s->d_own_mask, &need_part, &need_exact));
            if (need_exact) return ctx->fail_msg("fit: exact re-search did not resolve a tie");
        }
        if (need_part) CKR(session_partition_impl(ctx, s->d_own_mask));
    }
    int r = session_finish_impl(ctx, children_left, children_right, feature, threshold, value, n_node_samples,
                                capacity, node_count);
    session_free(ctx);
    return r;
}

int b200dt_fit(b200dt_ctx *ctx, const double *X, const double *y, int64_t N, int64_t F, int64_t ldx, int space,
               int criterion, int max_depth, double min_improvement, int64_t min_sample_leaf, int variant,
               int64_t *children_left, int64_t *children_right, int64_t *feature, double *threshold, double *value,
               int64_t capacity, int64_t *node_count) {
    return b200dt_fit_ex(ctx, X, y, nullptr, N, F, ldx, space, criterion, max_depth, min_improvement, min_sample_leaf,
                         variant, 0, children_left, children_right, feature, threshold, value, nullptr, capacity,
                         node_count);
}

// ---- single-node ops (parity tests of K1 / K2 / K4) --------------------------------------
int b200dt_argsort_columns(b200dt_ctx *ctx, const double *X, int64_t N, int64_t F, int64_t ldx, int x_space,
                           int64_t *orders_out, int out_space) {
    CK(cudaSetDevice(ctx->device));
    if (N <= 0 || F <= 0) return 0;
    const double *dX = X;
    int64_t ld = ldx;
    if (x_space == B200DT_HOST) {
        double *buf;
        CKR(scratch_get(ctx, SB_X, (size_t)N * F * 8, (void **)&buf));
        CK(cudaMemcpy2DAsync(buf, (size_t)F * 8, X, (size_t)ldx * 8, (size_t)F * 8, (size_t)N, cudaMemcpyHostToDevice,
                             ctx->stream));
        dX = buf;
        ld = F;
    }
    const int64_t stride = (N + 31) / 32 * 32;
    int *a, *b;
    CKR(scratch_get(ctx, SB_ORD_A, (size_t)F * stride * 4, (void **)&a));
    CKR(scratch_get(ctx, SB_ORD_B, (size_t)F * stride * 4, (void **)&b));
    CKR(sort_columns_device(ctx, dX, N, F, ld, a, b, stride));
    int64_t *d_out = orders_out;
    if (out_space == B200DT_HOST) CKR(scratch_get(ctx, SB_OUT, (size_t)N * F * 8, (void **)&d_out));
    {
        LaunchScope ls(ctx, KC_MISC, 12.0 * (double)N * F);
        dim3 grid((unsigned)((N + 31) / 32), (unsigned)((F + 31) / 32));
        orders_to_rowmajor_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(a, stride, N, F, d_out);
    }
    CK(cudaGetLastError());
    if (out_space == B200DT_HOST)
        CK(cudaMemcpyAsync(orders_out, d_out, (size_t)N * F * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// stage one node given as a host (n, F) int64 order matrix into a throw-away session
static int stage_single_node(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F, const double *X,
                             int64_t ldx, const double *y, int64_t N, int criterion, int variant) {
    if (n <= 0 || F <= 0 || N <= 0) return ctx->fail_msg("split_node: empty input");
    end_all_sessions(ctx);
    Session *s = new (std::nothrow) Session();
    if (!s) return ctx->fail_msg("out of host memory");
    ctx->session = s;
    s->N = N;
    s->F_local = s->F_global = (int)F;
    s->F_store = (int)F;
    s->total_col = (int)F - 1;
    s->criterion = criterion;
    s->variant = variant;
    s->max_depth = 1;
    s->col_stride = (n + 31) / 32 * 32;
    if (criterion == B200DT_L1) {  // extra column: the node's rows sorted by y (see l1.cu)
        s->ycol = (int)F;
        s->F_store = (int)F + 1;
    }
    std::vector<int> h((size_t)s->F_store * s->col_stride, 0);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t f = 0; f < F; ++f) {
            const int64_t r = orders[i * F + f];
            if (r < 0 || r >= N) return ctx->fail_msg("split_node: order entry out of range");
            h[(size_t)f * s->col_stride + i] = (int)r;
        }
    if (criterion == B200DT_L1) {
        int *yc = h.data() + (size_t)F * s->col_stride;
        for (int64_t i = 0; i < n; ++i) yc[i] = h[i];
        std::stable_sort(yc, yc + n, [&](int a, int b) { return y[a] < y[b]; });
    }
    CKR(scratch_get(ctx, SB_ORD_A, h.size() * 4, (void **)&s->ord[0]));
    CKR(scratch_get(ctx, SB_ORD_B, h.size() * 4, (void **)&s->ord[1]));
    CK(cudaMemcpyAsync(s->ord[0], h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    double *dX = nullptr, *dy;
    CKR(scratch_get(ctx, SB_Y, (size_t)N * 8, (void **)&dy));
    CK(cudaMemcpyAsync(dy, y, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (X) {
        CKR(scratch_get(ctx, SB_X, (size_t)N * F * 8, (void **)&dX));
        CK(cudaMemcpy2DAsync(dX, (size_t)F * 8, X, (size_t)ldx * 8, (size_t)F * 8, (size_t)N, cudaMemcpyHostToDevice,
                             ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));  // h goes out of scope
    s->dX = dX;
    s->ldx = F;
    s->dy = dy;
    CKR(scratch_get(ctx, SB_MASK, (size_t)((N + 31) / 32) * 4 + 64, (void **)&s->d_own_mask));
    HNode root;
    root.start = 0;
    root.n = (int)n;
    s->nodes.push_back(root);
    s->live.push_back(0);
    s->force_exact.assign(1, 0);
    return 0;
}

static int single_node_search(b200dt_ctx *ctx, int64_t *out_k, int64_t *out_f, double *out_score,
                              double *out_threshold) {
    Session *s = ctx->session;
    // the node total for the approximate path: sum of y over the node's rows (column 0 order)
    {
        // gather-sum on the host side of the staged rows is avoided: use the exact kernel's total
        s->force_exact[0] = (s->nodes[0].n <= EXACT_MAX_ROWS) ? 1 : 0;
    }
    b200dt_candidate *d_c;
    CKR(scratch_get(ctx, SB_RESULT, sizeof(b200dt_candidate) + 64, (void **)&d_c));
    b200dt_candidate h;
    for (int round = 0; round < 2; ++round) {
        CKR(session_search_impl(ctx, d_c, false));
        CK(cudaMemcpyAsync(&h, d_c, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (h.exact || !(h.gain2 >= h.gain - NEAR_TIE_EPS * std::fabs(h.gain))) break;
        s->force_exact[0] = 1;
    }
    *out_k = h.k;
    *out_f = h.f;
    *out_score = h.gain;
    *out_threshold = h.threshold;
    return 0;
}

int b200dt_l2_split_node(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F, const double *X, int64_t ldx,
                         const double *y, int64_t N, int variant, int64_t *out_k, int64_t *out_f, double *out_gain,
                         double *out_threshold) {
    CK(cudaSetDevice(ctx->device));
    if (n < 2) return ctx->fail_msg("l2_split_node: attempt to get argmax of an empty sequence (n < 2)");
    CKR(stage_single_node(ctx, orders, n, F, X, ldx, y, N, B200DT_L2, variant));
    Session *s = ctx->session;
    // approximate path needs the node total: sum y over the node's rows
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) total += y[orders[i * F + (F - 1)]];
    s->nodes[0].total = total;
    double score = 0.0;
    int r = single_node_search(ctx, out_k, out_f, &score, out_threshold);
    *out_gain = (variant == 1) ? score : score * score;
    session_free(ctx);
    return r;
}

int b200dt_l1_split_node(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F, const double *X, int64_t ldx,
                         const double *y, int64_t N, int64_t *out_k, int64_t *out_f, double *out_cost,
                         double *out_threshold) {
    CK(cudaSetDevice(ctx->device));
    if (n < 2) return ctx->fail_msg("l1_split_node: attempt to get argmin of an empty sequence (n < 2)");
    CKR(stage_single_node(ctx, orders, n, F, X, ldx, y, N, B200DT_L1, 1));
    double score = 0.0;
    int r = single_node_search(ctx, out_k, out_f, &score, out_threshold);
    *out_cost = -score;
    session_free(ctx);
    return r;
}

// replaces: best_l1_improvements[_v2](ys) as a stand-alone function of an (n, F) matrix whose columns need not be
// permutations of one another                                                     ref strategy_l1.py:49-122
// Every column is its own one-feature problem; all F of them are searched in ONE level of a throw-away session:
// row f*n + i holds ys[i, f], node f is the segment [f*n, (f+1)*n) of a single feature column in identity order.
int b200dt_l1_best_columns(b200dt_ctx *ctx, const double *ys, int64_t n, int64_t F, int64_t *out_k, int64_t *out_f,
                           double *out_cost) {
    if (!ctx) return 1;
    if (!ys || !out_k || !out_f || !out_cost) return ctx->fail_msg("l1_best_columns: null argument");
    if (n < 2 || F < 1) return ctx->fail_msg("l1_best_columns: attempt to get argmin of an empty sequence (n < 2)");
    const int64_t NN = n * F;
    if (NN >= (1ll << 30)) return ctx->fail_msg("l1_best_columns: n * F must be < 2^30");
    CK(cudaSetDevice(ctx->device));
    end_all_sessions(ctx);
    Session *s = new (std::nothrow) Session();
    if (!s) return ctx->fail_msg("out of host memory");
    ctx->session = s;
    s->N = NN;
    s->F_local = s->F_global = 1;
    s->F_store = 2;
    s->total_col = 0;
    s->ycol = 1;
    s->criterion = B200DT_L1;
    s->variant = 1;
    s->max_depth = 1;
    s->col_stride = (NN + 31) / 32 * 32;
    std::vector<double> yy((size_t)NN);
    std::vector<int> h((size_t)2 * s->col_stride, 0);
    for (int64_t f = 0; f < F; ++f) {
        double *yc = yy.data() + f * n;
        for (int64_t i = 0; i < n; ++i) yc[i] = ys[i * F + f];
        int *idc = h.data() + f * n, *ysorted = h.data() + s->col_stride + f * n;
        for (int64_t i = 0; i < n; ++i) idc[i] = ysorted[i] = (int)(f * n + i);
        std::stable_sort(ysorted, ysorted + n, [&](int a, int b) { return yy[a] < yy[b]; });
    }
    CKR(scratch_get(ctx, SB_ORD_A, h.size() * 4, (void **)&s->ord[0]));
    CKR(scratch_get(ctx, SB_ORD_B, h.size() * 4, (void **)&s->ord[1]));
    double *dy;
    CKR(scratch_get(ctx, SB_Y, (size_t)NN * 8, (void **)&dy));
    CK(cudaMemcpyAsync(s->ord[0], h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy, yy.data(), (size_t)NN * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s->dX = nullptr;
    s->ldx = 1;
    s->dy = dy;
    CKR(scratch_get(ctx, SB_MASK, (size_t)((NN + 31) / 32) * 4 + 64, (void **)&s->d_own_mask));
    for (int64_t f = 0; f < F; ++f) {
        HNode nd;
        nd.start = (int)(f * n);
        nd.n = (int)n;
        s->nodes.push_back(nd);
        s->live.push_back((int)f);
    }
    s->force_exact.assign(s->live.size(), 0);
    s->known_total.assign(s->live.size(), 0.0);
    s->has_total.assign(s->live.size(), 0);
    b200dt_candidate *d_c;
    CKR(scratch_get(ctx, SB_RESULT, (size_t)F * sizeof(b200dt_candidate) + 64, (void **)&d_c));
    std::vector<b200dt_candidate> hc((size_t)F);
    for (int round = 0; round < 2; ++round) {
        CKR(session_search_impl(ctx, d_c, round == 1));
        CK(cudaMemcpyAsync(hc.data(), d_c, (size_t)F * sizeof(b200dt_candidate), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        bool again = false;
        for (int64_t f = 0; f < F; ++f) {
            const b200dt_candidate &c = hc[f];
            if (!c.exact && c.gain2 >= c.gain - NEAR_TIE_EPS * std::fabs(c.gain)) {
                s->force_exact[f] = 1;   // a near tie of a chunked (parallel-order) cost sum: redo sequentially
                again = true;
            }
        }
        if (!again) break;
    }
    // flat row-major argmin of the reference (strategy_l1.py:120-121): lowest cost, then lowest row, then lowest column
    int64_t bf = -1;
    for (int64_t f = 0; f < F; ++f) {
        if (hc[f].k < 0) continue;
        if (bf < 0 || hc[f].gain > hc[bf].gain || (hc[f].gain == hc[bf].gain && hc[f].k < hc[bf].k)) bf = f;
    }
    if (bf < 0) {
        session_free(ctx);
        return ctx->fail_msg("l1_best_columns: no candidate");
    }
    *out_k = hc[bf].k;
    *out_f = bf;
    *out_cost = -hc[bf].gain;
    session_free(ctx);
    return 0;
}

int b200dt_split_orders(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F, const int64_t *left_rows,
                        int64_t n_left, int64_t n_samples, int64_t *left_out, int64_t *right_out) {
    CK(cudaSetDevice(ctx->device));
    if (n_left < 0 || n_left > n) return ctx->fail_msg("split_orders: bad n_left");
    std::vector<double> ydummy((size_t)n_samples, 0.0);
    CKR(stage_single_node(ctx, orders, n, F, nullptr, 0, ydummy.data(), n_samples, B200DT_L2, 1));
    Session *s = ctx->session;
    // side mask from the explicit row list (one.py:111-112)
    std::vector<u32> hmask((size_t)(n_samples + 31) / 32, 0u);
    for (int64_t i = 0; i < n_left; ++i) {
        const int64_t r = left_rows[i];
        if (r < 0 || r >= n_samples) return ctx->fail_msg("split_orders: left row out of range");
        hmask[r >> 5] |= 1u << (r & 31);
    }
    CK(cudaMemcpyAsync(s->d_own_mask, hmask.data(), hmask.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SegDev sg;
    sg.total = 0.0;
    sg.total_w = 0.0;
    sg.start = 0;
    sg.n = (int)n;
    sg.pbase = 0;
    sg.np = (int)((n + PT_TILE - 1) / PT_TILE);
    sg.nleft = (int)n_left;
    sg.exact = 0;
    sg.pstep = 1;
    sg.pside = 0;
    sg.vstart = -1;
    sg.cbase = 0;
    s->plan_segs.assign(1, sg);
    s->plan_tiles.clear();
    for (int t = 0; t < sg.np; ++t) s->plan_tiles.push_back(TileDesc{0, t});
    s->plan_elems = n;
    s->criterion = B200DT_L2;
    CKR(session_partition_impl(ctx, s->d_own_mask));
    std::vector<int> h((size_t)F * s->col_stride);
    CK(cudaMemcpyAsync(h.data(), s->ord[s->cur], h.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t n_right = n - n_left;
    for (int64_t f = 0; f < F; ++f) {
        const int *col = h.data() + (size_t)f * s->col_stride;
        for (int64_t i = 0; i < n_left; ++i) left_out[i * F + f] = col[i];
        for (int64_t i = 0; i < n_right; ++i) right_out[i * F + f] = col[n_left + i];
    }
    session_free(ctx);
    return 0;
}

int b200dt_cummedian(b200dt_ctx *ctx, const double *a, int64_t n, int a_space, double *out, int out_space) {
    CK(cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    const double *d_a = a;
    double *d_out = out;
    if (a_space == B200DT_HOST) {
        double *buf;
        CKR(scratch_get(ctx, SB_Y, (size_t)n * 8, (void **)&buf));
        CK(cudaMemcpyAsync(buf, a, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_a = buf;
    }
    if (out_space == B200DT_HOST) CKR(scratch_get(ctx, SB_OUT, (size_t)n * 16, (void **)&d_out));
    CKR(l1_cummedian_device(ctx, d_a, n, d_out));
    if (out_space == B200DT_HOST)
        CK(cudaMemcpyAsync(out, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200dt_predict(b200dt_ctx *ctx, const double *X, int64_t N, int64_t F, int64_t ldx, int x_space,
                   const int64_t *feature, const double *threshold, const int64_t *children_left,
                   const int64_t *children_right, const double *value, int64_t n_nodes, double *out_value,
                   int32_t *out_leaf, int out_space) {
    NvtxRange nvtx("b200dt: predict");
    CK(cudaSetDevice(ctx->device));
    if (N <= 0) return 0;
    if (n_nodes <= 0 || n_nodes > (1ll << 30)) return ctx->fail_msg("predict: bad node count");
    if (ctx->session || ctx->class_session || ctx->row_session)
        return ctx->fail_msg("predict: a fit session is in progress on this context (they share its scratch buffers); "
                             "finish it first or use another context");
    const double *dX = X;
    int64_t ld = ldx;
    if (x_space == B200DT_HOST) {
        double *buf;
        CKR(scratch_get(ctx, SB_X, (size_t)N * F * 8, (void **)&buf));
        CK(cudaMemcpy2DAsync(buf, (size_t)F * 8, X, (size_t)ldx * 8, (size_t)F * 8, (size_t)N, cudaMemcpyHostToDevice,
                             ctx->stream));
        dX = buf;
        ld = F;
    }
    // tree -> device, int32 node links; longest root-to-leaf path for the traffic estimate
    const size_t nn = (size_t)n_nodes;
    char *stage;
    const size_t tree_bytes = nn * (8 + 8 + 4 + 4 + 4) + 64;
    CKR(pinned_get(ctx, PIN_MISC, tree_bytes, (void **)&stage));
    double *h_thr = (double *)stage, *h_val = h_thr + nn;
    int *h_feat = (int *)(h_val + nn), *h_left = h_feat + nn, *h_right = h_left + nn;
    for (size_t i = 0; i < nn; ++i) {
        const int64_t l = children_left[i], r = children_right[i], f = feature[i];
        if (l < -1 || l >= n_nodes || r < -1 || r >= n_nodes) return ctx->fail_msg("predict: child id out of range");
        if (f >= F || f < -F) return ctx->fail_msg("predict: feature index out of range");
        h_thr[i] = threshold[i];
        h_val[i] = value[i];
        h_feat[i] = (int)f;
        h_left[i] = (int)l;
        h_right[i] = (int)r;
    }
    std::vector<int> depth(nn, 0);
    double path_sum = 0.0;
    int64_t leaves = 0;
    for (size_t i = 0; i < nn; ++i) {  // parents precede children in both pre-order and BFS ids
        if (h_left[i] > (int)i) depth[h_left[i]] = depth[i] + 1;
        if (h_right[i] > (int)i) depth[h_right[i]] = depth[i] + 1;
        if (h_left[i] == -1 && h_right[i] == -1 && (i == 0 || depth[i] > 0)) {
            path_sum += depth[i];
            ++leaves;
        }
    }
    const double avg_path = leaves ? path_sum / (double)leaves : 0.0;
    char *d_tree;
    CKR(scratch_get(ctx, SB_TREE, tree_bytes, (void **)&d_tree));
    CK(cudaMemcpyAsync(d_tree, stage, tree_bytes, cudaMemcpyHostToDevice, ctx->stream));
    double *d_thr = (double *)d_tree, *d_val = d_thr + nn;
    int *d_feat = (int *)(d_val + nn), *d_left = d_feat + nn, *d_right = d_left + nn;
    double *d_ov = out_value;
    int *d_ol = out_leaf;
    if (out_space == B200DT_HOST) {
        if (out_value) CKR(scratch_get(ctx, SB_OUT, (size_t)N * 8, (void **)&d_ov));
        if (out_leaf) CKR(scratch_get(ctx, SB_LEAF, (size_t)N * 4, (void **)&d_ol));
    }
    CKR(predict_device(ctx, dX, N, F, ld, d_feat, d_left, d_right, d_thr, d_val, (int)n_nodes, d_ov, d_ol, avg_path));
    if (out_space == B200DT_HOST) {
        if (out_value) CK(cudaMemcpyAsync(out_value, d_ov, (size_t)N * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_leaf) CK(cudaMemcpyAsync(out_leaf, d_ol, (size_t)N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- measurement ----------------------------------------------------------------------------
int b200dt_profile_enable(b200dt_ctx *ctx, int on) {
    ctx->prof_on = on != 0;
    return 0;
}

int b200dt_profile_reset(b200dt_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    prof_drain(ctx);
    for (int i = 0; i < KC_COUNT; ++i) {
        ctx->prof_ms[i] = 0.0;
        ctx->prof_launches[i] = 0;
        ctx->prof_bytes[i] = 0.0;
    }
    ctx->launches = 0;
    return 0;
}

int b200dt_profile_count(void) { return KC_COUNT; }

const char *b200dt_profile_name(int kc) { return (kc >= 0 && kc < KC_COUNT) ? kKernelClassNames[kc] : ""; }

int b200dt_profile_get(b200dt_ctx *ctx, int kc, double *ms, int64_t *launches, double *alg_bytes) {
    if (kc < 0 || kc >= KC_COUNT) return ctx->fail_msg("profile: bad kernel class");
    CK(cudaSetDevice(ctx->device));
    prof_drain(ctx);
    *ms = ctx->prof_ms[kc];
    *launches = ctx->prof_launches[kc];
    *alg_bytes = ctx->prof_bytes[kc];
    return 0;
}

int64_t b200dt_launch_count(const b200dt_ctx *ctx) { return ctx->launches; }

}  // extern "C"
