// Host-side state of one tree-growing session (shared by fit.cu and l1.cu).
#pragma once
#include <vector>

#include "tree_kernels.cuh"

struct HNode {
    int parent = -1;
    int is_left = 1;
    int depth = 0;
    int start = 0;
    int n = 0;
    double total = 0.0;
    double total_w = 0.0;   // weighted fits: sum of the weights of the node (total is then sum of y*w)
    int feature = -2;
    double thr = -2.0;
    int left = -1, right = -1;
    double value = 0.0;
    bool leaf = false;
    // candidates already produced for this node by the fused level kernel of its parent:
    // slots fp_base + 2 * tile + fp_side, tile < fp_np
    bool has_fused = false;
    int fp_base = 0, fp_np = 0, fp_side = 0;
    int vstart = -1;   // >= 0: this node's rows still lie inside its parent's segment, which starts here
};

// A split node of the first level of a pair whose children are searched before any data moves (children.cu)
struct PendParent {
    int node;          // index into Session::nodes
    int start, n, nleft;
};

struct Session {
    int64_t N = 0;
    int F_local = 0, F_store = 0, F_global = 0, col_offset = 0, total_col = 0, ycol = -1;
    int criterion = B200DT_L2, max_depth = 0, variant = 1;
    double min_improvement = 0.0;
    int64_t min_sample_leaf = 1;
    const double *dX = nullptr, *dy = nullptr;
    int64_t ldx = 0;
    int *ord[2] = {nullptr, nullptr};
    double *ypay[2] = {nullptr, nullptr};  // y travelling with the row ids (L2 fits); null = gather path
    const double *dw = nullptr;            // sample weights (four.py); dy is then y*w
    double *wpay[2] = {nullptr, nullptr};  // the weights travelling like ypay
    int cur = 0;
    int64_t col_stride = 0;
    std::vector<HNode> nodes;
    std::vector<int> live;         // node ids awaiting a search at the current depth
    std::vector<char> force_exact;  // per live node: re-search sequentially
    std::vector<SegDev> h_segs;    // segments of the live nodes (as uploaded)
    int depth = 0;
    // device state of the current level
    SegDev *d_segs = nullptr;
    Partial *d_partials = nullptr;
    double *d_exact_total = nullptr;
    b200dt_candidate *d_cand = nullptr;  // single-GPU candidate buffer
    u32 *d_own_mask = nullptr;
    // partition plan produced by decide
    std::vector<SegDev> plan_segs;
    std::vector<PlanSeg> plan_fused;   // same nodes, for the fused level kernel (payload sessions)
    std::vector<TileDesc> plan_tiles;
    int flags = 0;             // B200DT_NUMBER_BY_LEVEL | B200DT_LEAF_IF_LE (four.py conventions)
    int fused_slots = 0;       // candidate slots of the CURRENT level written by the fused kernel
    int next_fused_slots = 0;  // ... of the level being planned
    size_t next_partial_slots = 0;  // candidate slots the next level needs in total
    int64_t plan_elems = 0;
    bool exact_round_done = false;
    // row relabelling (relabel.cu): once per fit, at the first level whose children all fit an L1-resident
    // window of the side mask, the partition hands out node-contiguous row ids
    // feature-sharded L2 fits: the node totals the sequential-order search needs (the reference sums a node in the
    // order of the LAST global column, one.py:57) are computed by the rank that owns that column and travel inside
    // the candidate records (exact == 2) -- no rank keeps a shadow copy of the column
    bool sharded = false;
    std::vector<double> known_total;   // per live node
    std::vector<char> has_total;
    bool scan_bulk = false;          // B200DT_SCAN_BULK=1: the TMA double-buffered split search (l2.cu), shifted tile grid
    int64_t relabel_window = 0;      // rows; 0 = never relabel
    bool plan_relabel = false;       // the plan just decided relabels
    bool relabeled = false;
    int *d_rowmap = nullptr;         // row id -> original row, once relabelled
    // two levels per data movement (children.cu), B200DT_TWO_LEVEL
    bool two_level = false;        // enabled for this session
    bool plan_two_level = false;   // the plan just decided: search the children from the parents' order, move nothing yet
    bool pending = false;          // the live nodes' rows still lie inside their parents' segments
    std::vector<PendParent> plan_parents;   // parents of the plan being decided (same order as plan_segs)
    std::vector<PendParent> pend;           // ... of the pending level
    std::vector<SegDev> pend_segs;          // their binary-partition plan, kept for materialisation on a near tie
    std::vector<TileDesc> pend_tiles;
    int64_t pend_elems = 0;
    u32 *d_mask_a = nullptr;                // side mask of the first level of the pair
    // L1 leaves whose median is still to be read from the y-order column: those that live in the
    // current order buffer (unsplit nodes) and those created by the partition being applied
    std::vector<int> median_pending_cur, median_pending_next;
};


enum PinSlot { PIN_SEARCH = 0, PIN_PLAN, PIN_RESULT, PIN_MISC, PIN_L1, PIN_CLASS, PIN_WALK, PIN_BOUNCE };

void class_session_free(b200dt_ctx *ctx);  // classes.cu
void row_session_free(b200dt_ctx *ctx);    // rows.cu
void tree_session_free(b200dt_ctx *ctx);   // fit.cu
// The three session kinds (and predict) share the context's scratch buffers, so at most ONE may be live:
// beginning any kind ends whatever was in progress on the context.
static inline void end_all_sessions(b200dt_ctx *ctx) {
    tree_session_free(ctx);
    class_session_free(ctx);
    row_session_free(ctx);
}
