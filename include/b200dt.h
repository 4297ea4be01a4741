/* b200dt.h -- C ABI of libb200tree.so: the B200 (sm_100a) regression-tree hot path.
 *
 * This is the drop-in boundary for the ONE hot path of yupbank/np_decision_tree:
 * exhaustive L2 / L1 split search, node partition and batched inference.  The
 * reference has no FFI of its own (plain Python functions + one Cython module),
 * so each entry point cites the reference function(s) it replaces, as
 * path:line below /root/reference/.  The Python mirror of the reference API that
 * binds these symbols with ctypes is np_decision_tree_b200/decision_tree/.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     b200dt_last_error(ctx) (ctx may be NULL for errors of b200dt_create).
 *   - `space` arguments: B200DT_HOST pointers are plain host memory (numpy),
 *     B200DT_DEVICE pointers are device memory on the context's GPU (e.g. a
 *     torch CUDA tensor's data_ptr()).  Caller buffers are never freed or written
 *     unless documented as outputs.
 *   - there is NO CPU fallback: without a usable CUDA device b200dt_create fails.
 *   - a context is not re-entrant; one context = one host thread, one GPU, one stream.
 *   - row ids are int32 on the device (N < 2^31, like ref utils.py:19 / pyx:9).
 */
#ifndef B200DT_H
#define B200DT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DT_ABI_VERSION 1

#define B200DT_HOST 0
#define B200DT_DEVICE 1

#define B200DT_L2 0 /* MSE: ref decision_tree/one.py                  */
#define B200DT_L1 1 /* MAE: ref decision_tree/strategy_l1.py          */
#define B200DT_GINI 2   /* gini:           ref decision_tree/strategy.py:32-55,272-283  */
#define B200DT_P_AT_1 3 /* precision at 1: ref decision_tree/strategy.py:79-100,313-323 */

typedef struct b200dt_ctx b200dt_ctx;

/* ---- context ---------------------------------------------------------------- */
int b200dt_abi_version(void);
int b200dt_create(b200dt_ctx **out, int device_id);
void b200dt_destroy(b200dt_ctx *ctx);
const char *b200dt_last_error(const b200dt_ctx *ctx);
/* Launch on this cudaStream_t (e.g. torch's current stream) instead of the context's own
 * non-blocking stream.  NULL is CUDA's legacy default stream; (void*)-1 selects the own stream. */
int b200dt_set_stream(b200dt_ctx *ctx, void *cuda_stream);
int b200dt_synchronize(b200dt_ctx *ctx);

/* ---- K1: per-column argsort --------------------------------------------------
 * replaces: np.argsort(X, axis=0)     ref one.py:139, strategy_l1.py:147,177
 * X is (N, F) float64, row stride ldx elements.  Stable LSD radix sort on the
 * order-preserving uint64 image of the key (== np.argsort(kind='stable'); equal to
 * the reference's default kind whenever a column has distinct values).
 * orders_out: (N, F) int64 row-major, like the reference's `orders`. */
int b200dt_argsort_columns(b200dt_ctx *ctx, const double *X, int64_t N, int64_t F, int64_t ldx,
                           int x_space, int64_t *orders_out, int out_space);

/* ---- K2: L2 split search of ONE node -----------------------------------------
 * replaces: greedy_regression_split(orders, x, y, v)            ref one.py:96-107
 *           best_variance_improvements / _v2 + np_gene          ref one.py:52-63,81-93
 * orders: (n, F) int64 row-major host array of row ids into X / y (N rows).
 * variant: 1 = `v == 1`, anything else = the float32-ratio search (`v != 1`).
 * outputs: best sorted position k (left child = first k+1 rows of column f),
 * feature f, gain, threshold = X[orders[k,f], f].  Tie rule: max gain, then
 * lowest k, then lowest f (one.py:59,62). */
int b200dt_l2_split_node(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F,
                         const double *X, int64_t ldx, const double *y, int64_t N, int variant,
                         int64_t *out_k, int64_t *out_f, double *out_gain, double *out_threshold);

/* replaces: one.best_variance_improvements(ys) / _v0 / _v2 as stand-alone functions   ref one.py:56-63,66-78,81-93
 * cumsums: (n, F) float64 row-major HOST array of column-wise running sums (np.cumsum(y[orders], axis=0)).
 * variant 1: row-wise argmax of |S - (k+1) mean| over the features, then argmax over the rows of
 *            np_gene(n) * value^2 -> (row, feature, gain)                                  (one.py:56-63)
 * variant 0: flat row-major argmax of square(ratio_a * S - mean * ratio_b), float32 ratios  (one.py:66-78)
 * variant 2: flat argmax of |ratio_a * S - mean * ratio_b|, returns the SQUARE of the best  (one.py:81-93)
 * The arithmetic replays the reference's operation order with correctly rounded IEEE operations. */
int b200dt_l2_score_cumsums(b200dt_ctx *ctx, const double *cumsums, int64_t n, int64_t F, int variant,
                            int64_t *out_k, int64_t *out_f, double *out_score);

/* ---- K3: L1 split search of ONE node, and the running median ------------------
 * replaces: greedy_regression_split[_v2](orders, X, y)   ref strategy_l1.py:125-140
 *           best_l1_improvements[_v2]                     ref strategy_l1.py:49-122
 * out_cost = total L1 cost of the two children (minimised; ties lowest k, then f). */
int b200dt_l1_split_node(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F,
                         const double *X, int64_t ldx, const double *y, int64_t N,
                         int64_t *out_k, int64_t *out_f, double *out_cost, double *out_threshold);
/* replaces: best_l1_improvements / best_l1_improvements_v2(ys) as stand-alone functions   ref strategy_l1.py:49-122
 * ys: (n, F) float64 row-major HOST matrix whose columns are searched independently (they need not be permutations
 * of one another); all F columns go through one batched search.  Flat row-major argmin: lowest cost, then lowest
 * row, then lowest column. */
int b200dt_l1_best_columns(b200dt_ctx *ctx, const double *ys, int64_t n, int64_t F, int64_t *out_k, int64_t *out_f,
                           double *out_cost);
/* replaces: np_stream_median.cummedian(arr)               ref np_stream_median.pyx:8-39
 *           StreamMedian.__call__                          ref stream_median.py:90-92
 * out: (n, 2) float64 (lower, upper) median of a[0..i]. */
int b200dt_cummedian(b200dt_ctx *ctx, const double *a, int64_t n, int a_space, double *out,
                     int out_space);

/* ---- K4: stable node partition -------------------------------------------------
 * replaces: split_orders(orders, index, n_samples)        ref one.py:110-118
 * orders (n, F) int64 host; left_rows (n_left,) int64 host; outputs (n_left, F) and
 * (n - n_left, F) int64 row-major host arrays. */
int b200dt_split_orders(b200dt_ctx *ctx, const int64_t *orders, int64_t n, int64_t F,
                        const int64_t *left_rows, int64_t n_left, int64_t n_samples,
                        int64_t *left_out, int64_t *right_out);

/* ---- whole fit on one GPU --------------------------------------------------------
 * replaces: one.build_regression_tree / build_regression_tree_v1_5   ref one.py:135-192
 *           strategy_l1.build_regression_tree / _v2                  ref strategy_l1.py:143-200
 * Tree arrays are caller-allocated host arrays of length `capacity`
 * (= 2**(max_depth+1) - 1), pre-filled by the caller with the reference's sentinels
 * (-1, -1, -2, -2.0; ref one.py:17-20).  Node ids are pre-order, left first
 * (one.py:26-31,157-160).  `value` is written for every node (the reference only
 * defines it on leaves).  *node_count = max_node_ + 1. */
int b200dt_fit(b200dt_ctx *ctx, const double *X, const double *y, int64_t N, int64_t F,
               int64_t ldx, int space, int criterion, int max_depth, double min_improvement,
               int64_t min_sample_leaf, int variant, int64_t *children_left,
               int64_t *children_right, int64_t *feature, double *threshold, double *value,
               int64_t capacity, int64_t *node_count);

/* Same fit with the conventions of the reference's level-synchronous builder:
 * replaces: four.build_regression_tree(X, y, max_depth, min_improvement, min_sample_leaf)
 *                                                                    ref four.py:108-172
 * flags: B200DT_NUMBER_BY_LEVEL -- node ids level by level, inside a level all LEFT children of the
 *        previous level's split nodes (in their order) and then all RIGHT children (four.py:113-170:
 *        sizes = left_sizes + right_sizes, ids handed out in that order, four.py:84-89);
 *        B200DT_LEAF_IF_LE -- a node is a leaf when its best gain is <= min_improvement
 *        (four.py:143; one.py:150 uses <).
 * n_node_samples (may be NULL): rows of every node, int32 like the reference (four.py:69,91-92).
 * weight (may be NULL): (N,) float64 sample weights, same space as X / y (four.py:116-117, 21-30:
 *   y <- y*w once, gain = (sum_yw - sum_w * T_yw / T_w)^2 / (sum_w (T_w - sum_w)), flat (k, f) argmax,
 *   leaf = sum(y w) / sum(w)); L2, variant 1 only. */
#define B200DT_NUMBER_BY_LEVEL 1
#define B200DT_LEAF_IF_LE 2
int b200dt_fit_ex(b200dt_ctx *ctx, const double *X, const double *y, const double *weight,
                  int64_t N, int64_t F, int64_t ldx, int space, int criterion, int max_depth,
                  double min_improvement,
                  int64_t min_sample_leaf, int variant, int flags, int64_t *children_left,
                  int64_t *children_right, int64_t *feature, double *threshold, double *value,
                  int32_t *n_node_samples, int64_t capacity, int64_t *node_count);

/* ---- feature-sharded fit: one session per GPU, driven level by level --------------
 * The host loop of ref one.py:141-160, level-synchronous.  Each rank owns columns
 * [col_offset, col_offset + F_local) of a global (N, F_global) matrix; X points at the
 * rank's slice.  Per level:
 *   search  -> local best candidate per live node (device array the caller all-gathers)
 *   decide  -> every rank feeds the gathered candidates; the library picks each node's
 *              winner with the reference's tie rule and the owner marks the left rows in
 *              the side mask (device, N/32 words, caller sum-all-reduces it)
 *   partition -> stable partition of every local column by the reduced mask.
 * Nodes searched in the reference's sequential fp64 order (small nodes, near ties) need the node total in the
 * order of global column F_global-1 (one.py:57).  The rank that owns that column computes it and publishes it in
 * its candidate record (exact == 2, f == F_global-1, left_sum = total; other ranks emit exact == 2, f == -1);
 * `decide` then reports need_exact and the next search_exact round runs the search on every rank.  So a level may
 * take up to three search_exact / gather / decide rounds; levels without such nodes take none.
 * last_col / last_col_ld: accepted for compatibility, ignored (no rank keeps a shadow copy of that column). */
typedef struct b200dt_candidate {
    double gain;      /* the score the search MAXIMISES: L2 v==1: variance gain (one.py:61);
                         L2 v!=1: the un-squared ratio score (its square is the gain, one.py:93);
                         L1: minus the children cost (strategy_l1.py:119)                   */
    double aux;       /* L2: |S - (k+1)*mean| of the candidate (in-row tie rule)        */
    double left_sum;  /* L2: sum of y over the left rows                               */
    double gain2;     /* best score among the OTHER local candidates (near-tie check)  */
    double threshold; /* X[orders[k,f], f]                                             */
    int32_t k;        /* sorted position                                               */
    int32_t f;        /* GLOBAL feature index                                          */
    int32_t exact;    /* 1: scores follow the reference's sequential fp64 order; 2: no candidate, left_sum
                         carries the node total in the last global column's order (see below)  */
    int32_t pad;
} b200dt_candidate;

int b200dt_session_begin(b200dt_ctx *ctx, const double *X, const double *y, int64_t N,
                         int64_t F_local, int64_t ldx, int space, int criterion, int max_depth,
                         double min_improvement, int64_t min_sample_leaf, int variant,
                         int64_t col_offset, int64_t F_global, const double *last_col,
                         int64_t last_col_ld);
/* number of nodes awaiting a search at the current level (0 => the tree is complete) */
int b200dt_session_live(b200dt_ctx *ctx, int64_t *n_live);
/* d_out: device array of n_live candidates */
int b200dt_session_search(b200dt_ctx *ctx, b200dt_candidate *d_out);
/* d_gathered: device array [n_ranks][n_live]; d_side_mask: device, (N+31)/32 uint32, zeroed
 * by the library before marking.  *need_partition = 0 when no child needs a search. */
int b200dt_session_decide(b200dt_ctx *ctx, const b200dt_candidate *d_gathered, int n_ranks,
                          uint32_t *d_side_mask, int *need_partition, int *need_exact);
/* re-search, with the reference's sequential arithmetic, the nodes `decide` reported as
 * near-ties (need_exact != 0); afterwards gather and call decide again */
int b200dt_session_search_exact(b200dt_ctx *ctx, b200dt_candidate *d_out);
int b200dt_session_partition(b200dt_ctx *ctx, const uint32_t *d_side_mask);
int b200dt_session_finish(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right,
                          int64_t *feature, double *threshold, double *value, int64_t capacity,
                          int64_t *node_count);

/* ---- K5: batched inference -------------------------------------------------------
 * replaces: utils.inference / utils.inference_leaf        ref utils.py:15-31,34-48
 * Go left iff X[row, feature[node]] <= threshold[node]; stop where the chosen child is -1.
 * out_value (N,) float64 and/or out_leaf (N,) int32 may be NULL.  Tree arrays are host. */
int b200dt_predict(b200dt_ctx *ctx, const double *X, int64_t N, int64_t F, int64_t ldx,
                   int x_space, const int64_t *feature, const double *threshold,
                   const int64_t *children_left, const int64_t *children_right,
                   const double *value, int64_t n_nodes, double *out_value, int32_t *out_leaf,
                   int out_space);

/* ---- K6: classification trees (exhaustive gini / precision-at-1 search) -----------------
 * replaces: tree_builder.build_classification_tree(X, y, max_depth, max_feature, min_improvement,
 *           min_sample_leaf, split_method=greedy_classification[_p_at_k], max_classes)
 *           ref tree_builder.py:56-96,149-159; strategy.py:272-283,313-323; base.py:13-21
 * labels (N,) int32 in [0, n_classes), same memory space as X.  n_classes <= 256, N <= 2^26 and
 * n_classes * 2^ceil(log2 N) <= 2^32 (class and row id share one 32-bit entry).  Pre-order node
 * ids; a node is a leaf when rows <= 2 * min_sample_leaf, at max_depth, or when the best split's
 * improvement < min_improvement; rows go left by VALUE (x <= threshold, tree_builder.py:31-34);
 * leaf value = first most frequent class.  n_node_samples may be NULL.  Column sub-sampling
 * (F > max_feature, np.random) is the caller's business: pass the columns to search. */
int b200dt_fit_classes(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F,
                       int64_t ldx, int space, int n_classes, int criterion, int max_depth,
                       double min_improvement, int64_t min_sample_leaf, int64_t *children_left,
                       int64_t *children_right, int64_t *feature, double *threshold, double *value,
                       int32_t *n_node_samples, int64_t capacity, int64_t *node_count);
/* The same fit as a sequence of steps, for feature-sharded multi-GPU fits (one process per GPU,
 * every rank holds columns [col_offset, col_offset + F_local) of the F_global columns and all labels).
 * Per level:  search -> all-gather of the candidates -> decide -> sum-all-reduce of the side mask AND of
 * the left class counts (only the owner of a winning column contributes non-zeros) -> partition.
 * Candidates carry GLOBAL column ids; every rank takes the same decisions from the same gathered array
 * (score desc, position asc, column asc = the reference's flat argmax, strategy.py:277-278).
 * d_left_counts: device int32 [n_live][n_classes] (the first n_split rows are used).
 * driven by np_decision_tree_b200/sharded.py::fit_classes_sharded. */
typedef struct b200dt_class_candidate {
    double score;     /* best surrogate value of the node over this rank's columns            */
    double threshold; /* feature value at (k, f)                                               */
    int32_t k;        /* sorted position inside the node, -1 = none                            */
    int32_t f;        /* GLOBAL column id                                                      */
    int32_t k_ext;    /* last position whose value equals the threshold (rows go left by value) */
    int32_t pad;
} b200dt_class_candidate;
int b200dt_class_session_begin(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N,
                               int64_t F_local, int64_t ldx, int space, int n_classes, int criterion,
                               int max_depth, double min_improvement, int64_t min_sample_leaf,
                               int64_t col_offset, int64_t F_global);
int b200dt_class_session_live(b200dt_ctx *ctx, int64_t *n_live);
int b200dt_class_session_search(b200dt_ctx *ctx, b200dt_class_candidate *d_out);
int b200dt_class_session_decide(b200dt_ctx *ctx, const b200dt_class_candidate *d_gathered, int n_ranks,
                                uint32_t *d_side_mask, int32_t *d_left_counts, int64_t *n_split);
int b200dt_class_session_partition(b200dt_ctx *ctx, const uint32_t *d_side_mask,
                                   const int32_t *d_left_counts);
int b200dt_class_session_finish(b200dt_ctx *ctx, int64_t *children_left, int64_t *children_right,
                                int64_t *feature, double *threshold, double *value,
                                int32_t *n_node_samples, int64_t capacity, int64_t *node_count);
/* One node: best (column, sorted position, threshold, surrogate score) over all columns with the
 * reference's flat argmax (lowest position, then lowest column).
 * replaces: strategy.greedy_classification / greedy_classification_p_at_k   ref strategy.py:272-283,313-323 */
int b200dt_class_split_node(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t n, int64_t F,
                            int64_t ldx, int space, int n_classes, int criterion, int64_t *out_f,
                            int64_t *out_k, double *out_threshold, double *out_score);

/* ---- K7: node-at-a-time primitives for the random-threshold classification splitters -----------
 * replaces the array work of strategy.random_classify / random_classify_p_at_k (ref strategy.py:149-172)
 * under tree_builder.build_tree (ref tree_builder.py:24-41,43-52,56-96).  The reference draws its thresholds
 * (and its column sub-samples) from numpy's GLOBAL generator per node in depth-first order, so the host keeps
 * that loop and those np.random calls; per node the GPU supplies
 *   minmax:            np.min / np.max over the node's rows of the given columns        (strategy.py:150)
 *   threshold_counts:  class counts of the rows with x <= threshold, per column         (strategy.py:151-153)
 *   split:             stable split of the node's row list by (column, threshold)       (tree_builder.py:31-34)
 * A node is the segment [start, start + n) of row-id array `buffer` (0 / 1); `split` writes the children into
 * the same segment of the other array: left rows [start, start + n_left), right rows behind them.  The root is
 * segment [0, N) of buffer 0.  cols / thresholds / outputs are HOST arrays; out_counts is [n_cols][n_classes]. */
int b200dt_rows_begin(b200dt_ctx *ctx, const double *X, const int32_t *labels, int64_t N, int64_t F,
                      int64_t ldx, int space, int n_classes, int64_t *class_counts);
int b200dt_rows_minmax(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, const int32_t *cols,
                       int n_cols, double *out_min, double *out_max);
int b200dt_rows_threshold_counts(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n,
                                 const int32_t *cols, int n_cols, const double *thresholds,
                                 int64_t *out_counts);
/* regression with random thresholds (strategy.random_split_v2, ref strategy.py:136-146): register the float64
 * targets once, then per node: sum of y and row count left of every threshold (masks.T.dot(y), masks.sum(0)) and
 * out_node = {sum, min, max} of y over the node (np.mean for the leaf, np.unique(y).size <= 1 for base.py:16).
 * Floating-point sums in a FIXED order that depends on the node's rows only (per-thread shares, then the row groups
 * of a block in shared memory, then the blocks' slots): the sums are column-independent, so two columns whose
 * thresholds cut off the SAME rows get bit-identical sums and tie exactly, as they do in the reference (whose sums
 * go through BLAS, so their last bits are not defined by the reference itself). */
int b200dt_rows_set_targets(b200dt_ctx *ctx, const double *y, int space);
int b200dt_rows_threshold_sums(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, const int32_t *cols,
                               int n_cols, const double *thresholds, double *out_sums,
                               int64_t *out_counts, double *out_node);
/* n_left may be NULL: the split is then only queued (no host round trip); the caller already knows the count from
 * threshold_counts / threshold_sums of the same (column, threshold). */
int b200dt_rows_split(b200dt_ctx *ctx, int buffer, int64_t start, int64_t n, int32_t col,
                      double threshold, int64_t *n_left);
int b200dt_rows_end(b200dt_ctx *ctx);

/* ---- measurement -------------------------------------------------------------------
 * When enabled, every kernel launch is bracketed by CUDA events on the launch stream;
 * `get` synchronises and returns, per kernel class, the summed device time, the number of
 * launches and the ALGORITHMIC bytes those launches moved. */
int b200dt_profile_enable(b200dt_ctx *ctx, int on);
int b200dt_profile_reset(b200dt_ctx *ctx);
int b200dt_profile_count(void);
const char *b200dt_profile_name(int kernel_class);
int b200dt_profile_get(b200dt_ctx *ctx, int kernel_class, double *ms, int64_t *launches,
                       double *alg_bytes);
/* total kernel launches issued by the library on this context since the last reset */
int64_t b200dt_launch_count(const b200dt_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200DT_H */
