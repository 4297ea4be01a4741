"""numpy restatement of the reference hot path -- TEST INFRASTRUCTURE ONLY.

Not a product path: see ``oracle/__init__.py`` for who may import this.

All "ref:" citations are paths below /root/reference/ (yupbank/np_decision_tree).
The arithmetic (operation order, dtypes, tie rules) follows the reference
exactly; the code organisation is our own.  Parity pinning: replayed against
fixtures produced by the reference itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import heapq
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ref: decision_tree/one.py:5-6 -- the five parallel arrays of the tree.
TreeFields = namedtuple(
    "TreeFields", "children_left children_right feature threshold value")


class TreeArrays:
    """Tree container with the reference's layout (ref: decision_tree/one.py:9-24).

    ``tree_`` holds five arrays of length ``2**(max_depth+1) - 1``; unused
    slots keep -1 / -1 / -2 / -2.0; ``value`` is only defined on leaves.
    ``max_node_`` is the id of the last node created (node_count - 1).
    """

    def __init__(self, max_depth):
        slots = 2 ** (max_depth + 1) - 1
        self.tree_ = TreeFields(
            np.full(slots, -1, dtype=np.int64),
            np.full(slots, -1, dtype=np.int64),
            np.full(slots, -2, dtype=np.int64),
            np.full(slots, -2.0, dtype=np.float64),
            # the reference leaves this uninitialised (np.empty); NaN makes
            # accidental use of a non-leaf value visible in tests
            np.full(slots, np.nan, dtype=np.float64),
        )
        self.max_node_ = -1

    @property
    def node_count(self):
        return self.max_node_ + 1

    def leaf_mask(self):
        m = np.zeros(self.tree_.feature.shape[0], dtype=bool)
        m[: self.node_count] = self.tree_.children_left[: self.node_count] == -1
        return m


# --------------------------------------------------------------------------
# sort
# --------------------------------------------------------------------------
def argsort_columns(X, kind=None):
    """Per-column sort permutation (ref: decision_tree/one.py:139).

    The reference uses numpy's default (unstable) kind, so the order among
    equal feature values is unspecified there.  ``kind='stable'`` is what the
    CUDA radix sort produces; both agree on columns with distinct values.
    """
    return np.argsort(X, axis=0, kind=kind)


# --------------------------------------------------------------------------
# L2 (MSE) split search
# --------------------------------------------------------------------------
def l2_weights(n):
    """1 / (left_count * right_count) for split positions 0..n-2.

    ref: decision_tree/one.py:52-53 (np_gene).  left_count is float32,
    right_count int64; the product promotes to float64.
    """
    left_count = np.arange(1, n, dtype=np.float32)
    right_count = np.arange(n - 1, 0, -1)
    return np.reciprocal(left_count * right_count)


def l2_best_split(prefix):
    """Best (position, feature, gain) from column prefix sums.

    ref: decision_tree/one.py:56-63 (best_variance_improvements).
    ``prefix[k, f]`` = sum of y over the first k+1 rows of feature f's order.
    The node total is read from the LAST column (one.py:57).  Ties: lowest
    feature inside a position (one.py:59), then lowest position (one.py:62).
    """
    n = prefix.shape[0]
    node_mean = prefix[-1][-1] / n
    expected = np.arange(1, n + 1) * node_mean
    dev = np.abs(prefix - expected[:, np.newaxis])
    feat_of_pos = np.argmax(dev, axis=1)
    dev_of_pos = dev[np.arange(n), feat_of_pos]
    gain = l2_weights(n) * np.square(dev_of_pos)[:-1]
    k = np.argmax(gain)
    return k, feat_of_pos[k], gain[k]


def l2_best_split_ratio(prefix):
    """The ``v != 1`` search with float32 sqrt ratios.

    ref: decision_tree/one.py:81-93 (best_variance_improvements_v2).  Flat
    row-major argmax => lowest position, then lowest feature.
    """
    n = prefix.shape[0]
    node_mean = prefix[-1][-1] / n
    lc = np.arange(1, n, dtype=np.float32)
    rc = np.arange(n - 1, 0, -1, dtype=np.float32)
    scale_sum = np.sqrt(np.reciprocal(lc * rc))[:, np.newaxis]
    scale_mean = np.sqrt(lc / rc)[:, np.newaxis]
    score = np.abs(scale_sum * prefix[:-1] - node_mean * scale_mean)
    k, f = np.unravel_index(np.argmax(score), score.shape)
    return k, f, np.square(score[k, f])


def l2_split_node(orders, X, y, v=1):
    """ref: decision_tree/one.py:96-107 (greedy_regression_split)."""
    prefix = np.cumsum(y[orders], axis=0)
    search = l2_best_split if v == 1 else l2_best_split_ratio
    k, f, gain = search(prefix)
    threshold = X[orders[k, f], f]
    left_rows = orders[:, f][: k + 1]
    return threshold, f, gain, left_rows


# --------------------------------------------------------------------------
# node partition
# --------------------------------------------------------------------------
def partition_orders(orders, left_rows, n_samples):
    """Stable split of every column of ``orders`` into (left, right).

    ref: decision_tree/one.py:110-118 (split_orders).
    """
    goes_left = np.zeros(n_samples, dtype=bool)
    goes_left[left_rows] = True
    mask = goes_left[orders]
    n_left = left_rows.shape[0]
    n_right = orders.shape[0] - n_left
    left = orders.T[mask.T].reshape(-1, n_left).T
    right = orders.T[~mask.T].reshape(-1, n_right).T
    return left, right


# --------------------------------------------------------------------------
# tree growth (shared driver; ref: one.py:135-162, strategy_l1.py:143-200)
# --------------------------------------------------------------------------
def _grow(X, y, orders, max_depth, min_improvement, min_sample_leaf,
          split_fn, leaf_fn):
    tree = TreeArrays(max_depth)
    t = tree.tree_
    # LIFO work list of (orders, parent, is_left, depth): pre-order, left first
    # (ref: one.py:141-160; node ids from one.py:26-31)
    pending = [(orders, None, True, 0)]
    while pending:
        node_orders, parent, is_left, depth = pending.pop()
        tree.max_node_ += 1
        node = tree.max_node_
        if parent is not None:
            (t.children_left if is_left else t.children_right)[parent] = node
        # leaf test: rows <= min_sample_leaf or depth limit (one.py:40-46,145)
        if node_orders.shape[0] <= min_sample_leaf or depth >= max_depth:
            t.value[node] = leaf_fn(y[node_orders])
            continue
        threshold, f, score, left_rows = split_fn(node_orders, X, y)
        # one.py:150 / strategy_l1.py:158 -- same "<" test for both criteria
        if score < min_improvement:
            t.value[node] = leaf_fn(y[node_orders])
            continue
        t.threshold[node] = threshold
        t.feature[node] = f
        left, right = partition_orders(node_orders, left_rows, y.shape[0])
        pending.append((right, node, False, depth + 1))
        pending.append((left, node, True, depth + 1))
    return tree


def fit_l2(X, y, max_depth=2, max_feature=100, min_improvement=1e-7,
           min_sample_leaf=1, v=1, sort_kind=None):
    """ref: decision_tree/one.py:135-162 (build_regression_tree).

    ``max_feature`` is accepted and ignored, as in the reference.
    Leaf value: np.mean over the (n, F) gather (one.py:132,146).
    """
    orders = argsort_columns(X, kind=sort_kind)
    return _grow(X, y, orders, max_depth, min_improvement, min_sample_leaf,
                 lambda o, X_, y_: l2_split_node(o, X_, y_, v), np.mean)


# --------------------------------------------------------------------------
# running median (native in the reference)
# --------------------------------------------------------------------------
def running_median_pairs_py(seq):
    """Pure-Python (lower, upper) running median; small inputs only.

    ref: decision_tree/np_stream_median.pyx:8-39 (cummedian) and
    decision_tree/stream_median.py:33-64,90-92.  Two heaps: ``low`` is a
    max-heap (stored negated), ``high`` a min-heap.
    """
    n = len(seq)
    out = np.zeros((n, 2), dtype=np.float64)
    low, high = [], []
    for i in range(n):
        x = float(seq[i])
        if not low or x <= -low[0]:
            heapq.heappush(low, -x)
        else:
            heapq.heappush(high, x)
        if len(low) > len(high) + 1:
            heapq.heappush(high, -heapq.heappop(low))
        elif len(high) > len(low) + 1:
            heapq.heappush(low, -heapq.heappop(high))
        if len(low) == len(high):
            out[i, 0], out[i, 1] = -low[0], high[0]
        elif len(low) > len(high):
            out[i, 0] = out[i, 1] = -low[0]
        else:
            out[i, 0] = out[i, 1] = high[0]
    return out


_C_LIB = None


def build_c_oracle(force=False):
    """Compile oracle/running_median.c -> oracle/_build/liboracle_cpu.so."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle_cpu.so")
    src = os.path.join(_HERE, "running_median.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", so, src])
    return so


def _c_lib():
    global _C_LIB
    if _C_LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.oracle_running_median_pairs.restype = ctypes.c_int
        lib.oracle_running_median_pairs.argtypes = [
            ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _C_LIB = lib
    return _C_LIB


def running_median_pairs(seq):
    """(n, 2) array of (lower, upper) running medians (C restatement).

    ref: decision_tree/np_stream_median.pyx:8-39.  Prefix of length m:
    lower = order statistic floor((m-1)/2), upper = floor(m/2).
    """
    a = np.ascontiguousarray(seq, dtype=np.float64)
    n = a.shape[0]
    out = np.zeros((n, 2), dtype=np.float64)
    if n:
        rc = _c_lib().oracle_running_median_pairs(
            a.ctypes.data, n, 1, out.ctypes.data)
        if rc != 0:
            raise MemoryError("oracle_running_median_pairs failed")
    return out


def _median_pairs_columns(ys):
    """(n, F) -> two (n, F) arrays: lower / upper running median per column.

    ref: strategy_l1.py:88 (np.apply_along_axis(cummedian, 0, ys)).
    """
    n, F = ys.shape
    lo = np.empty((n, F), dtype=np.float64)
    hi = np.empty((n, F), dtype=np.float64)
    for f in range(F):
        pairs = running_median_pairs(ys[:, f])
        lo[:, f] = pairs[:, 0]
        hi[:, f] = pairs[:, 1]
    return lo, hi


# --------------------------------------------------------------------------
# L1 (MAE) split search
# --------------------------------------------------------------------------
def l1_best_split(ys):
    """Best (position, feature, children L1 cost) -- cost is MINIMISED.

    ref: decision_tree/strategy_l1.py:87-122 (best_l1_improvements_v2); the
    pure-Python twin strategy_l1.py:49-84 yields identical results.
    """
    n, F = ys.shape
    # prefix side: cost(prefix + y) = cost(prefix) + dist(y, [lo, hi]) with
    # the pair of the prefix BEFORE y joins (strategy_l1.py:89-100)
    lo_p, hi_p = _median_pairs_columns(ys)
    nxt = ys[1:]
    lo_b, hi_b = lo_p[:-1], hi_p[:-1]
    left_cost = np.cumsum(
        np.where(nxt <= lo_b, lo_b - nxt, 0) + np.where(nxt >= hi_b, nxt - hi_b, 0),
        axis=0)
    left_cost = np.concatenate([np.zeros((1, F)), left_cost])
    # suffix side: the pair already CONTAINS y (strategy_l1.py:103-117)
    lo_s, hi_s = _median_pairs_columns(ys[::-1])
    lo_s, hi_s = lo_s[::-1], hi_s[::-1]
    right_cost = np.cumsum(
        (np.where(ys <= lo_s, hi_s - ys, 0) + np.where(ys >= hi_s, ys - lo_s, 0))[::-1],
        axis=0)[::-1]
    # strategy_l1.py:119-121: total cost of the two children, flat argmin
    cost = right_cost[1:] + left_cost[:-1]
    k, f = np.unravel_index(np.argmin(cost), cost.shape)
    return k, f, cost[k, f]


def l1_split_node(orders, X, y):
    """ref: decision_tree/strategy_l1.py:134-140 (greedy_regression_split_v2)."""
    k, f, cost = l1_best_split(y[orders])
    threshold = X[orders[k, f], f]
    left_rows = orders[: k + 1, f]
    return threshold, f, cost, left_rows


def fit_l1(X, y, max_depth=2, max_feature=100, min_improvement=1e-7,
           min_sample_leaf=1, leaf_from_data=np.median, sort_kind=None):
    """ref: decision_tree/strategy_l1.py:173-200 (build_regression_tree_v2);
    strategy_l1.py:143-170 (build_regression_tree) builds the same tree."""
    orders = argsort_columns(X, kind=sort_kind)
    return _grow(X, y, orders, max_depth, min_improvement, min_sample_leaf,
                 l1_split_node, leaf_from_data)


# --------------------------------------------------------------------------
# inference
# --------------------------------------------------------------------------
def _descend(data, tree):
    """Level-synchronous traversal; returns the final node per row.

    ref: decision_tree/utils.py:15-30 / 34-47.  Left iff x <= threshold;
    a row stops where the chosen child is -1.  Shape quirk kept: if nothing
    ever moves (root-only tree) the result has shape (1,).
    """
    t = tree.tree_
    rows = np.arange(data.shape[0], dtype=np.int32)
    at = np.zeros(1, dtype=np.int32)
    while True:
        go_left = data[rows, t.feature[at]] <= t.threshold[at]
        nxt = np.where(go_left, t.children_left[at], t.children_right[at])
        moves = nxt != -1
        if not np.any(moves):
            return at
        at = np.where(moves, nxt, at)


def predict(data, tree):
    """ref: decision_tree/utils.py:15-31 (inference)."""
    return tree.tree_.value[_descend(data, tree)].ravel()


def predict_leaf(data, tree):
    """ref: decision_tree/utils.py:34-48 (inference_leaf)."""
    return _descend(data, tree)


# --------------------------------------------------------------------------
# level-synchronous builder with optional sample weights
# --------------------------------------------------------------------------
LevelTreeFields = namedtuple(
    "LevelTreeFields", "children_left children_right feature threshold value n_node_samples")


class LevelTreeArrays:
    """ref: decision_tree/four.py:61-98 -- the one.py container plus int32 n_node_samples (-2 = unused)."""

    def __init__(self, max_depth):
        slots = 2 ** (max_depth + 1) - 1
        self.tree_ = LevelTreeFields(
            np.full(slots, -1, dtype=np.int64), np.full(slots, -1, dtype=np.int64),
            np.full(slots, -2, dtype=np.int64), np.full(slots, -2.0, dtype=np.float64),
            np.full(slots, np.nan, dtype=np.float64), np.full(slots, -2, dtype=np.int32))
        self.max_node_ = -1

    @property
    def node_count(self):
        return self.max_node_ + 1

    def leaf_mask(self):
        m = np.zeros(self.tree_.feature.shape[0], dtype=bool)
        m[: self.node_count] = self.tree_.children_left[: self.node_count] == -1
        return m


def l2_best_split_weighted(ys, ws):
    """Weighted search: ys = (y*w)[orders], ws = w[orders] of one node.

    ref: decision_tree/four.py:21-30.  Flat row-major argmax => lowest position, then feature."""
    sum_y = np.cumsum(ys, axis=0)
    sum_w = np.cumsum(ws, axis=0)
    expected = sum_w * sum_y[-1][-1] / sum_w[-1][-1]
    sq = np.square(sum_y - expected)
    ratio = np.reciprocal(sum_w[:-1] * (sum_w[-1][-1] - sum_w[:-1]))
    gain = ratio * sq[:-1]
    k, f = np.unravel_index(np.argmax(gain), gain.shape)
    return k, f, gain[k, f]


def fit_l2_levelwise(X, y, max_depth=2, min_improvement=1e-7, min_sample_leaf=1, weight=None,
                     sort_kind=None):
    """ref: decision_tree/four.py:108-172 (build_regression_tree).

    Differences from one.py kept: node ids level by level -- all left children of a level, then
    all right children (four.py:113,162-170) --, leaf when gain <= min_improvement (four.py:143),
    n_node_samples, leaf = mean over column 0's order (four.py:130,145), weights multiply y once
    (four.py:116-117) and leaf = sum(y w) / sum(w) (four.py:132-133)."""
    tree = LevelTreeArrays(max_depth)
    t = tree.tree_
    n_samples = X.shape[0]
    yw = y * weight if weight is not None else y
    level = [(argsort_columns(X, kind=sort_kind), None, True)]
    for depth in range(max_depth + 1):
        lefts, rights = [], []
        for orders, parent, is_left in level:
            tree.max_node_ += 1
            node = tree.max_node_
            if parent is not None:
                (t.children_left if is_left else t.children_right)[parent] = node
            n = orders.shape[0]
            t.n_node_samples[node] = n
            ys = yw[orders]
            ws = weight[orders] if weight is not None else None

            def leaf():
                if ws is None:
                    t.value[node] = np.mean(ys[:, 0])
                else:
                    t.value[node] = np.sum(ys[:, 0]) / np.sum(ws[:, 0])

            if n <= min_sample_leaf or depth >= max_depth:
                leaf()
                continue
            if ws is None:
                k, f, gain = l2_best_split(np.cumsum(ys, axis=0))     # four.py:11-20 == one.py:56-63
            else:
                k, f, gain = l2_best_split_weighted(ys, ws)
            if gain <= min_improvement:
                leaf()
                continue
            t.feature[node] = f
            t.threshold[node] = X[orders[k, f], f]
            lo, ro = partition_orders(orders, orders[: k + 1, f], n_samples)
            lefts.append((lo, node, True))
            rights.append((ro, node, False))
        level = lefts + rights
    return tree


# --------------------------------------------------------------------------
# classification: exhaustive gini / precision-at-1 search   
# --------------------------------------------------------------------------
def class_prefix_counts(sorted_labels, n_classes):
    """(n, F, C) float64 running class counts down every sorted column.

    ref: decision_tree/strategy.py:274-275 -- ``np.cumsum(y[orders], axis=0)`` on the
    one-hot targets that tree_builder.py:157-158 builds with ``np.eye(max_classes)[y]``.
    The counts are whole numbers, so the summation order cannot change them.
    """
    return np.cumsum(np.eye(n_classes)[sorted_labels], axis=0)


def gini_surrogate(prefix):
    """(n-1, F) score of every split position: sum_c L_c^2 / n_left + sum_c R_c^2 / n_right.

    ref: decision_tree/strategy.py:32-55 (surrogate_gini_improvements called with two
    arguments: candidates are rows [:-1], the node totals are row [-1]).
    """
    n = prefix.shape[0]
    left, total = prefix[:-1], prefix[-1]
    n_left = np.arange(1, n)[:, np.newaxis]
    right = total - left
    return (np.square(left).sum(axis=2) / n_left
            + np.square(right).sum(axis=2) / (n - n_left))


def p_at_1_surrogate(prefix):
    """(n-1, F): largest class count on the left + largest class count on the right.

    ref: decision_tree/strategy.py:79-100 (surrogate_p_at_k_improvements; its ``k`` is unused).
    """
    left, total = prefix[:-1], prefix[-1]
    return np.max(left, axis=2) + np.max(total - left, axis=2)


def class_split_node(X_node, onehot_node, criterion="gini", sort_kind=None):
    """(column, threshold, improvement, split_row) of one node.

    ref: decision_tree/strategy.py:272-283 (greedy_classification) and :313-323
    (greedy_classification_p_at_k): per-node argsort, flat row-major argmax
    (lowest position, then lowest column), threshold = the feature value at the
    chosen position, improvement via surrogate_to_gini / surrogate_to_p_at_k
    (strategy.py:103-108).
    """
    order = np.argsort(X_node, axis=0, kind=sort_kind)
    labels = np.argmax(onehot_node, axis=1)
    prefix = class_prefix_counts(labels[order], onehot_node.shape[1])
    n = X_node.shape[0]
    class_totals = onehot_node.sum(axis=0)
    if criterion == "gini":
        score = gini_surrogate(prefix)
    elif criterion == "p_at_k":
        score = p_at_1_surrogate(prefix)
    else:
        raise ValueError(criterion)
    k, f = np.unravel_index(np.argmax(score), score.shape)
    threshold = X_node[order[k, f], f]
    if criterion == "gini":
        improvement = score[k, f] / n - np.sum(np.square(class_totals / onehot_node.sum()))
    else:
        improvement = (score[k, f] - np.max(class_totals)) / n
    return int(f), threshold, improvement, int(k)


def fit_classes(X, labels, n_classes, max_depth=2, max_feature=100, min_improvement=0.01,
                min_sample_leaf=1, criterion="gini", sort_kind=None):
    """Depth-first classification tree with the reference's node order and leaf rules.

    ref: decision_tree/tree_builder.py:56-96 (build_tree) as driven by
    build_classification_tree (:149-159) with split_method=greedy_classification[_p_at_k];
    decision_tree/base.py:13-21 (is_leaf: rows <= 2 * min_sample_leaf, or a single
    distinct value in the one-hot block); tree_builder.py:24-41 (split_mask: rows go
    left when value <= threshold, so equal values never straddle a split);
    tree_builder.py:20-21 (leaf = first most frequent class).
    Column sub-sampling (tree_builder.py:44-48, np.random) applies only when
    F > max_feature and is not restated.
    """
    X = np.asarray(X)
    labels = np.asarray(labels)
    if X.shape[1] > max_feature:
        raise NotImplementedError("random column sub-sampling (F > max_feature) is not restated")
    onehot = np.eye(n_classes)[labels]
    tree = TreeArrays(max_depth)
    t = tree.tree_
    pending = [(np.arange(X.shape[0]), -1, True, 0)]
    while pending:
        rows, parent, is_left, depth = pending.pop()
        tree.max_node_ += 1
        node = tree.max_node_
        if parent >= 0:
            (t.children_left if is_left else t.children_right)[parent] = node
        Xn, yn = X[rows], onehot[rows]
        leaf = rows.size <= 2 * min_sample_leaf or np.unique(yn).size <= 1 or depth >= max_depth
        if not leaf:
            f, threshold, improvement, _ = class_split_node(Xn, yn, criterion, sort_kind)
            leaf = improvement < min_improvement
        if leaf:
            t.value[node] = np.argmax(yn.sum(axis=0))
            continue
        t.feature[node] = f
        t.threshold[node] = threshold
        goes_left = Xn[:, f] <= threshold
        pending.append((rows[~goes_left], node, False, depth + 1))
        pending.append((rows[goes_left], node, True, depth + 1))
    return tree


# --------------------------------------------------------------------------
# classification with random thresholds (ExtraTree style)   
# --------------------------------------------------------------------------
def random_class_split_node(X_node, onehot_node, criterion="gini", rng=np.random):
    """(column, threshold, improvement) of one node with ONE random threshold per column.

    ref: decision_tree/strategy.py:149-159 (random_classify) and :162-172
    (random_classify_p_at_k).  The thresholds are ``uniform(column min, column max)``
    drawn from numpy's GLOBAL generator in column order; rows go left when
    ``x <= threshold``; the best column is the first argmax of the surrogate
    (a constant column gives 0/0 = nan, which np.argmax treats as the maximum).
    """
    n = X_node.shape[0]
    thresholds = rng.uniform(X_node.min(axis=0), X_node.max(axis=0))
    goes_left = X_node <= thresholds
    left = goes_left.T.astype(np.float64) @ onehot_node          # (F, C) whole numbers
    n_left = goes_left.sum(axis=0)
    total = onehot_node.sum(axis=0)
    right = total - left
    with np.errstate(divide="ignore", invalid="ignore"):
        if criterion == "gini":
            score = np.square(left).sum(axis=1) / n_left + np.square(right).sum(axis=1) / (n - n_left)
        elif criterion == "p_at_k":
            score = np.max(left, axis=1) + np.max(right, axis=1)
        else:
            raise ValueError(criterion)
    best = int(np.argmax(score))
    if criterion == "gini":
        improvement = score[best] / n - np.sum(np.square(total / onehot_node.sum()))
    else:
        improvement = (score[best] - np.max(total)) / n
    return best, thresholds[best], improvement


def fit_classes_random(X, labels, n_classes, max_depth=2, max_feature=100, min_improvement=0.01,
                       min_sample_leaf=1, criterion="gini", rng=np.random):
    """Depth-first classification tree with random thresholds and column sub-sampling.

    ref: decision_tree/tree_builder.py:43-52 (find_best_split: all columns when
    F <= max_feature, else the first max_feature of ``np.random.permutation(columns)``),
    :56-96 (build_tree), base.py:13-21.  Consumes ``rng`` in the reference's order:
    per searched node, in depth-first order, [permutation] then F' uniforms.
    """
    X = np.asarray(X)
    labels = np.asarray(labels)
    onehot = np.eye(n_classes)[labels]
    all_columns = np.arange(X.shape[1])
    tree = TreeArrays(max_depth)
    t = tree.tree_
    pending = [(np.arange(X.shape[0]), -1, True, 0)]
    while pending:
        rows, parent, is_left, depth = pending.pop()
        tree.max_node_ += 1
        node = tree.max_node_
        if parent >= 0:
            (t.children_left if is_left else t.children_right)[parent] = node
        Xn, yn = X[rows], onehot[rows]
        leaf = rows.size <= 2 * min_sample_leaf or np.unique(yn).size <= 1 or depth >= max_depth
        if not leaf:
            columns = all_columns if all_columns.size <= max_feature else rng.permutation(all_columns)[:max_feature]
            best, threshold, improvement = random_class_split_node(Xn[:, columns], yn, criterion, rng)
            f = columns[best]
            leaf = improvement < min_improvement
        if leaf:
            t.value[node] = np.argmax(yn.sum(axis=0))
            continue
        t.feature[node] = f
        t.threshold[node] = threshold
        goes_left = Xn[:, f] <= threshold
        pending.append((rows[~goes_left], node, False, depth + 1))
        pending.append((rows[goes_left], node, True, depth + 1))
    return tree


# --------------------------------------------------------------------------
# regression with random thresholds (ExtraTree style)   
# --------------------------------------------------------------------------
def random_l2_split_node(X_node, y_node, rng=np.random):
    """(column, threshold, improvement) of one node, one random threshold per column.

    ref: decision_tree/strategy.py:136-146 (random_split_v2) with
    surrogate_variance_improvements (:4-27, four-argument form) and
    surrogate_to_variance (:28-29).  ``masks.T.dot(y)`` goes through BLAS in the
    reference: the sums are only defined up to floating-point summation order.
    """
    n = X_node.shape[0]
    thresholds = rng.uniform(X_node.min(axis=0), X_node.max(axis=0))
    goes_left = X_node <= thresholds
    left_sum = goes_left.T.astype(np.float64) @ y_node
    n_left = goes_left.sum(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        score = np.square(left_sum) / n_left + np.square(y_node.sum() - left_sum) / (n - n_left)
    best = int(np.argmax(score))
    return best, thresholds[best], score[best] / n - np.square(np.mean(y_node))


def fit_l2_random(X, y, max_depth=2, max_feature=100, min_improvement=0.01, min_sample_leaf=1, rng=np.random):
    """ref: decision_tree/tree_builder.py:56-96 (build_tree) with split_method=random_split_v2 and the default
    leaf_from_data_regression (np.mean, :16-17); leaf rules of base.py:13-21 (rows <= 2 * min_sample_leaf, or
    all targets equal).  Consumes ``rng`` like the reference: per searched node, depth first, [column
    permutation] then one uniform per column."""
    X = np.asarray(X)
    y = np.asarray(y, dtype=np.float64)
    all_columns = np.arange(X.shape[1])
    tree = TreeArrays(max_depth)
    t = tree.tree_
    pending = [(np.arange(X.shape[0]), -1, True, 0)]
    while pending:
        rows, parent, is_left, depth = pending.pop()
        tree.max_node_ += 1
        node = tree.max_node_
        if parent >= 0:
            (t.children_left if is_left else t.children_right)[parent] = node
        Xn, yn = X[rows], y[rows]
        leaf = rows.size <= 2 * min_sample_leaf or np.unique(yn).size <= 1 or depth >= max_depth
        if not leaf:
            columns = all_columns if all_columns.size <= max_feature else rng.permutation(all_columns)[:max_feature]
            best, threshold, improvement = random_l2_split_node(Xn[:, columns], yn, rng)
            f = columns[best]
            leaf = improvement < min_improvement
        if leaf:
            t.value[node] = np.mean(yn)
            continue
        t.feature[node] = f
        t.threshold[node] = threshold
        goes_left = Xn[:, f] <= threshold
        pending.append((rows[~goes_left], node, False, depth + 1))
        pending.append((rows[goes_left], node, True, depth + 1))
    return tree
