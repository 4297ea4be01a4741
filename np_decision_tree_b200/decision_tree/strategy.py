"""Mirror of the exhaustive classification searches of ``decision_tree/strategy.py``.

``greedy_classification`` / ``greedy_classification_p_at_k`` keep the reference's signature
(``X`` (n, F), ``y`` one-hot (n, C)) and return value ``(column, threshold, improvement,
constant_attr_mask)``; the search itself -- per-column sort, running class counts, gini or
precision-at-1 surrogate of every split position, flat argmax -- runs on the GPU
(``b200dt_class_split_node``, csrc/classes.cu).  Citations are to
/root/reference/decision_tree/strategy.py.

``random_classify`` / ``random_classify_p_at_k`` (strategy.py:149-172) draw one threshold per column
from numpy's GLOBAL generator; the draw stays on the host (same ``np.random.uniform`` call, so a seeded
run reproduces the reference), the column min / max and the class counts left of the thresholds come
from the GPU (csrc/rows.cu).  ``random_split_v2`` (strategy.py:136-146) is the regression counterpart: the
GPU returns the sum of y and the row count left of every threshold.  ``random_split`` (rank-based cuts through
a per-node argsort, strategy.py:175-185) and ``greedy_split`` (strategy.py:188-208) run their argsort on the GPU
(K1) and keep the reference's numpy arithmetic for the small per-column reductions.
"""
import ctypes

import numpy as np

from .. import _lib
from ._common import context_for, matrix_arg


def surrogate_to_p_at_k(value, y):
    """strategy.py:103-104."""
    return (value - np.max(y.sum(axis=0))) / y.shape[0]


def surrogate_to_gini(value, y):
    """strategy.py:107-108."""
    return value / y.shape[0] - np.sum(np.square(y.sum(axis=0) / y.sum()))


def surrogate_variance_improvements(left_sums, left_counts, total_sums=None, total_count=None):
    """strategy.py:4-27: sum_left^2 / n_left + sum_right^2 / n_right for every candidate.  With two arguments the
    last row holds the totals."""
    if total_sums is None:
        left_sums, total_sums = left_sums[:-1], left_sums[-1]
    if total_count is None:
        left_counts, total_count = left_counts[:-1], left_counts[-1]
    return (np.square(left_sums) / left_counts
            + np.square(total_sums - left_sums) / (total_count - left_counts))


def surrogate_to_variance(value, total_count, y_mean):
    """strategy.py:28-29."""
    return value / total_count - np.square(y_mean)


def surrogate_gini_improvements(left_class_counts, left_counts, total_class_counts=None, total_count=None):
    """strategy.py:32-55: sum_c L_c^2 / n_left + sum_c R_c^2 / n_right for every candidate.  With two
    arguments the last row holds the totals (cumulative counts down a sorted column)."""
    if total_class_counts is None:
        left_class_counts, total_class_counts = left_class_counts[:-1], left_class_counts[-1]
    if total_count is None:
        left_counts, total_count = left_counts[:-1], left_counts[-1]
    right = total_class_counts - left_class_counts
    return (np.square(left_class_counts).sum(axis=2) / left_counts
            + np.square(right).sum(axis=2) / (total_count - left_counts))


def surrogate_p_at_k_improvements(left_class_counts, left_counts, total_class_counts=None, total_count=None, k=1):
    """strategy.py:79-100: largest class count left + largest class count right (``k`` is unused there too)."""
    if total_class_counts is None:
        left_class_counts, total_class_counts = left_class_counts[:-1], left_class_counts[-1]
    return np.max(left_class_counts, axis=2) + np.max(total_class_counts - left_class_counts, axis=2)


def _gini_from_counts(value, class_counts, n):
    """surrogate_to_gini (strategy.py:107-108) from the class totals of the node."""
    return value / n - np.sum(np.square(class_counts / class_counts.sum()))


def _p_at_k_from_counts(value, class_counts, n):
    """surrogate_to_p_at_k (strategy.py:103-104) from the class totals of the node."""
    return (value - np.max(class_counts)) / n


def _labels_of(y):
    """int32 class ids of a one-hot (n, C) block (tree_builder.py:157-158 builds it with np.eye)."""
    y = np.asarray(y)
    if y.ndim != 2:
        raise ValueError("y must be the one-hot (n, C) matrix the reference builds with np.eye(C)[labels]")
    if not np.array_equal(y.sum(axis=1), np.ones(y.shape[0])) or not np.isin(y, (0, 1)).all():
        raise ValueError("y must be one-hot: exactly one 1 per row")
    return np.ascontiguousarray(np.argmax(y, axis=1), dtype=np.int32), int(y.shape[1])


def _search(X, y, criterion):
    labels, C = _labels_of(y)
    ctx = context_for()
    xp, xs, xk, n, F, ldx = matrix_arg(X, "X")
    if xs != _lib.HOST:
        raise TypeError("the single-node searches take host arrays (like the reference)")
    if labels.shape[0] != n:
        raise ValueError(f"y has {labels.shape[0]} rows, X has {n}")
    if n < 2:
        raise ValueError("attempt to get argmax of an empty sequence")   # what np.argmax raises in the reference
    f, k = ctypes.c_int64(), ctypes.c_int64()
    thr, score = ctypes.c_double(), ctypes.c_double()
    ctx.check(ctx.lib.b200dt_class_split_node(
        ctx.handle, xp, labels.ctypes.data, n, F, ldx, xs, C, criterion, ctypes.byref(f), ctypes.byref(k),
        ctypes.byref(thr), ctypes.byref(score)))
    return int(f.value), float(thr.value), float(score.value), F


def greedy_classification(X, y):
    """strategy.py:272-283: exhaustive gini search of one node."""
    f, thr, score, F = _search(X, y, _lib.GINI)
    return f, thr, surrogate_to_gini(score, np.asarray(y)), np.zeros(F, dtype=bool)


def greedy_classification_p_at_k(X, y, k=1):
    """strategy.py:313-323: exhaustive precision-at-1 search (``k`` is unused there as well)."""
    f, thr, score, F = _search(X, y, _lib.P_AT_1)
    return f, thr, surrogate_to_p_at_k(score, np.asarray(y)), np.zeros(F, dtype=bool)


def _random_search(X, y, surrogate, to_improvement):
    from ._rows import RowSession, random_node_search
    labels, C = _labels_of(y)
    if np.asarray(X).shape[0] < 1:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    sess = RowSession(X, labels, C)
    try:
        best, thr, improvement, _ = random_node_search(sess, 0, 0, sess.N, np.arange(sess.F), sess.root_counts,
                                                       surrogate, to_improvement)
        return best, thr, improvement, np.zeros(sess.F, dtype=bool)
    finally:
        sess.close()


def random_classify(X, y):
    """strategy.py:149-159: one U(min, max) threshold per column (np.random), best column by gini."""
    return _random_search(X, y, surrogate_gini_improvements, _gini_from_counts)


def random_classify_p_at_k(X, y, k=1):
    """strategy.py:162-172: the same with the precision-at-1 surrogate."""
    return _random_search(X, y, surrogate_p_at_k_improvements, _p_at_k_from_counts)


def random_split_v2(X, y, cal_improvements=None):
    """strategy.py:136-146: one U(min, max) threshold per column (np.random), best column by variance reduction.
    ``cal_improvements`` is unused in the reference as well."""
    from ._rows import RowSession, random_regression_node_search
    y = np.ascontiguousarray(y, dtype=np.float64)
    if np.asarray(X).shape[0] < 1:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    sess = RowSession(X, np.zeros(y.shape[0], dtype=np.int32), 1)
    try:
        sess.set_targets(y)
        best, thr, improvement, _, _ = random_regression_node_search(
            sess, 0, 0, sess.N, np.arange(sess.F), surrogate_variance_improvements, surrogate_to_variance)
        return best, thr, improvement, np.zeros(sess.F, dtype=bool)
    finally:
        sess.close()


def cal_variance_improvements(y, masks):
    """strategy.py:108-117: variance reduction of every column's (rows x columns) boolean left mask."""
    total_count = y.size
    total_sum = y.sum()
    left_sums = masks.T.dot(y)
    left_counts = masks.sum(axis=0)
    return ((np.square(left_sums) / left_counts + np.square(total_sum - left_sums) / (total_count - left_counts))
            / total_count - np.square(np.mean(y)))


def cal_gini_improvements(y, masks):
    """strategy.py:120-132: the same for one-hot targets (gini)."""
    total_count = y.shape[0]
    totals = y.sum(axis=0)
    left = masks.T.dot(y)
    left_counts = masks.sum(axis=0)
    return (np.square(left).sum(axis=1) / left_counts
            + np.square(totals - left).sum(axis=1) / (total_count - left_counts)
            - np.square(totals).sum() / total_count) / total_count


def _argsort_columns(X):
    """np.argsort(X, axis=0) on the GPU (K1, b200dt_argsort_columns): (n, F) int64, stable."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be 2-D")
    out = np.empty(X.shape, dtype=np.int64)
    if X.size:
        ctx = context_for()
        ctx.check(ctx.lib.b200dt_argsort_columns(ctx.handle, X.ctypes.data, X.shape[0], X.shape[1], X.shape[1],
                                                 _lib.HOST, out.ctypes.data, _lib.HOST))
    return out


def random_split(X, y, cal_improvements=cal_variance_improvements):
    """strategy.py:175-185: one random cutting rank per column (``np.random.choice``, the same call as the
    reference, so a seeded run reproduces it), columns compared by ``cal_improvements`` of the masks
    ``argsort(X) <= rank`` exactly as the reference forms them; the threshold is the column's value at that rank.
    The per-column argsort -- the O(n log n) part -- runs on the GPU."""
    indexes = _argsort_columns(X)
    cutting_rank = np.random.choice(X.shape[0] - 1, X.shape[1])
    improvements = cal_improvements(y, indexes <= cutting_rank)
    best_index = np.argmax(improvements)
    data_row_index = indexes[cutting_rank[best_index], best_index]
    return best_index, X[data_row_index, best_index], improvements[best_index], np.zeros(X.shape[1], dtype=bool)


def greedy_split(X, y):
    """strategy.py:188-208: exhaustive variance search of one node with the flat (row, column) argmax of
    sum_left^2 / n_left + sum_right^2 / n_right (``one.build_regression_tree`` is the whole-tree GPU form of the
    same search).  GPU argsort, then the reference's column cumsums."""
    orders = _argsort_columns(X)
    total_sum = y.sum()
    total_count = X.shape[0]
    left_sums = np.cumsum(y[orders], axis=0)[:-1]
    left_counts = np.arange(1, total_count)[:, np.newaxis]
    surrogate = np.square(left_sums) / left_counts + np.square(total_sum - left_sums) / (total_count - left_counts)
    best_row, best_column = np.unravel_index(np.argmax(surrogate), surrogate.shape)
    best_threshold = X[orders[best_row, best_column], best_column]
    best_improvement = surrogate[best_row, best_column] / total_count - np.square(np.mean(y))
    return best_column, best_threshold, best_improvement, np.zeros(X.shape[1], dtype=bool)
