"""Mirror of the reference module ``decision_tree/one.py`` -- L2 (MSE) regression trees.

Same names, positional order and defaults as the reference; the numeric work runs on the GPU
through libb200tree.so (include/b200dt.h).  Citations are to /root/reference/decision_tree/one.py.
"""
import ctypes
from collections import namedtuple

import numpy as np

from .. import _lib
from ._common import context_for, fit_into, matrix_arg, orders_arg

# one.py:4-6
Task = namedtuple('Task', 'orders parent is_left depth')
Tree = namedtuple('DecisionTree', 'children_left children_right feature threshold value')


class DecisionTree:
    """Tree container with the reference's layout and methods (one.py:9-49).

    ``tree_`` = five arrays of length ``max_node - 1``: int64 children_left / children_right
    (-1 = none), int64 feature (-2 on leaves), float64 threshold (-2.0 on leaves), float64
    value.  The reference leaves ``value`` uninitialised on internal nodes; this build
    stores the node mean there.  ``max_node_`` = id of the last node.
    """

    def __init__(self, max_node):
        size = max_node - 1
        self.tree_ = Tree(np.full(size, -1, dtype=np.int64), np.full(size, -1, dtype=np.int64),
                          np.full(size, -2, dtype=np.int64), np.full(size, -2.0, dtype=np.float64),
                          np.zeros(size, dtype=np.float64))
        self.max_node_ = -1

    def new_node_from_task(self, task):
        self.max_node_ += 1
        if task.parent is not None:
            side = self.tree_.children_left if task.is_left else self.tree_.children_right
            side[task.parent] = self.max_node_
        return self.max_node_

    def add_leaf(self, node_id, value):
        self.tree_.value[node_id] = value

    def add_binary(self, node_id, col, threshold):
        self.tree_.threshold[node_id] = threshold
        self.tree_.feature[node_id] = col

    @staticmethod
    def is_leaf(orders, min_sample_per_leaf=10):
        return orders.shape[0] <= min_sample_per_leaf

    def final(self):
        return self


def np_gene(n):
    """Split weights 1/(left*right) (one.py:52-53).  Host helper kept for API parity; the
    search kernels compute the same weights on the fly."""
    return np.reciprocal(np.arange(1, n, dtype=np.float32) * np.arange(n - 1, 0, -1))


def _score_cumsums(ys, variant):
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    if ys.ndim != 2:
        raise ValueError("cumsums must be (n, F)")
    ctx = context_for()
    k, f, score = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    ctx.check(ctx.lib.b200dt_l2_score_cumsums(ctx.handle, ys.ctypes.data, ys.shape[0], ys.shape[1], variant,
                                              ctypes.byref(k), ctypes.byref(f), ctypes.byref(score)))
    return k.value, f.value, score.value


def best_variance_improvements(ys):
    """(row, column, gain) of the best split of the column cumsums ``ys`` (n, F) (one.py:56-63): per row the
    feature with the largest |S - (k+1) mean|, then the row with the largest weighted square."""
    return _score_cumsums(ys, 1)


def best_variance_improvements_v0(cumsums):
    """one.py:66-78: float32 ratio weights, flat argmax of the SQUARED scores."""
    return _score_cumsums(cumsums, 0)


def best_variance_improvements_v2(cumsums):
    """one.py:81-93: float32 ratio weights, flat argmax of the absolute scores; returns the square of the best."""
    return _score_cumsums(cumsums, 2)


def greedy_regression_split(orders, x, y, v):
    """Best split of one node (one.py:96-107) -> (threshold, column, gain, left_rows).

    ``orders`` is the (n, F) int64 matrix of per-column sorted row ids, as in the reference."""
    orders = orders_arg(orders)
    ctx = context_for()
    xp, xs, xk, N, F, ldx = matrix_arg(x, "x")
    y = np.ascontiguousarray(y, dtype=np.float64)
    if xs != _lib.HOST:
        raise TypeError("greedy_regression_split takes host arrays (like the reference)")
    if orders.shape[1] != F:
        raise ValueError("orders and x disagree on the number of features")
    k, f = ctypes.c_int64(), ctypes.c_int64()
    gain, thr = ctypes.c_double(), ctypes.c_double()
    ctx.check(ctx.lib.b200dt_l2_split_node(
        ctx.handle, orders.ctypes.data, orders.shape[0], F, xp, ldx, y.ctypes.data, y.shape[0],
        1 if v == 1 else 2, ctypes.byref(k), ctypes.byref(f), ctypes.byref(gain), ctypes.byref(thr)))
    left_rows = orders[:, f.value][:k.value + 1]
    return thr.value, f.value, gain.value, left_rows


def split_orders(orders, index, n_samples):
    """Stable per-column partition (one.py:110-118) -> (left_orders, right_orders)."""
    orders = orders_arg(orders)
    index = np.ascontiguousarray(index, dtype=np.int64)
    ctx = context_for()
    n, F = orders.shape
    n_left = index.shape[0]
    left = np.empty((n_left, F), dtype=np.int64)
    right = np.empty((n - n_left, F), dtype=np.int64)
    ctx.check(ctx.lib.b200dt_split_orders(ctx.handle, orders.ctypes.data, n, F, index.ctypes.data, n_left,
                                          int(n_samples), left.ctypes.data, right.ctypes.data))
    return left, right


def split_orders_v2(orders, index, n_samples):
    """one.py:121-129 (unused in the reference): the same partition through an argsort of the masks.  Its
    default-kind argsort is not stable and its reshape assumes ``orders.shape[1] == n_samples``, so the reference's
    own result is only defined through ``split_orders``; this delegates to it."""
    return split_orders(orders, index, n_samples)


leaf_from_data = np.mean   # one.py:132


def build_regression_tree(X, y, max_depth=2, max_feature=100, min_improvement=1e-7, min_sample_leaf=1, v=1):
    """Fit an L2 regression tree (one.py:135-162).  ``max_feature`` is accepted and ignored,
    as in the reference.  X, y: numpy arrays, or torch CUDA tensors (no host copy)."""
    tree = DecisionTree(2 ** (max_depth + 1))
    return fit_into(tree, X, y, _lib.L2, max_depth, min_improvement, min_sample_leaf,
                    1 if v == 1 else 2).final()


def build_regression_tree_v1_5(X, y, max_depth=2, max_feature=100, min_improvement=1e-7, min_sample_leaf=1, v=1):
    """The reference's deque twin (one.py:165-192) builds the same tree; same kernel path."""
    return build_regression_tree(X, y, max_depth, max_feature, min_improvement, min_sample_leaf, v)
