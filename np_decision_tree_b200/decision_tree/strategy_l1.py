"""Mirror of the reference module ``decision_tree/strategy_l1.py`` -- L1 (MAE) regression trees.

Citations are to /root/reference/decision_tree/strategy_l1.py.
"""
import ctypes

import numpy as np

from .. import _lib
from ._common import context_for, fit_into, matrix_arg, orders_arg
from .one import DecisionTree, Task, split_orders  # noqa: F401  (strategy_l1.py:3)
from .np_stream_median import cummedian  # noqa: F401  (strategy_l1.py:4)


def running_median_v2(a):
    """strategy_l1.py:11-13: the (lower, upper) running median of ``a`` -- one kernel call here."""
    return cummedian(np.asarray(a, dtype=np.float64))


def _leaves_from(tree, X, y, leaf_from_data):
    """Leaf values by an arbitrary reducer (strategy_l1.py:154,184 call ``leaf_from_data(y[task.orders])`` on the
    (n_node, F) gather): rows are routed to their leaves on the GPU (inference_leaf), the reducer sees each leaf's
    targets replicated over the F columns.  Exact for reducers that do not depend on the order inside a column."""
    from .utils import inference_leaf
    Xh = X.cpu().numpy() if hasattr(X, "is_cuda") else np.asarray(X)
    yh = y.cpu().numpy() if hasattr(y, "is_cuda") else np.asarray(y, dtype=np.float64)
    leaf = np.asarray(inference_leaf(Xh, tree)).ravel()
    if leaf.shape[0] != yh.shape[0]:            # root-only tree: the (1,) quirk of utils.py:37
        leaf = np.zeros(yh.shape[0], dtype=np.int64)
    order = np.argsort(leaf, kind="stable")
    ids, starts = np.unique(leaf[order], return_index=True)
    bounds = np.append(starts, leaf.shape[0])
    F = Xh.shape[1]
    for i, node in enumerate(ids):
        rows = order[bounds[i]:bounds[i + 1]]
        tree.tree_.value[node] = leaf_from_data(np.broadcast_to(yh[rows][:, None], (rows.shape[0], F)))
    return tree


def best_l1_improvements_v2(ys):
    """(row, column, children L1 cost) of the best split of ``ys`` (n, F) (strategy_l1.py:87-122).
    The cost is minimised; ties go to the lowest row, then the lowest column (flat argmin).

    The columns of an arbitrary ``ys`` need not be permutations of one another, so each column is its own
    one-feature problem; all F of them are searched in one batched call (``b200dt_l1_best_columns``)."""
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    if ys.ndim != 2:
        raise ValueError("ys must be (n, F)")
    n, F = ys.shape
    ctx = context_for()
    k, f, cost = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    ctx.check(ctx.lib.b200dt_l1_best_columns(ctx.handle, ys.ctypes.data, n, F, ctypes.byref(k), ctypes.byref(f),
                                             ctypes.byref(cost)))
    return k.value, f.value, cost.value


best_l1_improvements = best_l1_improvements_v2   # strategy_l1.py:49-84: same results, pure Python there


def greedy_regression_split_v2(orders, X, y):
    """strategy_l1.py:134-140 -> (threshold, column, cost, left_rows)."""
    orders = orders_arg(orders)
    ctx = context_for()
    xp, xs, xk, N, F, ldx = matrix_arg(X)
    y = np.ascontiguousarray(y, dtype=np.float64)
    k, f = ctypes.c_int64(), ctypes.c_int64()
    cost, thr = ctypes.c_double(), ctypes.c_double()
    ctx.check(ctx.lib.b200dt_l1_split_node(ctx.handle, orders.ctypes.data, orders.shape[0], F, xp, ldx,
                                           y.ctypes.data, y.shape[0], ctypes.byref(k), ctypes.byref(f),
                                           ctypes.byref(cost), ctypes.byref(thr)))
    return thr.value, f.value, cost.value, orders[:k.value + 1, f.value]


greedy_regression_split = greedy_regression_split_v2   # strategy_l1.py:125-131


def build_regression_tree_v2(X, y, max_depth=2, max_feature=100, min_improvement=1e-7, min_sample_leaf=1,
                             leaf_from_data=np.median):
    """Fit an L1 regression tree (strategy_l1.py:173-200).  A node becomes a leaf when its best
    children COST is below ``min_improvement`` (strategy_l1.py:188), as in the reference."""
    tree = DecisionTree(2 ** (max_depth + 1))
    fit_into(tree, X, y, _lib.L1, max_depth, min_improvement, min_sample_leaf, 1)
    if leaf_from_data is not np.median:     # np.median leaves come out of the fit itself (the y-sorted column)
        _leaves_from(tree, X, y, leaf_from_data)
    return tree.final()


def build_regression_tree(X, y, max_depth=2, max_feature=100, min_improvement=1e-7, min_sample_leaf=1,
                          leaf_from_data=np.median):
    """strategy_l1.py:143-170: the pure-Python-heap twin builds the same tree; same kernel path."""
    return build_regression_tree_v2(X, y, max_depth, max_feature, min_improvement, min_sample_leaf,
                                    leaf_from_data)
