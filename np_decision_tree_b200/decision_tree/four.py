"""Mirror of the reference module ``decision_tree/four.py`` -- the level-synchronous L2 builder.

Same search as ``one.py`` (four.py:11-20 == one.py:56-63) with the conventions of this module:
node ids level by level (all left children of a level, then all right children), a node is a
leaf when its best gain is ``<= min_improvement``, and the tree carries ``n_node_samples``.
Citations are to /root/reference/decision_tree/four.py.
"""
import ctypes
from collections import namedtuple

import numpy as np

from .. import _lib
from ._common import context_for, matrix_arg

Tree = namedtuple('DecisionTree', 'children_left children_right feature threshold value n_node_samples')


class DecisionTree:
    """four.py:61-98: the one.py container plus int32 ``n_node_samples`` (-2 in unused slots)."""

    def __init__(self, max_node):
        size = max_node - 1
        self.tree_ = Tree(np.full(size, -1, dtype=np.int64), np.full(size, -1, dtype=np.int64),
                          np.full(size, -2, dtype=np.int64), np.full(size, -2.0, dtype=np.float64),
                          np.zeros(size, dtype=np.float64), np.full(size, -2, dtype=np.int32))
        self.max_node_ = -1

    def add_node(self, parent=None, is_left=None):
        self.max_node_ += 1
        if parent is not None:
            side = self.tree_.children_left if is_left else self.tree_.children_right
            side[parent] = self.max_node_
        return self.max_node_

    def add_size(self, node_id, size):
        self.tree_.n_node_samples[node_id] = size

    def add_leaf(self, node_id, value):
        self.tree_.value[node_id] = value

    def add_binary(self, node_id, col, threshold):
        self.tree_.threshold[node_id] = threshold
        self.tree_.feature[node_id] = col


def build_regression_tree(X, y, max_depth=2, min_improvement=1e-7, min_sample_leaf=1, weight=None):
    """four.py:108-172.  ``weight``: optional (N,) sample weights (four.py:116-117, 21-30)."""
    tree = DecisionTree(2 ** (max_depth + 1))
    ctx = context_for(X, y)
    xp, xs, xk, N, F, ldx = matrix_arg(X)
    yp, ys, yk = _lib.as_arg(y, np.float64, "y")
    if xs != ys:
        raise ValueError("X and y must both be host arrays or both be CUDA tensors")
    wp = None
    if weight is not None:
        wp, ws, wk = _lib.as_arg(weight, np.float64, "weight")
        if ws != xs:
            raise ValueError("weight must live where X and y live")
        if int(np.prod(weight.shape)) != N:
            raise ValueError("weight must have one entry per row")
    t = tree.tree_
    count = ctypes.c_int64(0)
    ctx.check(ctx.lib.b200dt_fit_ex(
        ctx.handle, xp, yp, wp, N, F, ldx, xs, _lib.L2, int(max_depth), float(min_improvement), int(min_sample_leaf), 1,
        _lib.NUMBER_BY_LEVEL | _lib.LEAF_IF_LE, t.children_left.ctypes.data, t.children_right.ctypes.data,
        t.feature.ctypes.data, t.threshold.ctypes.data, t.value.ctypes.data, t.n_node_samples.ctypes.data,
        t.feature.shape[0], ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    return tree
