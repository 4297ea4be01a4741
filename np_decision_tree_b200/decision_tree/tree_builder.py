"""Mirror of ``decision_tree/tree_builder.py`` for the exhaustive classification trees.

``build_classification_tree`` keeps the reference's signature and defaults.  The whole fit --
one sort, then per level the class-count scans, the gini / precision-at-1 scores of every split
position of every live node, the decisions and the stable partition -- runs on the GPU
(``b200dt_fit_classes``, csrc/classes.cu) and returns the reference's ``DecisionTree`` with
pre-order node ids.  Citations are to /root/reference/decision_tree/tree_builder.py.
"""
import ctypes

import numpy as np

from .. import _lib
from ._common import context_for, matrix_arg
from .base import BestSplit, DecisionTree, Mask, Task, is_leaf
from . import strategy as _strategy
from .strategy import (greedy_classification, greedy_classification_p_at_k, greedy_split, random_classify,  # noqa: F401
                       random_classify_p_at_k, random_split, random_split_v2)


def leaf_from_data_regression(y):
    """tree_builder.py:16-17."""
    return np.mean(y)


def leaf_from_data_classification(y):
    """tree_builder.py:20-21: first most frequent class of a one-hot block."""
    return np.argmax(y.sum(axis=0))


def _checked_labels(y, max_classes):
    """int32 class ids with numpy's index semantics of ``np.eye(max_classes)[y]`` (tree_builder.py:157-158)."""
    lab = np.asarray(y)
    if lab.ndim != 1 or not np.issubdtype(lab.dtype, np.integer):
        raise IndexError("y must be a 1-D array of integer class ids (it indexes np.eye(max_classes))")
    if lab.size and (lab.min() < -max_classes or lab.max() >= max_classes):
        raise IndexError(f"index out of bounds for axis 0 with size {max_classes}")
    return np.ascontiguousarray(np.where(lab < 0, lab + max_classes, lab), dtype=np.int32)


_CRITERION = {greedy_classification: _lib.GINI, greedy_classification_p_at_k: _lib.P_AT_1}
_RANDOM = {random_classify: (_strategy.surrogate_gini_improvements, _strategy._gini_from_counts),
           random_classify_p_at_k: (_strategy.surrogate_p_at_k_improvements, _strategy._p_at_k_from_counts)}


def init_mask(n_row, n_col):
    """tree_builder.py:10-12."""
    return Mask(np.ones(n_row, dtype=bool), np.ones(n_col, dtype=bool))


def find_best_split(X, y, candidate_attributes, max_feature=100, split_method=random_split):
    """tree_builder.py:43-52: at most ``max_feature`` candidate columns (a ``np.random.permutation`` draw when there
    are more, the same call as the reference), searched by ``split_method``."""
    if candidate_attributes.size <= max_feature:
        attributes = candidate_attributes
    else:
        attributes = np.random.permutation(candidate_attributes)[:max_feature]
    best_index, best_threshold, best_improvement, constant_attr_mask = split_method(X[:, attributes], y)
    return BestSplit(attributes[best_index], best_threshold, attributes[constant_attr_mask], best_improvement)


def _grow_node_at_a_time(X, y, max_depth, max_feature, min_improvement, min_sample_leaf, split_method, leaf_from_data):
    """The reference's own control flow (tree_builder.py:56-96) for splitters that cannot be level-synchronous:
    anything that draws from ``np.random`` per node in depth-first order (column sub-sampling, ``random_split``),
    or an arbitrary user callable.  One node at a time, left child first; the splitter gets the node's rows and
    candidate columns and does its array work on the GPU (per-column argsort, exhaustive class search, ...); rows
    go left by value (``x <= threshold``, tree_builder.py:24-41).  Nodes are kept as row-index lists instead of the
    reference's full-length boolean masks."""
    X = np.asarray(X)
    y = np.asarray(y)
    tree = DecisionTree(2 ** (max_depth + 1))
    pending = [(np.arange(X.shape[0]), np.ones(X.shape[1], dtype=bool), None, True, 0)]
    while pending:
        rows, columns, parent, is_left, depth = pending.pop()
        node_id = tree.new_node_from_task(Task(None, parent, is_left, depth))
        Xf, yf = X[rows], y[rows]
        row_mask = Mask(np.ones(rows.shape[0], dtype=bool), columns)
        best = None
        if not (is_leaf(row_mask, yf, min_sample_leaf) or depth >= max_depth):
            best = find_best_split(Xf, yf, np.flatnonzero(columns), max_feature, split_method)
            if best.improvement < min_improvement:
                best = None
        if best is None:
            tree.add_leaf(node_id, leaf_from_data(yf))
            continue
        tree.add_binary(node_id, best)
        keep = columns.copy()
        keep[best.constant_attrs] = False
        goes_left = Xf[:, best.attribute] <= best.threshold
        pending.append((rows[~goes_left], keep, node_id, False, depth + 1))
        pending.append((rows[goes_left], keep, node_id, True, depth + 1))
    return tree.final()


def build_tree(X, y,
               max_depth=2,
               max_feature=100,
               min_improvement=0.01,
               min_sample_leaf=1,
               split_method=greedy_split,
               leaf_from_data=leaf_from_data_regression):
    """tree_builder.py:56-96.  Fast paths: a classification splitter with one-hot ``y`` and all columns in play is
    the level-synchronous GPU fit (``build_classification_tree``); ``random_split_v2`` with mean leaves runs over
    device row lists (csrc/rows.cu).  Everything else -- the reference's defaults ``greedy_split`` / ``random_split``,
    column sub-sampling under an exhaustive search, any user ``split_method`` or ``leaf_from_data`` -- takes the
    reference's node-at-a-time control flow with the splitter's array work on the GPU."""
    yy = np.asarray(y) if not hasattr(y, "is_cuda") else None
    n_cols = int(X.shape[1])
    if (split_method in _CRITERION or split_method in _RANDOM) and yy is not None and yy.ndim == 2 \
            and leaf_from_data is leaf_from_data_classification \
            and (split_method in _RANDOM or n_cols <= max_feature):
        return build_classification_tree(X, np.argmax(yy, axis=1), max_depth, max_feature, min_improvement,
                                         min_sample_leaf, split_method, leaf_from_data_classification, yy.shape[1])
    if split_method is random_split_v2 and leaf_from_data is leaf_from_data_regression:
        from ._rows import grow_random_regression
        return grow_random_regression(DecisionTree(2 ** (max_depth + 1)), X, y, int(max_depth), int(max_feature),
                                      float(min_improvement), int(min_sample_leaf),
                                      _strategy.surrogate_variance_improvements, _strategy.surrogate_to_variance).final()
    if hasattr(X, "is_cuda"):
        raise TypeError("the node-at-a-time builders take host arrays (like the reference)")
    return _grow_node_at_a_time(X, y, int(max_depth), int(max_feature), float(min_improvement), int(min_sample_leaf),
                                split_method, leaf_from_data)


def build_regression_tree(X, y,
                          max_depth=2,
                          max_feature=100,
                          min_improvement=0.01,
                          min_sample_leaf=1,
                          split_method=random_split,
                          leaf_from_data=leaf_from_data_regression):
    """tree_builder.py:138-146 (default splitter: ``random_split``, one random cutting rank per column)."""
    return build_tree(X, y, max_depth, max_feature, min_improvement, min_sample_leaf, split_method, leaf_from_data)


def build_classification_tree(X, y,
                              max_depth=2,
                              max_feature=100,
                              min_improvement=0.01,
                              min_sample_leaf=1,
                              split_method=random_classify,
                              leaf_from_data=leaf_from_data_classification,
                              max_classes=2):
    """tree_builder.py:149-159 -> build_tree (:56-96).  ``y`` holds integer class ids in
    ``[0, max_classes)``.  ``split_method``: this package's ``random_classify`` (the reference's default)
    / ``random_classify_p_at_k`` -- node-at-a-time, thresholds and column sub-samples drawn from
    ``np.random`` exactly like the reference -- or the exhaustive ``greedy_classification`` /
    ``greedy_classification_p_at_k`` (level-synchronous, needs ``max_feature >= X.shape[1]``)."""
    if (split_method not in _CRITERION and split_method not in _RANDOM) or \
            leaf_from_data is not leaf_from_data_classification:
        # tree_builder.py:157-159: one-hot targets, then build_tree with whatever the caller plugged in
        lab = _checked_labels(y, max_classes)
        return _grow_node_at_a_time(np.asarray(X), np.eye(max_classes)[lab], int(max_depth), int(max_feature),
                                    float(min_improvement), int(min_sample_leaf), split_method, leaf_from_data)
    if split_method in _RANDOM:
        from ._rows import grow_random
        lab = y if (hasattr(y, "data_ptr") and not isinstance(y, np.ndarray)) else _checked_labels(y, max_classes)
        surrogate, to_improvement = _RANDOM[split_method]
        return grow_random(DecisionTree(2 ** (max_depth + 1)), X, lab, int(max_classes), int(max_depth),
                           int(max_feature), float(min_improvement), int(min_sample_leaf), surrogate,
                           to_improvement).final()
    xp, xs, xk, N, F, ldx = matrix_arg(X)
    if F > max_feature:
        # column sub-sampling draws np.random.permutation per node in depth-first order (tree_builder.py:44-48):
        # keep that order; every node's exhaustive search still runs on the GPU (b200dt_class_split_node)
        if xs != _lib.HOST:
            raise TypeError("column sub-sampling (F > max_feature) takes host arrays (like the reference)")
        lab = _checked_labels(y, max_classes)
        return _grow_node_at_a_time(np.asarray(X), np.eye(max_classes)[lab], int(max_depth), int(max_feature),
                                    float(min_improvement), int(min_sample_leaf), split_method,
                                    leaf_from_data_classification)
    if hasattr(y, "data_ptr") and not isinstance(y, np.ndarray):
        import torch
        lab = y.to(torch.int32).contiguous()
        lp, ls, lk = _lib.as_arg(lab, np.int32, "y")
    else:
        lab = _checked_labels(y, max_classes)
        lp, ls, lk = lab.ctypes.data, _lib.HOST, lab
    if xs != ls:
        raise ValueError("X and y must both be host arrays or both be CUDA tensors")
    if int(lab.shape[0]) != N:
        raise ValueError(f"y has {int(lab.shape[0])} entries, X has {N} rows")
    ctx = context_for(X, y)
    tree = DecisionTree(2 ** (max_depth + 1))
    t = tree.tree_
    count = ctypes.c_int64(0)
    ctx.check(ctx.lib.b200dt_fit_classes(
        ctx.handle, xp, lp, N, F, ldx, xs, int(max_classes), _CRITERION[split_method], int(max_depth),
        float(min_improvement), int(min_sample_leaf), t.children_left.ctypes.data, t.children_right.ctypes.data,
        t.feature.ctypes.data, t.threshold.ctypes.data, t.value.ctypes.data, None, t.feature.shape[0],
        ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    return tree.final()
