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
from .base import DecisionTree
from . import strategy as _strategy
from .strategy import (greedy_classification, greedy_classification_p_at_k, random_classify,  # noqa: F401
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


def build_tree(X, y,
               max_depth=2,
               max_feature=100,
               min_improvement=0.01,
               min_sample_leaf=1,
               split_method=None,
               leaf_from_data=leaf_from_data_regression):
    """tree_builder.py:56-96.  Regression trees with random thresholds: ``split_method=random_split_v2`` (the
    reference's default here, ``greedy_split``, is the exhaustive MSE search that ``decision_tree.one`` supersedes:
    use ``one.build_regression_tree``).  One-hot ``y`` with a classification splitter is routed to
    ``build_classification_tree``."""
    if split_method in _CRITERION or split_method in _RANDOM:
        yy = np.asarray(y)
        if yy.ndim != 2:
            raise ValueError("classification splitters take the one-hot (n, C) targets build_classification_tree makes")
        return build_classification_tree(X, np.argmax(yy, axis=1), max_depth, max_feature, min_improvement,
                                         min_sample_leaf, split_method, leaf_from_data_classification, yy.shape[1])
    if split_method is not random_split_v2:
        raise NotImplementedError("build_tree: split_method must be this module's random_split_v2 (or a classification "
                                  "splitter with one-hot y); for the exhaustive MSE search use one.build_regression_tree")
    if leaf_from_data is not leaf_from_data_regression:
        raise NotImplementedError("regression leaves are the node mean (leaf_from_data_regression)")
    from ._rows import grow_random_regression
    return grow_random_regression(DecisionTree(2 ** (max_depth + 1)), X, y, int(max_depth), int(max_feature),
                                  float(min_improvement), int(min_sample_leaf), _strategy.surrogate_variance_improvements,
                                  _strategy.surrogate_to_variance).final()


def build_regression_tree(X, y,
                          max_depth=2,
                          max_feature=100,
                          min_improvement=0.01,
                          min_sample_leaf=1,
                          split_method=random_split,
                          leaf_from_data=leaf_from_data_regression):
    """tree_builder.py:138-146.  (The reference's default ``random_split`` is not accelerated: pass
    ``split_method=random_split_v2``.)"""
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
    if split_method not in _CRITERION and split_method not in _RANDOM:
        raise NotImplementedError("split_method must be one of this module's random_classify, "
                                  "random_classify_p_at_k, greedy_classification, greedy_classification_p_at_k")
    if leaf_from_data is not leaf_from_data_classification:
        raise NotImplementedError("classification leaves are the most frequent class (leaf_from_data_classification)")
    if split_method in _RANDOM:
        from ._rows import grow_random
        lab = y if (hasattr(y, "data_ptr") and not isinstance(y, np.ndarray)) else _checked_labels(y, max_classes)
        surrogate, to_improvement = _RANDOM[split_method]
        return grow_random(DecisionTree(2 ** (max_depth + 1)), X, lab, int(max_classes), int(max_depth),
                           int(max_feature), float(min_improvement), int(min_sample_leaf), surrogate,
                           to_improvement).final()
    xp, xs, xk, N, F, ldx = matrix_arg(X)
    if F > max_feature:
        raise NotImplementedError("F > max_feature samples columns with np.random per node "
                                  "(tree_builder.py:44-48); pass max_feature >= X.shape[1]")
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
