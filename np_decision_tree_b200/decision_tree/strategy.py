"""Mirror of the exhaustive classification searches of ``decision_tree/strategy.py``.

``greedy_classification`` / ``greedy_classification_p_at_k`` keep the reference's signature
(``X`` (n, F), ``y`` one-hot (n, C)) and return value ``(column, threshold, improvement,
constant_attr_mask)``; the search itself -- per-column sort, running class counts, gini or
precision-at-1 surrogate of every split position, flat argmax -- runs on the GPU
(``b200dt_class_split_node``, csrc/classes.cu).  Citations are to
/root/reference/decision_tree/strategy.py.

The random-threshold splitters (``random_classify`` ..., strategy.py:136-185) draw from
``np.random`` per node and are not part of the accelerated path.
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


def _not_accelerated(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"{name} draws random thresholds from np.random per node; only the exhaustive searches "
            "(greedy_classification, greedy_classification_p_at_k) run on the GPU path")
    fn.__name__ = name
    return fn


random_classify = _not_accelerated("random_classify")
random_classify_p_at_k = _not_accelerated("random_classify_p_at_k")
random_split = _not_accelerated("random_split")
random_split_v2 = _not_accelerated("random_split_v2")
