"""Host driver of the random-threshold classification splitters (csrc/rows.cu).

The reference draws from numpy's GLOBAL generator per node, in depth-first order
(ref decision_tree/strategy.py:150, tree_builder.py:44-48).  To stay a drop-in under
``np.random.seed`` this driver keeps that loop and makes the very same ``np.random`` calls with the
same arguments; the array work of every node (column min / max, class counts left of the thresholds,
the row split) runs on the GPU over the node's row list.
"""
import ctypes

import numpy as np

from .. import _lib
from ._common import context_for, matrix_arg


class RowSession:
    """Row lists of the nodes of one tree on the device (b200dt_rows_*)."""

    def __init__(self, X, labels, n_classes):
        self.ctx = context_for(X, labels)
        xp, xs, self._xk, self.N, self.F, ldx = matrix_arg(X)
        if hasattr(labels, "data_ptr") and not isinstance(labels, np.ndarray):
            import torch
            lab = labels.to(torch.int32).contiguous()
            lp, ls, self._lk = _lib.as_arg(lab, np.int32, "y")
        else:
            lab = np.ascontiguousarray(labels, dtype=np.int32)
            lp, ls, self._lk = lab.ctypes.data, _lib.HOST, lab
        if xs != ls:
            raise ValueError("X and y must both be host arrays or both be CUDA tensors")
        if int(lab.shape[0]) != self.N:
            raise ValueError(f"y has {int(lab.shape[0])} entries, X has {self.N} rows")
        self.C = int(n_classes)
        counts = np.zeros(self.C, dtype=np.int64)
        self.ctx.check(self.ctx.lib.b200dt_rows_begin(self.ctx.handle, xp, lp, self.N, self.F, ldx, xs, self.C,
                                                      counts.ctypes.data))
        self.root_counts = counts

    def minmax(self, buffer, start, n, cols):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        lo, hi = np.empty(cols.size), np.empty(cols.size)
        self.ctx.check(self.ctx.lib.b200dt_rows_minmax(self.ctx.handle, buffer, start, n, cols.ctypes.data, cols.size,
                                                       lo.ctypes.data, hi.ctypes.data))
        return lo, hi

    def threshold_counts(self, buffer, start, n, cols, thresholds):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        thr = np.ascontiguousarray(thresholds, dtype=np.float64)
        out = np.empty((cols.size, self.C), dtype=np.int64)
        self.ctx.check(self.ctx.lib.b200dt_rows_threshold_counts(self.ctx.handle, buffer, start, n, cols.ctypes.data,
                                                                 cols.size, thr.ctypes.data, out.ctypes.data))
        return out

    def set_targets(self, y):
        """Register the float64 regression targets (b200dt_rows_set_targets)."""
        yp, ys, self._yk = _lib.as_arg(y, np.float64, "y")
        self.ctx.check(self.ctx.lib.b200dt_rows_set_targets(self.ctx.handle, yp, ys))

    def threshold_sums(self, buffer, start, n, cols, thresholds):
        """(sum of y left of each threshold, rows left of it, (sum, min, max) of y over the node)."""
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        thr = np.ascontiguousarray(thresholds, dtype=np.float64)
        sums, counts, node = np.empty(cols.size), np.empty(cols.size, dtype=np.int64), np.empty(3)
        self.ctx.check(self.ctx.lib.b200dt_rows_threshold_sums(self.ctx.handle, buffer, start, n, cols.ctypes.data,
                                                               cols.size, thr.ctypes.data, sums.ctypes.data,
                                                               counts.ctypes.data, node.ctypes.data))
        return sums, counts, node

    def split(self, buffer, start, n, col, threshold, expected=None):
        """Stable split of the node's row list.  With ``expected`` (the left count the search already produced) the
        split is only queued: no host round trip per node."""
        if expected is not None:
            self.ctx.check(self.ctx.lib.b200dt_rows_split(self.ctx.handle, buffer, start, n, int(col), float(threshold),
                                                          None))
            return int(expected)
        n_left = ctypes.c_int64()
        self.ctx.check(self.ctx.lib.b200dt_rows_split(self.ctx.handle, buffer, start, n, int(col), float(threshold),
                                                      ctypes.byref(n_left)))
        return int(n_left.value)

    def close(self):
        self.ctx.check(self.ctx.lib.b200dt_rows_end(self.ctx.handle))


def random_node_search(sess, buffer, start, n, cols, class_counts, surrogate, to_improvement):
    """One node: thresholds ~ U(column min, column max) from np.random, first argmax of the surrogate.

    ref strategy.py:149-159 / :162-172.  Returns (index into cols, threshold, improvement, left class counts)."""
    lo, hi = sess.minmax(buffer, start, n, cols)
    thresholds = np.random.uniform(lo, hi)                                   # strategy.py:150
    left = sess.threshold_counts(buffer, start, n, cols, thresholds)        # == masks.T.dot(y), whole numbers
    with np.errstate(divide="ignore", invalid="ignore"):                     # a constant column gives 0/0 = nan
        scores = surrogate(left.astype(np.float64)[np.newaxis, :], left.sum(axis=1),
                           class_counts.astype(np.float64), n).ravel()
    best = int(np.argmax(scores))
    return best, thresholds[best], to_improvement(scores[best], class_counts.astype(np.float64), n), left[best]


def grow_random(tree, X, labels, n_classes, max_depth, max_feature, min_improvement, min_sample_leaf, surrogate,
                to_improvement):
    """tree_builder.py:56-96 (build_tree) with a random splitter: depth-first, left child first."""
    from .base import Task
    sess = RowSession(X, labels, n_classes)
    try:
        all_cols = np.arange(sess.F)
        pending = [(0, sess.N, 0, sess.root_counts, None, True, 0)]
        while pending:
            start, n, buffer, counts, parent, is_left, depth = pending.pop()
            node_id = tree.new_node_from_task(Task(None, parent, is_left, depth))
            # base.py:13-21: rows <= 2 * min_sample_leaf, or a one-hot block with a single distinct value
            leaf = n <= 2 * min_sample_leaf or n == 0 or n_classes == 1 or depth >= max_depth
            if not leaf:
                cols = all_cols if all_cols.size <= max_feature else np.random.permutation(all_cols)[:max_feature]
                best, threshold, improvement, left = random_node_search(sess, buffer, start, n, cols, counts,
                                                                        surrogate, to_improvement)
                leaf = improvement < min_improvement
            if leaf:
                tree.add_leaf(node_id, np.argmax(counts))
                continue
            tree.add_binary_v2(node_id, cols[best], threshold)
            n_left = sess.split(buffer, start, n, cols[best], threshold, expected=int(left.sum()))
            pending.append((start + n_left, n - n_left, buffer ^ 1, counts - left, node_id, False, depth + 1))
            pending.append((start, n_left, buffer ^ 1, left, node_id, True, depth + 1))
    finally:
        sess.close()
    return tree


def random_regression_node_search(sess, buffer, start, n, cols, surrogate, to_variance):
    """One node of the regression ExtraTree (ref strategy.py:136-146, random_split_v2): thresholds from np.random,
    left sums / counts from the GPU, the tiny surrogate + argmax in numpy like the reference.
    Returns (index into cols, threshold, improvement, node y sum, node y min, node y max, rows left of the winner)."""
    lo, hi = sess.minmax(buffer, start, n, cols)
    thresholds = np.random.uniform(lo, hi)
    left_sums, left_counts, node = sess.threshold_sums(buffer, start, n, cols, thresholds)
    with np.errstate(divide="ignore", invalid="ignore"):
        scores = surrogate(left_sums, left_counts, node[0], n)
    best = int(np.argmax(scores))
    return best, thresholds[best], to_variance(scores[best], n, node[0] / n), node, int(left_counts[best])


def grow_random_regression(tree, X, y, max_depth, max_feature, min_improvement, min_sample_leaf, surrogate,
                           to_variance):
    """tree_builder.py:56-96 (build_tree) with split_method=random_split_v2 and leaf_from_data_regression."""
    from .base import Task
    ny = int(np.prod(y.shape))
    sess = RowSession(X, np.zeros(ny, dtype=np.int32) if not (hasattr(y, "is_cuda") and y.is_cuda)
                      else __import__("torch").zeros(ny, dtype=__import__("torch").int32, device=y.device), 1)
    try:
        sess.set_targets(y)
        all_cols = np.arange(sess.F)
        first = np.array([0], dtype=np.int32)
        pending = [(0, sess.N, 0, None, True, 0)]
        while pending:
            start, n, buffer, parent, is_left, depth = pending.pop()
            node_id = tree.new_node_from_task(Task(None, parent, is_left, depth))
            if n == 0:   # an empty child (every row went to the other side): np.mean([]) in the reference
                tree.add_leaf(node_id, np.nan)
                continue
            stop = n <= 2 * min_sample_leaf or depth >= max_depth
            node = None
            if not stop:
                # base.py:16: a node whose targets are all equal is a leaf -- known from the node's min / max.  The
                # reference tests it BEFORE find_best_split draws the column permutation (tree_builder.py:70-78,
                # 44-48), so look first: a pure node must not consume a draw of the global generator
                _, _, node = sess.threshold_sums(buffer, start, n, first, np.array([np.inf]))
                if node[1] == node[2]:
                    stop = True
                else:
                    cols = (all_cols if all_cols.size <= max_feature
                            else np.random.permutation(all_cols)[:max_feature])
                    best, threshold, improvement, node, n_left_expected = random_regression_node_search(
                        sess, buffer, start, n, cols, surrogate, to_variance)
                    stop = improvement < min_improvement
            if stop:
                if node is None:
                    _, _, node = sess.threshold_sums(buffer, start, n, first, np.array([np.inf]))
                tree.add_leaf(node_id, node[0] / n)       # leaf_from_data_regression: np.mean (tree_builder.py:16-17)
                continue
            tree.add_binary_v2(node_id, cols[best], threshold)
            n_left = sess.split(buffer, start, n, cols[best], threshold, expected=n_left_expected)
            pending.append((start + n_left, n - n_left, buffer ^ 1, node_id, False, depth + 1))
            pending.append((start, n_left, buffer ^ 1, node_id, True, depth + 1))
    finally:
        sess.close()
    return tree
