"""Mirror of the reference module ``decision_tree/base.py``: the records and the tree container that the
classification builder (``tree_builder.py``) hands around.  Names, fields and semantics follow the
reference (citations: /root/reference/decision_tree/base.py); the implementation is this package's own.
"""
from collections import namedtuple

import numpy as np

# records (base.py:5-10) -- same type names and field order as the reference
Mask = namedtuple("DataMask", ["row", "column"])
Task = namedtuple("Task", ["mask", "parent", "is_left", "depth"])
BestSplit = namedtuple("BestSplit", ["attribute", "threshold", "constant_attrs", "improvement"])
Tree = namedtuple("DecisionTree", ["children_left", "children_right", "feature", "threshold", "value"])

_NO_CHILD, _NO_FEATURE, _NO_THRESHOLD = -1, -2, -2.0


def is_leaf(mask, y, min_sample_per_leaf=10):
    """base.py:13-21.  A node stops when it holds at most ``2 * min_sample_per_leaf`` rows, when its target
    block contains a single distinct value, or when no candidate column is left."""
    few_rows = np.count_nonzero(mask.row) <= 2 * min_sample_per_leaf
    return bool(few_rows or np.unique(y).size <= 1 or not np.any(mask.column))


def is_leaf_v2(orders, min_sample_per_leaf=10):
    """base.py:24-29: the same row test on an ``(n, F)`` order matrix."""
    return len(orders) <= 2 * min_sample_per_leaf


def _blank_arrays(size):
    """Five parallel arrays of an empty tree: no children, no feature / threshold; ``value`` is zero-filled
    (the reference leaves it uninitialised, base.py:36)."""
    child = np.full(size, _NO_CHILD, dtype=np.int64)
    return (child, child.copy(), np.full(size, _NO_FEATURE, dtype=np.int64),
            np.full(size, _NO_THRESHOLD, dtype=np.float64), np.zeros(size, dtype=np.float64))


class DecisionTree:
    """Tree container (base.py:32-69): ``tree_`` is a ``Tree`` of five arrays of length ``max_node - 1``,
    ``max_node_`` the id of the newest node.  Node ids are handed out in the order tasks are taken up."""

    def __init__(self, max_node):
        self.tree_ = Tree(*_blank_arrays(max_node - 1))
        self.max_node_ = -1

    def new_node_from_task(self, task):
        node = self.max_node_ = self.max_node_ + 1
        if task.parent is not None:   # hook the new node under its parent, on the side the task names
            side = "children_left" if task.is_left else "children_right"
            getattr(self.tree_, side)[task.parent] = node
        return node

    def add_leaf(self, node_id, value):
        self.tree_.value[node_id] = value

    def add_binary(self, node_id, best_split):
        self.add_binary_v2(node_id, best_split.attribute, best_split.threshold)

    def add_binary_v2(self, node_id, col, threshold):
        t = self.tree_
        t.feature[node_id], t.threshold[node_id] = col, threshold

    def final(self):
        return self
