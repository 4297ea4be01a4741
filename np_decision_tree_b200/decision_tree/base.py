"""Mirror of the reference module ``decision_tree/base.py`` -- containers of the classification
builder.  Citations are to /root/reference/decision_tree/base.py."""
from collections import namedtuple

import numpy as np

# base.py:5-10
Mask = namedtuple('DataMask', 'row column')
Task = namedtuple('Task', 'mask parent is_left depth')
BestSplit = namedtuple('BestSplit', 'attribute threshold constant_attrs improvement')
Tree = namedtuple('DecisionTree', 'children_left children_right feature threshold value')


def is_leaf(mask, y, min_sample_per_leaf=10):
    """base.py:13-21: at most ``2 * min_sample_per_leaf`` rows, a single distinct target value,
    or no candidate column left."""
    return bool(np.sum(mask.row) <= min_sample_per_leaf * 2 or np.unique(y).size <= 1
                or np.sum(mask.column) == 0)


def is_leaf_v2(orders, min_sample_per_leaf=10):
    """base.py:24-29."""
    return orders.shape[0] <= min_sample_per_leaf * 2


class DecisionTree:
    """Tree container with the reference's layout and methods (base.py:32-69): five arrays of
    length ``max_node - 1``; children -1, feature -2, threshold -2.0 where unused.  The
    reference leaves ``value`` uninitialised; this build zero-fills it."""

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

    def add_binary(self, node_id, best_split):
        self.tree_.threshold[node_id] = best_split.threshold
        self.tree_.feature[node_id] = best_split.attribute

    def add_binary_v2(self, node_id, col, threshold):
        self.tree_.threshold[node_id] = threshold
        self.tree_.feature[node_id] = col

    def final(self):
        return self
