"""sklearn-style wrapper of the classification trees, after the reference's ``Mree``
(ref run_classification.py:18-38): ``fit(x, y)`` / ``predict(x)`` / ``get_params()`` around
``tree_builder.build_classification_tree`` and ``utils.inference``.

The reference's class is a bare object; current scikit-learn needs ``__sklearn_tags__`` for
``cross_val_score``, so this one derives from ``BaseEstimator`` / ``ClassifierMixin`` (sklearn is only
imported here).  ``.params`` keeps the reference's dict view.
"""
from sklearn.base import BaseEstimator, ClassifierMixin

from .decision_tree.strategy import greedy_classification
from .decision_tree.tree_builder import build_classification_tree
from .decision_tree.utils import inference


class Mree(ClassifierMixin, BaseEstimator):
    def __init__(self, max_depth=10, max_feature=20, min_improvement=1e-7, max_classes=23,
                 split_method=greedy_classification):
        self.max_depth = max_depth
        self.max_feature = max_feature
        self.min_improvement = min_improvement
        self.max_classes = max_classes
        self.split_method = split_method

    @property
    def params(self):
        return self.get_params(deep=False)

    def fit(self, x, y):
        self.t = build_classification_tree(x, y, **self.params)
        return self

    def predict(self, x):
        return inference(x, self.t)
