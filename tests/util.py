"""Shared helpers for the parity tests."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_problem(seed, n, F, y_round=None, informative=3):
    """Same generator as tests/golden/make_golden.py (RandomState: frozen stream)."""
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, F))
    coef = np.zeros(F)
    informative = min(informative, F)
    coef[rs.permutation(F)[:informative]] = 100.0 * rs.uniform(size=informative)
    y = X @ coef + 0.1 * rs.standard_normal(n)
    if y_round is not None:
        y = np.round(y / y_round) * y_round
    return X, y


def assert_same_tree(got, want, leaf_rtol=1e-9, what=""):
    """Structure, features and thresholds bit-exact; leaf values to 1e-9 relative
    (+1e-9*max|value| absolute for leaves near zero).  ``want`` may be a tree
    object or a dict-like golden prefix view with the same field names."""
    g = got.tree_
    w = want.tree_ if hasattr(want, "tree_") else want
    wmax = want.max_node_ if hasattr(want, "max_node_") else int(w.max_node)
    assert got.max_node_ == wmax, f"{what}: node count {got.max_node_ + 1} != {wmax + 1}"
    for name in ("children_left", "children_right", "feature"):
        a, b = getattr(g, name), getattr(w, name)
        assert a.dtype == np.int64 and a.shape == b.shape, f"{what}: {name} dtype/shape"
        assert np.array_equal(a, b), f"{what}: {name} differs at {np.flatnonzero(a != b)[:8]}"
    assert g.threshold.dtype == np.float64
    assert np.array_equal(g.threshold, w.threshold), \
        f"{what}: threshold differs at {np.flatnonzero(g.threshold != w.threshold)[:8]}"
    n = wmax + 1
    leaf = np.zeros(g.feature.shape[0], dtype=bool)
    leaf[:n] = g.children_left[:n] == -1
    wv = np.asarray(w.value)
    scale = np.max(np.abs(wv[leaf])) if leaf.any() else 0.0
    np.testing.assert_allclose(g.value[leaf], wv[leaf], rtol=leaf_rtol, atol=leaf_rtol * scale,
                               err_msg=f"{what}: leaf values")


class GoldenTree:
    """View of one tree stored in a golden npz under ``prefix``."""

    def __init__(self, npz, prefix):
        for name in ("children_left", "children_right", "feature", "threshold", "value", "leaf"):
            setattr(self, name, npz[prefix + name])
        self.max_node = int(npz[prefix + "max_node"])
        self.tree_ = self
        self.max_node_ = self.max_node
