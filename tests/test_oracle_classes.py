"""The classification oracle (oracle.fit_classes / class_split_node) against trees grown by the
reference's own tree_builder.build_classification_tree (tests/golden/make_golden_classes.py)."""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree

CASES = ["g0", "g1", "g2", "g3", "g4", "g5", "gk", "p0", "p1", "p2", "pk"]


def case_args(g, name):
    p = f"cl_{name}_"
    return dict(X=g[p + "X"], labels=g[p + "labels"], n_classes=int(g[p + "classes"]),
                max_depth=int(g[p + "depth"]), min_improvement=float(g[p + "min_improvement"]),
                min_sample_leaf=int(g[p + "min_sample_leaf"]), criterion=str(g[p + "criterion"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_tree_matches_reference(golden_classes, name):
    a = case_args(golden_classes, name)
    t = oracle.fit_classes(**a)
    assert_same_tree(t, GoldenTree(golden_classes, f"cl_{name}_"), leaf_rtol=0.0, what=name)
    assert np.array_equal(oracle.predict(a["X"], t), golden_classes[f"cl_{name}_pred"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_root_search_matches_reference(golden_classes, name):
    a = case_args(golden_classes, name)
    f, thr, imp, _ = oracle.class_split_node(a["X"], np.eye(a["n_classes"])[a["labels"]], a["criterion"])
    want = golden_classes[f"cl_{name}_root"]
    assert f == int(want[0]) and thr == want[1]
    assert imp == want[2]          # same numpy expression: bit-identical


def test_stable_sort_changes_nothing_on_distinct_values(golden_classes):
    a = case_args(golden_classes, "g2")
    t = oracle.fit_classes(sort_kind="stable", **a)
    assert_same_tree(t, GoldenTree(golden_classes, "cl_g2_"), leaf_rtol=0.0)


def test_equal_feature_values_never_straddle_a_split():
    # tree_builder.py:31-34: rows go left by VALUE, so a run of equal values stays together
    rs = np.random.RandomState(5)
    X = rs.randint(0, 4, size=(200, 3)).astype(np.float64)
    labels = (X[:, 0] + rs.randint(0, 2, size=200) > 2).astype(np.int64)
    t = oracle.fit_classes(X, labels, 2, max_depth=4, min_improvement=1e-7, sort_kind="stable")
    leaf = oracle.predict_leaf(X, t)
    for node in np.unique(leaf):
        rows = X[leaf == node]
        assert rows.shape[0] > 0
    assert t.node_count >= 3


# ---- random-threshold splitters: the oracle must consume numpy's global generator like the reference ----
RANDOM_CASES = ["r0", "r1", "r2", "r3", "r4", "r5"]


def random_case_args(g, name):
    p = f"rc_{name}_"
    return dict(X=g[p + "X"], labels=g[p + "labels"], n_classes=int(g[p + "classes"]),
                max_depth=int(g[p + "depth"]), max_feature=int(g[p + "max_feature"]),
                min_improvement=float(g[p + "min_improvement"]),
                min_sample_leaf=int(g[p + "min_sample_leaf"]), criterion=str(g[p + "criterion"]))


@pytest.mark.parametrize("name", RANDOM_CASES)
def test_oracle_random_tree_matches_reference(golden_classes, name):
    a = random_case_args(golden_classes, name)
    np.random.seed(int(golden_classes[f"rc_{name}_seed"]))
    t = oracle.fit_classes_random(**a)
    assert_same_tree(t, GoldenTree(golden_classes, f"rc_{name}_"), leaf_rtol=0.0, what=name)
    assert np.array_equal(oracle.predict(a["X"], t), golden_classes[f"rc_{name}_pred"])


@pytest.mark.parametrize("name", RANDOM_CASES)
def test_oracle_random_root_search_matches_reference(golden_classes, name):
    a = random_case_args(golden_classes, name)
    np.random.seed(77)
    f, thr, imp = oracle.random_class_split_node(a["X"], np.eye(a["n_classes"])[a["labels"]], a["criterion"])
    want = golden_classes[f"rc_{name}_root"]
    assert f == int(want[0]) and thr == want[1]
    assert imp == want[2] or (np.isnan(imp) and np.isnan(want[2]))


# ---- regression with random thresholds (random_split_v2 under build_tree) ------------------------------------
@pytest.mark.parametrize("name", ["x0", "x1", "x2"])
def test_oracle_random_regression_tree_matches_reference(golden_classes, name):
    g, p = golden_classes, f"rx_{name}_"
    np.random.seed(int(g[p + "seed"]))
    t = oracle.fit_l2_random(g[p + "X"], g[p + "y"], max_depth=int(g[p + "depth"]),
                             max_feature=int(g[p + "max_feature"]), min_improvement=float(g[p + "min_improvement"]),
                             min_sample_leaf=int(g[p + "min_sample_leaf"]))
    # structure / thresholds exact (thresholds are np.random draws); leaf means: same numpy expression here
    assert_same_tree(t, GoldenTree(g, p), leaf_rtol=1e-12, what=name)
    np.testing.assert_allclose(oracle.predict(g[p + "X"], t), g[p + "pred"], rtol=1e-12)
