"""CPU: the oracle (numpy/C restatement) against fixtures produced by the reference itself.

This is what pins the oracle (the reference has no tests of its own).
"""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree, sha

L2_CASES = ["s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7"]


@pytest.mark.parametrize("name", L2_CASES)
def test_l2_tree_small(golden_small, name):
    g = golden_small
    X, y, depth = g[f"l2_{name}_X"], g[f"l2_{name}_y"], int(g[f"l2_{name}_depth"])
    t = oracle.fit_l2(X, y, max_depth=depth)
    assert_same_tree(t, GoldenTree(g, f"l2_{name}_"), leaf_rtol=0.0, what=name)
    assert np.array_equal(oracle.predict(X, t), g[f"l2_{name}_pred"])
    assert np.array_equal(oracle.predict_leaf(X, t), g[f"l2_{name}_leafid"])


@pytest.mark.parametrize("name", [c for c in L2_CASES if c not in ("s7",)])
def test_l2_tree_small_ratio_variant(golden_small, name):
    g = golden_small
    X, y, depth = g[f"l2_{name}_X"], g[f"l2_{name}_y"], int(g[f"l2_{name}_depth"])
    t = oracle.fit_l2(X, y, max_depth=depth, v=2)
    assert_same_tree(t, GoldenTree(g, f"l2v2_{name}_"), leaf_rtol=0.0, what=name)


def test_l2_knobs(golden_small):
    g = golden_small
    t = oracle.fit_l2(g["l2_knobs_X"], g["l2_knobs_y"], max_depth=6, min_sample_leaf=20,
                      min_improvement=50.0)
    assert_same_tree(t, GoldenTree(g, "l2_knobs_"), leaf_rtol=0.0)


def test_split_functions(golden_small):
    g = golden_small
    X, y, orders = g["fn_X"], g["fn_y"], g["fn_orders"]
    assert np.array_equal(oracle.argsort_columns(X), orders)
    assert np.array_equal(oracle.l2_weights(257), g["fn_np_gene"])
    cums = np.cumsum(y[orders], axis=0)
    k, f, gain = oracle.l2_best_split(cums)
    assert [k, f] == list(g["fn_best"]) and gain == g["fn_best_gain"]
    k, f, gain = oracle.l2_best_split_ratio(cums)
    assert [k, f] == list(g["fn_best_v2"]) and gain == g["fn_best_v2_gain"]
    thr, col, gain, left_rows = oracle.l2_split_node(orders, X, y, 1)
    assert thr == g["fn_split_thr"] and col == g["fn_split_col"] and gain == g["fn_split_gain"]
    assert np.array_equal(left_rows, g["fn_left_rows"])
    lo, ro = oracle.partition_orders(orders, left_rows, y.shape[0])
    assert np.array_equal(lo, g["fn_left_orders"]) and np.array_equal(ro, g["fn_right_orders"])


@pytest.mark.parametrize("name", ["rm_a", "rm_b", "rm_c", "rm_d", "rm_e"])
def test_running_median(golden_small, name):
    g = golden_small
    assert np.array_equal(oracle.running_median_pairs(g[name]), g[name + "_out"])
    assert np.array_equal(oracle.running_median_pairs_py(g[name]), g[name + "_out"])


def test_running_median_is_order_statistic():
    rs = np.random.RandomState(3)
    a = np.round(rs.standard_normal(300) * 4) / 2
    out = oracle.running_median_pairs(a)
    for m in range(1, a.size + 1):
        s = np.sort(a[:m])
        assert out[m - 1, 0] == s[(m - 1) // 2] and out[m - 1, 1] == s[m // 2]


def test_l1_search(golden_small):
    g = golden_small
    X, y = g["l1fn_X"], g["l1fn_y"]
    k, f, c = oracle.l1_best_split(y[np.argsort(X, axis=0)])
    assert [k, f] == list(g["l1fn_best"]) and c == g["l1fn_cost"]


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_l1_tree_small(golden_small, name):
    g = golden_small
    X, y, depth = g[f"l1_{name}_X"], g[f"l1_{name}_y"], int(g[f"l1_{name}_depth"])
    t = oracle.fit_l1(X, y, max_depth=depth)
    assert_same_tree(t, GoldenTree(g, f"l1_{name}_"), leaf_rtol=0.0, what=name)
    assert np.array_equal(oracle.predict(X, t), g[f"l1_{name}_pred"])


def test_config_c1_c2(golden_configs, c2_data):
    """BASELINE.json configs[0] and configs[1] (10 000 x 100, depth 1 and depth 10 + inference)."""
    g = golden_configs
    X, y = c2_data
    assert sha(oracle.argsort_columns(X).astype(np.int64)) == str(g["argsort_sha"])
    t1 = oracle.fit_l2(X, y, max_depth=1)
    assert_same_tree(t1, GoldenTree(g, "c1_"), leaf_rtol=0.0)
    assert sha(oracle.predict(X, t1)) == str(g["c1_pred_sha"])
    t10 = oracle.fit_l2(X, y, max_depth=10)
    assert_same_tree(t10, GoldenTree(g, "c2_"), leaf_rtol=0.0)
    assert t10.node_count == 1927          # SURVEY.md section 4 (= sklearn's node count)
    p = oracle.predict(X, t10)
    assert sha(p) == str(g["c2_pred_sha"])
    assert sha(oracle.predict_leaf(X, t10)) == str(g["c2_leafid_sha"])
    assert np.mean((p - y) ** 2) == g["c2_mse"]


def test_config_l1_depth1(golden_configs, c2_data):
    g = golden_configs
    X, y = c2_data
    t = oracle.fit_l1(X, y, max_depth=1)
    assert_same_tree(t, GoldenTree(g, "l1c_"), leaf_rtol=0.0)
    assert np.mean(np.abs(oracle.predict(X, t) - y)) == g["l1c_mae"]


def test_predict_root_only_shape_quirk(golden_small):
    """ref utils.py:20,26-31: a root-only tree yields shape (1,)."""
    g = golden_small
    t = oracle.fit_l2(g["l2_s7_X"], g["l2_s7_y"], max_depth=0)
    assert oracle.predict(g["l2_s7_X"], t).shape == (1,) == g["l2_s7_pred"].shape


LEVEL_CASES = ["u0", "u1", "u2", "u3", "w0", "w1", "w2"]


@pytest.mark.parametrize("name", LEVEL_CASES)
def test_levelwise_tree_small(golden_small, name):
    """ref four.py: level-synchronous builder, unweighted (u*) and weighted (w*)."""
    g = golden_small
    X, y, depth = g[f"lv_{name}_X"], g[f"lv_{name}_y"], int(g[f"lv_{name}_depth"])
    w = g[f"lv_{name}_w"] if f"lv_{name}_w" in g.files else None
    t = oracle.fit_l2_levelwise(X, y, max_depth=depth, weight=w)
    assert_same_tree(t, GoldenTree(g, f"lv_{name}_"), leaf_rtol=0.0, what=name)
    assert np.array_equal(t.tree_.n_node_samples, g[f"lv_{name}_n_node_samples"])
    assert t.tree_.n_node_samples.dtype == np.int32


def test_levelwise_knobs(golden_small):
    g = golden_small
    t = oracle.fit_l2_levelwise(g["lv_knobs_X"], g["lv_knobs_y"], max_depth=5, min_improvement=200.0,
                                min_sample_leaf=30)
    assert_same_tree(t, GoldenTree(g, "lv_knobs_"), leaf_rtol=0.0)
    assert np.array_equal(t.tree_.n_node_samples, g["lv_knobs_n_node_samples"])
