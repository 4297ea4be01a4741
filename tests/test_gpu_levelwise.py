"""GPU parity of the level-synchronous builder (ref decision_tree/four.py):
same split search as one.py, level-order node ids, `<=` leaf test, n_node_samples."""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree, small_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def four():
    from np_decision_tree_b200.decision_tree import four, utils

    class NS:
        pass
    ns = NS()
    ns.four, ns.utils = four, utils
    return ns


@pytest.mark.parametrize("name", ["u0", "u1", "u2", "u3"])
def test_levelwise_golden(four, golden_small, name):
    g = golden_small
    X, y, depth = g[f"lv_{name}_X"], g[f"lv_{name}_y"], int(g[f"lv_{name}_depth"])
    t = four.four.build_regression_tree(X, y, max_depth=depth)
    assert_same_tree(t, GoldenTree(g, f"lv_{name}_"), what=name)
    assert t.tree_.n_node_samples.dtype == np.int32
    assert np.array_equal(t.tree_.n_node_samples, g[f"lv_{name}_n_node_samples"])


def test_levelwise_knobs(four, golden_small):
    g = golden_small
    t = four.four.build_regression_tree(g["lv_knobs_X"], g["lv_knobs_y"], max_depth=5, min_improvement=200.0,
                                        min_sample_leaf=30)
    assert_same_tree(t, GoldenTree(g, "lv_knobs_"))
    assert np.array_equal(t.tree_.n_node_samples, g["lv_knobs_n_node_samples"])


@pytest.mark.parametrize("n,F,depth,seed,yr", [(50000, 9, 9, 1, None), (3000, 6, 11, 2, 10.0)])
def test_levelwise_vs_oracle(four, n, F, depth, seed, yr):
    X, y = small_problem(seed, n, F, yr, informative=4)
    t = four.four.build_regression_tree(X, y, max_depth=depth)
    o = oracle.fit_l2_levelwise(X, y, max_depth=depth, sort_kind="stable")
    assert_same_tree(t, o, what=f"{n}x{F}")
    assert np.array_equal(t.tree_.n_node_samples, o.tree_.n_node_samples)
    # the same rows reach the same leaves through the level-ordered ids
    assert np.array_equal(four.utils.inference_leaf(X, t), oracle.predict_leaf(X, o))


def test_levelwise_is_the_one_py_tree_renumbered(four):
    from np_decision_tree_b200.decision_tree import one
    X, y = small_problem(5, 20000, 7)
    a = one.build_regression_tree(X, y, max_depth=6)
    b = four.four.build_regression_tree(X, y, max_depth=6)
    assert a.max_node_ == b.max_node_
    assert sorted(a.tree_.threshold[: a.max_node_ + 1]) == sorted(b.tree_.threshold[: b.max_node_ + 1])
    np.testing.assert_allclose(four.utils.inference(X, a), four.utils.inference(X, b), rtol=1e-12)


@pytest.mark.parametrize("name", ["w0", "w1", "w2"])
def test_weighted_golden(four, golden_small, name):
    """sample weights: ref four.py:21-30, 116-117, 132-133"""
    g = golden_small
    X, y, w, depth = g[f"lv_{name}_X"], g[f"lv_{name}_y"], g[f"lv_{name}_w"], int(g[f"lv_{name}_depth"])
    t = four.four.build_regression_tree(X, y, max_depth=depth, weight=w)
    assert_same_tree(t, GoldenTree(g, f"lv_{name}_"), what=name)
    assert np.array_equal(t.tree_.n_node_samples, g[f"lv_{name}_n_node_samples"])


@pytest.mark.parametrize("n,F,depth,seed,yr", [(60000, 7, 8, 3, None), (2500, 5, 10, 4, 10.0), (2049, 3, 3, 5, None)])
def test_weighted_vs_oracle(four, n, F, depth, seed, yr):
    X, y = small_problem(seed, n, F, yr, informative=3)
    w = np.random.RandomState(seed + 7).uniform(0.1, 5.0, size=n)
    t = four.four.build_regression_tree(X, y, max_depth=depth, weight=w)
    o = oracle.fit_l2_levelwise(X, y, max_depth=depth, weight=w, sort_kind="stable")
    assert_same_tree(t, o, what=f"weighted {n}x{F}")
    assert np.array_equal(t.tree_.n_node_samples, o.tree_.n_node_samples)


def test_weighted_torch_inputs_and_unit_weights(four):
    import torch
    X, y = small_problem(8, 30000, 6)
    w = np.random.RandomState(1).uniform(0.5, 2.0, size=30000)
    t_host = four.four.build_regression_tree(X, y, max_depth=6, weight=w)
    t_dev = four.four.build_regression_tree(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), max_depth=6,
                                            weight=torch.from_numpy(w).cuda())
    assert_same_tree(t_dev, t_host, leaf_rtol=1e-12)
    # unit weights: same partition of the rows as the unweighted builder (formulas differ in rounding only)
    a = four.four.build_regression_tree(X, y, max_depth=5)
    b = four.four.build_regression_tree(X, y, max_depth=5, weight=np.ones(30000))
    assert np.array_equal(a.tree_.n_node_samples, b.tree_.n_node_samples)
