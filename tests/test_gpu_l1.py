"""GPU parity tests of the L1 (MAE) path through the C ABI: running median (cummedian), L1 split
search (K3), L1 fit -- against the reference-generated golden fixtures and the oracle.
Medians and split indices bit-exact; costs and leaf values within 1e-9 relative."""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree, small_problem

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def l1():
    from np_decision_tree_b200.decision_tree import strategy_l1, np_stream_median, utils

    class NS:
        pass
    ns = NS()
    ns.l1, ns.sm, ns.utils = strategy_l1, np_stream_median, utils
    return ns


@pytest.mark.parametrize("name", ["rm_a", "rm_b", "rm_c", "rm_d", "rm_e"])
def test_cummedian_golden(l1, golden_small, name):
    assert np.array_equal(l1.sm.cummedian(golden_small[name]), golden_small[name + "_out"])


@pytest.mark.parametrize("n,seed,rnd", [(2, 0, None), (3, 1, None), (33, 2, 1.0), (4097, 3, None),
                                        (100000, 4, 0.01), (1000003, 5, None)])
def test_cummedian_vs_oracle(l1, n, seed, rnd):
    a = np.random.RandomState(seed).standard_normal(n)
    if rnd:
        a = np.round(a / rnd) * rnd
    assert np.array_equal(l1.sm.cummedian(a), oracle.running_median_pairs(a))


def test_cummedian_torch(l1):
    import torch
    a = np.random.RandomState(9).standard_normal(5000)
    out = l1.sm.cummedian(torch.from_numpy(a).cuda())
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), oracle.running_median_pairs(a))


def test_l1_search_golden(l1, golden_small):
    g = golden_small
    X, y = g["l1fn_X"], g["l1fn_y"]
    orders = np.argsort(X, axis=0, kind="stable")
    k, f, c = l1.l1.best_l1_improvements_v2(y[orders])
    assert [k, f] == list(g["l1fn_best"]) and c == pytest.approx(float(g["l1fn_cost"]), rel=RTOL)
    thr, col, cost, left = l1.l1.greedy_regression_split_v2(orders, X, y)
    assert col == f and left.shape[0] == k + 1 and thr == X[orders[k, f], f]
    assert cost == pytest.approx(float(g["l1fn_cost"]), rel=RTOL)


@pytest.mark.parametrize("n,F,seed,yr", [(2, 1, 0, None), (3, 3, 1, None), (40, 5, 2, 10.0), (2049, 4, 3, None),
                                          (60000, 6, 4, None)])
def test_l1_split_node_vs_oracle(l1, n, F, seed, yr):
    X, y = small_problem(seed, n, F, yr)
    orders = np.argsort(X, axis=0, kind="stable")
    thr, col, cost, left = l1.l1.greedy_regression_split_v2(orders, X, y)
    othr, ocol, ocost, oleft = oracle.l1_split_node(orders, X, y)
    assert (col, thr) == (ocol, othr) and np.array_equal(left, oleft)
    assert cost == pytest.approx(float(ocost), rel=RTOL, abs=1e-12)


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_l1_fit_golden(l1, golden_small, name):
    g = golden_small
    X, y, depth = g[f"l1_{name}_X"], g[f"l1_{name}_y"], int(g[f"l1_{name}_depth"])
    for build in (l1.l1.build_regression_tree_v2, l1.l1.build_regression_tree):
        t = build(X, y, max_depth=depth)
        assert_same_tree(t, GoldenTree(g, f"l1_{name}_"), what=name)
    np.testing.assert_allclose(l1.utils.inference(X, t), g[f"l1_{name}_pred"], rtol=RTOL)


def test_l1_fit_config_depth1(l1, golden_configs, c2_data):
    """10 000 x 100 MAE depth 1 (README.md:107-131 / BASELINE.md section 2)."""
    g = golden_configs
    X, y = c2_data
    t = l1.l1.build_regression_tree_v2(X, y, max_depth=1)
    assert_same_tree(t, GoldenTree(g, "l1c_"), what="L1 depth 1")
    assert np.mean(np.abs(l1.utils.inference(X, t) - y)) == pytest.approx(float(g["l1c_mae"]), rel=RTOL)


@pytest.mark.parametrize("n,F,depth,seed,yr", [(5000, 8, 6, 1, None), (20000, 5, 4, 2, 5.0), (300, 3, 9, 3, 50.0),
                                                (100000, 20, 1, 4, None)])
def test_l1_fit_vs_oracle(l1, n, F, depth, seed, yr):
    X, y = small_problem(seed, n, F, yr, informative=3)
    t = l1.l1.build_regression_tree_v2(X, y, max_depth=depth)
    assert_same_tree(t, oracle.fit_l1(X, y, max_depth=depth, sort_kind="stable"), what=f"{n}x{F}")


@pytest.mark.parametrize("n,F,depth,seed,yr", [(150000, 6, 3, 21, None),    # several chunked nodes per level
                                                (80000, 4, 2, 22, 5.0),       # repeated targets: exact cost ties -> exact re-search
                                                (70000, 3, 2, 23, None)])
def test_l1_large_nodes_chunked_walks_and_costs(l1, n, F, depth, seed, yr):
    """Nodes >= 32768 rows: the median walks run as concurrent chunks started from selected order statistics and
    the cost sums are chunked (near ties fall back to the sequential sums): same tree as the oracle."""
    X, y = small_problem(seed, n, F, yr, informative=2)
    t = l1.l1.build_regression_tree_v2(X, y, max_depth=depth)
    assert_same_tree(t, oracle.fit_l1(X, y, max_depth=depth, sort_kind="stable"), what=f"{n}x{F} depth {depth}")


def test_l1_exact_env_matches(l1, monkeypatch):
    X, y = small_problem(24, 90000, 4, None, informative=2)
    a = l1.l1.build_regression_tree_v2(X, y, max_depth=2)
    monkeypatch.setenv("B200DT_L1_EXACT", "1")
    monkeypatch.setenv("B200DT_L1_CHUNKS", "1")
    b = l1.l1.build_regression_tree_v2(X, y, max_depth=2)
    assert_same_tree(a, b, what="chunked vs sequential L1")


def test_l1_leaf_rule_quirk(l1):
    """strategy_l1.py:158,188: a node is a leaf when its best children COST is < min_improvement,
    so every 2-row node is a leaf (cost 0) and other nodes split down to single rows."""
    X, y = small_problem(5, 64, 3)
    t = l1.l1.build_regression_tree_v2(X, y, max_depth=12)
    o = oracle.fit_l1(X, y, max_depth=12, sort_kind="stable")
    assert_same_tree(t, o)


@pytest.mark.parametrize("n,F,seed,yr", [(2, 3, 1, None), (257, 9, 2, 5.0), (5000, 4, 3, None), (70000, 3, 4, None)])
def test_best_l1_improvements_independent_columns(l1, n, F, seed, yr):
    """ys with ARBITRARY columns (not permutations of one another): every column is its own problem, searched in one
    batched call; the flat argmin (lowest cost, lowest row, lowest column) follows strategy_l1.py:119-121."""
    rs = np.random.RandomState(seed)
    ys = rs.standard_normal((n, F)) * rs.uniform(0.5, 20.0, size=(1, F))
    if yr:
        ys = np.round(ys * yr) / yr
    ys[:, F - 1] = ys[:, 0]                                     # two identical columns: the lower index must win ties
    k, f, c = l1.l1.best_l1_improvements_v2(ys)
    ok, of, oc = oracle.l1_best_split(ys)
    assert (k, f) == (ok, of)
    assert c == pytest.approx(float(oc), rel=RTOL, abs=1e-12)
