"""GPU parity tests of the L2 path through the C ABI: sort (K1), split search (K2),
partition (K4), whole fit, inference (K5) -- against the oracle and the reference-generated
golden fixtures.  Integer / index / threshold results must be bit-exact; gains and leaf values
within 1e-9 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree, sha, small_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def dt():
    from np_decision_tree_b200.decision_tree import one, utils
    from np_decision_tree_b200 import _lib

    class NS:
        pass
    ns = NS()
    ns.one, ns.utils, ns.lib = one, utils, _lib
    ns.ctx = _lib.default_context()
    return ns


def gpu_argsort(dt, X):
    X = np.ascontiguousarray(X, dtype=np.float64)
    out = np.empty(X.shape, dtype=np.int64)
    ctx = dt.ctx
    ctx.check(ctx.lib.b200dt_argsort_columns(ctx.handle, X.ctypes.data, X.shape[0], X.shape[1], X.shape[1],
                                             0, out.ctypes.data, 0))
    return out


# ---- K1 ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,F", [(1, 1), (2, 3), (31, 5), (4096, 2), (4097, 9), (10000, 100), (300001, 3)])
def test_argsort_matches_numpy_stable(dt, n, F):
    rs = np.random.RandomState(n + F)
    X = rs.standard_normal((n, F)) * np.exp(rs.uniform(-20, 20, size=(1, F)))
    assert np.array_equal(gpu_argsort(dt, X), np.argsort(X, axis=0, kind="stable"))


def test_argsort_special_values(dt):
    rs = np.random.RandomState(5)
    X = np.round(rs.standard_normal((5000, 4)) * 3) / 2          # heavy duplicates, +-0.0
    X[::7, 1] = -0.0
    X[::11, 2] = np.inf
    X[::13, 2] = -np.inf
    X[::17, 3] = np.nan
    X[5, 0] = 5e-324
    X[6, 0] = -5e-324
    X[7, 0] = np.finfo(np.float64).max
    assert np.array_equal(gpu_argsort(dt, X), np.argsort(X, axis=0, kind="stable"))


def test_argsort_sampled_range_outliers(dt):
    """>= 2^20 rows: the 32-bit image is quantised between a SAMPLED min / max (1 row group in 32).  Values the
    sample never saw clamp to the end slots: a few are repaired by the fix-up, many send the column to the
    64-bit path -- the result is the exact stable order either way."""
    n = (1 << 20) + 4097
    rs = np.random.RandomState(77)
    X = rs.standard_normal((n, 4))
    hidden = np.arange(n)[(np.arange(n) // 32) % 32 != 0]            # rows outside the sampled groups
    X[hidden[rs.permutation(hidden.size)[:20]], 0] = rs.uniform(50, 60, size=20)       # 20 high outliers
    X[hidden[rs.permutation(hidden.size)[:7]], 0] = -rs.uniform(50, 60, size=7)        # 7 low ones
    X[hidden[rs.permutation(hidden.size)[:500]], 1] = rs.uniform(1e3, 1e6, size=500)   # too many to repair
    X[hidden[:40], 2] = 1e300                                                          # 40 equal outliers
    X[:, 3] = np.where((np.arange(n) // 32) % 32 == 0, 1.5, rs.standard_normal(n))     # constant SAMPLE
    assert np.array_equal(gpu_argsort(dt, X), np.argsort(X, axis=0, kind="stable"))


def test_argsort_forced_64bit_path(dt, monkeypatch):
    monkeypatch.setenv("B200DT_SORT64", "1")      # read at every call
    rs = np.random.RandomState(78)
    X = np.round(rs.standard_normal((70001, 5)) * 4) / 8
    X[::9, 2] = -0.0
    X[::10, 3] = np.nan
    assert np.array_equal(gpu_argsort(dt, X), np.argsort(X, axis=0, kind="stable"))


def test_argsort_golden(dt, golden_configs, c2_data):
    X, _ = c2_data      # distinct values per column: equals the reference's default-kind argsort
    assert sha(gpu_argsort(dt, X)) == str(golden_configs["argsort_sha"])


# ---- K2 / K4 on one node ---------------------------------------------------------------------
def test_split_node_golden(dt, golden_small):
    g = golden_small
    X, y, orders = g["fn_X"], g["fn_y"], g["fn_orders"]
    thr, col, gain, left_rows = dt.one.greedy_regression_split(orders, X, y, 1)
    assert col == g["fn_split_col"] and thr == g["fn_split_thr"]
    assert np.array_equal(left_rows, g["fn_left_rows"])
    assert gain == pytest.approx(float(g["fn_split_gain"]), rel=RTOL)
    thr2, col2, gain2, left2 = dt.one.greedy_regression_split(orders, X, y, 2)
    k2, f2 = g["fn_best_v2"]
    assert col2 == f2 and left2.shape[0] == k2 + 1
    assert gain2 == pytest.approx(float(g["fn_best_v2_gain"]), rel=RTOL)


@pytest.mark.parametrize("n,F,seed", [(2, 1, 0), (3, 4, 1), (33, 7, 2), (2048, 5, 3), (2049, 5, 4),
                                      (50000, 13, 5), (200000, 4, 6)])
@pytest.mark.parametrize("v", [1, 2])
def test_split_node_vs_oracle(dt, n, F, seed, v):
    X, y = small_problem(seed, n, F)
    orders = np.argsort(X, axis=0, kind="stable")
    thr, col, gain, left_rows = dt.one.greedy_regression_split(orders, X, y, v)
    othr, ocol, ogain, oleft = oracle.l2_split_node(orders, X, y, v)
    assert (col, thr) == (ocol, othr)
    assert np.array_equal(left_rows, oleft)
    assert gain == pytest.approx(float(ogain), rel=RTOL)


def test_split_node_subset_of_rows(dt):
    """A node deep in the tree: orders index a subset of the N rows."""
    X, y = small_problem(9, 5000, 6)
    keep = np.flatnonzero(X[:, 2] > 0.3)
    orders = np.argsort(X, axis=0, kind="stable")
    node = np.stack([o[np.isin(o, keep)] for o in orders.T], axis=1)
    got = dt.one.greedy_regression_split(node, X, y, 1)
    want = oracle.l2_split_node(node, X, y, 1)
    assert got[0] == want[0] and got[1] == want[1] and np.array_equal(got[3], want[3])


def test_split_orders_golden(dt, golden_small):
    g = golden_small
    lo, ro = dt.one.split_orders(g["fn_orders"], g["fn_left_rows"], g["fn_y"].shape[0])
    assert np.array_equal(lo, g["fn_left_orders"]) and np.array_equal(ro, g["fn_right_orders"])


@pytest.mark.parametrize("n,F,frac", [(2, 2, 0.5), (100, 3, 0.01), (2048, 4, 0.5), (2049, 4, 0.99),
                                      (70000, 6, 0.37)])
def test_split_orders_vs_oracle(dt, n, F, frac):
    rs = np.random.RandomState(n)
    X = rs.standard_normal((n, F))
    orders = np.argsort(X, axis=0, kind="stable")
    n_left = min(max(1, int(n * frac)), n - 1)
    left_rows = orders[:n_left, F - 1].copy()
    lo, ro = dt.one.split_orders(orders, left_rows, n)
    olo, oro = oracle.partition_orders(orders, left_rows, n)
    assert np.array_equal(lo, olo) and np.array_equal(ro, oro)


# ---- whole fit ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7"])
def test_fit_golden_small(dt, golden_small, name):
    g = golden_small
    X, y, depth = g[f"l2_{name}_X"], g[f"l2_{name}_y"], int(g[f"l2_{name}_depth"])
    t = dt.one.build_regression_tree(X, y, max_depth=depth)
    assert_same_tree(t, GoldenTree(g, f"l2_{name}_"), what=name)
    assert np.allclose(dt.utils.inference(X, t), g[f"l2_{name}_pred"], rtol=RTOL, atol=1e-9 * np.abs(y).max())
    assert np.array_equal(dt.utils.inference_leaf(X, t), g[f"l2_{name}_leafid"])
    assert dt.utils.inference(X, t).shape == g[f"l2_{name}_pred"].shape   # (1,) quirk for depth 0


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3", "s4", "s5", "s6"])
def test_fit_golden_small_ratio_variant(dt, golden_small, name):
    g = golden_small
    X, y, depth = g[f"l2_{name}_X"], g[f"l2_{name}_y"], int(g[f"l2_{name}_depth"])
    t = dt.one.build_regression_tree(X, y, max_depth=depth, v=2)
    assert_same_tree(t, GoldenTree(g, f"l2v2_{name}_"), what=name)


def test_fit_knobs(dt, golden_small):
    g = golden_small
    t = dt.one.build_regression_tree(g["l2_knobs_X"], g["l2_knobs_y"], max_depth=6, min_sample_leaf=20,
                                     min_improvement=50.0)
    assert_same_tree(t, GoldenTree(g, "l2_knobs_"))


def test_fit_config_c1_c2(dt, golden_configs, c2_data):
    """BASELINE.json configs[0], configs[1]: 10 000 x 100, depth 1 and depth 10 + inference."""
    g = golden_configs
    X, y = c2_data
    t1 = dt.one.build_regression_tree(X, y, max_depth=1)
    assert_same_tree(t1, GoldenTree(g, "c1_"), what="C1")
    t10 = dt.one.build_regression_tree(X, y, max_depth=10)
    assert_same_tree(t10, GoldenTree(g, "c2_"), what="C2")
    assert t10.max_node_ + 1 == 1927
    assert sha(dt.utils.inference_leaf(X, t10)) == str(g["c2_leafid_sha"])
    p = dt.utils.inference(X, t10)
    np.testing.assert_allclose(p[:64], g["c2_pred_head"], rtol=RTOL)
    assert np.mean((p - y) ** 2) == pytest.approx(float(g["c2_mse"]), rel=1e-9)


@pytest.mark.parametrize("n,F,depth,seed,yr", [(200000, 16, 8, 1, None), (30000, 8, 12, 2, 5.0),
                                                (4097, 3, 12, 3, None), (100000, 5, 6, 4, 100.0)])
def test_fit_vs_oracle(dt, n, F, depth, seed, yr):
    X, y = small_problem(seed, n, F, yr, informative=4)
    t = dt.one.build_regression_tree(X, y, max_depth=depth)
    assert_same_tree(t, oracle.fit_l2(X, y, max_depth=depth, sort_kind="stable"), what=f"{n}x{F}")


def test_fit_duplicate_columns_structural_ties(dt):
    """Identical columns give bit-exact gain ties at every node: lowest feature must win,
    including at the root (large node -> near-tie detection -> sequential re-search)."""
    X, y = small_problem(7, 20000, 4)
    X = np.concatenate([X, X[:, [2, 0]], -X[:, [1]]], axis=1)
    t = dt.one.build_regression_tree(X, y, max_depth=7)
    assert_same_tree(t, oracle.fit_l2(X, y, max_depth=7, sort_kind="stable"))


def test_fit_constant_target(dt):
    X, _ = small_problem(8, 3000, 3)
    y = np.full(3000, 2.5)
    t = dt.one.build_regression_tree(X, y, max_depth=3)
    assert_same_tree(t, oracle.fit_l2(X, y, max_depth=3, sort_kind="stable"))


def test_fit_torch_cuda_inputs(dt):
    import torch
    X, y = small_problem(12, 50000, 10)
    t_host = dt.one.build_regression_tree(X, y, max_depth=6)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    t_dev = dt.one.build_regression_tree(Xd, yd, max_depth=6)
    assert_same_tree(t_dev, t_host, leaf_rtol=1e-12)
    p = dt.utils.inference(Xd, t_dev)     # device in, device out; same tree => same bits as the host path
    assert p.is_cuda and np.array_equal(p.cpu().numpy(), dt.utils.inference(X, t_dev))
    np.testing.assert_allclose(p.cpu().numpy(), dt.utils.inference(X, t_host), rtol=1e-12)


# ---- K5 ------------------------------------------------------------------------------------------
def test_inference_vs_oracle_and_sklearn(dt):
    from sklearn.tree import DecisionTreeRegressor
    X, y = small_problem(13, 20000, 9)
    t = oracle.fit_l2(X, y, max_depth=9, sort_kind="stable")
    t.tree_.value[np.isnan(t.tree_.value)] = 0.0
    Xq = np.random.RandomState(1).standard_normal((33333, 9))
    assert np.array_equal(dt.utils.inference(Xq, t), oracle.predict(Xq, t))
    assert np.array_equal(dt.utils.inference_leaf(Xq, t), oracle.predict_leaf(Xq, t))
    sk = DecisionTreeRegressor(max_depth=7, random_state=0).fit(X, y)     # ref README.md:47
    assert np.array_equal(dt.utils.inference(Xq, sk), sk.predict(Xq))


def test_no_cpu_fallback_symbols():
    """The product package must not import the oracle."""
    import sys
    import np_decision_tree_b200.decision_tree.one  # noqa: F401
    mods = [m for m in sys.modules if m.startswith("np_decision_tree_b200")]
    for m in mods:
        src = getattr(sys.modules[m], "__file__", "") or ""
        if src.endswith(".py"):
            assert "oracle" not in open(src).read().replace("no CPU fallback", ""), m


# ---- optional two-levels-per-data-movement mode (csrc/children.cu, B200DT_TWO_LEVEL=1) -------------------
@pytest.fixture
def two_level_env(monkeypatch):
    monkeypatch.setenv("B200DT_TWO_LEVEL", "1")   # read by the library at every fit


@pytest.mark.parametrize("seed,n,F,depth,yr,v", [(11, 60000, 6, 8, None, 1), (12, 30000, 5, 9, 50.0, 1),
                                                 (13, 40000, 4, 6, None, 2), (14, 5000, 3, 7, None, 1)])
def test_two_level_mode_builds_the_same_tree(two_level_env, seed, n, F, depth, yr, v):
    """Children searched from the parent's order + one 4-way partition per level pair: same tree as the oracle
    (large nodes take the pair path, small ones fall back to single levels, near ties materialise first)."""
    from np_decision_tree_b200.decision_tree import one
    X, y = small_problem(seed, n, F, yr)
    want = oracle.fit_l2(X, y, max_depth=depth, v=v, sort_kind="stable")
    got = one.build_regression_tree(X, y, max_depth=depth, v=v)
    assert_same_tree(got, want, what=f"two-level seed {seed}")


def test_two_level_mode_duplicate_columns(two_level_env):
    """Structural ties (identical columns) force the exact re-search on the second level of a pair."""
    from np_decision_tree_b200.decision_tree import one
    X, y = small_problem(15, 50000, 3)
    X = np.concatenate([X, X], axis=1)
    want = oracle.fit_l2(X, y, max_depth=6, sort_kind="stable")
    assert_same_tree(one.build_regression_tree(X, y, max_depth=6), want)


# ---- row relabelling (csrc/relabel.cu): node-contiguous row ids, handed out by one partition ----------
@pytest.mark.parametrize("window", [1, 700, 20000])
@pytest.mark.parametrize("seed,n,F,depth,yr,v", [(21, 60000, 6, 9, None, 1), (22, 30000, 5, 10, 50.0, 1),
                                                 (23, 40000, 4, 7, None, 2), (24, 5000, 3, 8, None, 1)])
def test_relabelled_rows_build_the_same_tree(monkeypatch, window, seed, n, F, depth, yr, v):
    """B200DT_RELABEL_ROWS makes these small problems relabel their rows (at the level whose children all
    have <= `window` rows; window 1 never does, 20000 does at the top, 700 in the middle of the tree):
    thresholds are fetched through the row map and the side masks use the new ids -- same tree."""
    from np_decision_tree_b200.decision_tree import one
    monkeypatch.setenv("B200DT_RELABEL_ROWS", str(window))
    X, y = small_problem(seed, n, F, yr)
    want = oracle.fit_l2(X, y, max_depth=depth, v=v, sort_kind="stable")
    got = one.build_regression_tree(X, y, max_depth=depth, v=v)
    assert_same_tree(got, want, what=f"relabel window {window} seed {seed}")


def test_relabelled_rows_duplicate_columns_and_weights(monkeypatch):
    """Structural ties after the relabelling (exact re-search on relabelled rows), and the weighted builder."""
    from np_decision_tree_b200.decision_tree import one, four
    monkeypatch.setenv("B200DT_RELABEL_ROWS", "9000")
    X, y = small_problem(25, 50000, 3)
    X2 = np.concatenate([X, X], axis=1)
    assert_same_tree(one.build_regression_tree(X2, y, max_depth=8), oracle.fit_l2(X2, y, max_depth=8, sort_kind="stable"))
    w = np.random.RandomState(26).uniform(0.25, 4.0, size=y.shape[0])
    got = four.build_regression_tree(X, y, max_depth=7, weight=w)
    monkeypatch.setenv("B200DT_RELABEL_ROWS", "0")
    want = four.build_regression_tree(X, y, max_depth=7, weight=w)
    for name in ("children_left", "children_right", "feature", "threshold", "n_node_samples"):
        assert np.array_equal(getattr(got.tree_, name), getattr(want.tree_, name)), name
    # node sums come out of a look-back scan whose grouping depends on timing: last bits may differ
    np.testing.assert_allclose(got.tree_.value, want.tree_.value, rtol=RTOL)


# ---- the TMA (cp.async.bulk + mbarrier) double-buffered split search, B200DT_SCAN_BULK=1 -------------------------
@pytest.mark.parametrize("seed,n,F,depth,yr,v", [(31, 60001, 6, 9, None, 1), (32, 30000, 5, 10, 50.0, 1),
                                                 (33, 40000, 4, 7, None, 2), (34, 5003, 3, 8, None, 1)])
def test_bulk_copy_scan_builds_the_same_tree(monkeypatch, seed, n, F, depth, yr, v):
    """Tiles copied by per-lane bulk asynchronous copies into padded rows, odd node starts shifted to 16-byte
    alignment, partial first / last tiles: same tree as the oracle."""
    from np_decision_tree_b200.decision_tree import one
    monkeypatch.setenv("B200DT_SCAN_BULK", "1")
    X, y = small_problem(seed, n, F, yr)
    want = oracle.fit_l2(X, y, max_depth=depth, v=v, sort_kind="stable")
    assert_same_tree(one.build_regression_tree(X, y, max_depth=depth, v=v), want, what=f"bulk scan seed {seed}")


def test_wide_sort_batches_use_the_big_radix_tile(monkeypatch):
    """B200DT_SORT_BATCH >= 128 sorts 128+ columns per batch with the 512-thread / 8192-key radix shape (the default
    batch of 32 columns uses 4096-key tiles): partial last tile, more than one tile per column, same tree."""
    from np_decision_tree_b200.decision_tree import one
    monkeypatch.setenv("B200DT_SORT_BATCH", "256")
    X, y = small_problem(77, 20011, 132, None)
    want = oracle.fit_l2(X, y, max_depth=3, v=1, sort_kind="stable")
    assert_same_tree(one.build_regression_tree(X, y, max_depth=3, v=1), want, what="wide sort batch")


def test_pageable_strided_host_input_takes_the_bounce_pipeline(dt):
    """>= 64 MB of ordinary (pageable) host memory with a row stride larger than F: packed into pinned slots by host
    threads, column batch by column batch (csrc/fit.cu: bounce_start) -- same tree as the contiguous / device paths."""
    import ctypes
    import torch
    from np_decision_tree_b200.decision_tree.one import DecisionTree
    rs = np.random.RandomState(91)
    wide = rs.standard_normal((300000, 48))
    F, ldx, depth = 40, 48, 6
    Xv = wide[:, 5:5 + F]                       # a view: rows 48 doubles apart
    y = Xv[:, 3] * 20.0 - Xv[:, 17] * 7.0 + 0.5 * rs.standard_normal(300000)
    want = dt.one.build_regression_tree(torch.from_numpy(np.ascontiguousarray(Xv)).cuda(), torch.from_numpy(y).cuda(),
                                        max_depth=depth)
    tree = DecisionTree(2 ** (depth + 1))
    t = tree.tree_
    count = ctypes.c_int64()
    ctx = dt.ctx
    ctx.check(ctx.lib.b200dt_fit(ctx.handle, wide.ctypes.data + 5 * 8, y.ctypes.data, 300000, F, ldx, 0, 0, depth, 1e-7, 1, 1,
                                 t.children_left.ctypes.data, t.children_right.ctypes.data, t.feature.ctypes.data,
                                 t.threshold.ctypes.data, t.value.ctypes.data, t.feature.shape[0], ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    assert_same_tree(tree, want, leaf_rtol=1e-12, what="strided pageable input")
