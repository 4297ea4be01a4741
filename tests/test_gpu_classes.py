"""Classification trees on the GPU (csrc/classes.cu through b200dt_fit_classes /
b200dt_class_split_node) against the reference-generated goldens and the numpy oracle.
Integer class statistics: every comparison is bit-exact."""
import numpy as np
import pytest

import oracle
from util import GoldenTree, assert_same_tree

pytestmark = pytest.mark.gpu

CASES = ["g0", "g1", "g2", "g3", "g4", "g5", "gk", "p0", "p1", "p2", "pk"]


@pytest.fixture(scope="module")
def mods():
    from np_decision_tree_b200.decision_tree import strategy, tree_builder, utils
    return strategy, tree_builder, utils


def method_of(strategy, criterion):
    return strategy.greedy_classification if criterion == "gini" else strategy.greedy_classification_p_at_k


def class_problem(seed, n, F, C, noise=0.5, integer_features=None):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, F))
    W = rs.standard_normal((F, C)) * (rs.uniform(size=(F, 1)) < 0.6)
    labels = np.argmax(X @ W + noise * rs.standard_normal((n, C)), axis=1).astype(np.int64)
    if integer_features:
        X = np.round(X * integer_features)          # heavy ties inside every column
    return X, labels


@pytest.mark.parametrize("name", CASES)
def test_tree_matches_reference_golden(golden_classes, mods, name):
    strategy, tree_builder, utils = mods
    g, p = golden_classes, f"cl_{name}_"
    C, F = int(g[p + "classes"]), g[p + "X"].shape[1]
    t = tree_builder.build_classification_tree(
        g[p + "X"], g[p + "labels"], max_depth=int(g[p + "depth"]), max_feature=F,
        min_improvement=float(g[p + "min_improvement"]), min_sample_leaf=int(g[p + "min_sample_leaf"]),
        split_method=method_of(strategy, str(g[p + "criterion"])), max_classes=C)
    assert_same_tree(t, GoldenTree(g, p), leaf_rtol=0.0, what=name)
    assert np.array_equal(utils.inference(g[p + "X"], t), g[p + "pred"])


@pytest.mark.parametrize("name", ["g0", "g2", "g5", "p1"])
def test_three_pass_search_matches_reference_golden(golden_classes, mods, name, monkeypatch):
    """B200DT_CLASS_3PASS=1: the histogram + scan + score kernels (the path > 32 classes and precision-at-1 with
    many classes take anyway) forced for small class counts: same reference trees."""
    monkeypatch.setenv("B200DT_CLASS_3PASS", "1")
    test_tree_matches_reference_golden(golden_classes, mods, name)


@pytest.mark.parametrize("name", CASES)
def test_root_search_matches_reference_golden(golden_classes, mods, name):
    strategy, _, _ = mods
    g, p = golden_classes, f"cl_{name}_"
    C = int(g[p + "classes"])
    col, thr, imp, const = method_of(strategy, str(g[p + "criterion"]))(g[p + "X"], np.eye(C)[g[p + "labels"]])
    want = g[p + "root"]
    assert (col, thr) == (int(want[0]), want[1])
    assert imp == want[2]
    assert const.dtype == bool and const.shape == (g[p + "X"].shape[1],) and not const.any()


@pytest.mark.parametrize("criterion", ["gini", "p_at_k"])
@pytest.mark.parametrize("seed,n,F,C,depth,ties", [
    (1, 5000, 16, 7, 10, None),        # several 1024-row tiles per node
    (2, 3000, 5, 64, 8, None),         # two 32-class chunks
    (3, 2500, 4, 200, 6, None),        # seven chunks
    (4, 4000, 6, 3, 9, 3.0),           # repeated feature values: value-based split extends the tie run
    (5, 1025, 3, 2, 12, None),         # one row past a tile boundary
    (6, 2048, 2, 4, 5, 1.0),
])
def test_tree_matches_oracle(mods, criterion, seed, n, F, C, depth, ties):
    strategy, tree_builder, utils = mods
    X, labels = class_problem(seed, n, F, C, integer_features=ties)
    kw = dict(max_depth=depth, max_feature=F, min_improvement=1e-7, min_sample_leaf=1)
    want = oracle.fit_classes(X, labels, C, criterion=criterion, sort_kind="stable", **kw)
    got = tree_builder.build_classification_tree(X, labels, split_method=method_of(strategy, criterion),
                                                 max_classes=C, **kw)
    assert_same_tree(got, want, leaf_rtol=0.0, what=f"{criterion} seed {seed}")
    assert np.array_equal(utils.inference(X, got), oracle.predict(X, want))


def test_larger_node_counts_match_oracle(mods):
    strategy, tree_builder, _ = mods
    X, labels = class_problem(11, 200_000, 8, 5)
    kw = dict(max_depth=6, max_feature=8, min_improvement=1e-7)
    want = oracle.fit_classes(X, labels, 5, **kw)
    got = tree_builder.build_classification_tree(X, labels, split_method=strategy.greedy_classification,
                                                 max_classes=5, **kw)
    assert_same_tree(got, want, leaf_rtol=0.0)


def test_pure_node_is_a_leaf_and_depth_zero(mods):
    strategy, tree_builder, _ = mods
    X = np.random.RandomState(0).standard_normal((50, 3))
    t = tree_builder.build_classification_tree(X, np.full(50, 2), split_method=strategy.greedy_classification,
                                               max_classes=4)
    assert t.max_node_ == 0 and t.tree_.value[0] == 2.0 and t.tree_.children_left[0] == -1
    labels = (X[:, 1] > 0).astype(np.int64)
    t = tree_builder.build_classification_tree(X, labels, max_depth=0, split_method=strategy.greedy_classification)
    assert t.max_node_ == 0 and t.tree_.value[0] == float(np.argmax(np.bincount(labels)))
    # min_improvement = 0 lets a pure node split at the first position of column 0, like the reference
    want = oracle.fit_classes(X, np.zeros(50, dtype=np.int64), 2, max_depth=2, min_improvement=0.0)
    got = tree_builder.build_classification_tree(X, np.zeros(50, dtype=np.int64), max_depth=2, min_improvement=0.0,
                                                 split_method=strategy.greedy_classification)
    assert_same_tree(got, want, leaf_rtol=0.0)


def test_all_rows_on_one_side(mods):
    # a constant column chosen at position 0 sends every row left; the right child is an empty leaf of class 0
    strategy, tree_builder, _ = mods
    X = np.ones((40, 1))
    labels = np.arange(40) % 2
    kw = dict(max_depth=2, max_feature=1, min_improvement=-1.0)
    want = oracle.fit_classes(X, labels, 2, sort_kind="stable", **kw)
    got = tree_builder.build_classification_tree(X, labels, split_method=strategy.greedy_classification, **kw)
    assert_same_tree(got, want, leaf_rtol=0.0)


def test_torch_cuda_inputs(mods):
    import torch
    strategy, tree_builder, _ = mods
    X, labels = class_problem(21, 3000, 6, 4)
    kw = dict(max_depth=6, max_feature=6, min_improvement=1e-7, split_method=strategy.greedy_classification,
              max_classes=4)
    host = tree_builder.build_classification_tree(X, labels, **kw)
    dev = tree_builder.build_classification_tree(torch.from_numpy(X).cuda(), torch.from_numpy(labels).cuda(), **kw)
    assert_same_tree(dev, host, leaf_rtol=0.0)


def test_errors(mods):
    strategy, tree_builder, _ = mods
    from np_decision_tree_b200._lib import B200Error
    X, labels = class_problem(31, 100, 4, 3)
    with pytest.raises(B200Error, match="n_classes"):
        tree_builder.build_classification_tree(X, labels, split_method=strategy.greedy_classification,
                                               max_classes=300)
    with pytest.raises(B200Error, match="single row"):
        tree_builder.build_classification_tree(X, labels, max_depth=12, min_sample_leaf=0, min_improvement=-1.0,
                                               split_method=strategy.greedy_classification, max_classes=3)


# ---- random-threshold splitters (csrc/rows.cu): seeded np.random must reproduce the reference ----------
RANDOM_CASES = ["r0", "r1", "r2", "r3", "r4", "r5"]


def random_method_of(strategy, criterion):
    return strategy.random_classify if criterion == "gini" else strategy.random_classify_p_at_k


@pytest.mark.parametrize("name", RANDOM_CASES)
def test_random_tree_matches_reference_golden(golden_classes, mods, name):
    strategy, tree_builder, utils = mods
    g, p = golden_classes, f"rc_{name}_"
    np.random.seed(int(g[p + "seed"]))
    t = tree_builder.build_classification_tree(
        g[p + "X"], g[p + "labels"], max_depth=int(g[p + "depth"]), max_feature=int(g[p + "max_feature"]),
        min_improvement=float(g[p + "min_improvement"]), min_sample_leaf=int(g[p + "min_sample_leaf"]),
        split_method=random_method_of(strategy, str(g[p + "criterion"])), max_classes=int(g[p + "classes"]))
    assert_same_tree(t, GoldenTree(g, p), leaf_rtol=0.0, what=name)
    assert np.array_equal(utils.inference(g[p + "X"], t), g[p + "pred"])


@pytest.mark.parametrize("name", RANDOM_CASES)
def test_random_root_search_matches_reference_golden(golden_classes, mods, name):
    strategy, _, _ = mods
    g, p = golden_classes, f"rc_{name}_"
    C = int(g[p + "classes"])
    np.random.seed(77)
    col, thr, imp, const = random_method_of(strategy, str(g[p + "criterion"]))(g[p + "X"], np.eye(C)[g[p + "labels"]])
    want = g[p + "root"]
    assert (col, thr) == (int(want[0]), want[1])
    assert imp == want[2] or (np.isnan(imp) and np.isnan(want[2]))
    assert const.dtype == bool and not const.any()


@pytest.mark.parametrize("criterion", ["gini", "p_at_k"])
@pytest.mark.parametrize("seed,n,F,C,depth,max_feature", [(51, 20000, 24, 6, 10, 24), (52, 5000, 300, 3, 7, 40),
                                                           (53, 3000, 5, 40, 9, 5)])
def test_random_tree_matches_oracle(mods, criterion, seed, n, F, C, depth, max_feature):
    strategy, tree_builder, utils = mods
    X, labels = class_problem(seed, n, F, C)
    kw = dict(max_depth=depth, max_feature=max_feature, min_improvement=1e-7, min_sample_leaf=1)
    np.random.seed(seed)
    want = oracle.fit_classes_random(X, labels, C, criterion=criterion, **kw)
    np.random.seed(seed)
    got = tree_builder.build_classification_tree(X, labels, split_method=random_method_of(strategy, criterion),
                                                 max_classes=C, **kw)
    assert_same_tree(got, want, leaf_rtol=0.0, what=f"random {criterion} seed {seed}")
    assert np.array_equal(utils.inference(X, got), oracle.predict(X, want))


def test_reference_default_split_method_runs(mods):
    """tree_builder.build_classification_tree's default splitter is random_classify (tree_builder.py:153)."""
    _, tree_builder, _ = mods
    X, labels = class_problem(54, 2000, 8, 2)
    np.random.seed(5)
    want = oracle.fit_classes_random(X, labels, 2)
    np.random.seed(5)
    got = tree_builder.build_classification_tree(X, labels)
    assert_same_tree(got, want, leaf_rtol=0.0)


def test_random_tree_torch_cuda_inputs(mods):
    import torch
    strategy, tree_builder, _ = mods
    X, labels = class_problem(55, 4000, 10, 4)
    kw = dict(max_depth=6, max_feature=6, min_improvement=1e-7, split_method=strategy.random_classify, max_classes=4)
    np.random.seed(9)
    host = tree_builder.build_classification_tree(X, labels, **kw)
    np.random.seed(9)
    dev = tree_builder.build_classification_tree(torch.from_numpy(X).cuda(), torch.from_numpy(labels).cuda(), **kw)
    assert_same_tree(dev, host, leaf_rtol=0.0)


def test_estimator_wrapper_cross_val():
    """run_classification.py:18-38,43-52: the Mree estimator under sklearn's cross_val_score."""
    from sklearn.model_selection import cross_val_score
    from np_decision_tree_b200.estimator import Mree
    from np_decision_tree_b200.decision_tree import strategy
    X, labels = class_problem(61, 3000, 6, 3, noise=0.2)

    def scorer(est, x, y):
        return float(np.mean(est.predict(x) == y))

    for method, mf in ((strategy.greedy_classification, 6), (strategy.random_classify, 4)):
        np.random.seed(3)
        scores = cross_val_score(Mree(max_depth=8, max_feature=mf, max_classes=3, split_method=method), X, labels,
                                 cv=3, scoring=scorer)
        assert scores.shape == (3,) and scores.min() > 0.5, (method.__name__, scores)


# ---- regression with random thresholds (random_split_v2 under build_tree; csrc/rows.cu) --------------------
@pytest.mark.parametrize("name", ["x0", "x1", "x2"])
def test_random_regression_tree_matches_reference_golden(golden_classes, mods, name):
    """Structure and thresholds exact under the same seed; leaf means to 1e-9 (floating-point sums: the
    reference's go through BLAS, ours through atomics)."""
    strategy, tree_builder, utils = mods
    g, p = golden_classes, f"rx_{name}_"
    np.random.seed(int(g[p + "seed"]))
    t = tree_builder.build_tree(g[p + "X"], g[p + "y"], max_depth=int(g[p + "depth"]),
                                max_feature=int(g[p + "max_feature"]), min_improvement=float(g[p + "min_improvement"]),
                                min_sample_leaf=int(g[p + "min_sample_leaf"]), split_method=strategy.random_split_v2)
    assert_same_tree(t, GoldenTree(g, p), leaf_rtol=1e-9, what=name)
    np.testing.assert_allclose(utils.inference(g[p + "X"], t), g[p + "pred"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed,n,F,depth,max_feature,msl", [(71, 30000, 16, 10, 16, 100), (72, 4000, 200, 8, 30, 50)])
def test_random_regression_tree_matches_oracle(mods, seed, n, F, depth, max_feature, msl):
    """Nodes stop at 2 * msl rows: in nodes of a handful of rows two columns can cut COMPLEMENTARY row sets, whose
    scores are mathematically equal; which one wins then hangs on the last bit of the reference's BLAS dot product
    and pairwise y.sum(), which no other summation order reproduces."""
    from util import small_problem
    strategy, tree_builder, utils = mods
    X, y = small_problem(seed, n, F, None, informative=4)
    kw = dict(max_depth=depth, max_feature=max_feature, min_improvement=1e-6, min_sample_leaf=msl)
    np.random.seed(seed)
    want = oracle.fit_l2_random(X, y, **kw)
    np.random.seed(seed)
    got = tree_builder.build_regression_tree(X, y, split_method=strategy.random_split_v2, **kw)
    assert_same_tree(got, want, leaf_rtol=1e-9, what=f"random regression seed {seed}")


def test_random_regression_deep_tree_quality(mods):
    """Down to single rows the tree may differ from the oracle's on tied candidates only: same size class, same fit."""
    from util import small_problem
    strategy, tree_builder, utils = mods
    X, y = small_problem(71, 30000, 16, None, informative=4)
    kw = dict(max_depth=10, max_feature=16, min_improvement=1e-6)
    np.random.seed(71)
    want = oracle.fit_l2_random(X, y, **kw)
    np.random.seed(71)
    got = tree_builder.build_regression_tree(X, y, split_method=strategy.random_split_v2, **kw)
    mse_w = np.mean((oracle.predict(X, want) - y) ** 2)
    mse_g = np.mean((utils.inference(X, got) - y) ** 2)
    assert abs(got.max_node_ + 1 - want.node_count) <= 0.05 * want.node_count
    assert mse_g == pytest.approx(mse_w, rel=0.05)
    top = slice(0, 16)   # the first nodes (large) are identical
    assert np.array_equal(got.tree_.feature[top], want.tree_.feature[top])
    assert np.array_equal(got.tree_.threshold[top], want.tree_.threshold[top])


def test_random_split_v2_single_node(mods):
    from util import small_problem
    strategy, _, _ = mods
    X, y = small_problem(73, 5000, 7)
    np.random.seed(11)
    wf, wthr, wimp = oracle.random_l2_split_node(X, y)
    np.random.seed(11)
    f, thr, imp, const = strategy.random_split_v2(X, y)
    assert (f, thr) == (wf, wthr) and imp == pytest.approx(wimp, rel=1e-9)
    assert const.shape == (7,) and not const.any()


def test_random_regression_constant_targets_is_a_leaf(mods):
    strategy, tree_builder, _ = mods
    X = np.random.RandomState(0).standard_normal((100, 3))
    t = tree_builder.build_tree(X, np.full(100, 2.5), max_depth=3, split_method=strategy.random_split_v2)
    assert t.max_node_ == 0 and t.tree_.value[0] == 2.5
