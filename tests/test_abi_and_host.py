"""CPU: the C-ABI library loads and exports every symbol include/b200dt.h declares; the host-side
mirror keeps the reference's names, defaults and tree layout; no compute is attempted without a GPU."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from np_decision_tree_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200dt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200dt.h but not exported"


def test_python_binding_covers_header():
    from np_decision_tree_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_candidate_struct_layout():
    from np_decision_tree_b200 import _lib
    assert ctypes.sizeof(_lib.Candidate) == 56
    assert _lib.Candidate.k.offset == 40 and _lib.Candidate.f.offset == 44 and _lib.Candidate.exact.offset == 48


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from np_decision_tree_b200 import _lib
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        _lib.Context(0)
    from np_decision_tree_b200.decision_tree import one
    with pytest.raises(_lib.B200Error):
        one.build_regression_tree(np.zeros((4, 2)), np.zeros(4), max_depth=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "np_decision_tree_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "np_tree_oracle" not in src, fn


def test_mirror_signatures_match_the_reference():
    """Names, positional order and defaults of ref one.py:135,165; strategy_l1.py:143,173; utils.py:15,34."""
    from np_decision_tree_b200.decision_tree import one, strategy_l1, utils, np_stream_median

    def sig(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    want_l2 = [("X", inspect._empty), ("y", inspect._empty), ("max_depth", 2), ("max_feature", 100),
               ("min_improvement", 1e-7), ("min_sample_leaf", 1), ("v", 1)]
    assert sig(one.build_regression_tree) == want_l2
    assert sig(one.build_regression_tree_v1_5) == want_l2
    want_l1 = want_l2[:-1] + [("leaf_from_data", np.median)]
    assert sig(strategy_l1.build_regression_tree) == want_l1
    assert sig(strategy_l1.build_regression_tree_v2) == want_l1
    assert [n for n, _ in sig(utils.inference)] == ["data", "clf"]
    assert [n for n, _ in sig(utils.inference_leaf)] == ["data", "clf"]
    assert [n for n, _ in sig(np_stream_median.cummedian)] == ["arr"]
    assert [n for n, _ in sig(one.greedy_regression_split)] == ["orders", "x", "y", "v"]
    assert [n for n, _ in sig(one.split_orders)] == ["orders", "index", "n_samples"]


def test_tree_container_layout():
    """ref one.py:9-24: five arrays of length max_node-1, sentinels -1/-1/-2/-2.0, max_node_ = -1."""
    from np_decision_tree_b200.decision_tree.one import DecisionTree, Task
    t = DecisionTree(2 ** (10 + 1))
    tr = t.tree_
    assert tr._fields == ("children_left", "children_right", "feature", "threshold", "value")
    assert all(a.shape == (2047,) for a in tr)
    assert [a.dtype for a in tr] == [np.int64, np.int64, np.int64, np.float64, np.float64]
    assert (tr.children_left == -1).all() and (tr.children_right == -1).all()
    assert (tr.feature == -2).all() and (tr.threshold == -2.0).all() and t.max_node_ == -1
    # the reference's builder methods (one.py:26-38)
    root = t.new_node_from_task(Task(None, None, True, 0))
    left = t.new_node_from_task(Task(None, root, True, 1))
    right = t.new_node_from_task(Task(None, root, False, 1))
    t.add_binary(root, 3, 0.5)
    t.add_leaf(left, 1.25)
    assert (root, left, right) == (0, 1, 2) and tr.children_left[0] == 1 and tr.children_right[0] == 2
    assert tr.feature[0] == 3 and tr.threshold[0] == 0.5 and tr.value[1] == 1.25 and t.final() is t
    assert DecisionTree.is_leaf(np.zeros((1, 4)), 1) and not DecisionTree.is_leaf(np.zeros((2, 4)), 1)


def test_install_aliases():
    import sys
    import np_decision_tree_b200 as pkg
    saved = {k: sys.modules.get(k) for k in ("decision_tree", "decision_tree.one", "decision_tree.strategy_l1",
                                              "decision_tree.utils", "np_stream_median")}
    try:
        pkg.install_aliases()
        from decision_tree.one import build_regression_tree   # noqa: F401  (the reference's import path)
        from decision_tree.utils import inference              # noqa: F401
        from np_stream_median import cummedian                 # noqa: F401
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_shard_columns():
    from np_decision_tree_b200.sharded import shard_columns
    for F in (1, 7, 100, 256, 257):
        for G in (1, 2, 3, 4, 8):
            spans = [shard_columns(F, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_classification_mirror_host_side_checks():
    """Host-side argument checks and helpers of the tree_builder / strategy mirrors run before any GPU work."""
    import numpy as np
    import pytest
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.decision_tree import strategy, tree_builder
    X = np.random.RandomState(0).standard_normal((20, 5))
    y = np.arange(20) % 2
    # every splitter of the reference is now routed somewhere; without a GPU the first kernel call fails loudly
    with pytest.raises((_lib.B200Error, OSError)):
        strategy.random_split(X, y.astype(float))
    with pytest.raises((_lib.B200Error, OSError)):
        tree_builder.build_regression_tree(X, y.astype(float))             # reference default: random_split
    # the mask-based improvement helpers are plain numpy, like the reference's (strategy.py:108-132)
    masks = X <= np.median(X, axis=0)
    yy = X[:, 0] * 2.0
    want = (np.square(masks.T.dot(yy)) / masks.sum(0) + np.square(yy.sum() - masks.T.dot(yy)) / (20 - masks.sum(0))) / 20 \
        - np.square(np.mean(yy))
    assert np.allclose(strategy.cal_variance_improvements(yy, masks), want)
    assert strategy.cal_gini_improvements(np.eye(2)[y], masks).shape == (5,)
    assert strategy.surrogate_to_variance(8.0, 4, 1.0) == 1.0
    ls = np.array([1.0, 3.0])
    assert np.array_equal(strategy.surrogate_variance_improvements(ls, np.array([1, 2]), 6.0, 3),
                          np.square(ls) / np.array([1, 2]) + np.square(6.0 - ls) / np.array([2, 1]))
    with pytest.raises(IndexError):
        tree_builder.build_classification_tree(X, y + 5, split_method=strategy.greedy_classification)
    with pytest.raises(ValueError, match="one-hot"):
        strategy.greedy_classification(X, np.ones((20, 2)))
    assert strategy.surrogate_to_gini(30.0, np.eye(2)[y]) == 30.0 / 20 - 0.5
    # host helpers with the reference's two calling conventions (strategy.py:32-55)
    cum = np.cumsum(np.eye(2)[y][:, np.newaxis, :], axis=0)               # (n, 1, C) running class counts
    a = strategy.surrogate_gini_improvements(cum, np.arange(1, 21)[:, np.newaxis])
    b = strategy.surrogate_gini_improvements(cum[:-1], np.arange(1, 20)[:, np.newaxis], cum[-1], 20)
    assert a.shape == (19, 1) and np.array_equal(a, b)


def test_variant_library_override(monkeypatch):
    """B200DT_LIB (tuning builds, scripts/build_variant.sh) replaces the default library and fails loudly when the
    file is missing -- never a silent fall back to the default build."""
    from np_decision_tree_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("B200DT_LIB", "/nonexistent/libb200tree_variant.so")
    with pytest.raises(_lib.B200Error):
        _lib.load()
    monkeypatch.setenv("B200DT_LIB", _lib.LIB_PATH)
    monkeypatch.setattr(_lib, "_lib", None)
    lib = _lib.load()
    assert hasattr(lib, "b200dt_session_begin") and lib.b200dt_abi_version() > 0
    monkeypatch.setattr(_lib, "_lib", None)   # the next user loads the default again
