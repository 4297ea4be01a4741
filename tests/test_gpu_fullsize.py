"""GPU: size-independent properties at BASELINE.json's full row count (10 000 000 rows), where the
CPU oracle cannot run (> 100 GB of host RAM for the reference at 10 M x 256).

Columns are reduced (8 instead of 256) so the test takes seconds; every kernel still runs at the
10 M-row tile counts, look-back chain lengths and index ranges of the real config.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, F, DEPTH = 10_000_000, 8, 10


@pytest.fixture(scope="module")
def problem():
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    X = torch.randn((N, F), dtype=torch.float64, device="cuda", generator=g)
    y = 80.0 * X[:, 1] - 35.0 * X[:, 4] + 12.0 * X[:, 6] * X[:, 1] + 0.5 * torch.randn(N, dtype=torch.float64,
                                                                                        device="cuda", generator=g)
    return X, y


def test_argsort_is_sorted_permutation(problem):
    """sortedness + permutation (checksum) of every column's order at 10 M rows"""
    import torch
    from np_decision_tree_b200 import _lib
    X, _ = problem
    ctx = _lib.default_context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    out = torch.empty((N, F), dtype=torch.int64, device="cuda")
    ctx.check(ctx.lib.b200dt_argsort_columns(ctx.handle, X.data_ptr(), N, F, F, 1, out.data_ptr(), 1))
    for f in range(F):
        o = out[:, f]
        v = X[o, f]
        assert bool((v[1:] >= v[:-1]).all()), f"column {f} not sorted"
        # stable: equal neighbours keep ascending row ids
        eq = v[1:] == v[:-1]
        assert bool((o[1:][eq] > o[:-1][eq]).all())
        assert int(o.sum()) == N * (N - 1) // 2 and int(o.min()) == 0 and int(o.max()) == N - 1
        assert int(torch.bincount(o[:1_000_000] % 1000, minlength=1000).sum()) == 1_000_000


def test_fit_invariants_and_inference_consistency(problem):
    import torch
    from np_decision_tree_b200.decision_tree import one, utils
    X, y = problem
    prev_mse = None
    for depth in (1, 4, DEPTH):
        t = one.build_regression_tree(X, y, max_depth=depth)
        tr = t.tree_
        n = t.max_node_ + 1
        assert n == 2 ** (depth + 1) - 1, "continuous data at 10 M rows must fill the tree"
        internal = tr.children_left[:n] != -1
        # pre-order ids: left child = id + 1, children ids inside the array, leaves carry sentinels
        ids = np.arange(n)
        assert np.array_equal(tr.children_left[:n][internal], ids[internal] + 1)
        assert (tr.children_right[:n][internal] > tr.children_left[:n][internal]).all()
        assert (tr.feature[:n][~internal] == -2).all() and (tr.threshold[:n][~internal] == -2.0).all()
        assert ((tr.feature[:n][internal] >= 0) & (tr.feature[:n][internal] < F)).all()
        # every threshold is a data value of its feature (ref one.py:104-105)
        for v in np.flatnonzero(internal)[:64]:
            f = int(tr.feature[v])
            assert bool((X[:, f] == float(tr.threshold[v])).any()), f"threshold of node {v} is not a data value"
        # inference routes rows exactly as the fit partitioned them: leaf value == mean of the rows there
        leaf = utils.inference_leaf(X, t)
        pred = utils.inference(X, t)
        counts = torch.bincount(leaf, minlength=n)
        sums = torch.zeros(n, dtype=torch.float64, device="cuda").index_add_(0, leaf, y)
        leaves = np.flatnonzero(~internal)
        assert int(counts[torch.from_numpy(leaves).cuda()].sum()) == N and bool((counts[torch.from_numpy(np.flatnonzero(internal)).cuda()] == 0).all())
        means = (sums / counts.clamp(min=1)).cpu().numpy()
        np.testing.assert_allclose(tr.value[leaves], means[leaves], rtol=1e-9, atol=1e-9 * float(y.abs().max()))
        mse = float(((pred - y) ** 2).mean())
        if prev_mse is not None:
            assert mse < prev_mse, "train MSE must fall with depth"
        prev_mse = mse
    # root split == exhaustive search done with torch on the GPU (independent of our kernels)
    best = None
    for f in range(F):
        o = torch.argsort(X[:, f], stable=True)
        s = torch.cumsum(y[o], 0)[:-1]
        k = torch.arange(1, N, dtype=torch.float64, device="cuda")
        gain = (s - k * (float(y.sum()) / N)) ** 2 / (k * (N - k))
        kk = int(torch.argmax(gain))
        cand = (float(gain[kk]), f, float(X[o[kk], f]))
        if best is None or cand[0] > best[0]:
            best = cand
    t1 = one.build_regression_tree(X, y, max_depth=1)
    assert int(t1.tree_.feature[0]) == best[1] and float(t1.tree_.threshold[0]) == best[2]


def test_fit_same_tree_on_two_ranks(problem):
    """1 rank vs 2 emulated ranks (4 + 4 columns) at 10 M rows: identical tree."""
    import torch
    from np_decision_tree_b200.decision_tree import one
    from test_gpu_sharded import fit_emulated
    from util import assert_same_tree
    X, y = problem
    Xh, yh = X.cpu().numpy(), y.cpu().numpy()
    del X, y
    torch.cuda.empty_cache()
    t1 = one.build_regression_tree(Xh, yh, max_depth=DEPTH)      # host-buffer path (overlapped H2D)
    for t in fit_emulated(Xh, yh, 2, DEPTH):
        assert_same_tree(t, t1, leaf_rtol=1e-9, what="2 ranks vs 1 rank at 10 M rows")
