"""GPU: the feature-sharded fit must build the IDENTICAL tree for any number of ranks.

* emulated ranks: G sessions (one context each) on one GPU, driven level by level exactly like
  sharded.fit_sharded, with the all-gather / all-reduce replaced by torch.cat / sum.  Runs on the
  1-GPU box and covers the whole host + kernel logic of sharding (shadow last column, global
  feature ids, owner marking, uneven slices).
* real NCCL: two processes, two GPUs (skipped when fewer than 2 GPUs are visible).
"""
import ctypes
import os
import socket

import numpy as np
import pytest

import oracle
from util import assert_same_tree, small_problem

pytestmark = pytest.mark.gpu


def fit_emulated(X, y, world, max_depth, criterion=0, v=1):
    import torch
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.decision_tree.one import DecisionTree
    from np_decision_tree_b200.sharded import shard_columns
    dev = torch.device("cuda", 0)
    N, F = X.shape
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    yd = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    last = Xd[:, F - 1].contiguous()
    ctxs, slices, masks = [], [], []
    stream = torch.cuda.current_stream(dev).cuda_stream
    for r in range(world):
        lo, hi = shard_columns(F, world, r)
        ctx = _lib.Context(0)
        ctx.set_stream(stream)
        Xl = Xd[:, lo:hi].contiguous()
        slices.append(Xl)
        ctx.check(ctx.lib.b200dt_session_begin(ctx.handle, Xl.data_ptr(), yd.data_ptr(), N, hi - lo, hi - lo,
                                               _lib.DEVICE, criterion, max_depth, 1e-7, 1, v, lo, F, None, 0))
        ctxs.append(ctx)
        masks.append(torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev))
    n_live = ctypes.c_int64()
    while True:
        ctxs[0].check(ctxs[0].lib.b200dt_session_live(ctxs[0].handle, ctypes.byref(n_live)))
        if n_live.value == 0:
            break
        cands = [torch.empty(n_live.value * _lib.CANDIDATE_BYTES, dtype=torch.uint8, device=dev) for _ in ctxs]
        for ctx, c in zip(ctxs, cands):
            ctx.check(ctx.lib.b200dt_session_search(ctx.handle, c.data_ptr()))
        gathered = torch.cat(cands)                      # == all_gather_into_tensor layout [rank][node]
        flags = []
        for ctx, m in zip(ctxs, masks):
            p, e = ctypes.c_int(), ctypes.c_int()
            ctx.check(ctx.lib.b200dt_session_decide(ctx.handle, gathered.data_ptr(), world, m.data_ptr(),
                                                    ctypes.byref(p), ctypes.byref(e)))
            flags.append((p.value, e.value))
        assert len(set(flags)) == 1, "ranks disagree"
        rounds = 0
        while flags[0][1]:      # totals round(s) and / or the sequential-order search (see sharded.fit_sharded)
            rounds += 1
            assert rounds <= 4
            for ctx, c in zip(ctxs, cands):
                ctx.check(ctx.lib.b200dt_session_search_exact(ctx.handle, c.data_ptr()))
            gathered = torch.cat(cands)
            flags = []
            for ctx, m in zip(ctxs, masks):
                p, e = ctypes.c_int(), ctypes.c_int()
                ctx.check(ctx.lib.b200dt_session_decide(ctx.handle, gathered.data_ptr(), world, m.data_ptr(),
                                                        ctypes.byref(p), ctypes.byref(e)))
                flags.append((p.value, e.value))
            assert len(set(flags)) == 1, "ranks disagree"
        if flags[0][0]:
            total = torch.stack(masks).sum(dim=0, dtype=torch.int32)   # == all_reduce(SUM)
            for ctx, m in zip(ctxs, masks):
                m.copy_(total)
                ctx.check(ctx.lib.b200dt_session_partition(ctx.handle, m.data_ptr()))
    trees = []
    for ctx in ctxs:
        tree = DecisionTree(2 ** (max_depth + 1))
        t = tree.tree_
        count = ctypes.c_int64()
        ctx.check(ctx.lib.b200dt_session_finish(ctx.handle, t.children_left.ctypes.data,
                                                t.children_right.ctypes.data, t.feature.ctypes.data,
                                                t.threshold.ctypes.data, t.value.ctypes.data, t.feature.shape[0],
                                                ctypes.byref(count)))
        tree.max_node_ = int(count.value) - 1
        trees.append(tree)
        ctx.close()
    return trees


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("n,F,depth,seed,yr", [(20000, 16, 8, 1, None), (3000, 10, 10, 2, 20.0)])
def test_emulated_ranks_same_tree(world, n, F, depth, seed, yr):
    X, y = small_problem(seed, n, F, yr, informative=4)
    want = oracle.fit_l2(X, y, max_depth=depth, sort_kind="stable")
    trees = fit_emulated(X, y, world, depth)
    for r, t in enumerate(trees):
        assert_same_tree(t, want, what=f"rank {r} of {world}")


def test_emulated_ranks_duplicate_columns_across_ranks():
    """Exact ties between columns owned by DIFFERENT ranks: lowest global feature id must win."""
    X, y = small_problem(3, 12000, 4)
    X = np.concatenate([X, X], axis=1)          # column f and f+4 identical, on different ranks for world=2
    want = oracle.fit_l2(X, y, max_depth=6, sort_kind="stable")
    for t in fit_emulated(X, y, 2, 6):
        assert_same_tree(t, want)


def test_emulated_ranks_config_c2(golden_configs, c2_data):
    from util import GoldenTree
    X, y = c2_data
    for t in fit_emulated(X, y, 4, 10):
        assert_same_tree(t, GoldenTree(golden_configs, "c2_"), what="C2 on 4 emulated ranks")


def _nccl_worker(rank, world, port, X, y, depth, out):
    import torch
    import torch.distributed as dist
    from np_decision_tree_b200.sharded import fit_sharded, shard_columns
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        F = X.shape[1]
        lo, hi = shard_columns(F, world, rank)
        Xl = torch.from_numpy(np.ascontiguousarray(X[:, lo:hi])).cuda()
        yd = torch.from_numpy(y).cuda()
        last = torch.from_numpy(np.ascontiguousarray(X[:, F - 1])).cuda()
        t = fit_sharded(Xl, yd, lo, F, last_col=last, max_depth=depth)
        out[rank] = {k: getattr(t.tree_, k).copy() for k in t.tree_._fields} | {"max_node": t.max_node_}
    finally:
        dist.destroy_process_group()


def test_nccl_two_gpus_same_tree():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    X, y = small_problem(4, 50000, 12, None, informative=4)
    depth = 8
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, X, y, depth, out), nprocs=2, join=True)
    want = oracle.fit_l2(X, y, max_depth=depth, sort_kind="stable")
    for r in (0, 1):
        got = out[r]
        assert got["max_node"] == want.max_node_
        for k in ("children_left", "children_right", "feature", "threshold"):
            assert np.array_equal(got[k], getattr(want.tree_, k)), (r, k)


# ---------------------------------------------------------------------------------------------
# classification (csrc/classes.cu sessions): same idea, one more exchanged array (left class counts)
# ---------------------------------------------------------------------------------------------
def class_problem(seed, n, F, C, noise=0.5):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, F))
    W = rs.standard_normal((F, C)) * (rs.uniform(size=(F, 1)) < 0.6)
    labels = np.argmax(X @ W + noise * rs.standard_normal((n, C)), axis=1).astype(np.int64)
    return X, labels


def fit_classes_emulated(X, labels, C, world, max_depth, criterion, min_improvement=1e-7, min_sample_leaf=1):
    import torch
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.decision_tree.base import DecisionTree
    from np_decision_tree_b200.sharded import shard_columns
    dev = torch.device("cuda", 0)
    N, F = X.shape
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    ld = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ctxs, keep, masks = [], [], []
    for r in range(world):
        lo, hi = shard_columns(F, world, r)
        ctx = _lib.Context(0)
        ctx.set_stream(stream)
        Xl = Xd[:, lo:hi].contiguous()
        keep.append(Xl)
        ctx.check(ctx.lib.b200dt_class_session_begin(ctx.handle, Xl.data_ptr(), ld.data_ptr(), N, hi - lo, hi - lo,
                                                     _lib.DEVICE, C, criterion, max_depth, min_improvement,
                                                     min_sample_leaf, lo, F))
        ctxs.append(ctx)
        masks.append(torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev))
    n_live = ctypes.c_int64()
    while True:
        ctxs[0].check(ctxs[0].lib.b200dt_class_session_live(ctxs[0].handle, ctypes.byref(n_live)))
        if n_live.value == 0:
            break
        cands = [torch.empty(n_live.value * _lib.CLASS_CANDIDATE_BYTES, dtype=torch.uint8, device=dev) for _ in ctxs]
        counts = [torch.zeros(n_live.value * C, dtype=torch.int32, device=dev) for _ in ctxs]
        for ctx, c in zip(ctxs, cands):
            ctx.check(ctx.lib.b200dt_class_session_search(ctx.handle, c.data_ptr()))
        gathered = torch.cat(cands)
        splits = []
        for ctx, m, lc in zip(ctxs, masks, counts):
            ns = ctypes.c_int64()
            ctx.check(ctx.lib.b200dt_class_session_decide(ctx.handle, gathered.data_ptr(), world, m.data_ptr(),
                                                          lc.data_ptr(), ctypes.byref(ns)))
            splits.append(ns.value)
        assert len(set(splits)) == 1, "ranks disagree"
        if splits[0]:
            mask_sum = torch.stack(masks).sum(dim=0, dtype=torch.int32)     # == all_reduce(SUM)
            count_sum = torch.stack(counts).sum(dim=0, dtype=torch.int32)
            for ctx, m, lc in zip(ctxs, masks, counts):
                m.copy_(mask_sum)
                lc.copy_(count_sum)
                ctx.check(ctx.lib.b200dt_class_session_partition(ctx.handle, m.data_ptr(), lc.data_ptr()))
    trees = []
    for ctx in ctxs:
        tree = DecisionTree(2 ** (max_depth + 1))
        t = tree.tree_
        count = ctypes.c_int64()
        ctx.check(ctx.lib.b200dt_class_session_finish(ctx.handle, t.children_left.ctypes.data,
                                                      t.children_right.ctypes.data, t.feature.ctypes.data,
                                                      t.threshold.ctypes.data, t.value.ctypes.data, None,
                                                      t.feature.shape[0], ctypes.byref(count)))
        tree.max_node_ = int(count.value) - 1
        trees.append(tree)
        ctx.close()
    return trees


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("criterion", ["gini", "p_at_k"])
def test_emulated_ranks_same_classification_tree(world, criterion):
    from np_decision_tree_b200 import _lib
    X, labels = class_problem(41, 6000, 16, 5)
    want = oracle.fit_classes(X, labels, 5, max_depth=8, max_feature=16, min_improvement=1e-7, criterion=criterion)
    crit = _lib.GINI if criterion == "gini" else _lib.P_AT_1
    for r, t in enumerate(fit_classes_emulated(X, labels, 5, world, 8, crit)):
        assert_same_tree(t, want, leaf_rtol=0.0, what=f"{criterion}: rank {r} of {world}")


def test_emulated_ranks_classification_duplicate_columns_and_ties():
    """Identical columns on different ranks and repeated feature values: the lowest GLOBAL column id wins
    and the value-based split extension is applied by the owner."""
    from np_decision_tree_b200 import _lib
    X, labels = class_problem(42, 4000, 3, 3)
    X = np.round(X * 4.0)
    X = np.concatenate([X, X], axis=1)
    want = oracle.fit_classes(X, labels, 3, max_depth=6, max_feature=6, min_improvement=1e-7, sort_kind="stable")
    for t in fit_classes_emulated(X, labels, 3, 2, 6, _lib.GINI):
        assert_same_tree(t, want, leaf_rtol=0.0)


def test_emulated_ranks_classification_golden(golden_classes):
    from np_decision_tree_b200 import _lib
    from util import GoldenTree
    g, p = golden_classes, "cl_g5_"
    trees = fit_classes_emulated(g[p + "X"], g[p + "labels"], int(g[p + "classes"]), 4, int(g[p + "depth"]), _lib.GINI,
                                 float(g[p + "min_improvement"]), int(g[p + "min_sample_leaf"]))
    for t in trees:
        assert_same_tree(t, GoldenTree(g, p), leaf_rtol=0.0, what="g5 on 4 emulated ranks")


def test_fit_classes_sharded_single_process():
    """The public sharded entry point with world size 1 (no process group) equals the plain fit."""
    import torch
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.sharded import fit_classes_sharded
    X, labels = class_problem(43, 5000, 8, 4)
    want = oracle.fit_classes(X, labels, 4, max_depth=7, max_feature=8, min_improvement=1e-7)
    t = fit_classes_sharded(torch.from_numpy(X).cuda(), torch.from_numpy(labels).cuda(), 0, 8, 4,
                            criterion=_lib.GINI, max_depth=7, min_improvement=1e-7)
    assert_same_tree(t, want, leaf_rtol=0.0)


def _nccl_class_worker(rank, world, port, X, labels, C, depth, out):
    import torch
    import torch.distributed as dist
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.sharded import fit_classes_sharded, shard_columns
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        F = X.shape[1]
        lo, hi = shard_columns(F, world, rank)
        Xl = torch.from_numpy(np.ascontiguousarray(X[:, lo:hi])).cuda()
        t = fit_classes_sharded(Xl, torch.from_numpy(labels).cuda(), lo, F, C, criterion=_lib.GINI, max_depth=depth,
                                min_improvement=1e-7)
        out[rank] = {k: getattr(t.tree_, k).copy() for k in t.tree_._fields} | {"max_node": t.max_node_}
    finally:
        dist.destroy_process_group()


def test_nccl_two_gpus_same_classification_tree():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    X, labels = class_problem(44, 40000, 12, 6)
    depth = 8
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_class_worker, args=(2, port, X, labels, 6, depth, out), nprocs=2, join=True)
    want = oracle.fit_classes(X, labels, 6, max_depth=depth, max_feature=12, min_improvement=1e-7)
    for r in (0, 1):
        got = out[r]
        assert got["max_node"] == want.max_node_
        for k in ("children_left", "children_right", "feature", "threshold"):
            assert np.array_equal(got[k], getattr(want.tree_, k)), (r, k)
