"""Feature-sharded fit: one process per GPU, columns split across ranks.

Every column's sort, scan and partition are independent; only the per-node winner and the
resulting left-row set couple the columns (ref decision_tree/one.py:59-62,106,111-113).  So
each rank runs a libb200tree *session* on its column slice and the ranks exchange, per level,

  1. an all-gather of one 56-byte candidate per live node (<= 512 nodes at depth 10), after
     which every rank applies the reference's tie rule (gain desc, position asc, feature asc)
     to the same gathered array and so reaches the same decision, and
  2. a sum-all-reduce of the N-bit left-row mask (the owner of each winning column sets the
     bits of its nodes; bits of different nodes are disjoint, so SUM == OR).

Both run over NCCL (NVLink 5 / NVSwitch) on torch's current stream, the same stream the
kernels are launched on.  The tree is identical for any number of ranks.
"""
import ctypes

import numpy as np

from . import _lib
from .decision_tree.one import DecisionTree


def shard_columns(n_features, world_size, rank):
    """Contiguous column slice [lo, hi) of ``rank``; the first F % G ranks get one extra."""
    base, extra = divmod(int(n_features), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_candidates(cand_bytes, world_size, group=None):
    """All-gather the per-node candidate records: returns a uint8 tensor laid out
    [rank][node][56 bytes], the layout b200dt_session_decide expects."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return cand_bytes
    out = torch.empty(world_size * cand_bytes.numel(), dtype=torch.uint8, device=cand_bytes.device)
    dist.all_gather_into_tensor(out, cand_bytes, group=group)
    return out


def reduce_side_mask(mask_words, world_size, group=None):
    """Combine the ranks' left-row masks in place (int32 words; disjoint bits => SUM == OR)."""
    import torch.distributed as dist
    if world_size > 1:
        dist.all_reduce(mask_words, op=dist.ReduceOp.SUM, group=group)
    return mask_words


def fit_sharded(X_local, y, col_offset, n_features, last_col=None, criterion=_lib.L2, max_depth=2,
                min_improvement=1e-7, min_sample_leaf=1, v=1, group=None, ctx=None):
    """Grow one tree over columns sharded across the ranks of ``group``.

    X_local: (N, F_local) float64 torch tensor, CUDA or (pinned) host -- this rank's columns
             [col_offset, col_offset + F_local) of the global (N, n_features) matrix.
    y:       (N,) float64 CUDA tensor, replicated.
    last_col: accepted for compatibility and ignored (round 1 kept a shadow copy of the last global column on
             every rank; the node totals the reference takes in that column's order, one.py:57, now travel
             inside the candidate records of the rank that owns the column).
    Returns a DecisionTree identical on every rank.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if X_local.is_cuda != y.is_cuda:
        raise TypeError("X_local and y must both be CUDA tensors or both be host tensors")
    on_device = X_local.is_cuda
    if on_device:
        dev = X_local.device
    else:
        # host (ideally pinned) tensors: the library copies the column batches on a second stream
        # while it sorts the ones already there
        dev = torch.device("cuda", torch.cuda.current_device())
    ctx = ctx or _lib.default_context(dev.index)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    X_local = X_local.contiguous()
    y = y.contiguous()
    N, F_local = X_local.shape
    space = _lib.DEVICE if on_device else _lib.HOST
    lib, h = ctx.lib, ctx.handle
    ctx.check(lib.b200dt_session_begin(h, X_local.data_ptr(), y.data_ptr(), N, F_local, X_local.stride(0),
                                       space, criterion, int(max_depth), float(min_improvement),
                                       int(min_sample_leaf), 1 if v == 1 else 2, int(col_offset),
                                       int(n_features), None, 0))
    mask = torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev)
    n_live = ctypes.c_int64()
    need_part, need_exact = ctypes.c_int(), ctypes.c_int()
    while True:
        ctx.check(lib.b200dt_session_live(h, ctypes.byref(n_live)))
        if n_live.value == 0:
            break
        cand = torch.empty(n_live.value * _lib.CANDIDATE_BYTES, dtype=torch.uint8, device=dev)
        ctx.check(lib.b200dt_session_search(h, cand.data_ptr()))
        gathered = gather_candidates(cand, world, group)
        ctx.check(lib.b200dt_session_decide(h, gathered.data_ptr(), world, mask.data_ptr(),
                                            ctypes.byref(need_part), ctypes.byref(need_exact)))
        rounds = 0
        while need_exact.value:
            # some nodes must be searched in the reference's sequential fp64 order (small nodes, or a near-tie
            # between two candidates of a parallel-scan search).  That search needs the node total in the order
            # of the LAST global column (one.py:57): the rank owning that column publishes it in its candidate
            # record first (one round), then every rank searches (next round), then the decision stands.
            rounds += 1
            if rounds > 4:
                raise _lib.B200Error("exact re-search did not resolve a tie")
            ctx.check(lib.b200dt_session_search_exact(h, cand.data_ptr()))
            gathered = gather_candidates(cand, world, group)
            ctx.check(lib.b200dt_session_decide(h, gathered.data_ptr(), world, mask.data_ptr(),
                                                ctypes.byref(need_part), ctypes.byref(need_exact)))
        if need_part.value:
            reduce_side_mask(mask, world, group)
            ctx.check(lib.b200dt_session_partition(h, mask.data_ptr()))
    tree = DecisionTree(2 ** (max_depth + 1))
    t = tree.tree_
    count = ctypes.c_int64()
    ctx.check(lib.b200dt_session_finish(h, t.children_left.ctypes.data, t.children_right.ctypes.data,
                                        t.feature.ctypes.data, t.threshold.ctypes.data, t.value.ctypes.data,
                                        t.feature.shape[0], ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    return tree


def fit_classes_sharded(X_local, labels, col_offset, n_features, n_classes, criterion=_lib.GINI, max_depth=2,
                        min_improvement=0.01, min_sample_leaf=1, group=None, ctx=None):
    """Grow one classification tree (exhaustive gini / precision-at-1 search, ref
    decision_tree/tree_builder.py:56-96,149-159) over columns sharded across the ranks of ``group``.

    X_local: (N, F_local) float64 CUDA tensor -- this rank's columns [col_offset, col_offset + F_local).
    labels:  (N,) int32 CUDA tensor of class ids in [0, n_classes), replicated.
    Per level: all-gather of one 32-byte candidate per live node, then a sum-all-reduce of the N-bit
    left-row mask and of the left children's class counts (both written only by the owner of the
    winning column).  Returns a base.DecisionTree identical on every rank and to the 1-GPU tree.
    """
    import torch
    import torch.distributed as dist
    from .decision_tree.base import DecisionTree as ClassTree
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if not (X_local.is_cuda and labels.is_cuda):
        raise TypeError("fit_classes_sharded takes CUDA tensors")
    dev = X_local.device
    ctx = ctx or _lib.default_context(dev.index)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    X_local = X_local.contiguous()
    labels = labels.to(torch.int32).contiguous()
    N, F_local = X_local.shape
    lib, h = ctx.lib, ctx.handle
    ctx.check(lib.b200dt_class_session_begin(h, X_local.data_ptr(), labels.data_ptr(), N, F_local, X_local.stride(0),
                                             _lib.DEVICE, int(n_classes), int(criterion), int(max_depth),
                                             float(min_improvement), int(min_sample_leaf), int(col_offset),
                                             int(n_features)))
    mask = torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev)
    n_live, n_split = ctypes.c_int64(), ctypes.c_int64()
    while True:
        ctx.check(lib.b200dt_class_session_live(h, ctypes.byref(n_live)))
        if n_live.value == 0:
            break
        cand = torch.empty(n_live.value * _lib.CLASS_CANDIDATE_BYTES, dtype=torch.uint8, device=dev)
        left_counts = torch.zeros(n_live.value * int(n_classes), dtype=torch.int32, device=dev)
        ctx.check(lib.b200dt_class_session_search(h, cand.data_ptr()))
        gathered = gather_candidates(cand, world, group)
        ctx.check(lib.b200dt_class_session_decide(h, gathered.data_ptr(), world, mask.data_ptr(),
                                                  left_counts.data_ptr(), ctypes.byref(n_split)))
        if n_split.value:
            reduce_side_mask(mask, world, group)
            reduce_side_mask(left_counts, world, group)
            ctx.check(lib.b200dt_class_session_partition(h, mask.data_ptr(), left_counts.data_ptr()))
    tree = ClassTree(2 ** (max_depth + 1))
    t = tree.tree_
    count = ctypes.c_int64()
    ctx.check(lib.b200dt_class_session_finish(h, t.children_left.ctypes.data, t.children_right.ctypes.data,
                                              t.feature.ctypes.data, t.threshold.ctypes.data, t.value.ctypes.data,
                                              None, t.feature.shape[0], ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    return tree


def predict_sharded_rows(X_rows, tree, ctx=None):
    """Inference shards by rows with no exchange: each rank predicts its own rows."""
    from .decision_tree.utils import inference
    return inference(X_rows, tree)
