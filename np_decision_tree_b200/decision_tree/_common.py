"""Shared host glue of the mirror modules: context / stream selection and the fit call."""
import ctypes

import numpy as np

from .. import _lib


def context_for(*arrays):
    """Context of the device the inputs live on (torch CUDA tensors) or the default one."""
    for a in arrays:
        if hasattr(a, "is_cuda") and a.is_cuda:
            import torch
            ctx = _lib.default_context(a.device.index)
            ctx.set_stream(torch.cuda.current_stream(a.device).cuda_stream)
            return ctx
    return _lib.default_context()


def matrix_arg(X, what="X"):
    """(ptr, space, keepalive, N, F, ldx) of a 2-D float64 matrix (numpy or torch)."""
    if hasattr(X, "data_ptr") and not isinstance(X, np.ndarray):
        if X.dim() != 2:
            raise ValueError(f"{what} must be 2-D")
        ptr, space, keep = _lib.as_arg(X, np.float64, what)
        return ptr, space, keep, int(X.shape[0]), int(X.shape[1]), int(X.shape[1])
    arr = np.ascontiguousarray(X, dtype=np.float64)
    if arr.ndim != 2:
        raise ValueError(f"{what} must be 2-D")
    return arr.ctypes.data, _lib.HOST, arr, arr.shape[0], arr.shape[1], arr.shape[1]


def fit_into(tree, X, y, criterion, max_depth, min_improvement, min_sample_leaf, variant):
    """Run b200dt_fit and fill ``tree`` (a DecisionTree with the reference's array layout)."""
    ctx = context_for(X, y)
    xp, xs, xk, N, F, ldx = matrix_arg(X)
    yp, ys, yk = _lib.as_arg(y, np.float64, "y")
    if xs != ys:
        raise ValueError("X and y must both be host arrays or both be CUDA tensors")
    ny = int(np.prod(y.shape))
    if ny != N:
        raise ValueError(f"y has {ny} entries, X has {N} rows")
    t = tree.tree_
    count = ctypes.c_int64(0)
    ctx.check(ctx.lib.b200dt_fit(
        ctx.handle, xp, yp, N, F, ldx, xs, criterion, int(max_depth), float(min_improvement),
        int(min_sample_leaf), int(variant), t.children_left.ctypes.data, t.children_right.ctypes.data,
        t.feature.ctypes.data, t.threshold.ctypes.data, t.value.ctypes.data, t.feature.shape[0],
        ctypes.byref(count)))
    tree.max_node_ = int(count.value) - 1
    return tree


def orders_arg(orders):
    arr = np.ascontiguousarray(orders, dtype=np.int64)
    if arr.ndim != 2:
        raise ValueError("orders must be (n, F)")
    return arr
