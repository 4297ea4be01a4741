"""Mirror of the reference module ``decision_tree/utils.py`` -- batched inference."""
import time

import numpy as np

from .. import _lib
from ._common import context_for, matrix_arg


def timeit(func):
    """utils.py:5-12."""
    def _(*args, **kwargs):
        start = time.time()
        res = func(*args, **kwargs)
        print('Func: %s, runtime: %.6f' % (func.__name__, time.time() - start))
        return res
    return _


def _tree_arrays(clf):
    t = clf.tree_
    feature = np.ascontiguousarray(t.feature, dtype=np.int64)
    n = feature.shape[0]
    return (feature, np.ascontiguousarray(t.threshold, dtype=np.float64),
            np.ascontiguousarray(t.children_left, dtype=np.int64),
            np.ascontiguousarray(t.children_right, dtype=np.int64),
            np.ascontiguousarray(np.asarray(t.value, dtype=np.float64).reshape(n, -1)[:, 0]), n)


def _traverse(data, clf, want_value, want_leaf):
    ctx = context_for(data)
    xp, xs, xk, N, F, ldx = matrix_arg(data, "data")
    feature, threshold, left, right, value, n_nodes = _tree_arrays(clf)
    on_device = xs == _lib.DEVICE
    out_v = out_l = None
    if on_device:
        import torch
        if want_value:
            out_v = torch.empty(N, dtype=torch.float64, device=data.device)
        if want_leaf:
            out_l = torch.empty(N, dtype=torch.int32, device=data.device)
        pv = out_v.data_ptr() if want_value else None
        pl = out_l.data_ptr() if want_leaf else None
    else:
        if want_value:
            out_v = np.empty(N, dtype=np.float64)
        if want_leaf:
            out_l = np.empty(N, dtype=np.int32)
        pv = out_v.ctypes.data if want_value else None
        pl = out_l.ctypes.data if want_leaf else None
    ctx.check(ctx.lib.b200dt_predict(
        ctx.handle, xp, N, F, ldx, xs, feature.ctypes.data, threshold.ctypes.data, left.ctypes.data,
        right.ctypes.data, value.ctypes.data, n_nodes, pv, pl, xs))
    # utils.py:20,26-31: when no row ever leaves the root the reference returns shape (1,)
    root_only = left[0] == -1 and right[0] == -1
    return out_v, out_l, root_only


def inference(data, clf):
    """Predict (utils.py:15-31): left iff data[row, feature] <= threshold.  ``clf`` may be a
    tree from this package, from the reference, or a fitted sklearn tree."""
    out, _, root_only = _traverse(data, clf, True, False)
    return out[:1] if root_only else out


def inference_leaf(data, clf):
    """Leaf id per row (utils.py:34-48); int64 like the reference (int32 for a root-only tree)."""
    _, out, root_only = _traverse(data, clf, False, True)
    if root_only:
        return out[:1]          # the reference's np.zeros(1, dtype=np.int32) (utils.py:37)
    # the reference's node ids become int64 after the first np.where with the int64 child arrays
    return out.astype(np.int64) if isinstance(out, np.ndarray) else out.long()
