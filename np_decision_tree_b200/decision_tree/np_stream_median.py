"""Mirror of the reference's native module ``np_stream_median`` (Cython, pyx:8-39)."""
import numpy as np

from .. import _lib
from ._common import context_for


def cummedian(arr):
    """(n, 2) array of (lower, upper) running medians of ``arr`` -- GPU order-statistic scan."""
    if hasattr(arr, "is_cuda") and arr.is_cuda:
        import torch
        ctx = context_for(arr)
        a = arr.contiguous()
        out = torch.empty((a.shape[0], 2), dtype=torch.float64, device=a.device)
        ctx.check(ctx.lib.b200dt_cummedian(ctx.handle, a.data_ptr(), a.shape[0], _lib.DEVICE,
                                           out.data_ptr(), _lib.DEVICE))
        return out
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("cummedian takes a 1-D array")
    ctx = context_for()
    out = np.zeros((a.shape[0], 2), dtype=np.float64)
    ctx.check(ctx.lib.b200dt_cummedian(ctx.handle, a.ctypes.data, a.shape[0], _lib.HOST,
                                       out.ctypes.data, _lib.HOST))
    return out
