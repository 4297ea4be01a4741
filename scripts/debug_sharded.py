import sys, os, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import small_problem
from np_decision_tree_b200 import _lib
from np_decision_tree_b200.sharded import shard_columns
DT = np.dtype([("gain", "f8"), ("aux", "f8"), ("left_sum", "f8"), ("gain2", "f8"), ("threshold", "f8"), ("k", "i4"), ("f", "i4"), ("exact", "i4"), ("pad", "i4")])
X, y = small_problem(3, 12000, 4)
X = np.concatenate([X, X], axis=1)
world, max_depth = 2, 6
dev = torch.device("cuda", 0)
N, F = X.shape
Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev); yd = torch.from_numpy(y).to(dev)
last = Xd[:, F - 1].contiguous()
ctxs, keep, masks = [], [], []
for r in range(world):
    lo, hi = shard_columns(F, world, r)
    ctx = _lib.Context(0); ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    Xl = Xd[:, lo:hi].contiguous(); keep.append(Xl)
    ctx.check(ctx.lib.b200dt_session_begin(ctx.handle, Xl.data_ptr(), yd.data_ptr(), N, hi - lo, hi - lo, 1, 0, max_depth, 1e-7, 1, 1, lo, F, None if hi == F else last.data_ptr(), 1))
    ctxs.append(ctx); masks.append(torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev))
n_live = ctypes.c_int64()
ctxs[0].check(ctxs[0].lib.b200dt_session_live(ctxs[0].handle, ctypes.byref(n_live)))
cands = [torch.zeros(n_live.value * 56, dtype=torch.uint8, device=dev) for _ in ctxs]
for rnd in range(2):
    for ctx, c in zip(ctxs, cands):
        fn = ctx.lib.b200dt_session_search if rnd == 0 else ctx.lib.b200dt_session_search_exact
        ctx.check(fn(ctx.handle, c.data_ptr()))
    g = torch.cat(cands)
    print("round", rnd, g.cpu().numpy().view(DT))
    for ctx, m in zip(ctxs, masks):
        p, e = ctypes.c_int(), ctypes.c_int()
        ctx.check(ctx.lib.b200dt_session_decide(ctx.handle, g.data_ptr(), world, m.data_ptr(), ctypes.byref(p), ctypes.byref(e)))
        print("   decide ->", p.value, e.value)
