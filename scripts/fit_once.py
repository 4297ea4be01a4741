"""Run the C4 fit a few times on device-resident synthetic data (for ncu / quick timing).

    python scripts/fit_once.py [rows] [features] [depth] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from np_decision_tree_b200 import _lib  # noqa: E402
from np_decision_tree_b200.sharded import fit_sharded  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
D = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
ctx = _lib.default_context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
X, y, last = bench.gen_problem_torch(N, F, 0, F, dev)
torch.cuda.synchronize()
for i in range(reps):
    ctx.profile_enable(True)
    ctx.profile_reset()
    t0 = time.perf_counter()
    tree = fit_sharded(X, y, 0, F, last_col=last, max_depth=D, ctx=ctx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = ctx.profile()
    print(f"fit {i}: {dt * 1e3:.1f} ms, {tree.max_node_ + 1} nodes; " +
          ", ".join(f"{k} {v['ms']:.1f}" for k, v in prof.items() if v["ms"] > 0.05))
