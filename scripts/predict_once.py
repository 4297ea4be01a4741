"""Fit a depth-10 tree on a small problem, then run utils.inference over device-resident rows (for ncu / timing).

    python scripts/predict_once.py [rows] [features] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from np_decision_tree_b200 import _lib  # noqa: E402
from np_decision_tree_b200.decision_tree import one, utils  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
ctx = _lib.default_context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
Xf, yf, _ = bench.gen_problem_torch(2_000_000, F, 0, F, dev)
tree = one.build_regression_tree(Xf, yf, max_depth=10)
del Xf, yf
g = torch.Generator(device=dev)
g.manual_seed(5)
X = torch.randn((N, F), dtype=torch.float64, device=dev, generator=g)
torch.cuda.synchronize()
for i in range(reps):
    ctx.profile_enable(True)
    ctx.profile_reset()
    t0 = time.perf_counter()
    out = utils.inference(X, tree)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    p = ctx.profile()["predict"]
    print(f"predict {i}: call {dt * 1e3:.2f} ms, kernel {p['ms']:.3f} ms, {p['alg_bytes'] / N:.0f} B/row, "
          f"{p['alg_bytes'] / p['ms'] / 1e6:.0f} GB/s, checksum {float(out.sum()):.6f}")
