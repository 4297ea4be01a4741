"""Feature-sharded gini classification fit (sharded.fit_classes_sharded), 10 M x 256, 8 classes, depth 10,
device-resident, columns split across the ranks.  Launch under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_cls_multi.py

Every rank generates the same data stream column block by column block (seeded per 32-column block), keeps its
own columns, and computes the labels from the first 8 columns + noise (also generated identically on every rank).
Time = max over ranks, CUDA events, best of 3 after one warm-up fit.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from np_decision_tree_b200 import _lib  # noqa: E402
from np_decision_tree_b200.sharded import fit_classes_sharded, shard_columns  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(os.environ.get("CLS_ROWS", 10_000_000))
F = int(os.environ.get("CLS_FEATURES", 256))
C = int(os.environ.get("CLS_CLASSES", 8))
lo, hi = shard_columns(F, world, rank)
g = torch.Generator(device="cuda")
X = torch.empty((N, hi - lo), dtype=torch.float64, device="cuda")
first = None
for c0 in range(0, F, 32):
    g.manual_seed(100 + c0)
    blk = torch.randn((N, min(32, F - c0)), dtype=torch.float64, device="cuda", generator=g)
    if c0 == 0:
        first = blk[:, :C].clone()
    a, b = max(lo, c0), min(hi, c0 + blk.shape[1])
    if a < b:
        X[:, a - lo:b - lo] = blk[:, a - c0:b - c0]
    del blk
g.manual_seed(7)
labels = (first + 0.5 * torch.randn((N, C), dtype=torch.float64, device="cuda", generator=g)).argmax(dim=1).to(torch.int32)
del first
kw = dict(criterion=_lib.GINI, max_depth=10, min_improvement=1e-7)
tree = fit_classes_sharded(X, labels, lo, F, C, **kw)
best = None
for _ in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tree = fit_classes_sharded(X, labels, lo, F, C, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    best = float(ms.item()) if best is None else min(best, float(ms.item()))
if rank == 0:
    print(json.dumps({"config": "gini classification fit, feature-sharded, device-resident", "n_gpus": world, "rows": N,
                      "features": F, "classes": C, "max_depth": 10, "nodes": tree.max_node_ + 1, "ms_per_fit": best,
                      "rows_features_per_s": N * F / (best / 1e3),
                      "root": [int(tree.tree_.feature[0]), float(tree.tree_.threshold[0])]}))
if world > 1:
    dist.destroy_process_group()
