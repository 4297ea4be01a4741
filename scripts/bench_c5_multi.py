"""BASELINE.json configs[4] (C5): inference of a depth-10 tree over 100 M x 256 rows, ROW-sharded over the
GPUs of one box with no exchange.  Launch under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_c5_multi.py

Every rank generates its own rows on the device (C5_ROWS total, default 100 M), fits the same depth-10 tree
on the same 1 M-row sample (same seed), predicts its shard K times; time = max over ranks (CUDA events).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from np_decision_tree_b200.decision_tree import one, utils  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total_rows = int(os.environ.get("C5_ROWS", 100_000_000))
F, steps = 256, 5
rows = total_rows // world
g = torch.Generator(device="cuda")
g.manual_seed(0)
Xs = torch.randn((1_000_000, F), dtype=torch.float64, device="cuda", generator=g)
cols = torch.arange(0, 250, 25, device="cuda")
ys = (Xs[:, cols] * torch.linspace(10, 100, 10, dtype=torch.float64, device="cuda")).sum(1)
tree = one.build_regression_tree(Xs, ys, max_depth=10)
del Xs, ys
g.manual_seed(1000 + rank)
X = torch.empty((rows, F), dtype=torch.float64, device="cuda")
for r0 in range(0, rows, 1_000_000):
    r1 = min(rows, r0 + 1_000_000)
    X[r0:r1] = torch.randn((r1 - r0, F), dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    p = utils.inference(X, tree)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    p = utils.inference(X, tree)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms.item()) / 1e3
    print(json.dumps({"config": "C5 inference, depth-10 tree, row-sharded, device-resident", "n_gpus": world,
                      "rows_total": rows * world, "rows_per_gpu": rows, "features": F, "nodes": tree.max_node_ + 1,
                      "ms_per_pass": t * 1e3, "rows_per_s": rows * world / t,
                      "rows_features_per_s": rows * world * F / t}))
if world > 1:
    dist.destroy_process_group()
