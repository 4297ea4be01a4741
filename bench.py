#!/usr/bin/env python
"""bench.py -- depth-10 MSE fit throughput (rows*features/s), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one whole fit (argsort of every column + 10 levels of split search and partition +
tree read-back) of the synthetic 10 000 000 x 256 float64 regression problem of
BASELINE.json configs[3] (make_regression's recipe: N(0,1) features, 10 informative columns
with coefficients 100*U(0,1), no noise).  With N > 1 the columns of the SAME problem are
sharded over the ranks (strong scaling).

  value : N*F / t, inputs resident in HBM when the timed region starts (CUDA events on the
          launch stream, barrier + synchronize on both sides, max over ranks).
  e2e   : the same fit through the public API with HOST buffers -- every step copies X and y
          host->device and the tree device->host inside the timed region.  Measured twice: with an
          ordinary pageable numpy array (what a user of the reference passes; this is `e2e.value`)
          and with a pinned buffer (`pinned_ms_per_step`).
  inference : utils.inference of the depth-10 tree of the timed fit over 12 500 000 x F device-resident
          rows PER GPU (row-sharded, no exchange; at 8 GPUs this is BASELINE.json configs[4],
          100 M x 256), rows/s and the fraction of the HBM roofline on 32 B x path + 8 B per row.
  tree_sha : SHA-256 over children / feature / threshold up to max_node_ (bit-exact quantities) -- identical at
          every GPU count; value_sha: the node values rounded to 9 significant digits.
  roofline : the kernel class with the largest share of the step; achieved = algorithmic
          bytes / summed CUDA-event time of its launches in the timed
          region; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the oracle (numpy port of the reference, single thread like the reference)
          on a bounded row sample of the same generator, rank 0, N=1 only.

--impl reference times the reference's CPU algorithm on a bounded sample per step: the oracle port
(`kind: "port"`).  The reference is Python; its sources are never copied into this tree (not even
into a git-ignored directory), so the unmodified reference itself only runs in the build container,
where tests/golden/make_golden*.py use it to generate the fixtures the oracle is pinned to.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth-10 MSE fit throughput"
UNIT = "rows*features/s"
N_INFORMATIVE = 10
SEED = 0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--features", type=int, default=256)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--cpu-rows", type=int, default=0, help="row sample of the CPU arms (0 = sized to the time budget)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--inference-rows", type=int, default=12_500_000, help="rows per GPU of the inference block")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic problem: identical for any number of ranks (columns are generated one by one)
# ---------------------------------------------------------------------------------------------
def informative_columns(F):
    rs = np.random.RandomState(SEED)
    cols = np.sort(rs.permutation(F)[:min(N_INFORMATIVE, F)])
    coef = 100.0 * rs.uniform(size=cols.shape[0])
    return cols, coef


def gen_column_torch(N, c, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED * 1_000_003 + 7919 * c + 1)
    return torch.randn(N, dtype=torch.float64, device=device, generator=g)


def gen_problem_torch(N, F, lo, hi, device):
    """Columns [lo, hi) of X and the replicated y."""
    import torch
    X = torch.empty((N, hi - lo), dtype=torch.float64, device=device)
    for c in range(lo, hi):
        X[:, c - lo] = gen_column_torch(N, c, device)
    cols, coef = informative_columns(F)
    y = torch.zeros(N, dtype=torch.float64, device=device)
    for c, w in zip(cols, coef):
        col = X[:, c - lo] if lo <= c < hi else gen_column_torch(N, int(c), device)
        y += w * col
    return X, y, None      # (third slot: the global last column, which ranks no longer need a copy of)


def gen_problem_numpy(N, F):
    """CPU-side sample of the same recipe (numpy stream; used only for the CPU arms)."""
    rs = np.random.RandomState(SEED + 1)
    X = rs.standard_normal((N, F))
    cols, coef = informative_columns(F)
    return X, X[:, cols] @ coef


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # "under load": samples in the upper half of what was seen (idle gaps sit at low clocks)
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU arms (oracle port; single thread like the reference)
# ---------------------------------------------------------------------------------------------
def cpu_fit_seconds(rows, F, depth, repeats=1):
    import oracle
    X, y = gen_problem_numpy(rows, F)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.fit_l2(X, y, max_depth=depth)
        best = min(best, time.perf_counter() - t0)
    return best


CPU_SECONDS_PER_ROW = 6.4e-5 / 256   # oracle port, per (row, feature) of a depth-10 fit (measured: 100 000 x 256 in 6.4 s)


def reference_rows(args, fits):
    """Rows of the CPU sample such that `fits` whole fits take about two minutes (50 000 ... 1 000 000)."""
    if args.cpu_rows > 0:
        return args.cpu_rows
    rows = 120.0 / max(1, fits) / (CPU_SECONDS_PER_ROW * args.features)
    return int(min(1_000_000, max(50_000, rows)) // 1000 * 1000)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F = args.features
    rows = reference_rows(args, args.steps + min(args.warmup, 1))
    import oracle
    X, y = gen_problem_numpy(rows, F)
    for _ in range(min(args.warmup, 1)):     # numpy has nothing to warm up beyond the first touch of the buffers
        oracle.fit_l2(X, y, max_depth=args.depth)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.fit_l2(X, y, max_depth=args.depth)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows * F / dt
    sample = (f"{rows} x {F} row sample of the same generator (sized so that {args.steps} fits take about two "
              f"minutes), whole depth-{args.depth} fit per step, 1 of {os.cpu_count()} host cores (the reference is "
              f"single-threaded numpy)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"L2 regression-tree fit, max_depth={args.depth}, {args.rows} x {F} f64 "
                               f"(BASELINE.json configs[3]); CPU arm on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from np_decision_tree_b200 import _lib
    from np_decision_tree_b200.sharded import fit_sharded, shard_columns
    from np_decision_tree_b200.decision_tree import one

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N, F, D = args.rows, args.features, args.depth
    lo, hi = shard_columns(F, world, rank)
    ctx = _lib.default_context(local)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    X, y, last = gen_problem_torch(N, F, lo, hi, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return fit_sharded(X, y, lo, F, last_col=last, max_depth=D, ctx=ctx)

    for _ in range(max(args.warmup, 3)):
        tree = step()
    ctx.profile_enable(True)
    ctx.profile_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tree = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = ctx.profile()
    launches = ctx.launch_count()
    ctx.profile_enable(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = N * F / (ms_per_step * 1e-3)
    tree_sha, value_sha = tree_digest(tree)

    # ---- inference of that tree, row-sharded, device-resident (BASELINE.json configs[4] at 8 GPUs) ----
    inference = None
    if not args.no_inference:
        inference = measure_inference(args, torch, dist, ctx, tree, world, rank, dev, barrier)

    # ---- e2e: host (pinned) buffers through the public API, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e(args, torch, dist, one, fit_sharded, ctx, X, y, last, lo, hi, world, rank, dev, barrier)
        except Exception as ex:  # e.g. not enough host memory to pin the matrix
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak()
        kernels = {}
        for name, p in prof.items():
            per_step_ms = p["ms"] / args.steps
            ach = p["alg_bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 and p["alg_bytes"] > 0 else None
            # bytes the launches of this class really moved in one step (ncu dram__bytes, 1-GPU capture of this
            # workload; scaled by this rank's share of the columns)
            traffic = ncu_traffic(name)
            tr = None
            if traffic and p["ms"] > 0:
                tr = traffic * ((hi - lo) / F) / (per_step_ms * 1e-3) / 1e9
            kernels[name] = {"ms_per_step": round(per_step_ms, 4), "launches_per_step": p["launches"] / args.steps,
                             "share_of_step": round(per_step_ms / ms_per_step, 4),
                             "achieved_gbs": None if ach is None else round(ach, 1),
                             "frac_alg": None if ach is None else round(ach / peak, 4),
                             "traffic_gbs": None if tr is None else round(tr, 1),
                             "frac_traffic": None if tr is None else round(tr / peak, 4)}
        timed = {k: v for k, v in prof.items() if v["alg_bytes"] > 0}
        dom = max(timed, key=lambda k: timed[k]["ms"]) if timed else None
        roofline = None
        if dom:
            p = prof[dom]
            ach = p["alg_bytes"] / (p["ms"] * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "traffic": ncu_traffic(dom, per_launch=True), "peak_source": peak_src,
                        "alg_bytes_per_launch": p["alg_bytes"] / p["launches"],
                        "avg_launch_ms": p["ms"] / p["launches"]}
            notes = {
                "partition": "algorithmic bytes = SURVEY 8d's 9 B per (row, feature) (4 in + 1 side flag + 4 out); by design "
                             "the kernel also moves the 8-byte y payload in and out (24 B per element, DESIGN.md section 3), "
                             "which is what lets the split search stream instead of gather; see kernels.partition."
                             "frac_traffic for the fraction of peak on the bytes it really moves",
                "l2_scan_search": "algorithmic bytes = SURVEY 8d's 12 B per (row, feature); the streaming kernel reads 8 B "
                                  "(frac_traffic)",
                "sort_radix_pass": "4-byte keys + 4-byte row ids, 4 passes of 8-bit digits",
            }
            if dom in notes:
                roofline["note"] = notes[dom]
        cpu = None
        if world == 1 and not args.no_cpu:
            rows = reference_rows(args, 6)      # about 20 s of CPU work
            secs = cpu_fit_seconds(rows, F, D)
            cpu = {"value": rows * F / secs, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{rows} x {F} row sample of the same recipe, whole depth-{D} fit, "
                             f"{secs:.1f} s on 1 of {os.cpu_count()} host cores (the reference is single-threaded)"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"L2 regression-tree fit, max_depth={D}, {N} x {F} f64 "
                                   "(BASELINE.json configs[3]), columns sharded over the GPUs",
                       "rows": N, "features": F, "max_depth": D, "nodes": tree.max_node_ + 1,
                       "l2_flush": "inputs (X %.1f GB per GPU) exceed the 126 MB L2" % (X.numel() * 8 / 1e9)},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
            "inference": inference, "tree_sha": tree_sha, "value_sha": value_sha,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(kernel, per_launch=False):
    """dram bytes (read + write) of the launches of one kernel class in ONE step of the 1-GPU C4 fit, from the
    committed ncu capture (profiles/ncu_traffic.json: {"class": {"bytes_per_step": .., "launches_per_step": ..}})."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p)).get(kernel)
        if not isinstance(rec, dict):
            return None
        return rec["bytes_per_step"] / rec["launches_per_step"] if per_launch else rec["bytes_per_step"]
    except Exception:
        return None


def tree_digest(tree):
    """(structure digest, value digest).  The first is SHA-256 over children_left, children_right, feature and
    threshold up to max_node_ -- bit-exact quantities that must not depend on the number of GPUs.  Node values are
    sums that a look-back scan may group differently from run to run (last bits), so they are hashed separately,
    rounded to 9 significant digits (the north-star tolerance is 1e-9)."""
    import hashlib
    n = tree.max_node_ + 1
    t = tree.tree_
    h = hashlib.sha256()
    for name in ("children_left", "children_right", "feature", "threshold"):
        h.update(np.ascontiguousarray(getattr(t, name)[:n]).tobytes())
    v = np.asarray(t.value[:n], dtype=np.float64)
    hv = hashlib.sha256(np.array([float("%.8e" % x) for x in v]).tobytes())
    return h.hexdigest(), hv.hexdigest()


def measure_inference(args, torch, dist, ctx, tree, world, rank, dev, barrier):
    """utils.inference over args.inference_rows device-resident rows per GPU (weak scaling over rows)."""
    from np_decision_tree_b200.decision_tree import utils
    rows, F = args.inference_rows, args.features
    free_b, _ = torch.cuda.mem_get_info(dev)
    while rows * F * 8 > 0.6 * free_b and rows > 1_000_000:
        rows //= 2
    g = torch.Generator(device=dev)
    g.manual_seed(SEED * 7 + 1000 + rank)
    Xq = torch.randn((rows, F), dtype=torch.float64, device=dev, generator=g)
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    for _ in range(3):
        out = utils.inference(Xq, tree)
    ctx.profile_reset()
    barrier()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = utils.inference(Xq, tree)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    kern = ctx.profile().get("predict", {"ms": 0.0, "alg_bytes": 0.0, "launches": 0})
    ctx.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    peak, _ = measured_peak()
    kern_ms = kern["ms"] / max(1, kern["launches"])
    alg = kern["alg_bytes"] / max(1, kern["launches"])
    res = {"rows_per_gpu": rows, "rows": rows * world, "features": F, "ms_per_call": ms,
           "value": rows * world / (ms * 1e-3), "unit": "rows/s",
           "rows_features_per_s": rows * world * F / (ms * 1e-3),
           "kernel_ms": kern_ms, "alg_bytes_per_row": alg / rows if rows else None,
           "achieved_gbs": alg / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None,
           "frac": alg / (kern_ms * 1e-3) / 1e9 / peak if kern_ms > 0 else None,
           "accounting": "32 B x mean path length + 8 B per row (SURVEY 8d), this rank's kernel; ms_per_call is the "
                         "whole utils.inference call (tree upload + kernel), max over ranks",
           "checksum": float(out.sum().item())}
    del Xq, out
    torch.cuda.empty_cache()
    return res


def bind_to_gpu_numa_node(dev_index):
    """Run this rank's host thread on the CPUs local to its GPU, so that the pinned host buffers it allocates
    next (first touch) sit on the NUMA node the GPU's PCIe root hangs off.  One process per GPU => every rank
    streams from its own socket instead of all ranks sharing one.  Best effort: returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[dev_index]) if visible and visible.replace(",", "").isdigit() else dev_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"host thread bound to the CPUs local to GPU {phys} ({len(os.sched_getaffinity(0))} cpus)"
    except Exception as ex:   # no NVML, no permission, single-node box: keep the default placement
        return f"default placement ({type(ex).__name__})"


def measure_e2e(args, torch, dist, one, fit_sharded, ctx, X, y, last, lo, hi, world, rank, dev, barrier):
    N, F, D = args.rows, args.features, args.depth
    Fl = hi - lo
    placement = bind_to_gpu_numa_node(dev.index if dev.index is not None else 0)
    # host copies of this rank's inputs in pinned memory (outside the timed region)
    hX = torch.empty((N, Fl), dtype=torch.float64, pin_memory=True)
    hy = torch.empty(N, dtype=torch.float64, pin_memory=True)
    hX.copy_(X)
    hy.copy_(y)
    hlast = None
    if last is not None:
        hlast = torch.empty(N, dtype=torch.float64, pin_memory=True)
        hlast.copy_(last)
    torch.cuda.synchronize()
    h2d = (N * Fl + N + (N if last is not None else 0)) * 8
    tree_bytes = (2 ** (D + 1) - 1) * 40
    api = ("decision_tree.one.build_regression_tree(numpy X, numpy y)" if world == 1
           else "sharded.fit_sharded(host X slice, host y)")
    del X
    torch.cuda.empty_cache()

    def timed(step):
        step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.e2e_steps):
            step()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) / args.e2e_steps
        t = torch.tensor([max(e0.elapsed_time(e1) / args.e2e_steps * 1e-3, wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # (1) pinned host buffers
    if world == 1:
        Xn, yn = hX.numpy(), hy.numpy()
        pinned = timed(lambda: one.build_regression_tree(Xn, yn, max_depth=D))   # the call a user of the reference makes
    else:
        # host tensors in: the library pipelines the H2D column batches with the sort
        pinned = timed(lambda: fit_sharded(hX, hy, lo, F, last_col=hlast, max_depth=D, ctx=ctx))
    # (2) ordinary pageable memory, what np.random / np.load hand a user of the reference
    pageable = None
    try:
        pX = torch.empty((N, Fl), dtype=torch.float64)
        pX.copy_(hX)
        py = hy.clone()
        plast = hlast.clone() if hlast is not None else None
        if world == 1:
            pXn, pyn = pX.numpy(), py.numpy()
            pageable = timed(lambda: one.build_regression_tree(pXn, pyn, max_depth=D))
        else:
            pageable = timed(lambda: fit_sharded(pX, py, lo, F, last_col=plast, max_depth=D, ctx=ctx))
    except MemoryError:
        pass
    return {"value": N * F / pinned, "unit": UNIT, "ms_per_step": pinned * 1e3, "steps": args.e2e_steps,
            "pinned_ms_per_step": pinned * 1e3,
            "pageable_ms_per_step": None if pageable is None else pageable * 1e3,
            "pageable_value": None if pageable is None else N * F / pageable,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": tree_bytes, "api": api,
            "note": "value = pinned host buffers (the contract's definition); pageable_* = the same call on an ordinary "
                    "numpy array, staged through the library's pinned bounce buffers",
            "host_placement": placement}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
