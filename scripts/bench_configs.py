"""Side measurements of the other BASELINE.json configs (bench.py carries only the headline, C4).

    python scripts/bench_configs.py [c2] [c3] [c5] [cls] [rreg] [--cpu]

C2: make_regression 10 000 x 100, L2 depth 10 fit + inference (README.md:60-64: 390 ms reference)
C3: L1 build_regression_tree_v2 depth 1, 1 M x 100 (cummedian path; BASELINE.md: 138 s on one core)
C5: inference of a depth-10 tree, N x 256 device-resident rows on one GPU (row-sharded across GPUs
    with no exchange, so N per GPU is what matters; default 12.5 M = the 8-GPU share of 100 M)
cls: exhaustive gini (CLS_CRIT=p_at_k: precision-at-1) classification fit, depth 10, CLS_ROWS x CLS_FEATURES
    device-resident rows, CLS_CLASSES classes (defaults 10 M x 256, 8)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from np_decision_tree_b200 import _lib  # noqa: E402
from np_decision_tree_b200.decision_tree import one, strategy_l1, utils  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c3", "c5"]
with_cpu = "--cpu" in sys.argv
out = {}


def best_of(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


if "c2" in args:
    from sklearn.datasets import make_regression
    X, y = make_regression(n_samples=10000, n_features=100, random_state=0)
    one.build_regression_tree(X, y, max_depth=10)
    t_fit, tree = best_of(lambda: one.build_regression_tree(X, y, max_depth=10))
    t_inf, p = best_of(lambda: utils.inference(X, tree))
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    t_fit_d, _ = best_of(lambda: one.build_regression_tree(Xd, yd, max_depth=10))
    out["c2"] = {"fit_ms_numpy_in": t_fit * 1e3, "fit_ms_device_in": t_fit_d * 1e3, "inference_ms_numpy_in": t_inf * 1e3,
                 "nodes": tree.max_node_ + 1, "train_mse": float(np.mean((p - y) ** 2)),
                 "rows_features_per_s": 1e6 / t_fit_d}
    if with_cpu:
        import oracle
        t0 = time.perf_counter()
        o = oracle.fit_l2(X, y, max_depth=10)
        out["c2"]["cpu_port_fit_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        oracle.predict(X, o)
        out["c2"]["cpu_port_inference_ms"] = (time.perf_counter() - t0) * 1e3

if "c3" in args:
    rs = np.random.RandomState(0)
    N, F = 1_000_000, 100
    X = rs.standard_normal((N, F))
    coef = np.zeros(F)
    coef[rs.permutation(F)[:10]] = 100 * rs.uniform(size=10)
    y = X @ coef
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    strategy_l1.build_regression_tree_v2(Xd, yd, max_depth=1)
    ctx = _lib.default_context(0)
    ctx.profile_enable(True)
    ctx.profile_reset()
    t_fit, tree = best_of(lambda: strategy_l1.build_regression_tree_v2(Xd, yd, max_depth=1), n=3)
    prof = ctx.profile()
    ctx.profile_enable(False)
    out["c3"] = {"fit_ms_device_in": t_fit * 1e3, "rows_features_per_s": N * F / t_fit,
                 "root_feature": int(tree.tree_.feature[0]), "root_threshold": float(tree.tree_.threshold[0]),
                 "kernels_ms_per_fit": {k: v["ms"] / 3 for k, v in prof.items()}}
    del Xd, yd

if "c5" in args:
    N = int(os.environ.get("C5_ROWS", 12_500_000))
    F = 256
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    # a depth-10 tree fitted on a 1 M-row sample of the same distribution
    Xs = torch.randn((1_000_000, F), dtype=torch.float64, device="cuda", generator=g)
    cols = torch.arange(0, 250, 25, device="cuda")
    ys = (Xs[:, cols] * torch.linspace(10, 100, 10, dtype=torch.float64, device="cuda")).sum(1)
    tree = one.build_regression_tree(Xs, ys, max_depth=10)
    del Xs, ys
    X = torch.randn((N, F), dtype=torch.float64, device="cuda", generator=g)
    utils.inference(X, tree)
    ctx = _lib.default_context(0)
    ctx.profile_enable(True)
    ctx.profile_reset()
    t_inf, p = best_of(lambda: utils.inference(X, tree), n=5)
    prof = ctx.profile()
    ctx.profile_enable(False)
    k = prof["predict"]
    leaf = utils.inference_leaf(X[:100000], tree)
    out["c5"] = {"rows": N, "features": F, "nodes": tree.max_node_ + 1, "predict_ms": t_inf * 1e3,
                 "kernel_ms": k["ms"] / k["launches"], "rows_per_s": N / t_inf, "rows_features_per_s": N * F / t_inf,
                 "alg_bytes_per_row": k["alg_bytes"] / k["launches"] / N,
                 "alg_gbs": k["alg_bytes"] / k["ms"] / 1e6, "full_row_stream_gbs": N * F * 8 / (k["ms"] / k["launches"]) / 1e6}
    if with_cpu:
        import oracle
        Xc = X[:2_000_000].cpu().numpy()
        t0 = time.perf_counter()
        oracle.predict(Xc, tree)
        dt = time.perf_counter() - t0
        out["c5"]["cpu_port_rows_per_s_on_2M_rows"] = 2_000_000 / dt

if "cls" in args:
    # classification: exhaustive gini fit, device-resident, depth 10
    from np_decision_tree_b200.decision_tree import strategy, tree_builder
    N = int(os.environ.get("CLS_ROWS", 10_000_000))
    F = int(os.environ.get("CLS_FEATURES", 256))
    C = int(os.environ.get("CLS_CLASSES", 8))
    crit = {"p_at_k": strategy.greedy_classification_p_at_k, "random": strategy.random_classify,
            "random_p_at_k": strategy.random_classify_p_at_k}.get(os.environ.get("CLS_CRIT"), strategy.greedy_classification)
    is_random = crit in (strategy.random_classify, strategy.random_classify_p_at_k)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    X = torch.empty((N, F), dtype=torch.float64, device="cuda")
    for c0 in range(0, F, 32):
        X[:, c0:c0 + 32] = torch.randn((N, min(32, F - c0)), dtype=torch.float64, device="cuda", generator=g)
    score = X[:, :C] + 0.5 * torch.randn((N, C), dtype=torch.float64, device="cuda", generator=g)
    labels = score.argmax(dim=1).to(torch.int32)
    del score
    kw = dict(max_depth=10, max_feature=F, min_improvement=1e-7, split_method=crit, max_classes=C)
    tree_builder.build_classification_tree(X, labels, **kw)
    ctx = _lib.default_context(0)
    ctx.profile_enable(True)
    ctx.profile_reset()
    np.random.seed(0)
    t_fit, tree = best_of(lambda: tree_builder.build_classification_tree(X, labels, **kw), n=3)
    prof = ctx.profile()
    ctx.profile_enable(False)
    pred = utils.inference(X[:1_000_000], tree)
    out["cls"] = {"rows": N, "features": F, "classes": C, "criterion": crit.__name__, "fit_ms_device_in": t_fit * 1e3,
                  "rows_features_per_s": N * F / t_fit, "nodes": tree.max_node_ + 1,
                  "train_accuracy_1M": float((torch.as_tensor(pred, device="cuda") == labels[:1_000_000]).double().mean()),
                  "kernels_ms_per_fit": {k: round(v["ms"] / 3, 3) for k, v in prof.items() if v["launches"]},
                  "kernels_alg_gbs": {k: round(v["alg_bytes"] / v["ms"] / 1e6, 1) for k, v in prof.items()
                                      if v["launches"] and v["alg_bytes"] > 0}}
    if with_cpu:
        import oracle
        n_s, f_s = 100_000, min(F, 32)
        Xc = X[:n_s, :f_s].cpu().numpy()
        lc = labels[:n_s].cpu().numpy().astype(np.int64)
        t0 = time.perf_counter()
        cname = "p_at_k" if "p_at_k" in crit.__name__ else "gini"
        np.random.seed(0)
        if is_random:
            o = oracle.fit_classes_random(Xc, lc, C, max_depth=10, max_feature=f_s, min_improvement=1e-7, criterion=cname)
        else:
            o = oracle.fit_classes(Xc, lc, C, max_depth=10, max_feature=f_s, min_improvement=1e-7, criterion=cname)
        dt = time.perf_counter() - t0
        np.random.seed(0)
        g = tree_builder.build_classification_tree(Xc, lc, max_depth=10, max_feature=f_s, min_improvement=1e-7,
                                                   split_method=crit, max_classes=C)
        same = all(np.array_equal(getattr(g.tree_, k), getattr(o.tree_, k))
                   for k in ("children_left", "children_right", "feature", "threshold"))
        out["cls"]["cpu_port"] = {"sample": f"{n_s} x {f_s}", "fit_s": dt, "rows_features_per_s": n_s * f_s / dt,
                                  "nodes": o.node_count, "gpu_tree_identical": bool(same)}
    del X, labels

if "rreg" in args:
    # regression ExtraTree (strategy.random_split_v2 under tree_builder.build_tree), depth 10, device-resident
    from np_decision_tree_b200.decision_tree import strategy, tree_builder
    N = int(os.environ.get("CLS_ROWS", 10_000_000))
    F = int(os.environ.get("CLS_FEATURES", 256))
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    X = torch.empty((N, F), dtype=torch.float64, device="cuda")
    for c0 in range(0, F, 32):
        X[:, c0:c0 + 32] = torch.randn((N, min(32, F - c0)), dtype=torch.float64, device="cuda", generator=g)
    cols = torch.arange(0, 250, 25, device="cuda")
    y = (X[:, cols] * torch.linspace(10, 100, 10, dtype=torch.float64, device="cuda")).sum(1)
    kw = dict(max_depth=10, max_feature=F, min_improvement=1e-6, split_method=strategy.random_split_v2)
    np.random.seed(0)
    tree_builder.build_tree(X, y, **kw)
    ctx = _lib.default_context(0)
    ctx.profile_enable(True)
    ctx.profile_reset()
    np.random.seed(0)
    t_fit, tree = best_of(lambda: tree_builder.build_tree(X, y, **kw), n=3)
    prof = ctx.profile()
    ctx.profile_enable(False)
    pred = utils.inference(X[:1_000_000], tree)
    out["rreg"] = {"rows": N, "features": F, "fit_ms_device_in": t_fit * 1e3, "rows_features_per_s": N * F / t_fit,
                   "nodes": tree.max_node_ + 1,
                   "train_mse_1M": float(((torch.as_tensor(pred, device="cuda") - y[:1_000_000]) ** 2).mean()),
                   "y_variance": float(y[:1_000_000].var()),
                   "kernels_ms_per_fit": {k: round(v["ms"] / 3, 3) for k, v in prof.items() if v["launches"]}}
    if with_cpu:
        import oracle
        n_s, f_s = 100_000, min(F, 32)
        Xc, yc = X[:n_s, :f_s].cpu().numpy(), y[:n_s].cpu().numpy()
        np.random.seed(0)
        t0 = time.perf_counter()
        o = oracle.fit_l2_random(Xc, yc, max_depth=10, max_feature=f_s, min_improvement=1e-6, min_sample_leaf=50)
        dt = time.perf_counter() - t0
        np.random.seed(0)
        gt = tree_builder.build_tree(Xc, yc, max_depth=10, max_feature=f_s, min_improvement=1e-6, min_sample_leaf=50,
                                     split_method=strategy.random_split_v2)
        same = all(np.array_equal(getattr(gt.tree_, k), getattr(o.tree_, k))
                   for k in ("children_left", "children_right", "feature", "threshold"))
        out["rreg"]["cpu_port"] = {"sample": f"{n_s} x {f_s}", "fit_s": dt, "rows_features_per_s": n_s * f_s / dt,
                                   "nodes": o.node_count, "gpu_tree_identical": bool(same)}
    del X, y

print(json.dumps(out, indent=1))
