"""Generate tests/golden/reference_classes.npz by running the UNMODIFIED reference's
classification builder (decision_tree/tree_builder.py + decision_tree/strategy.py).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_classes.py

Feature values are continuous (distinct per column): the reference sorts each node with numpy's
default unstable argsort, so only inputs without repeated feature values have a defined result.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from decision_tree import tree_builder, strategy, utils  # noqa: E402  (the reference)
from make_golden import tree_dict  # noqa: E402


def class_problem(seed, n, F, C, noise=0.5):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, F))
    W = rs.standard_normal((F, C)) * (rs.uniform(size=(F, 1)) < 0.6)
    labels = np.argmax(X @ W + noise * rs.standard_normal((n, C)), axis=1).astype(np.int64)
    return X, labels


METHODS = {"gini": strategy.greedy_classification, "p_at_k": strategy.greedy_classification_p_at_k}


def main():
    out = {}
    cases = [("g0", "gini", 70, 300, 6, 3, 4, {}), ("g1", "gini", 71, 97, 4, 2, 6, {}),
             ("g2", "gini", 72, 1000, 10, 10, 8, dict(min_improvement=1e-7)),
             ("g3", "gini", 73, 513, 3, 23, 9, dict(min_improvement=1e-7)),
             ("g4", "gini", 74, 5, 2, 2, 2, dict(min_improvement=0.0)),
             ("g5", "gini", 75, 3000, 8, 5, 10, dict(min_improvement=1e-7)),
             ("gk", "gini", 76, 400, 5, 4, 6, dict(min_improvement=0.02, min_sample_leaf=10)),
             ("p0", "p_at_k", 80, 300, 6, 3, 4, {}), ("p1", "p_at_k", 81, 1000, 10, 10, 8, dict(min_improvement=1e-7)),
             ("p2", "p_at_k", 82, 513, 3, 23, 9, dict(min_improvement=1e-7)),
             ("pk", "p_at_k", 83, 400, 5, 4, 6, dict(min_improvement=0.02, min_sample_leaf=10))]
    for name, crit, seed, n, F, C, depth, kw in cases:
        X, labels = class_problem(seed, n, F, C)
        t = tree_builder.build_classification_tree(X, labels, max_depth=depth, max_feature=F,
                                                   split_method=METHODS[crit], max_classes=C, **kw)
        out.update({f"cl_{name}_X": X, f"cl_{name}_labels": labels, f"cl_{name}_depth": np.int64(depth),
                    f"cl_{name}_classes": np.int64(C), f"cl_{name}_criterion": np.array(crit),
                    f"cl_{name}_min_improvement": np.float64(kw.get("min_improvement", 0.01)),
                    f"cl_{name}_min_sample_leaf": np.int64(kw.get("min_sample_leaf", 1))})
        out.update(tree_dict(f"cl_{name}_", t))
        out[f"cl_{name}_pred"] = utils.inference(X, t)
        # the root's single-node search (strategy.py:272-283 / :313-323)
        col, thr, imp, _ = METHODS[crit](X, np.eye(C)[labels])
        out[f"cl_{name}_root"] = np.array([col, thr, imp], dtype=np.float64)
    # ---- random-threshold splitters (strategy.py:149-172), seeded global generator ---------------------
    RANDOM = {"gini": strategy.random_classify, "p_at_k": strategy.random_classify_p_at_k}
    rcases = [("r0", "gini", 90, 300, 6, 3, 5, 6, {}), ("r1", "gini", 91, 1000, 10, 10, 9, 10, dict(min_improvement=1e-7)),
              ("r2", "gini", 92, 800, 12, 4, 8, 5, dict(min_improvement=1e-7)),          # F > max_feature: permutation
              ("r3", "p_at_k", 93, 600, 8, 5, 8, 8, dict(min_improvement=1e-7)),
              ("r4", "p_at_k", 94, 500, 9, 3, 7, 4, dict(min_improvement=0.001, min_sample_leaf=5)),
              ("r5", "gini", 95, 200, 4, 2, 6, 4, dict(min_improvement=-1.0))]            # + a constant column (nan score)
    import warnings
    for name, crit, seed, n, F, C, depth, max_feature, kw in rcases:
        X, labels = class_problem(seed, n, F, C)
        if name == "r5":
            X[:, 2] = 1.5
        np.random.seed(1000 + seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = tree_builder.build_classification_tree(X, labels, max_depth=depth, max_feature=max_feature,
                                                       split_method=RANDOM[crit], max_classes=C, **kw)
        out.update({f"rc_{name}_X": X, f"rc_{name}_labels": labels, f"rc_{name}_depth": np.int64(depth),
                    f"rc_{name}_classes": np.int64(C), f"rc_{name}_criterion": np.array(crit),
                    f"rc_{name}_max_feature": np.int64(max_feature), f"rc_{name}_seed": np.int64(1000 + seed),
                    f"rc_{name}_min_improvement": np.float64(kw.get("min_improvement", 0.01)),
                    f"rc_{name}_min_sample_leaf": np.int64(kw.get("min_sample_leaf", 1))})
        out.update(tree_dict(f"rc_{name}_", t))
        out[f"rc_{name}_pred"] = utils.inference(X, t)
        np.random.seed(77)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            col, thr, imp, _ = RANDOM[crit](X, np.eye(C)[labels])
        out[f"rc_{name}_root"] = np.array([col, thr, imp], dtype=np.float64)
    # ---- regression with random thresholds (strategy.py:136-146 under tree_builder.build_tree) ---------------
    from make_golden import small_problem
    for name, seed, n, F, depth, max_feature, kw in [("x0", 100, 400, 6, 5, 6, dict(min_improvement=1e-6)),
                                                     ("x1", 101, 2000, 12, 9, 5, dict(min_improvement=1e-6)),
                                                     ("x2", 102, 300, 4, 6, 4, dict(min_improvement=0.5, min_sample_leaf=8))]:
        X, y = small_problem(seed, n, F)
        np.random.seed(2000 + seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = tree_builder.build_tree(X, y, max_depth=depth, max_feature=max_feature,
                                        split_method=strategy.random_split_v2, **kw)
        out.update({f"rx_{name}_X": X, f"rx_{name}_y": y, f"rx_{name}_depth": np.int64(depth),
                    f"rx_{name}_max_feature": np.int64(max_feature), f"rx_{name}_seed": np.int64(2000 + seed),
                    f"rx_{name}_min_improvement": np.float64(kw.get("min_improvement", 0.01)),
                    f"rx_{name}_min_sample_leaf": np.int64(kw.get("min_sample_leaf", 1))})
        out.update(tree_dict(f"rx_{name}_", t))
        out[f"rx_{name}_pred"] = utils.inference(X, t)
    np.savez_compressed(os.path.join(HERE, "reference_classes.npz"), **out)
    print("reference_classes.npz", os.path.getsize(os.path.join(HERE, "reference_classes.npz")), "bytes",
          {k[:-9]: int(out[k]) + 1 for k in out if k.endswith("_max_node")})


if __name__ == "__main__":
    main()
