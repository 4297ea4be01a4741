"""Generate tests/golden/reference_builders.npz by running the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_builders.py

Fixtures for the parts of the drop-in boundary that are plain functions around the kernels:
``one.best_variance_improvements`` / ``_v0`` / ``_v2`` on caller-supplied cumsums, ``StreamMedian``,
``strategy.random_split`` / ``greedy_split`` and the node-at-a-time builders of ``tree_builder`` that draw from
numpy's global generator (seeded here), column sub-sampling under the exhaustive gini search, and an L1 tree
with ``leaf_from_data=np.mean``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.build_ref import build as build_ref, REF  # noqa: E402

so = build_ref()
sys.path.insert(0, os.path.dirname(so))
sys.path.insert(0, REF)

from decision_tree import one, strategy, strategy_l1, tree_builder, utils  # noqa: E402  (the reference)
from decision_tree.stream_median import StreamMedian  # noqa: E402
from make_golden import small_problem, tree_dict  # noqa: E402
from make_golden_classes import class_problem  # noqa: E402


def main():
    out = {}
    # ---- stand-alone cumsum scorers (one.py:56-93) -------------------------------------------------------
    for name, seed, n, F, yr in [("a", 21, 257, 9, 10.0), ("b", 22, 1000, 4, None), ("c", 23, 2, 3, None),
                                 ("d", 24, 40, 6, 50.0)]:
        X, y = small_problem(seed, n, F, yr)
        cums = np.cumsum(y[np.argsort(X, axis=0, kind="stable")], axis=0)
        out[f"bvi_{name}_cums"] = cums
        for tag, fn in (("v1", one.best_variance_improvements), ("v0", one.best_variance_improvements_v0),
                        ("v2", one.best_variance_improvements_v2)):
            k, f, g = fn(cums)
            out[f"bvi_{name}_{tag}"] = np.array([k, f], dtype=np.int64)
            out[f"bvi_{name}_{tag}_gain"] = np.float64(g)
    # ---- StreamMedian (stream_median.py:28-92) --------------------------------------------------------------
    rs = np.random.RandomState(31)
    seq = np.round(rs.standard_normal(301) * 3) / 2
    sm = StreamMedian()
    pairs, meds = [], []
    for x in seq:
        pairs.append(sm(x))
        meds.append(sm.median())
    out["sm_seq"], out["sm_pairs"], out["sm_median"] = seq, np.array(pairs), np.array(meds)
    # ---- strategy.random_split / greedy_split, single node ------------------------------------------------------
    X, y = small_problem(71, 500, 6)
    np.random.seed(7)
    col, thr, imp, _ = strategy.random_split(X, y)
    out.update({"rs_X": X, "rs_y": y, "rs_out": np.array([col, thr, imp], dtype=np.float64)})
    col, thr, imp, _ = strategy.greedy_split(X, y)
    out["gs_out"] = np.array([col, thr, imp], dtype=np.float64)
    # ---- node-at-a-time builders under a seeded global generator -----------------------------------------------
    np.random.seed(101)
    t = tree_builder.build_regression_tree(X, y, max_depth=4)                       # default: random_split
    out.update(tree_dict("rsplit_", t))
    np.random.seed(102)
    t = tree_builder.build_regression_tree(X, y, max_depth=5, max_feature=3, min_sample_leaf=5)   # + column sub-sampling
    out.update(tree_dict("rsplit_sub_", t))
    t = tree_builder.build_tree(X, y, max_depth=5)                                  # default: greedy_split
    out.update(tree_dict("gsplit_", t))
    Xc, labels = class_problem(84, 600, 8, 3)
    np.random.seed(103)
    t = tree_builder.build_classification_tree(Xc, labels, max_depth=5, max_feature=3, min_improvement=1e-7,
                                               split_method=strategy.greedy_classification, max_classes=3)
    out.update({"gsub_X": Xc, "gsub_labels": labels})
    out.update(tree_dict("gsub_", t))
    # ---- L1 tree with mean leaves (strategy_l1.py:143,173 leaf_from_data) ----------------------------------------
    Xl, yl = small_problem(56, 300, 5)
    t = strategy_l1.build_regression_tree_v2(Xl, yl, max_depth=3, leaf_from_data=np.mean)
    out.update({"l1mean_X": Xl, "l1mean_y": yl})
    out.update(tree_dict("l1mean_", t))
    fn = os.path.join(HERE, "reference_builders.npz")
    np.savez_compressed(fn, **out)
    print(fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
