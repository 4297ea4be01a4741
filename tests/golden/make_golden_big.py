"""Generate tests/golden/reference_big.npz by running the UNMODIFIED reference at the sizes
BASELINE.json names (build container only: needs /root/reference, ~12 GB RAM, ~6 minutes):

    python tests/golden/make_golden_big.py

* C3  -- strategy_l1.build_regression_tree_v2(max_depth=1) on
         make_regression(n_samples=1_000_000, n_features=100, random_state=0)   (configs[2])
* C4s -- one.build_regression_tree(max_depth=10) on a 1 000 000 x 256 matrix of the C4 shape
         (configs[3] has 10 M rows, which the reference cannot hold in this container's RAM),
         then utils.inference / inference_leaf over the same rows (configs[4]'s tree depth).

Only the trees, a few scalars and SHA-256 digests are stored; the tests regenerate the inputs
with `big_inputs.py` (numpy RandomState streams: frozen) and check the digests first.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.build_ref import build as build_ref, REF  # noqa: E402

so = build_ref()
sys.path.insert(0, os.path.dirname(so))
sys.path.insert(0, REF)

from decision_tree import one, strategy_l1, utils  # noqa: E402  (the reference)
from big_inputs import c3_inputs, c4_shape_inputs, sha  # noqa: E402
from make_golden import tree_dict  # noqa: E402


def main():
    out = {}
    t0 = time.time()
    X, y = c4_shape_inputs(1_000_000, 256)
    out["c4s_X_sha"] = np.array(sha(X))
    out["c4s_y_sha"] = np.array(sha(y))
    t = one.build_regression_tree(X, y, max_depth=10)
    print("C4-shape reference fit: %.1f s, %d nodes" % (time.time() - t0, t.max_node_ + 1), flush=True)
    out.update(tree_dict("c4s_", t))
    p = utils.inference(X, t)
    out["c4s_pred_sha"] = np.array(sha(p))
    out["c4s_pred_head"] = p[:64]
    out["c4s_leafid_sha"] = np.array(sha(utils.inference_leaf(X, t).astype(np.int64)))
    out["c4s_mse"] = np.float64(np.mean((p - y) ** 2))
    del X, y, t, p

    t0 = time.time()
    X, y = c3_inputs()
    out["c3_X_sha"] = np.array(sha(X))
    out["c3_y_sha"] = np.array(sha(y))
    t = strategy_l1.build_regression_tree_v2(X, y, max_depth=1)
    print("C3 reference fit: %.1f s, root feature %d threshold %r" %
          (time.time() - t0, t.tree_.feature[0], t.tree_.threshold[0]), flush=True)
    out.update(tree_dict("c3_", t))
    out["c3_mae"] = np.float64(np.mean(np.abs(utils.inference(X, t) - y)))
    fn = os.path.join(HERE, "reference_big.npz")
    np.savez_compressed(fn, **out)
    print(fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
