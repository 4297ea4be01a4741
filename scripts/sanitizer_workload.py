"""Small shapes that reach every hot kernel, for `compute-sanitizer --tool memcheck|racecheck`:
radix passes (32- and 64-bit paths), fix-up, streaming split search (register-staged and bulk-copy variants),
exact search, partition (plain and relabelling), relabel, finalize, predict, the gini lane kernel, the L1 rank /
chunked median walks / cost scans, cummedian, the row-list kernels.

    compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_decision_tree_b200.decision_tree import one, strategy, strategy_l1, tree_builder, utils  # noqa: E402
from np_decision_tree_b200.decision_tree.np_stream_median import cummedian  # noqa: E402

rs = np.random.RandomState(0)
X = rs.standard_normal((60000, 6))
y = X[:, 1] * 30 - X[:, 4] * 12 + rs.standard_normal(60000)
os.environ["B200DT_RELABEL_ROWS"] = "5000"
t = one.build_regression_tree(X, y, max_depth=8)
print("L2 fit (relabelled):", t.max_node_ + 1, "nodes")
os.environ["B200DT_SCAN_BULK"] = "1"
t2 = one.build_regression_tree(X, y, max_depth=8)
assert np.array_equal(t.tree_.feature, t2.tree_.feature) and np.array_equal(t.tree_.threshold, t2.tree_.threshold)
del os.environ["B200DT_SCAN_BULK"]
os.environ["B200DT_SORT64"] = "1"
t3 = one.build_regression_tree(np.round(X * 4) / 4, y, max_depth=5)
del os.environ["B200DT_SORT64"]
print("L2 fit (64-bit sort):", t3.max_node_ + 1, "nodes")
p = utils.inference(X, t)
print("inference:", float(np.mean((p - y) ** 2)))
labels = np.argmax(X[:, :4] + 0.3 * rs.standard_normal((60000, 4)), axis=1)
ct = tree_builder.build_classification_tree(X[:30000], labels[:30000], max_depth=6, max_feature=6, min_improvement=1e-7,
                                            split_method=strategy.greedy_classification, max_classes=4)
print("gini fit:", ct.max_node_ + 1, "nodes")
np.random.seed(1)
rt = tree_builder.build_tree(X[:20000], y[:20000], max_depth=4, split_method=strategy.random_split_v2)
print("random-threshold fit:", rt.max_node_ + 1, "nodes")
lt = strategy_l1.build_regression_tree_v2(X[:40000, :3], y[:40000], max_depth=2)
print("L1 fit (chunked walks):", lt.max_node_ + 1, "nodes")
print("cummedian:", cummedian(y[:5000])[-1])
print("sanitizer workload done")
