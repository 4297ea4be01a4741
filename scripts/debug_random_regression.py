import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle
from util import small_problem
from np_decision_tree_b200.decision_tree import strategy, tree_builder
X, y = small_problem(71, 30000, 16, None, informative=4)
kw = dict(max_depth=10, max_feature=16, min_improvement=1e-6)
np.random.seed(71); want = oracle.fit_l2_random(X, y, **kw)
np.random.seed(71); got = tree_builder.build_regression_tree(X, y, split_method=strategy.random_split_v2, **kw)
w, g = want.tree_, got.tree_
print("nodes", want.node_count, got.max_node_ + 1)
for i in range(min(want.node_count, got.max_node_ + 1)):
    a = (w.children_left[i], w.children_right[i], w.feature[i], w.threshold[i]); b = (g.children_left[i], g.children_right[i], g.feature[i], g.threshold[i])
    if a != b:
        print("first diff at node", i, "want", a, "got", b, "values", w.value[i], g.value[i]); break
# rows reaching that node in the oracle tree
leaf = oracle.predict_leaf(X, want)
def rows_of(tree, node):
    t = tree.tree_; idx = np.arange(X.shape[0]); path=[]
    # walk from root collecting rows that pass through `node`
    cur = np.zeros(X.shape[0], dtype=np.int64)
    active = np.ones(X.shape[0], dtype=bool); through = np.zeros(X.shape[0], dtype=bool)
    for _ in range(12):
        through |= active & (cur == node)
        f = t.feature[cur]; isleaf = t.children_left[cur] == -1
        go_left = X[idx, np.where(isleaf, 0, f)] <= t.threshold[cur]
        nxt = np.where(go_left, t.children_left[cur], t.children_right[cur])
        active &= ~isleaf
        cur = np.where(active, nxt, cur)
    return np.flatnonzero(through)
r = rows_of(want, i)
print("rows in node (oracle tree):", r.size, "y unique", np.unique(y[r]).size, "y range", y[r].min(), y[r].max())
