"""Dev helper: fit on the GPU and with the oracle, report the first node that differs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from np_decision_tree_b200.decision_tree import one
from util import small_problem


def node_rows(tree, X, node_target):
    # rows reaching node_target
    t = tree.tree_
    rows = {0: np.arange(X.shape[0])}
    stack = [0]
    while stack:
        v = stack.pop()
        if v == node_target:
            return rows[v]
        if t.children_left[v] != -1:
            r = rows[v]
            m = X[r, t.feature[v]] <= t.threshold[v]
            rows[t.children_left[v]] = r[m]
            rows[t.children_right[v]] = r[~m]
            stack += [t.children_right[v], t.children_left[v]]
    return None


def depth_of(tree, node):
    t = tree.tree_
    d = {0: 0}
    for v in range(tree.max_node_ + 1):
        for c in (t.children_left[v], t.children_right[v]):
            if c != -1:
                d[c] = d[v] + 1
    return d.get(node)


def compare(X, y, depth, label):
    g = one.build_regression_tree(X, y, max_depth=depth)
    o = oracle.fit_l2(X, y, max_depth=depth, sort_kind="stable")
    print(f"[{label}] gpu nodes {g.max_node_+1} oracle nodes {o.max_node_+1}")
    n = min(g.max_node_, o.max_node_) + 1
    for v in range(n):
        a = (g.tree_.feature[v], g.tree_.threshold[v], g.tree_.children_left[v], g.tree_.children_right[v])
        b = (o.tree_.feature[v], o.tree_.threshold[v], o.tree_.children_left[v], o.tree_.children_right[v])
        if a != b:
            rows = node_rows(o, X, v)
            print(f"  first diff at node {v}: depth {depth_of(o, v)} rows {None if rows is None else rows.size}")
            print(f"    gpu    feature {a[0]} thr {a[1]!r} children {a[2]},{a[3]}")
            print(f"    oracle feature {b[0]} thr {b[1]!r} children {b[2]},{b[3]}")
            if rows is not None and rows.size >= 2:
                orders = np.argsort(X, axis=0, kind="stable")
                node_orders = np.stack([c[np.isin(c, rows)] for c in orders.T], axis=1)
                thr, col, gain, left = oracle.l2_split_node(node_orders, X, y, 1)
                gthr, gcol, ggain, gleft = one.greedy_regression_split(node_orders, X, y, 1)
                print(f"    oracle split of that node: col {col} thr {thr!r} gain {gain!r} nleft {left.size}")
                print(f"    gpu single-node split:     col {gcol} thr {gthr!r} gain {ggain!r} nleft {gleft.size}")
            return False
    print("  identical structure")
    return True


if __name__ == "__main__":
    from sklearn.datasets import make_regression
    X, y = make_regression(n_samples=10000, n_features=100, random_state=0)
    for d in (2, 3, 4, 6, 8, 10):
        if not compare(X, y, d, f"C2 depth {d}"):
            break
    for (n, F, depth, seed) in [(4097, 3, 12, 3), (30000, 8, 12, 2), (10000, 100, 10, 5), (10000, 50, 10, 6)]:
        Xs, ys = small_problem(seed, n, F, None, informative=4)
        compare(Xs, ys, depth, f"{n}x{F} d{depth}")


def golden_check():
    import hashlib
    from sklearn.datasets import make_regression
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_configs.npz"))
    X, y = make_regression(n_samples=10000, n_features=100, random_state=0)
    print("X flags", X.flags["C_CONTIGUOUS"], "sha ok", hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() == str(g["X_sha"]))
    o = oracle.fit_l2(X, y, max_depth=10)
    print("oracle(default) == golden:", np.array_equal(o.tree_.children_left, g["c2_children_left"]),
          np.array_equal(o.tree_.threshold, g["c2_threshold"]))
    trees = []
    one.build_regression_tree(X, y, max_depth=1)
    for rep in range(6):
        t = one.build_regression_tree(X, y, max_depth=10)
        trees.append(t)
        same = np.array_equal(t.tree_.children_left, g["c2_children_left"]) and np.array_equal(t.tree_.threshold, g["c2_threshold"])
        print(f"  rep {rep}: nodes {t.max_node_+1} equal-to-golden {same}")
        if not same:
            bad = np.flatnonzero((t.tree_.children_left != g["c2_children_left"]) | (t.tree_.threshold != g["c2_threshold"]) | (t.tree_.feature != g["c2_feature"]))
            v = bad[0]
            print("    first bad node", v, "gpu", t.tree_.feature[v], repr(t.tree_.threshold[v]), t.tree_.children_left[v], t.tree_.children_right[v],
                  "golden", g["c2_feature"][v], repr(g["c2_threshold"][v]), g["c2_children_left"][v], g["c2_children_right"][v])
            # find parent
            par = np.flatnonzero((g["c2_children_left"] == v) | (g["c2_children_right"] == v))
            print("    parent", par, "bad list", bad[:10])
            rows = node_rows(o, X, v)
            print("    rows in node (oracle):", None if rows is None else rows.size, "depth", depth_of(o, v))
            if rows is not None and rows.size >= 2:
                orders = np.argsort(X, axis=0, kind="stable")
                node_orders = np.stack([c[np.isin(c, rows)] for c in orders.T], axis=1)
                thr, col, gain, left = oracle.l2_split_node(node_orders, X, y, 1)
                gthr, gcol, ggain, gleft = one.greedy_regression_split(node_orders, X, y, 1)
                print(f"    oracle split: col {col} thr {thr!r} gain {gain!r} nleft {left.size}")
                print(f"    gpu 1-node  : col {gcol} thr {gthr!r} gain {ggain!r} nleft {gleft.size}")


if __name__ == "__main__":
    golden_check()
