"""CPU oracle for the regression-tree hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in numpy (+ one small C file for the running median),
the algorithm of yupbank/np_decision_tree's hot path (decision_tree/one.py,
decision_tree/strategy_l1.py, decision_tree/np_stream_median.pyx,
decision_tree/utils.py, decision_tree/four.py, and the greedy classification search of
decision_tree/strategy.py + decision_tree/tree_builder.py).  Every function cites the reference file:line it
follows.  It is the checker the CUDA path is compared against.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under
``np_decision_tree_b200/`` imports it; the product path has no CPU fallback.

Pinning: the reference ships no tests or golden vectors.
The oracle is pinned against outputs of the reference itself, run in the build
container on seeded inputs by ``tests/golden/make_golden.py``; the resulting
fixtures are committed under ``tests/golden/`` and replayed by
``tests/test_oracle_golden.py``.
"""
from .np_tree_oracle import (  # noqa: F401
    TreeArrays,
    argsort_columns,
    l2_weights,
    l2_best_split,
    l2_best_split_ratio,
    l2_split_node,
    partition_orders,
    fit_l2,
    running_median_pairs,
    running_median_pairs_py,
    l1_best_split,
    l1_split_node,
    fit_l1,
    predict,
    predict_leaf,
    LevelTreeArrays,
    l2_best_split_weighted,
    fit_l2_levelwise,
    class_prefix_counts,
    gini_surrogate,
    p_at_1_surrogate,
    class_split_node,
    fit_classes,
    random_class_split_node,
    fit_classes_random,
    random_l2_split_node,
    fit_l2_random,
)
