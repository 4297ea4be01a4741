"""np_decision_tree_b200 -- the regression-tree hot path of yupbank/np_decision_tree on B200.

Drop-in for ONE path: exhaustive L2 / L1 split search, node partition and batched inference,
behind the reference's own Python API:

    from np_decision_tree_b200.decision_tree.one import build_regression_tree
    from np_decision_tree_b200.decision_tree.strategy_l1 import build_regression_tree_v2
    from np_decision_tree_b200.decision_tree.utils import inference

or, to keep the reference's import paths unchanged (``from decision_tree.one import ...``,
``from np_stream_median import cummedian``), call :func:`install_aliases` first.

All numeric work runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/b200dt.h`` (``libb200tree.so``); there is no CPU fallback.
"""
import sys

__version__ = "0.1.0"


def install_aliases():
    """Register this package's mirrors under the reference's module names."""
    from . import decision_tree as dt
    from .decision_tree import (one, strategy_l1, utils, np_stream_median, four, base, strategy, tree_builder,
                                stream_median)
    sys.modules["decision_tree"] = dt
    sys.modules["decision_tree.one"] = one
    sys.modules["decision_tree.strategy_l1"] = strategy_l1
    sys.modules["decision_tree.utils"] = utils
    sys.modules["decision_tree.four"] = four
    sys.modules["decision_tree.base"] = base
    sys.modules["decision_tree.strategy"] = strategy
    sys.modules["decision_tree.tree_builder"] = tree_builder
    sys.modules["decision_tree.stream_median"] = stream_median
    sys.modules["np_stream_median"] = np_stream_median   # ref strategy_l1.py:4 imports it top-level
