"""Mirror of the reference package ``decision_tree`` (hot-path modules only)."""
from . import one, strategy_l1, utils, np_stream_median, four  # noqa: F401
