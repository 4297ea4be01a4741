"""Mirror of the reference module ``decision_tree/stream_median.py`` -- the running (lower, upper) median.

The reference keeps two heaps and answers after every pushed number (stream_median.py:28-92); its native
counterpart is ``np_stream_median.cummedian`` (np_stream_median.pyx:8-39), which this package computes on the
GPU (``b200dt_cummedian``).  ``StreamMedian`` keeps the reference's call protocol -- ``add`` / ``median`` /
``double_median`` / ``cummedian(num)`` / ``__call__(num)`` -- on top of that kernel: the numbers seen so far
are buffered and the answer for the current prefix is the last pair of the kernel's output.  A caller that
owns the whole sequence should pass it to ``np_stream_median.cummedian`` (or ``feed``) in one call: that is one
kernel launch instead of one per element.  There is no CPU fallback here either.
"""
import numpy as np

from .np_stream_median import cummedian as _cummedian


class StreamMedian(object):
    def __init__(self):
        self._seen = []
        self._pair = None        # (lower, upper) median of everything seen, None = stale

    def add(self, num):
        """stream_median.py:33-47."""
        self._seen.append(float(num))
        self._pair = None

    def _current(self):
        if not self._seen:
            raise IndexError("list index out of range")     # the reference's heap lookup on an empty stream
        if self._pair is None:
            out = _cummedian(np.asarray(self._seen, dtype=np.float64))
            self._pair = (float(out[-1, 0]), float(out[-1, 1]))
        return self._pair

    def median(self):
        """stream_median.py:49-55: mean of the pair for an even count, else the middle value."""
        lo, hi = self._current()
        return (hi + lo) / 2 if len(self._seen) % 2 == 0 else lo

    def double_median(self):
        """stream_median.py:57-65."""
        return self._current()

    def cummedian(self, num):
        """stream_median.py:67-88: push, then the pair."""
        self.add(num)
        return self._current()

    def feed(self, numbers):
        """All pairs of a whole sequence in one kernel call -> (len, 2) array; the stream then continues after it."""
        numbers = np.asarray(numbers, dtype=np.float64).ravel()
        start = len(self._seen)
        self._seen.extend(numbers.tolist())
        out = _cummedian(np.asarray(self._seen, dtype=np.float64))
        self._pair = (float(out[-1, 0]), float(out[-1, 1]))
        return out[start:]

    def __call__(self, num):
        """stream_median.py:90-92."""
        self.add(num)
        return self.double_median()
