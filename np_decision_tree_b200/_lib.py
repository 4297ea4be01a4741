"""ctypes binding of libb200tree.so (include/b200dt.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable,
every entry point raises.  Set B200DT_BUILD=1 to let the import build the library with
nvcc when it is missing (``__graft_entry__.build()`` does this explicitly).
"""
import ctypes
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200tree.so")

HOST, DEVICE = 0, 1
L2, L1, GINI, P_AT_1 = 0, 1, 2, 3
NUMBER_BY_LEVEL, LEAF_IF_LE = 1, 2

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_dbl = ctypes.c_double


class Candidate(ctypes.Structure):
    """b200dt_candidate (include/b200dt.h)."""
    _fields_ = [("gain", c_dbl), ("aux", c_dbl), ("left_sum", c_dbl), ("gain2", c_dbl),
                ("threshold", c_dbl), ("k", ctypes.c_int32), ("f", ctypes.c_int32),
                ("exact", ctypes.c_int32), ("pad", ctypes.c_int32)]


CANDIDATE_BYTES = ctypes.sizeof(Candidate)
assert CANDIDATE_BYTES == 56
CLASS_CANDIDATE_BYTES = 32   # b200dt_class_candidate

_SIGNATURES = {
    "b200dt_abi_version": (c_int, []),
    "b200dt_create": (c_int, [ctypes.POINTER(c_vp), c_int]),
    "b200dt_destroy": (None, [c_vp]),
    "b200dt_last_error": (ctypes.c_char_p, [c_vp]),
    "b200dt_set_stream": (c_int, [c_vp, c_vp]),
    "b200dt_synchronize": (c_int, [c_vp]),
    "b200dt_argsort_columns": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_int]),
    "b200dt_l2_split_node": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_int,
                                     ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                     ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    "b200dt_l2_score_cumsums": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                        ctypes.POINTER(c_dbl)]),
    "b200dt_l1_split_node": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64,
                                     ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                     ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    "b200dt_l1_best_columns": (c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                       ctypes.POINTER(c_dbl)]),
    "b200dt_cummedian": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int]),
    "b200dt_split_orders": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "b200dt_fit": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_dbl, c_i64,
                           c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "b200dt_fit_ex": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_dbl, c_i64,
                              c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "b200dt_session_begin": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_dbl,
                                     c_i64, c_int, c_i64, c_i64, c_vp, c_i64]),
    "b200dt_session_live": (c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "b200dt_session_search": (c_int, [c_vp, c_vp]),
    "b200dt_session_search_exact": (c_int, [c_vp, c_vp]),
    "b200dt_session_decide": (c_int, [c_vp, c_vp, c_int, c_vp, ctypes.POINTER(c_int),
                                      ctypes.POINTER(c_int)]),
    "b200dt_session_partition": (c_int, [c_vp, c_vp]),
    "b200dt_session_finish": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "b200dt_predict": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_i64, c_vp, c_vp, c_int]),
    "b200dt_fit_classes": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_dbl, c_i64,
                                   c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "b200dt_class_session_begin": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_dbl,
                                           c_i64, c_i64, c_i64]),
    "b200dt_class_session_live": (c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "b200dt_class_session_search": (c_int, [c_vp, c_vp]),
    "b200dt_class_session_decide": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, ctypes.POINTER(c_i64)]),
    "b200dt_class_session_partition": (c_int, [c_vp, c_vp, c_vp]),
    "b200dt_class_session_finish": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "b200dt_class_split_node": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int,
                                        ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl),
                                        ctypes.POINTER(c_dbl)]),
    "b200dt_rows_begin": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_vp]),
    "b200dt_rows_minmax": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_vp]),
    "b200dt_rows_threshold_counts": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_vp]),
    "b200dt_rows_set_targets": (c_int, [c_vp, c_vp, c_int]),
    "b200dt_rows_threshold_sums": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "b200dt_rows_split": (c_int, [c_vp, c_int, c_i64, c_i64, ctypes.c_int32, c_dbl, ctypes.POINTER(c_i64)]),
    "b200dt_rows_end": (c_int, [c_vp]),
    "b200dt_profile_enable": (c_int, [c_vp, c_int]),
    "b200dt_profile_reset": (c_int, [c_vp]),
    "b200dt_profile_count": (c_int, []),
    "b200dt_profile_name": (ctypes.c_char_p, [c_int]),
    "b200dt_profile_get": (c_int, [c_vp, c_int, ctypes.POINTER(c_dbl), ctypes.POINTER(c_i64),
                                   ctypes.POINTER(c_dbl)]),
    "b200dt_launch_count": (c_i64, [c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


class B200Error(RuntimeError):
    pass


def load():
    """Load libb200tree.so; raise (never fall back) when it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = LIB_PATH
        variant = os.environ.get("B200DT_LIB")
        if variant:
            # developer override: another build of the same sources (kernel tuning experiments, scripts/build_variant.sh)
            if not os.path.exists(variant):
                raise B200Error(f"B200DT_LIB={variant} does not exist")
            lib = ctypes.CDLL(variant)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            return lib
        if os.environ.get("B200DT_CHECKED") == "1":
            # the bounds-checked build (`python -m np_decision_tree_b200._build --checked`): every hot kernel asserts
            # its indices on the device and traps with a message; used where compute-sanitizer is not available
            path = LIB_PATH.replace("libb200tree.so", "libb200tree_checked.so")
            if not os.path.exists(path):
                raise B200Error(f"{path} is missing: build it with `python -m np_decision_tree_b200._build --checked`")
            lib = ctypes.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            return lib
        if not os.path.exists(LIB_PATH):
            if os.environ.get("B200DT_BUILD") == "1":
                from . import _build
                _build.build()
            else:
                raise B200Error(
                    f"{LIB_PATH} is missing: build it with `python -m np_decision_tree_b200._build` "
                    "(needs nvcc).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.b200dt_abi_version() != 1:
            raise B200Error("libb200tree.so ABI version mismatch; rebuild it")
        _lib = lib
        return lib


class Context:
    """One b200dt_ctx = one GPU, one stream, one host thread."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_vp()
        if self.lib.b200dt_create(ctypes.byref(h), int(device)) != 0:
            raise B200Error(self.lib.b200dt_last_error(None).decode())
        self.handle = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise B200Error(self.lib.b200dt_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200dt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t handle (0 = legacy default stream), or None for the context's own."""
        handle = c_vp(-1) if cuda_stream is None else c_vp(int(cuda_stream))
        self.check(self.lib.b200dt_set_stream(self.handle, handle))

    def synchronize(self):
        self.check(self.lib.b200dt_synchronize(self.handle))

    # -- measurement ---------------------------------------------------------------
    def profile_enable(self, on=True):
        self.check(self.lib.b200dt_profile_enable(self.handle, 1 if on else 0))

    def profile_reset(self):
        self.check(self.lib.b200dt_profile_reset(self.handle))

    def profile(self):
        out = {}
        for kc in range(self.lib.b200dt_profile_count()):
            ms, n, b = c_dbl(), c_i64(), c_dbl()
            self.check(self.lib.b200dt_profile_get(self.handle, kc, ctypes.byref(ms), ctypes.byref(n),
                                                   ctypes.byref(b)))
            if n.value:
                out[self.lib.b200dt_profile_name(kc).decode()] = {
                    "ms": ms.value, "launches": n.value, "alg_bytes": b.value}
        return out

    def launch_count(self):
        return int(self.lib.b200dt_launch_count(self.handle))


_contexts = {}


def default_context(device=None):
    """Process-wide context per device (created on first use)."""
    if device is None:
        device = int(os.environ.get("B200DT_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def as_arg(a, dtype, what):
    """(pointer, space, keepalive) of a numpy array or a torch CUDA tensor."""
    if isinstance(a, np.ndarray) or not hasattr(a, "data_ptr"):
        arr = np.ascontiguousarray(a, dtype=dtype)
        return arr.ctypes.data, HOST, arr
    import torch
    want = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
    if a.dtype != want:
        raise TypeError(f"{what}: expected a {want} tensor, got {a.dtype}")
    if not a.is_contiguous():
        a = a.contiguous()
    if a.is_cuda:
        return a.data_ptr(), DEVICE, a
    arr = a.numpy()
    return arr.ctypes.data, HOST, arr
