"""ctypes binding of ``libquantik_b200.so`` (C ABI: ``include/quantik_b200.h``).

There is deliberately no fallback: if the shared library is missing, or no sm_100
device can be initialised, every compute call raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` or
``quantik-core-py_b200/csrc/build.sh``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_int, c_int8, c_size_t, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUANTIK_B200_LIB") or os.path.join(_HERE, "_lib", "libquantik_b200.so")

ABI_VERSION = 1
MAX_DEPTH = 16


class QuantikB200Error(RuntimeError):
    """A C-ABI call returned a negative QK_E_* code."""

    def __init__(self, code: int, text: str):
        super().__init__(f"libquantik_b200: {text or 'error'} (code {code})")
        self.code = code


# name -> (restype, argtypes); kept in one table so tests can check it against the header
PROTOTYPES = {
    "qk_abi_version": (c_int, []),
    "qk_init": (c_int, [c_int]),
    "qk_shutdown": (c_int, [c_int]),
    "qk_last_error": (c_char_p, []),
    "qk_set_option": (c_int, [c_char_p, ctypes.c_longlong]),
    "qk_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int)]),
    "qk_movegen_masks": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_apply_moves": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_win_status": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_canonical": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_orbit_size": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_apply_symmetry": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_apply_symmetry_to_moves": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_canon_dedup": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_size_t,
                               POINTER(c_size_t), c_int, c_void_p]),
    "qk_bfs_level": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_size_t),
                             c_void_p, c_int, c_void_p]),
    "qk_book_level": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_size_t), c_void_p,
                              c_int, c_void_p]),
    "qk_inbox_bytes": (c_int, [c_int, c_size_t, POINTER(c_size_t)]),
    "qk_payload_owner": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_int, c_void_p]),
    "qk_canon_dedup_scatter": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_void_p), c_int, c_int, c_size_t,
                                       c_void_p, c_int, c_void_p]),
    "qk_bfs_level_scatter": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_size_t, POINTER(c_void_p), c_int, c_int,
                                     c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_inbox_merge": (c_int, [c_void_p, c_int, c_size_t, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_size_t),
                               c_int, c_void_p]),
    "qk_playouts": (c_int, [c_void_p, c_size_t, c_uint32, c_uint64, c_int, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_int, c_void_p]),
    "qk_playouts_ex": (c_int, [c_void_p, c_size_t, c_uint32, c_uint64, c_uint64, c_uint32, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_eval_features": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_playouts_guided": (c_int, [c_void_p, c_size_t, c_uint32, c_uint64, c_uint64, c_uint32, c_int, c_double, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_solve": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "qk_random_roots": (c_int, [c_uint64, c_uint64, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p]),
    "qk_perft_probe": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "qk_perft": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_int, c_void_p]),
    "qk_int_issue_peak": (c_int, [c_int, POINTER(c_double), POINTER(c_double)]),
    "qk_pipe_rates": (c_int, [c_int, POINTER(c_double)]),
}

_lib = None
_lock = threading.Lock()
_inited = set()


def load() -> ctypes.CDLL:
    """dlopen the library and attach prototypes (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise QuantikB200Error(
                        -3, f"{LIB_PATH} not found: build it first (__graft_entry__.build()); "
                        "there is no CPU fallback")
                lib = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                if lib.qk_abi_version() != ABI_VERSION:
                    raise QuantikB200Error(-1, "ABI version mismatch between header and library")
                _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise QuantikB200Error(rc, (load().qk_last_error() or b"").decode("utf-8", "replace"))


def init(dev: int = 0) -> ctypes.CDLL:
    """qk_init(dev) once per process and device; raises when no sm_100 GPU is usable."""
    lib = load()
    if dev not in _inited:
        with _lock:
            if dev not in _inited:
                check(lib.qk_init(dev))
                _inited.add(dev)
    return lib


def set_option(name: str, value: int) -> None:
    """qk_set_option: tuning / testing switch (see include/quantik_b200.h); 0 restores the default."""
    check(load().qk_set_option(name.encode(), int(value)))


def shutdown(dev: int = 0) -> None:
    """qk_shutdown(dev); the next call on that device initialises it again."""
    with _lock:
        check(load().qk_shutdown(dev))
        _inited.discard(dev)


def device_info(dev: int = 0):
    lib = init(dev)
    sms, khz = c_int(), c_int()
    check(lib.qk_device_info(dev, ctypes.byref(sms), ctypes.byref(khz)))
    return {"sm_count": sms.value, "sm_clock_khz": khz.value}
