"""ctypes binding of libseacells_b200.so (include/seacells_b200.h).  No fallback: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseacells_b200.so")

SC_OK = 0
_ERRNAMES = {1: "CUDA", 2: "ARG", 3: "STATE", 4: "CAPACITY", 5: "NCCL"}


class SeacellsB200Error(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); kept in one table so tests can check it against the header
_vp, _i, _ll, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_double
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "sc_ctx_create": (_i, [_i, _pp]),
    "sc_ctx_destroy": (None, [_vp]),
    "sc_last_error": (C.c_char_p, [_vp]),
    "sc_launch_count": (C.c_ulonglong, [_vp]),
    "sc_knn_fallback_rows": (_i, [_vp]),
    "sc_knn_profile": (_i, [_vp, _vp]),
    "sc_sync": (_i, [_vp]),
    "sc_stream": (_vp, [_vp]),
    "sc_version": (C.c_char_p, []),
    "sc_nccl_load_library": (_i, [C.c_char_p]),
    "sc_nccl_unique_id": (_i, [_vp]),
    "sc_ctx_init_nccl": (_i, [_vp, _i, _i, _vp]),
    "sc_local_group_new": (_i, [_i, _pp]),
    "sc_local_group_free": (None, [_vp]),
    "sc_local_group_abort": (None, [_vp]),
    "sc_ctx_attach_local": (_i, [_vp, _vp, _i]),
    "sc_ctx_rank": (_i, [_vp]),
    "sc_ctx_world": (_i, [_vp]),
    "sc_knn": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sc_knn_tc_probe": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sc_knn_tc_error_factor": (_d, [_i]),
    "sc_kernel_build": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _pp, C.POINTER(_ll)]),
    "sc_kernel_build_from_knn": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _pp, C.POINTER(_ll)]),
    "sc_kernel_export": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sc_kernel_free": (None, [_vp]),
    "sc_spmm_csr_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sc_summarize": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp]),
    "sc_greedy_centers": (_i, [_vp, _i, _vp, _vp, _vp, _ll, _i, _vp]),
    "sc_fit_create": (_i, [_vp, _i, _i, _i, _pp]),
    "sc_fit_destroy": (None, [_vp]),
    "sc_fit_set_kernel": (_i, [_vp, _vp, _vp, _vp, _ll]),
    "sc_fit_set_kernel_from": (_i, [_vp, _vp]),
    "sc_fit_set_archetypes": (_i, [_vp, _vp]),
    "sc_fit_set_A_csc": (_i, [_vp, _vp, _vp, _vp]),
    "sc_fit_set_A_slots": (_i, [_vp, _vp, _vp, _vp]),
    "sc_fit_set_B_slots": (_i, [_vp, _vp, _vp, _vp]),
    "sc_fit_update_A": (_i, [_vp, _d, _vp]),
    "sc_fit_update_B": (_i, [_vp, _vp]),
    "sc_fit_rss": (_i, [_vp, C.POINTER(_d)]),
    "sc_fit_shard": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "sc_fit_run": (_i, [_vp, _i, _i, _d, _d, C.POINTER(_d), C.POINTER(_i), C.POINTER(_i), _vp, _i, C.POINTER(_i)]),
    "sc_fit_get_KB": (_i, [_vp, C.POINTER(_ll), _vp, _vp, _vp]),
    "sc_fit_slots": (_i, [_vp]),
    "sc_fit_get_A": (_i, [_vp, _vp, _vp, _vp]),
    "sc_fit_get_B": (_i, [_vp, _vp, _vp, _vp]),
    "sc_fit_hard_assignments": (_i, [_vp, _vp]),
    "sc_fit_hard_archetypes": (_i, [_vp, _vp]),
    "sc_fit_soft_assignments": (_i, [_vp, _i, _vp, _vp]),
    "sc_fit_stats": (_i, [_vp, _vp]),
    "sc_fit_profile": (_i, [_vp, _vp]),
}


def load():
    """Load the shared library (built in-tree by `__graft_entry__.build()` / `make -C seacells_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeacellsB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  seacells_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Address of a numpy array, a torch tensor (host or CUDA), a raw int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array passed to the C ABI must be C-contiguous")
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch.Tensor hand-off
        if not a.is_contiguous():
            raise ValueError("tensor passed to the C ABI must be contiguous")
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


class Context:
    """One per process/GPU: owns the CUDA stream the library launches on."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.sc_ctx_create(int(device), C.byref(h))
        if rc != SC_OK:
            raise SeacellsB200Error(
                f"sc_ctx_create(device={device}) failed ({_ERRNAMES.get(rc, rc)}): no usable CUDA device. "
                "seacells_b200 runs on B200 (sm_100a) only and has no CPU fallback.")
        self.h = h
        self.device = device

    def check(self, rc: int, what: str = ""):
        if rc != SC_OK:
            msg = self.lib.sc_last_error(self.h)
            raise SeacellsB200Error(f"{what} failed ({_ERRNAMES.get(rc, rc)}): {msg.decode() if msg else ''}")

    @property
    def rank(self) -> int:
        return int(self.lib.sc_ctx_rank(self.h))

    @property
    def world(self) -> int:
        return int(self.lib.sc_ctx_world(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.sc_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.sc_stream(self.h) or 0)

    def sync(self):
        self.check(self.lib.sc_sync(self.h), "sc_sync")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
