"""ctypes binding of libvael_b200.so (include/vael_b200.h).

The library is the product path: if it is missing this module raises at import of the
first compute call — there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libvael_b200.so"

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class CircuitDesc(C.Structure):
    _fields_ = [
        ("n_facts", C.c_int32), ("n_nodes", C.c_int32), ("n_levels", C.c_int32), ("n_edges", C.c_int32),
        ("n_worlds", C.c_int32), ("n_queries", C.c_int32), ("n_consts", C.c_int32), ("semiring", C.c_int32),
        ("norm_node", C.c_int32),
        ("node_kind", c_u8p), ("node_arg", c_i32p), ("level_ptr", c_i32p), ("child_ptr", c_i32p),
        ("child_idx", c_i32p), ("const_val", c_f32p), ("world_node", c_i32p), ("query_node", c_i32p),
    ]


class VaelError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vael_b200 error {code}: {msg}")
        self.code = code


_P = C.c_void_p
_PROTOTYPES = {
    "vael_abi_version": (C.c_int, []),
    "vael_last_error": (C.c_char_p, []),
    "vael_launch_count": (C.c_int64, []),
    "vael_last_kernel": (C.c_char_p, []),
    "vael_circuit_create": (C.c_int, [C.POINTER(CircuitDesc), C.POINTER(_P)]),
    "vael_circuit_destroy": (None, [_P]),
    "vael_circuit_fast_kind": (C.c_int, [_P]),
    "vael_circuit_hash": (C.c_uint64, [_P]),
    "vael_circuit_jit_status": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "vael_jit_compile_check": (C.c_int, [C.POINTER(CircuitDesc), C.c_char_p, C.c_int32]),
    "vael_circuit_forward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P, C.c_uint32, _P]),
    "vael_circuit_backward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, C.c_uint32, _P]),
    "vael_circuit_queries_all": (C.c_int, [_P, _P, C.c_int64, _P, C.c_uint32, _P]),
    "vael_circuit_forward_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P, C.c_uint32]),
    "vael_circuit_fwdbwd_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, C.c_uint32]),
    "vael_circuit_query_grad_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P, C.c_uint32]),
    "vael_facts_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "vael_facts_backward": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, _P]),
    "vael_facts_clip_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P]),
    "vael_facts_clip_backward": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P]),
    "vael_gumbel_world_forward": (C.c_int, [_P, _P, C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P,
                                            C.c_int32, _P, _P, _P, _P]),
    "vael_gumbel_world_backward": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, C.c_int32,
                                             _P, _P]),
    "vael_logic_fused_supported": (C.c_int, [_P]),
    "vael_logic_train_forward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_float, C.c_float, _P, C.c_uint64, _P,
                                           _P, _P, _P, _P, _P, _P]),
    "vael_logic_train_backward": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_float, C.c_float, _P, C.c_uint64, _P,
                                            _P, _P, _P, _P, _P]),
    "vael_elbo_workspace_bytes": (C.c_int64, []),
    "vael_elbo_terms": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int32, _P, C.c_float, C.c_float,
                                  C.c_float, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
ABI_VERSION = 2


def load() -> C.CDLL:
    """dlopen the in-tree library (building is `python -m vael_b200.build` / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("VAEL_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(f"{path} not found: build it with `python -m vael_b200.build` "
                          "(vael_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.vael_abi_version() != ABI_VERSION:
        raise ImportError("libvael_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise VaelError(rc, load().vael_last_error().decode())


def launch_count() -> int:
    return int(load().vael_launch_count())


def last_kernel() -> str:
    return load().vael_last_kernel().decode()


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ) if a.size else C.cast(None, typ)


def make_desc(circ) -> CircuitDesc:
    """vael_circuit_desc_t over the numpy arrays of a vael_b200.circuit.Circuit (no copy;
    keep `circ` alive while the descriptor is in use)."""
    d = CircuitDesc()
    d.n_facts, d.n_nodes, d.n_levels, d.n_edges = circ.n_facts, circ.n_nodes, circ.n_levels, circ.n_edges
    d.n_worlds, d.n_queries, d.n_consts = circ.n_worlds, circ.n_queries, len(circ.const_val)
    d.semiring, d.norm_node = circ.semiring, circ.norm_node
    d.node_kind = _ptr(circ.node_kind, c_u8p)
    d.node_arg = _ptr(circ.node_arg, c_i32p)
    d.level_ptr = _ptr(circ.level_ptr, c_i32p)
    d.child_ptr = _ptr(circ.child_ptr, c_i32p)
    d.child_idx = _ptr(circ.child_idx, c_i32p)
    d.const_val = _ptr(circ.const_val, c_f32p)
    d.world_node = _ptr(circ.world_node, c_i32p)
    d.query_node = _ptr(circ.query_node, c_i32p)
    return d
