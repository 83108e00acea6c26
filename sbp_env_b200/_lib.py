"""ctypes binding of ``libsbpgpu.so`` (C ABI declared in ``include/sbp_gpu.h``).

There is deliberately no fallback: if the CUDA library is missing or no B200 is
visible, every entry point of this package raises.  Nothing here (or anywhere in
the package) imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsbpgpu.so")

SBP_OK = 0
FLAG_NAN = 1
FLAG_OVERFLOW = 2
FLAG_PEER_TIMEOUT = 1024
FREE, BLOCKED, AMBIGUOUS = 1, 0, 2
METRIC = {"full": 0, "xy": 1}


class SbpGpuError(RuntimeError):
    """A libsbpgpu call failed (or the library / a CUDA device is missing)."""


_lib = None

_i32, _i64, _dbl, _vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
_pp = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int32 status unless listed in _RESTYPE
_SIGNATURES = {
    "sbp_abi_version": [],
    "sbp_last_error": [],
    "sbp_ctx_create": [_i32, _vp, _pp],
    "sbp_ctx_destroy": [_vp],
    "sbp_ctx_set_stream": [_vp, _vp],
    "sbp_ctx_sync": [_vp],
    "sbp_ctx_info": [_vp, C.POINTER(_i32), C.POINTER(_i64)],
    "sbp_map_create": [_vp, _vp, _i32, _i32, _pp],
    "sbp_map_create_device": [_vp, _vp, _i32, _i32, _pp],
    "sbp_map_destroy": [_vp],
    "sbp_map_info": [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i32)],
    "sbp_feasible2d": [_vp, _vp, _i64, _vp, _vp],
    "sbp_visible2d": [_vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "sbp_first_blocked2d": [_vp, _vp, _vp, _i64, _vp, _vp],
    "sbp_arm_feasible": [_vp, _dbl, _dbl, _vp, _i64, _vp, _vp],
    "sbp_arm_visible": [_vp, _dbl, _dbl, _vp, _vp, _i64, _vp, _vp, _vp],
    "sbp_arm_resolve_host": [_vp, _dbl, _dbl, _vp, _vp, _i64, _vp, C.POINTER(_i64)],
    "sbp_feasible2d_host": [_vp, _vp, _i64, _vp, C.POINTER(_i32)],
    "sbp_visible2d_host": [_vp, _vp, _vp, _i64, _vp, C.POINTER(_i32)],
    "sbp_first_blocked2d_host": [_vp, _vp, _vp, _i64, _vp, C.POINTER(_i32)],
    "sbp_arm_feasible_host": [_vp, _dbl, _dbl, _vp, _i64, _vp, C.POINTER(_i32), C.POINTER(_i64)],
    "sbp_arm_visible_host": [_vp, _dbl, _dbl, _vp, _vp, _i64, _vp, C.POINTER(_i32), C.POINTER(_i64)],
    "sbp_nodes_create": [_vp, _i32, _i64, _pp],
    "sbp_nodes_destroy": [_vp],
    "sbp_nodes_append": [_vp, _vp, _i64, _i32],
    "sbp_nodes_size": [_vp, C.POINTER(_i64)],
    "sbp_nodes_truncate": [_vp, _i64],
    "sbp_nodes_read": [_vp, _i64, _i64, _vp],
    "sbp_nodes_rows_device": [_vp, _i64, _i64, _vp],
    "sbp_nodes_knn": [_vp, _vp, _i64, _i32, _i32, _vp, _vp],
    "sbp_nodes_nearest": [_vp, _vp, _i64, _vp, _vp],
    "sbp_nodes_nearest_host": [_vp, _vp, _i64, _vp, _vp],
    "sbp_nodes_nearest1_host": [_vp, _vp, C.POINTER(_i64), C.POINTER(_dbl)],
    "sbp_nodes_near": [_vp, _vp, _i64, _dbl, _i32, _i32, _vp, _vp, _i64, _vp],
    "sbp_nodes_knn_host": [_vp, _vp, _i64, _i32, _i32, _vp, _vp],
    "sbp_nodes_near_host": [_vp, _vp, _i64, _dbl, _i32, _i32, _vp, _vp, _i64, C.POINTER(_i64)],
    "sbp_nodes_gather_edges": [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp],
    "sbp_nodes_knn_edges": [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    "sbp_knn_edges_visible2d": [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "sbp_prm_connect2d": [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp],
    "sbp_prm_connect2d_sorted": [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp],
    "sbp_nodes_near_visible": [_vp, _vp, _i32, _dbl, _dbl, _vp, _i64, _dbl, _i32, _i32, _i32, _vp, _vp,
                               _i32, _vp, _vp],
    "sbp_nodes_near_visible_host": [_vp, _vp, _i32, _dbl, _dbl, _vp, _dbl, _i32, _i32, _i32, _vp, _vp,
                                    _i32, C.POINTER(_i32), C.POINTER(_i32)],
    "sbp_tree_create": [_vp, C.POINTER(_vp)],
    "sbp_tree_destroy": [_vp],
    "sbp_tree_set": [_vp, _i64, _dbl, _i64],
    "sbp_tree_read": [_vp, _i64, _i64, _vp, _vp],
    "sbp_tree_write": [_vp, _i64, _i64, _vp, _vp],
    "sbp_tree_extend_host": [_vp, _vp, _vp, _i64, _dbl, _i32, _i32, _vp, _vp, _i32, _vp],
    "sbp_graph_create": [_vp, _i64, C.POINTER(_vp)],
    "sbp_graph_destroy": [_vp],
    "sbp_graph_build_knn": [_vp, _vp, _vp, _i64, _i32, _i64, _i32],
    "sbp_graph_build_packed": [_vp, _vp, _i64, _i32, _i64, _i32],
    "sbp_graph_build_packed_rows": [_vp, _vp, _vp, _i64, _i32, _i32],
    "sbp_graph_build_edges": [_vp, _vp, _vp, _vp, _i64, _i32],
    "sbp_graph_components": [_vp, _vp, C.POINTER(_i64)],
    "sbp_graph_info": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "sbp_graph_csr": [_vp, _vp, _vp, _i64],
    "sbp_graph_bfs": [_vp, _vp, _vp, _i64, _vp, _vp, _i32],
    "sbp_adjacency_peer_scatter": [_vp, _vp, _vp, _i64, _vp, _i32, _i64],
    "sbp_adjacency_multicast": [_vp, _vp, _vp, _i64, _vp, _i64],
    "sbp_ctx_tune": [_vp, C.c_char_p, _i64],
    "sbp_ctx_scratch_generation": [_vp, C.POINTER(_i64)],
    "sbp_microbench": [_vp, _i32, _i64, _i32, C.POINTER(_dbl), C.POINTER(_dbl)],
    "sbp_ctx_kernel_time": [_vp, C.POINTER(_dbl), C.POINTER(_i64)],
}
_RESTYPE = {"sbp_last_error": C.c_char_p}

EXTEND_OK, EXTEND_AMBIGUOUS, EXTEND_NO_PARENT = 0, 1, 2
EXTEND_MAX_FIXES, EXTEND_MAX_REWIRED = 24, 64


class ExtendResult(C.Structure):
    """``sbp_extend_result`` of include/sbp_gpu.h."""
    _fields_ = [("status", _i32), ("n_candidates", _i32), ("n_rewired", _i32), ("n_rewire_checked", _i32),
                ("new_idx", _i64), ("parent", _i64), ("cost", _dbl), ("c_parent", _dbl),
                ("rewired", _i32 * EXTEND_MAX_REWIRED), ("rewired_c", _dbl * EXTEND_MAX_REWIRED)]

#: every entry point the library exports is declared in include/sbp_gpu.h (the tuning /
#: measurement hooks under "instrumentation")
ABI_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load libsbpgpu.so; raise ``SbpGpuError`` (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbpGpuError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  sbp_env_b200 has no CPU implementation."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise SbpGpuError(f"cannot load {LIB_PATH}: {e}") from e
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, _i32)
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().sbp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != SBP_OK:
        raise SbpGpuError(f"{what or 'libsbpgpu'} failed (status {rc}): {last_error()}")
