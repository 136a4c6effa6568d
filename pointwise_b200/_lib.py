"""ctypes binding of libconv3p_b200.so (C ABI in include/conv3p_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library cannot be loaded every operator call
raises.  The library is built in-tree by ``pointwise_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_BUFFER_TOO_SMALL = 2
ERR_CUDA = 3
ERR_UNSUPPORTED = 4
ERR_PAIR_OVERFLOW = 5
ERR_NO_BACKWARD_LISTS = 6

# every symbol include/conv3p_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "conv3p_plan_bytes", "conv3p_plan_layout", "conv3p_plan_build_f32", "conv3p_plan_build_backward",
    "conv3p_plan_stats", "conv3p_plan_publish_stats", "conv3p_scratch_bytes", "conv3p_backward_scratch_bytes", "conv3p_forward_f32", "conv3p_forward_ex_f32", "conv3p_selu_backward_f32", "conv3p_backward_f32",
    "conv3p_op_workspace_bytes", "conv3p_op_workspace_bytes_ex", "conv3p_op_backward_workspace_bytes", "conv3p_op_forward_f32", "conv3p_op_backward_f32",
    "conv3p_op_workspace_bytes_f64", "conv3p_op_forward_f64", "conv3p_op_backward_f64",
    "conv3p_host_workspace_bytes", "conv3p_host_forward_f32", "conv3p_host_backward_f32",
    "conv3p_status_string", "conv3p_last_cuda_error", "conv3p_abi_version", "conv3p_launch_count",
    "conv3p_set_engine", "conv3p_profile_enable", "conv3p_profile_read",
    "conv3p_selftest_tc", "conv3p_selftest_tc_mn", "conv3p_debug_phase_cycles", "conv3p_debug_cta_cycles", "conv3p_debug_w2_cycles", "conv3p_debug_mma_rate",
    "conv3p_augment_rotate_jitter_f32", "conv3p_xyz_sort_workspace_bytes", "conv3p_xyz_sort_f32",
]


class Geom(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("stride", C.c_int * 3), ("voxel_size", C.c_float),
                ("pair_capacity", C.c_longlong)]


class PlanStats(C.Structure):
    _fields_ = [("total_pairs", C.c_longlong), ("backward_pairs", C.c_longlong),
                ("overflow", C.c_int), ("has_backward", C.c_int)]


class PlanLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "header", "cloud_meta", "sorted_key", "sorted_xyzi", "count_table", "pair_begin", "pair_len",
        "pair_row", "bwd_count", "bwd_row", "bwd_weight", "sort_tmp", "cell_start", "total_bytes")]


class Conv3pError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"conv3p_b200 status {status}: {message}")
        self.status = status


_lock = threading.Lock()
_lib = None


def _declare(L):
    vp, sz, ll, i, f = C.c_void_p, C.c_size_t, C.c_longlong, C.c_int, C.c_float
    gp = C.POINTER(Geom)
    i3 = C.POINTER(C.c_int)
    L.conv3p_plan_bytes.argtypes = [gp]
    L.conv3p_plan_bytes.restype = sz
    L.conv3p_plan_layout.argtypes = [gp, C.POINTER(PlanLayout)]
    L.conv3p_plan_build_f32.argtypes = [gp, vp, vp, sz, vp]
    L.conv3p_plan_build_backward.argtypes = [gp, vp, vp, sz, vp]
    L.conv3p_plan_stats.argtypes = [gp, vp, C.POINTER(PlanStats), vp]
    L.conv3p_plan_publish_stats.argtypes = [gp, vp, vp, vp]
    L.conv3p_scratch_bytes.argtypes = [gp, i, i]
    L.conv3p_scratch_bytes.restype = sz
    L.conv3p_backward_scratch_bytes.argtypes = [gp, i, i]
    L.conv3p_backward_scratch_bytes.restype = sz
    L.conv3p_forward_f32.argtypes = [gp, vp, vp, vp, i, i, vp, vp, sz, vp]
    L.conv3p_forward_ex_f32.argtypes = [gp, vp, vp, ll, vp, i, i, vp, ll, i, vp, sz, vp]
    L.conv3p_selu_backward_f32.argtypes = [vp, ll, vp, ll, vp, ll, i, vp]
    L.conv3p_backward_f32.argtypes = [gp, vp, vp, vp, vp, i, i, vp, vp, vp, sz, vp]
    L.conv3p_op_workspace_bytes.argtypes = [gp, i, i]
    L.conv3p_op_workspace_bytes.restype = sz
    L.conv3p_op_workspace_bytes_ex.argtypes = [gp, i3, i, i, i]
    L.conv3p_op_workspace_bytes_ex.restype = sz
    L.conv3p_op_backward_workspace_bytes.argtypes = [gp, i, i]
    L.conv3p_op_backward_workspace_bytes.restype = sz
    L.conv3p_op_forward_f32.argtypes = [vp, vp, vp, i3, i3, f, i, i, i, i, ll, vp, vp, sz, vp]
    L.conv3p_op_backward_f32.argtypes = [vp, vp, vp, vp, i3, i3, f, i, i, i, i, ll, vp, vp, vp, sz, vp]
    d = C.c_double
    L.conv3p_op_workspace_bytes_f64.argtypes = [gp, i3, i, i]
    L.conv3p_op_workspace_bytes_f64.restype = sz
    L.conv3p_op_forward_f64.argtypes = [vp, vp, vp, i3, i3, d, i, i, i, i, ll, vp, vp, sz, vp]
    L.conv3p_op_backward_f64.argtypes = [vp, vp, vp, vp, i3, i3, d, i, i, i, i, ll, vp, vp, vp, sz, vp]
    L.conv3p_host_workspace_bytes.argtypes = [gp, i, i]
    L.conv3p_host_workspace_bytes.restype = sz
    L.conv3p_host_forward_f32.argtypes = [vp, vp, vp, i3, f, i, i, i, i, ll, vp, vp, sz, vp]
    L.conv3p_host_backward_f32.argtypes = [vp, vp, vp, vp, i3, f, i, i, i, i, ll, vp, vp, vp, sz, vp]
    L.conv3p_status_string.argtypes = [i]
    L.conv3p_status_string.restype = C.c_char_p
    L.conv3p_last_cuda_error.restype = C.c_char_p
    L.conv3p_abi_version.restype = i
    L.conv3p_launch_count.argtypes = [i]
    L.conv3p_launch_count.restype = ll
    L.conv3p_set_engine.argtypes = [i]
    L.conv3p_set_engine.restype = i
    L.conv3p_profile_enable.argtypes = [i]
    L.conv3p_profile_enable.restype = i
    L.conv3p_profile_read.argtypes = [C.c_char_p, sz]
    L.conv3p_profile_read.restype = ll
    L.conv3p_debug_phase_cycles.argtypes = [vp]
    L.conv3p_debug_cta_cycles.argtypes = [vp]
    L.conv3p_debug_w2_cycles.argtypes = [vp]
    L.conv3p_debug_mma_rate.argtypes = [i, i, i, vp, vp]
    L.conv3p_selftest_tc.argtypes = [vp, vp, vp, i, i, i, vp]
    L.conv3p_selftest_tc_mn.argtypes = [vp, vp, vp, i, i, i, vp]
    L.conv3p_augment_rotate_jitter_f32.argtypes = [vp, vp, vp, C.c_double, C.c_double, i, i, vp, vp]
    L.conv3p_xyz_sort_workspace_bytes.argtypes = [i, i]
    L.conv3p_xyz_sort_workspace_bytes.restype = sz
    L.conv3p_xyz_sort_f32.argtypes = [vp, i, vp, i, i, i, vp, vp, vp, vp, sz, vp]


def lib():
    """The loaded library.  Raises (never falls back) when it is missing and cannot be built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = os.environ.get("CONV3P_LIB") or _build.LIB_PATH      # CONV3P_LIB: an experiment variant
                if not os.path.exists(path):
                    try:
                        _build.build()
                    except Exception as e:  # pragma: no cover - depends on the toolchain
                        raise RuntimeError(
                            f"libconv3p_b200.so is missing and could not be built ({e}); "
                            "pointwise_b200 has no CPU fallback") from e
                L = C.CDLL(path)
                _declare(L)
                _lib = L
    return _lib


def check(status: int) -> None:
    if status == OK:
        return
    L = lib()
    msg = L.conv3p_status_string(status).decode()
    if status == ERR_CUDA:
        msg += ": " + L.conv3p_last_cuda_error().decode()
    raise Conv3pError(status, msg)


def make_geom(B: int, N: int, stride, voxel_size: float, capacity: int) -> Geom:
    g = Geom()
    g.B, g.N = int(B), int(N)
    g.stride[0], g.stride[1], g.stride[2] = (int(s) for s in stride)
    g.voxel_size = float(voxel_size)
    g.pair_capacity = int(capacity)
    return g
