"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU and
exports every symbol include/conv3p_b200.h declares; size/layout queries and argument checks (host
logic only -- no compute call is made here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from pointwise_b200 import _lib
    return _lib.lib()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "conv3p_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(conv3p_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(L):
    from pointwise_b200 import _lib
    names = declared_functions()
    assert len(names) >= 23
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/conv3p_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"


def test_library_is_sm100a_native():
    import subprocess
    from pointwise_b200 import build
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", build.LIB_PATH],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sizes_and_layout(L):
    from pointwise_b200 import _lib
    g = _lib.make_geom(16, 4096, (1, 1, 1), 0.1, 16 * 4096 * 64)
    lay = _lib.PlanLayout()
    assert L.conv3p_plan_layout(g, lay) == 0
    assert lay.total_bytes == L.conv3p_plan_bytes(g) > 0
    offs = [getattr(lay, n) for n, _ in _lib.PlanLayout._fields_][:-1]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert lay.pair_begin - lay.count_table >= 16 * 4096 * 27 * 4
    assert L.conv3p_op_workspace_bytes(g, 64, 128) >= lay.total_bytes + 27 * 64 * 128 * 4
    assert L.conv3p_host_workspace_bytes(g, 64, 128) > L.conv3p_op_workspace_bytes(g, 64, 128)
    # the G store (27 * Cout floats per point) only where both gradient kernels are the tensor-core ones
    extra = L.conv3p_backward_scratch_bytes(g, 64, 128) - L.conv3p_scratch_bytes(g, 64, 128)
    assert extra >= 16 * 4096 * 27 * 128 * 4
    assert L.conv3p_op_backward_workspace_bytes(g, 64, 128) - L.conv3p_op_workspace_bytes(g, 64, 128) == extra
    assert L.conv3p_backward_scratch_bytes(g, 9, 9) == L.conv3p_scratch_bytes(g, 9, 9)
    g0 = _lib.make_geom(0, 0, (1, 1, 1), 0.1, 0)
    assert L.conv3p_plan_bytes(g0) > 0


@pytest.mark.parametrize("stride,voxel,cap", [((0, 1, 1), 0.1, 10), ((1, 1, 1), 0.0, 10),
                                              ((1, 1, 1), -1.0, 10), ((1, 1, 1), 0.1, -5)])
def test_bad_geometry_is_a_status_not_a_crash(L, stride, voxel, cap):
    from pointwise_b200 import _lib
    g = _lib.make_geom(2, 8, stride, voxel, cap)
    assert L.conv3p_plan_bytes(g) == 0
    assert L.conv3p_plan_build_f32(g, None, None, 0, None) == _lib.ERR_INVALID_ARGUMENT


def test_buffer_checks_precede_any_cuda_call(L):
    from pointwise_b200 import _lib
    g = _lib.make_geom(2, 8, (1, 1, 1), 0.1, 1000)
    assert L.conv3p_plan_build_f32(g, None, None, 0, None) == _lib.ERR_BUFFER_TOO_SMALL
    i3 = C.c_int * 3
    st = L.conv3p_op_forward_f32(None, None, None, i3(9, 9, 9), i3(1, 1, 1), 0.1, 2, 8, 4, 4, 100, None,
                                 None, 0, None)
    assert st == _lib.ERR_UNSUPPORTED           # more than 512 cells
    st = L.conv3p_op_forward_f32(None, None, None, i3(3, 3, 5), i3(1, 1, 1), 0.1, 2, 8, 4, 4, 100, None,
                                 None, 0, None)
    assert st == _lib.ERR_BUFFER_TOO_SMALL      # general filter path: workspace checked before any CUDA call
    assert L.conv3p_op_workspace_bytes_ex(g, i3(3, 3, 5), 4, 4, 0) > L.conv3p_plan_bytes(g)
    assert L.conv3p_op_workspace_bytes_ex(g, i3(3, 3, 3), 4, 4, 1) == L.conv3p_op_backward_workspace_bytes(g, 4, 4)
    assert L.conv3p_op_workspace_bytes_ex(g, i3(9, 9, 9), 4, 4, 0) == 0
    st = L.conv3p_op_forward_f32(None, None, None, i3(3, 3, 3), i3(1, 1, 1), 0.1, 2, 8, 4, 4, 100, None,
                                 None, 0, None)
    assert st == _lib.ERR_BUFFER_TOO_SMALL
    assert b"too small" in L.conv3p_status_string(_lib.ERR_BUFFER_TOO_SMALL)
    assert L.conv3p_abi_version() == 1


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing in the product imports the oracle."""
    import torch
    from pointwise_b200 import conv3p
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv3p(torch.zeros(1, 4, 3), torch.zeros(1, 4, 2), torch.zeros(3, 3, 3, 2, 2), [1, 1, 1], [0.1])
    pkg = os.path.join(ROOT, "pointwise_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f


def test_host_argument_parsing():
    import numpy as np
    import torch
    from pointwise_b200.ops import parse_stride, parse_voxel
    assert parse_stride([1, 2, 3]) == (1, 2, 3)
    assert parse_stride(torch.tensor([2, 2, 2], dtype=torch.int32)) == (2, 2, 2)
    assert parse_stride(np.array([4, 4, 4], np.int32)) == (4, 4, 4)
    assert parse_stride(3) == (3, 3, 3)
    assert parse_voxel([0.1]) == pytest.approx(0.1)
    assert parse_voxel(torch.tensor([0.1])) == pytest.approx(0.1)
    assert parse_voxel(0.25) == 0.25
    for bad in ([1, 1], [1, 1, 1, 1], torch.tensor([1])):
        with pytest.raises(ValueError, match="stride tensor to have size 3"):
            parse_stride(bad)
    with pytest.raises(ValueError, match="voxel tensor to have dimension 1"):
        parse_voxel([0.1, 0.2])
    with pytest.raises(ValueError):
        parse_stride([1, 0, 1])
    with pytest.raises(ValueError):
        parse_voxel([0.0])


def test_row_stride_of_channel_slices():
    """Host logic of the strided-row forward (conv3p_forward_ex_f32): which tensor layouts are passed through as
    (pointer, row stride) and which are made contiguous."""
    import torch
    from pointwise_b200.ops import row_stride_of
    B, N, W = 3, 5, 36
    buf = torch.zeros(B, N, W)
    def rs(t):
        return row_stride_of(tuple(t.shape), t.stride(), B, N)
    assert rs(buf) == W
    assert rs(buf[:, :, 9:18]) == W                      # a channel slice of the concat buffer
    assert rs(buf[:, :, 35:36]) == W                     # one channel
    assert rs(torch.zeros(B, N, 9)) == 9                 # dense
    assert rs(buf[:, ::2, :9]) is None                   # wrong N (and rows not uniformly spaced for this N)
    assert rs(buf.transpose(1, 2)) is None               # shape mismatch
    assert rs(torch.zeros(B, N, 9).transpose(0, 1).contiguous().transpose(0, 1)) is None   # batch stride not N * row
    assert rs(buf[:, :, ::2]) is None                    # channels not contiguous
    assert row_stride_of((1, N, 9), buf[:1, :, :9].stride(), 1, N) == W      # B == 1: batch stride is free
    assert row_stride_of((B, N, 0), (0, 0, 1), B, N) is None                 # empty
