"""CPU tests of host-side logic added in round 2: the benchmark's byte / flop models and workload tables, the plan
layout with the bin-offset table, workspace queries of the general filter path, and the collective helpers when no
process group exists.  No GPU, no compute calls."""
import ctypes as C
import math

import pytest

import bench
from pointwise_b200 import _lib, distributed, nets


def test_sweep_is_the_nine_point_grid_of_baseline_configs3():
    assert len(bench.SWEEP) == 9
    seen = set()
    for name in bench.SWEEP:
        clouds, N, Cin, Cout, stride, dist = bench.WORKLOADS[name]
        assert clouds * N == 1 << 18 and Cin == Cout and stride == (1, 1, 1)
        seen.add((N, Cin))
    assert seen == {(n, c) for n in (1024, 4096, 16384) for c in (9, 64, 256)}
    assert bench.WORKLOADS["headline"][:4] == (64, 4096, 64, 128)
    assert bench.WORKLOADS["headline_b16"][0] == 16          # BASELINE configs[4]: 128 clouds over 8 GPUs
    assert set(bench.NETS) == {"seg_net", "cls_net"}


def test_byte_models_follow_survey_8d():
    pts, Cin, Cout, K = 1000, 64, 128, 47.6
    fwd = bench.algorithmic_bytes("k_forward_tc", pts, Cin, Cout, K, K)
    assert math.isclose(fwd, pts * (K * (4 * Cin + 4) + 27 * 4 + 24 + 4 * Cout) + 27 * Cin * Cout * 4)
    bwd = bench.algorithmic_bytes("k_backward_input_tc", pts, Cin, Cout, K, K)
    shared = bench.algorithmic_bytes("k_backward_input_tc", pts, Cin, Cout, K, K, nbins_b=11.2, shared=True)
    assert math.isclose(shared - bwd, pts * 11.2 * 4 * Cout)          # the G-store rows it writes
    store = bench.algorithmic_bytes("k_backward_filter_tc", pts, Cin, Cout, K, K, nbins_b=11.2, shared=True)
    assert store < bench.algorithmic_bytes("k_backward_filter_tc", pts, Cin, Cout, K, K)   # rows instead of lists
    assert bench.compulsory_bytes("k_forward_tc", pts, Cin, Cout) == 4 * (3 + Cin + Cout) * pts + 27 * Cin * Cout * 4
    assert bench.is_contraction("k_small_backward_filter") and not bench.is_contraction("k_neighbor_search")


def test_roofline_block_reports_three_labelled_fractions():
    kern = {"k_forward_tc": (20, 14.0), "k_neighbor_search": (20, 6.0)}
    pk = dict(hbm_gbs=6552.6, bf16_tflops=1674.4, source="measured", tf32_tflops=767.0, tf32_source="cublas")
    traffic = {"k_forward_tc": {"dram_bytes_per_launch": 2.7e8, "binder": "l1/shared-memory data pipe (85 % busy under ncu)"}}
    r = bench.roofline_block(kern, 262144, 64, 128, 47.6, 47.6, 11.2, 11.2, False, pk, traffic)
    assert r["kernel"] == "k_forward_tc" and r["bound"].startswith("l1/shared")
    assert r["frac"] == r["frac_G"] and "gather model" in r["frac_label"]
    assert 0 < r["frac_A"] < r["frac_G"] and 0 < r["frac_F"] < 1
    assert "frac_fp32_simt_peak_74.4" not in r


def test_plan_layout_holds_the_bin_offset_table():
    L = _lib.lib()
    lay = _lib.PlanLayout()
    for B, N, cap_cells in [(4, 100, 4096), (4, 1024, 16384), (2, 4096, 65536), (1, 100000, 65536)]:
        g = _lib.make_geom(B, N, (1, 1, 1), 0.1, 64 * B * N)
        assert L.conv3p_plan_layout(g, lay) == _lib.OK
        assert lay.cell_start > lay.sort_tmp and lay.cell_start % 256 == 0
        assert lay.total_bytes - lay.cell_start >= 4 * B * (cap_cells + 1)
        assert lay.total_bytes == L.conv3p_plan_bytes(g)


def test_general_filter_workspace_queries():
    L = _lib.lib()
    i3 = C.c_int * 3
    g = _lib.make_geom(2, 300, (1, 2, 3), 0.1, 5000)
    small = L.conv3p_op_workspace_bytes_ex(g, i3(1, 1, 1), 4, 4, 0)
    big = L.conv3p_op_workspace_bytes_ex(g, i3(5, 5, 5), 4, 4, 0)
    assert 0 < small < big
    assert L.conv3p_op_workspace_bytes_ex(g, i3(8, 8, 9), 4, 4, 0) == 0          # 576 cells > 512
    assert L.conv3p_op_workspace_bytes_ex(g, i3(8, 8, 8), 4, 4, 0) > 0           # exactly 512
    assert L.conv3p_op_workspace_bytes_ex(g, i3(0, 3, 3), 4, 4, 0) == 0


def test_double_and_padded_scratch_queries():
    """Host-side size queries only (no GPU): the double operator's workspace, and the scratch of shapes that run
    zero-padded on the tensor-core kernels (api.cu, pad_channels) covers the padded copies and the padded shape's own
    scratch; engine flag 2048 (no padding) gives the smaller figure back."""
    L = _lib.lib()
    i3 = C.c_int * 3
    g = _lib.make_geom(16, 4096, (1, 1, 1), 0.1, 4_000_000)
    f32 = L.conv3p_op_workspace_bytes_ex(g, i3(3, 3, 3), 9, 9, 1)
    f64 = L.conv3p_op_workspace_bytes_f64(g, i3(3, 3, 3), 9, 9)
    assert f64 > 0 and f32 > 0
    assert L.conv3p_op_workspace_bytes_f64(g, i3(9, 9, 9), 9, 9) == 0
    pts = 16 * 4096
    padded = L.conv3p_scratch_bytes(g, 36, 13)
    assert padded >= L.conv3p_scratch_bytes(g, 64, 32) + 4 * pts * (32 + 64 + 64)      # g_pad, x_pad, gi_pad + inner
    assert L.conv3p_backward_scratch_bytes(g, 36, 13) >= padded + 4 * pts * 27 * 32     # + the padded shape's G store
    big = _lib.make_geom(64, 4096, (1, 1, 1), 0.1, 16_000_000)
    tiny_small, tiny_big = L.conv3p_scratch_bytes(g, 9, 9), L.conv3p_scratch_bytes(big, 9, 9)
    prev = L.conv3p_set_engine(2048)
    try:
        assert L.conv3p_scratch_bytes(g, 36, 13) < padded
        # tiny shapes are padded only from 128k points up
        assert L.conv3p_scratch_bytes(g, 9, 9) == tiny_small
        assert L.conv3p_scratch_bytes(big, 9, 9) < tiny_big
    finally:
        L.conv3p_set_engine(prev)
    assert tiny_big >= 4 * 64 * 4096 * (32 + 32 + 32)


def test_collective_helpers_without_a_process_group():
    import torch
    t = torch.ones(3)
    assert distributed.allreduce_grad_filter(t) is None
    assert distributed.allreduce_grad_filter_overlapped(t) is None
    assert torch.equal(t, torch.ones(3))
    assert nets.SHARE_PLANS is True


def test_exports_cover_the_header():
    """Every function the header declares is in _lib.EXPORTS (tests/test_abi.py checks they are exported)."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "conv3p_b200.h")).read()
    declared = set(re.findall(r"\b(conv3p_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"conv3p_b200"}
    missing = sorted(d for d in declared if d not in _lib.EXPORTS)
    assert not missing, missing
