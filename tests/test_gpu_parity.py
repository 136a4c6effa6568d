"""GPU parity tests: the CUDA path (through the C ABI) against the CPU checker on identical inputs.

Bar (BASELINE.json north_star): neighbour indices BIT-EXACT; features within a stated fp32 tolerance.
Feature tolerance used here:  |got - truth| <= ATOL + RTOL * sum|terms|  with truth and sum|terms|
accumulated in float64 by the oracle over the reference's own pair set.  RTOL = 1e-5 is the size of
the reference's own fp32 summation error (it adds ~K*Cin rounded terms sequentially), so "within
tolerance" means "as close to the exact sum as the reference itself".
"""
import numpy as np
import pytest
import torch

from helpers import assert_close_scaled, pair_sets
from pointwise_b200.synth import make_points, make_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL = 1e-7
V = 0.1


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_lists(plan):
    """-> count[B,N,27], and per cloud (off, j_local, f) reconstructed from the cell-grouped lists."""
    cnt = plan.count_table.cpu().numpy()
    begin = plan.pair_begin.cpu().numpy()
    length = plan.pair_len.cpu().numpy()
    rows = plan.pair_row.cpu().numpy()
    B, N = plan.B, plan.N
    clouds = []
    for b in range(B):
        off = np.zeros(N + 1, np.int64)
        js, fs = [], []
        for i in range(N):
            seg = rows[begin[b, i]:begin[b, i] + length[b, i]]
            assert length[b, i] == cnt[b, i].sum()
            assert ((seg >= b * N) & (seg < (b + 1) * N)).all(), "neighbour from another cloud"
            js.append(seg - b * N)
            fs.append(np.repeat(np.arange(27), cnt[b, i]))
            off[i + 1] = off[i] + length[b, i]
        clouds.append((off, np.concatenate(js) if js else np.zeros(0, np.int64),
                       np.concatenate(fs) if fs else np.zeros(0, np.int64)))
    return cnt, clouds


INDEX_CASES = [
    # B, N, stride, dist, quantise
    (4, 1024, (1, 1, 1), "sphere", None),
    (2, 1024, (2, 2, 2), "sphere", None),
    (2, 1000, (3, 3, 3), "room", None),
    (2, 777, (4, 4, 4), "room", None),
    (2, 2048, (1, 1, 1), "room", 0.05),     # bin-edge ties: asymmetry stress
    (2, 1024, (2, 2, 2), "cube", 0.05),
    (1, 4096, (1, 1, 1), "room", None),
    (2, 600, (1, 2, 3), "room", None),      # anisotropic stride
    (1, 1, (1, 1, 1), "cube", None),        # single point
    (3, 33, (4, 4, 4), "cube", 0.1),
]


@pytest.mark.parametrize("B,N,stride,dist,q", INDEX_CASES)
def test_neighbor_indices_bit_exact(checker, B, N, stride, dist, q):
    from pointwise_b200 import NeighborPlan
    pts = make_points(B, N, dist, seed=3, quantise=q)
    plan = NeighborPlan(dev(pts), stride, V)
    cnt, clouds = gpu_lists(plan)
    assert plan.stats.total_pairs == cnt.sum()
    for b in range(B):
        want_cnt = checker.neighbor_count(pts[b], stride, V)
        assert np.array_equal(cnt[b], want_cnt), f"count table differs in cloud {b}"
        off, j, f = checker.neighbors(pts[b], stride, V)
        goff, gj, gf = clouds[b]
        assert np.array_equal(off, goff)
        for i, (a, c) in enumerate(zip(pair_sets(off, j, f, N), pair_sets(goff, gj, gf, N))):
            assert np.array_equal(a, c), f"pair set differs at cloud {b} point {i}"


def test_dense_neighbourhood_uses_resweep(checker):
    """600 identical points + 400 spread ones: K = 600 per point exceeds the shared-memory stash."""
    from pointwise_b200 import NeighborPlan
    rng = np.random.default_rng(0)
    pts = np.concatenate([np.full((600, 3), 0.25, np.float32),
                          rng.uniform(-1, 1, (400, 3)).astype(np.float32)])[None]
    plan = NeighborPlan(dev(pts), 1, V)
    cnt, clouds = gpu_lists(plan)
    assert np.array_equal(cnt[0], checker.neighbor_count(pts[0], 1, V))
    off, j, f = checker.neighbors(pts[0], 1, V)
    for a, c in zip(pair_sets(off, j, f, 1000), pair_sets(*clouds[0], 1000)):
        assert np.array_equal(a, c)
    assert cnt[0, :600, 13].min() >= 600


def test_far_from_origin_and_tiny_voxel(checker):
    from pointwise_b200 import NeighborPlan
    pts = make_points(2, 500, "cube", seed=9) * 0.02 + np.float32(100.0)
    for voxel, stride in [(0.001, 1), (0.003, 2)]:
        plan = NeighborPlan(dev(pts), stride, voxel)
        cnt = plan.count_table.cpu().numpy()
        for b in range(2):
            assert np.array_equal(cnt[b], checker.neighbor_count(pts[b], stride, voxel))


def test_backward_lists_follow_reference_rule(port):
    """Backward pair multiset == {(j, ii, f') : ii in N(j), f' not a hole, count(ii,f') > 0}."""
    from pointwise_b200 import NeighborPlan
    B, N, stride = 2, 1024, (1, 1, 1)
    pts = make_points(B, N, "room", seed=5, quantise=0.05)
    plan = NeighborPlan(dev(pts), stride, V).ensure_backward()
    bc = plan.backward_count_table.cpu().numpy()
    begin = plan.pair_begin.cpu().numpy()
    brow = plan.backward_row.cpu().numpy()
    bw = plan.backward_weight.cpu().numpy()
    n_asym = 0
    for b in range(B):
        cnt = port.neighbor_count(pts[b], stride, V)
        off, j, f = port.neighbors(pts[b], stride, V)
        for jj in range(N):
            want = []
            for ii in j[off[jj]:off[jj + 1]]:
                lo = (pts[b, ii].astype(np.float64) - 3 * 0.5 * np.float64(np.float32(V))).astype(np.float32)
                c = np.minimum(2, ((pts[b, jj] - lo) / np.float32(V)).astype(np.int32))
                fp = (c[2] * 3 + c[1]) * 3 + c[0]
                if cnt[ii, fp] > 0:
                    want.append((fp, ii, cnt[ii, fp]))
                else:
                    n_asym += 1
            k = bc[b, jj].sum()
            seg = brow[begin[b, jj]:begin[b, jj] + k] - b * N
            wts = bw[begin[b, jj]:begin[b, jj] + k]
            fs = np.repeat(np.arange(27), bc[b, jj])
            got = sorted(zip(fs.tolist(), seg.tolist(), np.rint(1.0 / wts).astype(int).tolist()))
            assert got == sorted((int(a), int(c), int(d)) for a, c, d in want), (b, jj)
    assert plan.read_stats().backward_pairs == bc.sum()
    assert n_asym > 0, "quantised cloud should exercise the count==0 skip"


FEATURE_CASES = [
    # B, N, Cin, Cout, stride, dist, quantise
    (4, 1024, 9, 9, (1, 1, 1), "sphere", None),     # BASELINE config 1
    (2, 1024, 3, 9, (1, 1, 1), "sphere", None),     # ModelNet40 layer 1
    (2, 1024, 9, 9, (2, 2, 2), "sphere", None),
    (2, 1024, 9, 9, (3, 3, 3), "room", None),
    (2, 1024, 9, 9, (4, 4, 4), "room", None),
    (2, 2048, 36, 13, (1, 1, 1), "room", None),     # S3DIS layer 5
    (2, 2048, 64, 128, (1, 1, 1), "room", None),    # headline shape
    (1, 1500, 64, 128, (1, 1, 1), "room", 0.05),    # asymmetric pairs, N not a tile multiple
    (2, 512, 5, 7, (1, 2, 3), "cube", 0.05),
    (1, 300, 40, 300, (1, 1, 1), "room", None),     # Cout beyond one column block
    (1, 16384, 9, 9, (1, 1, 1), "room", None),      # BASELINE sweep: N=16k (K ~ 190 per point)
    (1, 4096, 64, 64, (2, 2, 2), "room", None),     # BASELINE sweep: C=64, dilated
    (1, 1024, 256, 256, (1, 1, 1), "room", None),   # BASELINE sweep: C=256
    (1, 1, 4, 4, (1, 1, 1), "cube", None),
    # BASELINE configs[1] at full size: the ModelNet40 classification layers, B=32 x 1024 (pointcnn2_acsd.py:48-67)
    (32, 1024, 3, 9, (1, 1, 1), "sphere", None),
    (32, 1024, 9, 9, (2, 2, 2), "sphere", None),
    (32, 1024, 9, 9, (3, 3, 3), "sphere", None),
    (32, 1024, 9, 9, (4, 4, 4), "sphere", None),
    # BASELINE configs[2] at full size: the S3DIS segmentation layers, B=16 x 4096 (pointcnn_scene_seg_acsd.py:51-57)
    (16, 4096, 9, 9, (1, 1, 1), "room", None),
    (16, 4096, 9, 9, (2, 2, 2), "room", None),
    (16, 4096, 9, 9, (3, 3, 3), "room", None),
    (16, 4096, 9, 9, (4, 4, 4), "room", None),
    (16, 4096, 36, 13, (1, 1, 1), "room", None),
]


@pytest.mark.parametrize("B,N,Cin,Cout,stride,dist,q", FEATURE_CASES)
def test_forward_backward_parity(port, checker, B, N, Cin, Cout, stride, dist, q):
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    pr = make_problem(B, N, Cin, Cout, dist, seed=2, quantise=q)
    plan = NeighborPlan(dev(pr["points"]), stride, V)
    out = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"])).cpu().numpy()
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    assert_close_scaled(out, o64, oabs, RTOL, ATOL, "forward")
    assert_close_scaled(o32, o64, oabs, RTOL, ATOL, "oracle fp32 vs fp64 (tolerance sanity)")

    gi, gf = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    assert_close_scaled(gi.cpu().numpy(), r[2], r[3], RTOL, ATOL, "grad_input")
    assert_close_scaled(gf.cpu().numpy(), r[4], r[5], RTOL, ATOL, "grad_filter")
    if checker.kind == "reference" and B * N * Cin * Cout <= 2048 * 36 * 13 * 2:
        # the reference's own object code agrees with the port it pins
        assert np.array_equal(checker.forward(pr["points"], pr["input"], pr["filter"], stride, V), o32)


KATS = [
    # points (x,y,z), input, grad, stride -> out, grad_in, {cell: grad_W}     
    ([(0, 0, 0)], [2], [1], 1, [26], [13], {13: 2}),
    ([(0, 0, 0), (0.1, 0, 0)], [2, 3], [1, 10], 1, [68, 63], [133, 144], {12: 20, 13: 32, 14: 3}),
    ([(0, 0, 0), (0.15, 0, 0)], [2, 3], [1, 10], 1, [68, 39], [13, 130], {13: 32}),
    ([(0, 0, 0), (0.16, 0, 0)], [2, 3], [1, 10], 1, [26, 39], None, None),
    ([(0, 0, 0), (0.06, 0, 0), (0.07, 0, 0)], [2, 3, 5], [1, 10, 100], 1, [82, 76, 76],
     [1333, 722, 722], {12: 220, 13: 442, 14: 4}),
    ([(0, 0, 0), (0.1, 0, 0), (0.2, 0, 0)], [2, 3, 5], [1, 10, 100], 2, [96, 39, 89],
     [1213, 130, 1314], {12: 200, 13: 532, 14: 5}),
    ([(0, 0, 0), (0.1, 0.1, 0.1)], [2, 3], [1, 10], 1, [104, 39], None, {0: 20, 13: 32, 26: 3}),
]


@pytest.mark.parametrize("k", range(len(KATS)))
def test_known_answers(k):
    from pointwise_b200 import conv3p
    pts, inp, g, stride, out, gin, gw = KATS[k]
    P = torch.tensor([pts], dtype=torch.float32).cuda()
    X = torch.tensor(inp, dtype=torch.float32).view(1, -1, 1).cuda().requires_grad_()
    W = torch.arange(27, dtype=torch.float32).view(3, 3, 3, 1, 1).cuda().requires_grad_()
    y = conv3p(P, X, W, [stride] * 3, [0.1])
    assert y.flatten().tolist() == [float(v) for v in out]
    y.backward(torch.tensor(g, dtype=torch.float32).view(1, -1, 1).cuda())
    if gin is not None:
        assert X.grad.flatten().tolist() == [float(v) for v in gin]
    if gw is not None:
        want = np.zeros(27)
        for c, v in gw.items():
            want[c] = v
        assert W.grad.flatten().tolist() == want.tolist()


def test_autograd_signature_and_gradients(port):
    from pointwise_b200 import conv3p
    pr = make_problem(2, 512, 9, 9, "sphere", seed=4)
    P = dev(pr["points"]).requires_grad_()
    X = dev(pr["input"]).requires_grad_()
    W = dev(pr["filter"]).requires_grad_()
    stride = torch.tensor([2, 2, 2], dtype=torch.int32)      # host tensors, as tf.constant in the models
    voxel = torch.tensor([0.1])
    y = conv3p(P, X, W, stride, voxel)
    assert y.shape == (2, 512, 9)
    y.backward(dev(pr["grad_out"]))
    assert P.grad is None                                      # no gradient to points (pointcnn2_acsd.py:31)
    gi, gf, gi64, gia, gf64, gfa = port.backward(pr["grad_out"], pr["points"], pr["input"],
                                                 pr["filter"], 2, V, with64=True)
    assert_close_scaled(X.grad.cpu().numpy(), gi64, gia, RTOL, ATOL, "autograd grad_input")
    assert_close_scaled(W.grad.cpu().numpy(), gf64, gfa, RTOL, ATOL, "autograd grad_filter")


def test_plan_reuse_and_determinism():
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    pr = make_problem(2, 2048, 16, 16, "room", seed=6)
    P, X, W, G = (dev(pr[k]) for k in ("points", "input", "filter", "grad_out"))
    a = NeighborPlan(P, 1, V)
    b = NeighborPlan(P, 1, V)
    ya, yb = conv3p_forward(a, X, W), conv3p_forward(b, X, W)
    assert torch.equal(ya, yb), "forward must not depend on where the lists landed in memory"
    ga, gb = conv3p_backward(a, G, X, W), conv3p_backward(b, G, X, W)
    assert torch.equal(ga[0], gb[0]) and torch.equal(ga[1], gb[1]), "backward must be deterministic"


def test_reference_error_messages():
    from pointwise_b200 import conv3p
    P = torch.zeros(2, 8, 3).cuda()
    X = torch.zeros(2, 8, 4).cuda()
    W = torch.zeros(3, 3, 3, 4, 5).cuda()
    with pytest.raises(ValueError, match="points shape"):
        conv3p(P.view(16, 3), X, W, [1, 1, 1], [0.1])
    with pytest.raises(ValueError, match="same batch size"):
        conv3p(P, X[:1], W, [1, 1, 1], [0.1])
    with pytest.raises(ValueError, match="same number of points"):
        conv3p(P, X[:, :4], W, [1, 1, 1], [0.1])
    with pytest.raises(ValueError, match="filter channels"):
        conv3p(P, X, torch.zeros(3, 3, 3, 3, 5).cuda(), [1, 1, 1], [0.1])
    with pytest.raises(ValueError, match="stride tensor to have size 3"):
        conv3p(P, X, W, [1, 1], [0.1])
    with pytest.raises(ValueError, match="voxel tensor to have dimension 1"):
        conv3p(P, X, W, [1, 1, 1], [0.1, 0.2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv3p(P.cpu(), X, W, [1, 1, 1], [0.1])


def test_explicit_capacity_overflow_is_reported():
    from pointwise_b200 import Conv3pError, NeighborPlan
    pts = dev(make_points(1, 512, "sphere", seed=1))
    with pytest.raises(Conv3pError, match="capacity"):
        NeighborPlan(pts, 1, V, capacity=600)
    # unchecked plans poison instead of raising
    from pointwise_b200 import conv3p_forward
    plan = NeighborPlan(pts, 1, V, capacity=600, check=False)
    y = conv3p_forward(plan, torch.ones(1, 512, 4).cuda(), torch.ones(3, 3, 3, 4, 4).cuda())
    assert torch.isnan(y).any() and not torch.isnan(y).all()


def test_full_size_properties():
    """BASELINE headline size (N=4096, 64->128): size-independent properties instead of the oracle."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    B, N, Cin, Cout = 16, 4096, 64, 128
    pr = make_problem(B, N, Cin, Cout, "room", seed=0)
    P, X, W, G = (dev(pr[k]) for k in ("points", "input", "filter", "grad_out"))
    plan = NeighborPlan(P, 1, V)
    cnt = plan.count_table
    assert int(cnt.sum()) == plan.stats.total_pairs
    assert int(cnt[:, :, 13].min()) >= 1, "every point is its own neighbour in the centre cell"
    # mirror symmetry of the counts: pairs in cell f of i <-> cell 26-f of j (up to edge rounding)
    per_cell = cnt.sum(dim=(0, 1)).double()
    assert torch.allclose(per_cell, per_cell.flip(0), rtol=1e-3)
    # linearity in the input and in the filter
    y1 = conv3p_forward(plan, X, W)
    X2 = torch.randn_like(X)
    y2 = conv3p_forward(plan, X2, W)
    y12 = conv3p_forward(plan, 0.5 * X + 2.0 * X2, W)
    err = (y12 - (0.5 * y1 + 2.0 * y2)).abs().max()
    assert err < 1e-5 * ((0.5 * y1).abs() + (2.0 * y2).abs()).max() * 27, float(err)
    # a constant input and a filter that is constant over (k) reproduces sum_f W[f] over non-empty cells
    ones = torch.ones_like(X)
    Wc = torch.randn(27, 1, Cout, device="cuda").expand(27, Cin, Cout).contiguous().view(3, 3, 3, Cin, Cout)
    yc = conv3p_forward(plan, ones, Wc)
    want = torch.einsum("bnf,fc->bnc", (cnt > 0).float(), Wc.view(27, Cin, Cout).sum(1))
    assert float((yc - want).abs().max()) < 1e-4 * float(want.abs().max())
    # adjoint identities (exact up to the rare non-symmetric edge pairs of continuous data)
    gi, gf = conv3p_backward(plan, G, X, W)
    lhs = (G.double() * y1.double()).sum()
    assert abs(lhs - (gi.double() * X.double()).sum()) <= 1e-3 * abs(lhs) + 1.0
    assert abs(lhs - (gf.double() * W.double()).sum()) <= 1e-3 * abs(lhs) + 1.0


def test_full_size_three_clouds_against_oracle(port):
    """The benchmarked configuration itself (clouds of 4096 points, 64->128): three clouds of a 16-cloud batch
    compared with the oracle -- rows of the full-batch output / grad_input, and grad_filter of a call on exactly
    those clouds (clouds are independent: tf_conv3p_atrous.cpp:456, 622)."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    B, N, Cin, Cout = 16, 4096, 64, 128
    pr = make_problem(B, N, Cin, Cout, "room", seed=0)
    P, X, W, G = (dev(pr[k]) for k in ("points", "input", "filter", "grad_out"))
    plan = NeighborPlan(P, 1, V)
    y = conv3p_forward(plan, X, W)
    gi, _ = conv3p_backward(plan, G, X, W)
    idx = [0, 7, 15]
    sub = {k: (v[idx] if k != "filter" else v) for k, v in pr.items()}
    o32, o64, oabs = port.forward(sub["points"], sub["input"], sub["filter"], 1, V, with64=True)
    r = port.backward(sub["grad_out"], sub["points"], sub["input"], sub["filter"], 1, V, with64=True)
    assert_close_scaled(y[idx].cpu().numpy(), o64, oabs, RTOL, ATOL, "forward (full-size batch)")
    assert_close_scaled(gi[idx].cpu().numpy(), r[2], r[3], RTOL, ATOL, "grad_input (full-size batch)")
    plan3 = NeighborPlan(dev(sub["points"]), 1, V)
    _, gf3 = conv3p_backward(plan3, dev(sub["grad_out"]), dev(sub["input"]), W)
    assert_close_scaled(gf3.cpu().numpy(), r[4], r[5], RTOL, ATOL, "grad_filter (three full-size clouds)")


def test_c_abi_one_shot_and_host_calls(port):
    """The reference-facing C entry points: device one-shot calls and host-buffer calls."""
    import ctypes as C
    from pointwise_b200 import _lib
    L = _lib.lib()
    B, N, Cin, Cout, stride = 2, 700, 9, 13, (2, 2, 2)
    pr = make_problem(B, N, Cin, Cout, "room", seed=8)
    cap = 64 * B * N
    g = _lib.make_geom(B, N, stride, V, cap)
    i3 = (C.c_int * 3)
    # host-buffer calls
    ws = torch.empty(L.conv3p_host_workspace_bytes(g, Cin, Cout), dtype=torch.uint8, device="cuda")
    out = np.zeros((B, N, Cout), np.float32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(L.conv3p_host_forward_f32(p(pr["points"]), p(pr["input"]), p(pr["filter"]), i3(*stride),
                                         V, B, N, Cin, Cout, cap, p(out), C.c_void_p(ws.data_ptr()),
                                         ws.numel(), None))
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    assert_close_scaled(out, o64, oabs, RTOL, ATOL, "host forward")
    gi = np.zeros((B, N, Cin), np.float32)
    gf = np.zeros((3, 3, 3, Cin, Cout), np.float32)
    _lib.check(L.conv3p_host_backward_f32(p(pr["grad_out"]), p(pr["points"]), p(pr["input"]),
                                          p(pr["filter"]), i3(*stride), V, B, N, Cin, Cout, cap, p(gi),
                                          p(gf), C.c_void_p(ws.data_ptr()), ws.numel(), None))
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    assert_close_scaled(gi, r[2], r[3], RTOL, ATOL, "host grad_input")
    assert_close_scaled(gf, r[4], r[5], RTOL, ATOL, "host grad_filter")
    # capacity too small -> status, not a crash
    st = L.conv3p_host_forward_f32(p(pr["points"]), p(pr["input"]), p(pr["filter"]), i3(*stride), V, B,
                                   N, Cin, Cout, 100, p(out), C.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert st == _lib.ERR_PAIR_OVERFLOW
    # unsupported filter size (more than 512 cells) -> status
    P, X, W = dev(pr["points"]), dev(pr["input"]), dev(pr["filter"])
    Y = torch.empty(B, N, Cout, device="cuda")
    ws2 = torch.empty(L.conv3p_op_workspace_bytes(g, Cin, Cout), dtype=torch.uint8, device="cuda")
    st = L.conv3p_op_forward_f32(P.data_ptr(), X.data_ptr(), W.data_ptr(), i3(9, 9, 9), i3(*stride), V, B,
                                 N, Cin, Cout, cap, Y.data_ptr(), ws2.data_ptr(), ws2.numel(), None)
    assert st == _lib.ERR_UNSUPPORTED          # more than 512 cells (other shapes: tests/test_gpu_generic.py)
    _lib.check(L.conv3p_op_forward_f32(P.data_ptr(), X.data_ptr(), W.data_ptr(), i3(3, 3, 3), i3(*stride),
                                       V, B, N, Cin, Cout, cap, Y.data_ptr(), ws2.data_ptr(),
                                       ws2.numel(), None))
    torch.cuda.synchronize()
    assert_close_scaled(Y.cpu().numpy(), o64, oabs, RTOL, ATOL, "one-shot forward")


def test_c_abi_one_shot_backward_with_and_without_shared_gather(port):
    """conv3p_op_backward_f32 at a tensor-core shape: the minimum workspace (each gradient kernel gathers) and the
    larger one of conv3p_op_backward_workspace_bytes (one gather, G store) give the same bits, both within tolerance
    of the oracle."""
    import ctypes as C
    from pointwise_b200 import _lib
    L = _lib.lib()
    B, N, Cin, Cout, stride = 2, 900, 64, 128, (1, 1, 1)
    pr = make_problem(B, N, Cin, Cout, "room", seed=18)
    cap = 96 * B * N
    g = _lib.make_geom(B, N, stride, V, cap)
    i3 = (C.c_int * 3)
    P, X, W, G = dev(pr["points"]), dev(pr["input"]), dev(pr["filter"]), dev(pr["grad_out"])
    res = []
    small, big = L.conv3p_op_workspace_bytes(g, Cin, Cout), L.conv3p_op_backward_workspace_bytes(g, Cin, Cout)
    assert big > small
    for nbytes in (small, big):
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        gi = torch.full((B, N, Cin), float("nan"), device="cuda")
        gf = torch.full((3, 3, 3, Cin, Cout), float("nan"), device="cuda")
        _lib.check(L.conv3p_op_backward_f32(G.data_ptr(), P.data_ptr(), X.data_ptr(), W.data_ptr(), i3(3, 3, 3),
                                            i3(*stride), V, B, N, Cin, Cout, cap, gi.data_ptr(), gf.data_ptr(),
                                            ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        res.append((gi.cpu().numpy(), gf.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    assert_close_scaled(res[1][0], r[2], r[3], RTOL, ATOL, "one-shot grad_input")
    assert_close_scaled(res[1][1], r[4], r[5], RTOL, ATOL, "one-shot grad_filter")


TC_SHAPES = [(64, 128), (32, 64), (64, 64), (32, 128), (128, 32), (96, 48), (128, 128), (256, 256), (64, 256),
             (32, 32), (256, 64)]


def tc_filter_shape(Cin, Cout):
    """Shapes whose weight gradient runs on tensor cores (backward_filter2.cu: M = 128 lanes = Cout block or stacked
    cells, N = Cin in 32-wide panels)."""
    return Cout in (32, 64, 128, 256) and Cin % 32 == 0 and 32 <= Cin <= 256


@pytest.mark.parametrize("Cin,Cout", TC_SHAPES)
def test_tensor_core_engine_matches_oracle_and_simt(port, Cin, Cout):
    """The tcgen05 3xTF32 engine and the fp32 SIMT engine both sit inside the same tolerance (forward, input
    gradient and weight gradient), and the tensor-core engine is really the one selected."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward, set_engine
    B, N, stride = 2, 1100, (1, 1, 1)          # 2200 points: several tiles, a ragged tail
    if Cin * Cout >= 256 * 256:
        N = 500                                # keeps the CPU checker in seconds
    pr = make_problem(B, N, Cin, Cout, "room", seed=12, quantise=0.05 if Cin == 64 and Cout == 64 else None)
    plan = NeighborPlan(dev(pr["points"]), stride, V)
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    outs, gins, gfs = {}, {}, {}
    for eng in ("simt", "tc", "tile"):
        prev = set_engine(eng)
        try:
            outs[eng] = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"])).cpu().numpy()
            gi_, gf_ = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
            gins[eng], gfs[eng] = gi_.cpu().numpy(), gf_.cpu().numpy()
        finally:
            set_engine(prev)
        assert assert_close_scaled(outs[eng], o64, oabs, RTOL, ATOL, f"forward[{eng}]") < RTOL
        assert assert_close_scaled(gins[eng], r[2], r[3], RTOL, ATOL, f"grad_input[{eng}]") < RTOL
        assert assert_close_scaled(gfs[eng], r[4], r[5], RTOL, ATOL, f"grad_filter[{eng}]") < RTOL
    assert not np.array_equal(outs["simt"], outs["tc"]), "tensor-core engine was not selected (forward)"
    assert not np.array_equal(outs["tile"], outs["tc"]), "engine 'tile' must not use tensor cores"
    # (96, 48) reaches the gradient kernels through channel padding (Cout -> 64), the others as they are
    assert not np.array_equal(gins["simt"], gins["tc"]), "tensor-core engine was not selected (grad_input)"
    assert not np.array_equal(gfs["simt"], gfs["tc"]), "tensor-core engine was not selected (grad_filter)"


# Shapes the tensor-core kernels do not take as they are: zero-padded channels (api.cu, pad_channels).  36 -> 13 is the
# segmentation network's last layer (scene_seg/pointcnn_scene_seg_acsd.py:56-57).
# (Cin, Cout, forward padded, backward padded): 64 -> 48 is a tensor-core forward shape as it is; 33 -> 8 pads to
# 64 x 16 forward but would need 64 x 32 backward, more than 6x the real product.
PADDED_SHAPES = [(36, 13, True, True), (40, 24, True, True), (20, 70, True, True), (64, 48, False, True),
                 (100, 100, True, True), (17, 33, True, True), (33, 8, True, False)]


@pytest.mark.parametrize("Cin,Cout,fwd_padded,bwd_padded", PADDED_SHAPES)
def test_channel_padding_onto_tensor_cores_matches_oracle(port, Cin, Cout, fwd_padded, bwd_padded):
    """Padded shapes stay inside the operator's tolerance, really take the tensor-core kernels (results differ from
    the fp32 engines', which engine flag 2048 selects), and the padding leaks nothing into the results."""
    from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward
    B, N, stride = 2, 1100, (1, 2, 1)
    pr = make_problem(B, N, Cin, Cout, "room", seed=21)
    plan = NeighborPlan(dev(pr["points"]), stride, V)
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    L = _lib.lib()
    res = {}
    for flags in (0, 2048):
        prev = L.conv3p_set_engine(flags)
        try:
            y = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"]))
            gi, gf = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
            gi_only, none = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]),
                                            need_filter_grad=False)
            res[flags] = (y.cpu().numpy(), gi.cpu().numpy(), gf.cpu().numpy())
            assert none is None and torch.equal(gi_only, gi)
        finally:
            L.conv3p_set_engine(prev)
        assert assert_close_scaled(res[flags][0], o64, oabs, RTOL, ATOL, f"forward[{flags}]") < RTOL
        assert assert_close_scaled(res[flags][1], r[2], r[3], RTOL, ATOL, f"grad_input[{flags}]") < RTOL
        assert assert_close_scaled(res[flags][2], r[4], r[5], RTOL, ATOL, f"grad_filter[{flags}]") < RTOL
    for a, b, what, padded in zip(res[0], res[2048], ("forward", "grad_input", "grad_filter"),
                                  (fwd_padded, bwd_padded, bwd_padded)):
        assert a.shape == b.shape
        assert np.array_equal(a, b) != padded, f"{what}: padded tensor-core path expected = {padded}"


def test_host_pipeline_matches_direct_calls():
    """HostConv3p (pinned host in/out, overlapped copies) returns exactly what the direct calls return."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    from pointwise_b200.host_api import HostConv3p
    B, N, Cin, Cout, stride = 3, 900, 9, 13, (2, 2, 2)
    pipe = HostConv3p(B, N, Cin, Cout, stride, V)
    probs = [make_problem(B, N, Cin, Cout, "room", seed=30 + i) for i in range(4)]
    host = [{k: torch.from_numpy(v).pin_memory() for k, v in pr.items()} for pr in probs]
    tickets, results = [], []
    for h in host:
        tickets.append(pipe.submit(h["points"], h["input"], h["filter"], h["grad_out"]))
        if len(tickets) >= 2:
            results.append([t.clone() for t in pipe.fetch(tickets[-2])])
    results.append([t.clone() for t in pipe.fetch(tickets[-1])])
    for pr, (y, gi, gf) in zip(probs, results):
        plan = NeighborPlan(dev(pr["points"]), stride, V)
        y0 = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"]))
        gi0, gf0 = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
        assert torch.equal(y, y0.cpu()) and torch.equal(gi, gi0.cpu()) and torch.equal(gf, gf0.cpu())


@pytest.mark.parametrize("B,N", [(0, 16), (3, 0), (0, 0)])
def test_empty_inputs(B, N):
    """Empty batches / empty clouds: empty outputs, zero weight gradient, no crash."""
    from pointwise_b200 import conv3p
    P = torch.zeros(B, N, 3, device="cuda")
    X = torch.zeros(B, N, 5, device="cuda", requires_grad=True)
    W = torch.randn(3, 3, 3, 5, 7, device="cuda", requires_grad=True)
    y = conv3p(P, X, W, [1, 1, 1], [0.1])
    assert y.shape == (B, N, 7)
    y.sum().backward()
    assert X.grad.shape == (B, N, 5)
    assert W.grad.shape == W.shape and float(W.grad.abs().max()) == 0.0


def test_ragged_cloud_sizes_through_tiles(port):
    """Cloud sizes that straddle tile boundaries of both engines (N = 1, 127, 129, 513 in one sweep)."""
    from pointwise_b200 import NeighborPlan, conv3p_forward
    for N in (1, 127, 129, 513):
        pr = make_problem(3, N, 32, 32, "cube", seed=N)
        pr["points"] *= 0.3                       # dense enough to have neighbours
        plan = NeighborPlan(dev(pr["points"]), 1, V)
        out = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"])).cpu().numpy()
        o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], 1, V, with64=True)
        assert_close_scaled(out, o64, oabs, RTOL, ATOL, f"forward N={N}")


@pytest.mark.parametrize("Cin,Cout,stride", [(9, 9, 1), (3, 9, 1), (36, 13, 1), (13, 36, 2), (16, 16, 1), (1, 1, 3)])
def test_small_channel_engine_and_tile_engine_match_oracle(port, Cin, Cout, stride):
    """The reference models' layer shapes: the warp-per-point engine ("simt"/auto) and the generic tile
    engine ("tile") both meet the tolerance, forward and backward."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward, set_engine
    B, N = 3, 700
    pr = make_problem(B, N, Cin, Cout, "room", seed=17, quantise=0.05 if Cin == 9 else None)
    plan = NeighborPlan(dev(pr["points"]), stride, V)
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V, with64=True)
    res = {}
    for eng in ("simt", "tile"):
        prev = set_engine(eng)
        try:
            y = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"])).cpu().numpy()
            gi, gf = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
        finally:
            set_engine(prev)
        assert_close_scaled(y, o64, oabs, RTOL, ATOL, f"forward[{eng}]")
        assert_close_scaled(gi.cpu().numpy(), r[2], r[3], RTOL, ATOL, f"grad_input[{eng}]")
        assert_close_scaled(gf.cpu().numpy(), r[4], r[5], RTOL, ATOL, f"grad_filter[{eng}]")
        res[eng] = y
    assert res["simt"].shape == res["tile"].shape


@pytest.mark.parametrize("B,N,stride,dist,q,Cin,Cout", [
    (3, 1500, 1, "room", 0.0, 64, 128), (2, 2048, 2, "room", 0.05, 64, 128), (4, 700, 1, "sphere", 0.0, 64, 128),
    (1, 129, 3, "cube", 0.0, 64, 128), (2, 900, 1, "room", 0.0, 64, 64), (2, 600, 1, "room", 0.05, 32, 32),
    (1, 700, 1, "room", 0.0, 256, 256), (2, 800, 2, "room", 0.0, 128, 64)])
def test_shared_gather_of_the_two_gradients_is_bit_identical(B, N, stride, dist, q, Cin, Cout):
    """With both gradients requested the grad_input kernel leaves its per-(point, cell) aggregates of grad_output in
    the G store and the grad_filter kernel reads them back instead of walking the backward lists again: same
    members, same order of additions -> bit-identical to the un-shared kernels (engine bit 256) and to a call that
    asks for grad_filter alone."""
    from pointwise_b200 import NeighborPlan, _lib, conv3p_backward
    pr = make_problem(B, N, Cin, Cout, dist, seed=77, quantise=q or None)
    plan = NeighborPlan(dev(pr["points"]), (stride,) * 3, V).ensure_backward()
    g, x, w = dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"])
    L = _lib.lib()
    assert L.conv3p_backward_scratch_bytes(plan.geom, Cin, Cout) > L.conv3p_scratch_bytes(plan.geom, Cin, Cout)
    gi_s, gf_s = conv3p_backward(plan, g, x, w)
    prev = L.conv3p_set_engine(256)
    try:
        gi_u, gf_u = conv3p_backward(plan, g, x, w)
    finally:
        L.conv3p_set_engine(prev)
    _, gf_only = conv3p_backward(plan, g, x, w, need_input_grad=False)
    torch.cuda.synchronize()
    assert torch.equal(gi_s, gi_u)
    assert torch.equal(gf_s, gf_u)
    assert torch.equal(gf_s, gf_only)
    assert float(gf_s.abs().max()) > 0


def test_overflow_poisons_grad_filter_too():
    """An unchecked plan whose pair capacity is too small: every affected output is NaN, including the weight
    gradient (it would otherwise silently miss the overflowed points' terms), on every engine."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, set_engine
    for Cin, Cout, eng in [(4, 4, "simt"), (9, 9, "simt"), (9, 9, "tile"), (64, 128, "tc"), (64, 64, "tc")]:
        pr = make_problem(1, 512, Cin, Cout, "sphere", seed=1)
        plan = NeighborPlan(dev(pr["points"]), 1, V, capacity=600, check=False)
        prev = set_engine(eng)
        try:
            gi, gf = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
        finally:
            set_engine(prev)
        assert torch.isnan(gf).all(), (Cin, Cout, eng)
        assert torch.isnan(gi).any() and not torch.isnan(gi).all()


def test_deferred_overflow_check_has_no_sync_in_steady_state_and_reports_late():
    """Second plan of a shape: built from the learned estimate, counters copied asynchronously; verify() / stats read
    them later.  An explicit, too small capacity with check="deferred" raises at verify(), not at construction."""
    from pointwise_b200 import Conv3pError, NeighborPlan
    pts = dev(make_points(2, 640, "sphere", seed=21))
    first = NeighborPlan(pts, 1, V)                 # learning build: synchronous
    assert first._stats_event is None and first.stats.total_pairs > 0
    second = NeighborPlan(pts, 1, V)                # steady state: deferred
    assert second._stats_event is not None and second._stats is None
    assert second.verify(block=True) and second.stats.total_pairs == first.stats.total_pairs
    assert torch.equal(second.count_table, first.count_table)
    late = NeighborPlan(pts, 1, V, capacity=700, check="deferred")
    with pytest.raises(Conv3pError, match="capacity"):
        late.verify(block=True)


def test_backward_without_lists_is_an_error_status():
    """C ABI: conv3p_backward_f32 on a plan whose backward lists were never built -> CONV3P_ERR_NO_BACKWARD_LISTS."""
    from pointwise_b200 import _lib
    L = _lib.lib()
    B, N, Cin, Cout = 1, 300, 4, 4
    pr = make_problem(B, N, Cin, Cout, "cube", seed=3)
    g = _lib.make_geom(B, N, (1, 1, 1), V, 64 * N)
    plan = torch.empty(L.conv3p_plan_bytes(g), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(L.conv3p_scratch_bytes(g, Cin, Cout), dtype=torch.uint8, device="cuda")
    P, X, W, G = (dev(pr[k]) for k in ("points", "input", "filter", "grad_out"))
    gi, gf = torch.empty_like(X), torch.empty_like(W)
    _lib.check(L.conv3p_plan_build_f32(g, P.data_ptr(), plan.data_ptr(), plan.numel(), None))
    args = (g, plan.data_ptr(), G.data_ptr(), X.data_ptr(), W.data_ptr(), Cin, Cout, gi.data_ptr(), gf.data_ptr(),
            scratch.data_ptr(), scratch.numel(), None)
    assert L.conv3p_backward_f32(*args) == _lib.ERR_NO_BACKWARD_LISTS
    _lib.check(L.conv3p_plan_build_backward(g, P.data_ptr(), plan.data_ptr(), plan.numel(), None))
    assert L.conv3p_backward_f32(*args) == _lib.OK
    torch.cuda.synchronize()
    assert torch.isfinite(gi).all() and torch.isfinite(gf).all()


@pytest.mark.parametrize("Cin,Cout", [(64, 128), (64, 64), (96, 32), (256, 256)])
def test_bf16_correction_split_against_three_tf32_products(port, Cin, Cout):
    """Tensor-core precision modes: the production split (TF32 main product + one BF16 chain for the two correction
    terms) and plain 3xTF32 (engine flag 512) both meet the operator's tolerance against the float64 oracle; they are
    different arithmetic (not bit-equal), and the production split stays within a small factor of 3xTF32's error."""
    from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward
    B, N = (1, 500) if Cin * Cout >= 65536 else (2, 1100)
    pr = make_problem(B, N, Cin, Cout, "room", seed=41)
    plan = NeighborPlan(dev(pr["points"]), 1, V).ensure_backward()
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], 1, V, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], 1, V, with64=True)
    L = _lib.lib()
    errs, outs = {}, {}
    for flag in (0, 512):
        prev = L.conv3p_set_engine(flag)
        try:
            y = conv3p_forward(plan, dev(pr["input"]), dev(pr["filter"])).cpu().numpy()
            gi, gf = conv3p_backward(plan, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
        finally:
            L.conv3p_set_engine(prev)
        errs[flag] = (assert_close_scaled(y, o64, oabs, RTOL, ATOL, f"forward[{flag}]"),
                      assert_close_scaled(gi.cpu().numpy(), r[2], r[3], RTOL, ATOL, f"grad_input[{flag}]"),
                      assert_close_scaled(gf.cpu().numpy(), r[4], r[5], RTOL, ATOL, f"grad_filter[{flag}]"))
        outs[flag] = y
    print("max err / sum|terms| (fwd, grad_input, grad_filter): production", errs[0], "3xTF32", errs[512])
    assert not np.array_equal(outs[0], outs[512])
    assert max(errs[0]) < 3e-6, errs          # 3x head-room under the 1e-5 tolerance


def test_whole_step_is_cuda_graph_capturable():
    """Plan build + forward + backward enqueue nothing but stream-ordered work on the caller's stream: the whole step can be
    captured into a CUDA graph (any host synchronisation or device allocation by the library would abort the capture,
    SURVEY 8b) and the replay gives bit-identical results."""
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    pr = make_problem(2, 1500, 64, 128, "room", seed=61)
    P, X, W, G = (dev(pr[k]) for k in ("points", "input", "filter", "grad_out"))
    cap = int(NeighborPlan(P, 1, V, check="sync").stats.total_pairs * 1.05) + 1024

    def step():
        plan = NeighborPlan(P, 1, V, check=False, capacity=cap)
        y = conv3p_forward(plan, X, W)
        gi, gf = conv3p_backward(plan, G, X, W)
        return y, gi, gf

    want = step()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        got = step()
    for t in got:
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(want, got))


def test_channel_padding_with_strided_rows_selu_and_overflow_poison():
    """The padded tensor-core path behind conv3p_forward_ex_f32: input and output as channel slices of wider buffers,
    SELU in the epilogue, and NaN poisoning when the plan's lists overflowed -- same results as the fp32 engines
    (engine flag 2048) up to the operator's tolerance, untouched channels outside the slices."""
    from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward
    B, N, Cin, Cout = 2, 700, 36, 13
    pr = make_problem(B, N, Cin, Cout, "room", seed=33)
    plan = NeighborPlan(dev(pr["points"]), (1, 1, 1), V)
    wide_in = torch.full((B, N, 50), 7.0, device="cuda")
    wide_in[..., 5:5 + Cin] = dev(pr["input"])
    L = _lib.lib()
    outs = {}
    for flags in (0, 2048):
        prev = L.conv3p_set_engine(flags)
        try:
            wide_out = torch.full((B, N, 20), -3.0, device="cuda")
            conv3p_forward(plan, wide_in[..., 5:5 + Cin], dev(pr["filter"]), activation="selu", out=wide_out[..., 2:2 + Cout])
            outs[flags] = wide_out.cpu().numpy()
        finally:
            L.conv3p_set_engine(prev)
        assert (outs[flags][..., :2] == -3.0).all() and (outs[flags][..., 2 + Cout:] == -3.0).all()
    assert not np.array_equal(outs[0], outs[2048])
    np.testing.assert_allclose(outs[0], outs[2048], rtol=2e-4, atol=2e-5)
    # overflow: a plan built with too small a capacity poisons the padded path's outputs too
    small = NeighborPlan(dev(pr["points"]), (1, 1, 1), V, capacity=2000, check=False)
    y = conv3p_forward(small, dev(pr["input"]), dev(pr["filter"]))
    gi, gf = conv3p_backward(small, dev(pr["grad_out"]), dev(pr["input"]), dev(pr["filter"]))
    assert torch.isnan(y).any() and torch.isnan(gi).any() and torch.isnan(gf).all()
    assert not torch.isnan(y).all()          # rows whose lists fit are computed
