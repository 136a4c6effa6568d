"""GPU tests of the T = double instantiation of the operator.  The reference registers Conv3p / Conv3pGrad for double as
well (register_op.cpp:45, 64; CPU kernels tf_conv3p_atrous.cpp:516, 727): every tensor in double and the neighbour
predicate evaluated in double.  Checked against golden vectors generated from the reference's double kernels
(tests/golden/make_golden.py, DOUBLE) -- count tables bit-identical, sums within 1e-12 of scale -- and, where the
reference's object code is present, against it directly."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "f64_*.npz")))
RTOL = 1e-12     # of the result's scale: double sums of at most a few hundred terms, order of additions differs


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want, what):
    assert got.dtype == np.float64, f"{what}: result is {got.dtype}"
    scale = float(np.abs(want).max()) + 1e-300
    err = float(np.abs(got - want).max())
    assert err <= RTOL * scale, f"{what}: max |err| {err:.3e} at scale {scale:.3e}"


def count_table(P, dims, stride, voxel, Cin, Cout, cap):
    """The count table [B, N, cells] the one-shot double call leaves in its workspace."""
    from pointwise_b200 import _lib
    L = _lib.lib()
    B, N = P.shape[:2]
    cells = dims[0] * dims[1] * dims[2]
    geom = _lib.make_geom(B, N, stride, voxel, cap)
    i3 = C.c_int * 3
    nbytes = L.conv3p_op_workspace_bytes_f64(geom, i3(*dims), Cin, Cout)
    assert nbytes > 0
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    X = torch.zeros(B, N, Cin, dtype=torch.float64, device="cuda")
    W = torch.zeros(*dims, Cin, Cout, dtype=torch.float64, device="cuda")
    out = torch.empty(B, N, Cout, dtype=torch.float64, device="cuda")
    _lib.check(L.conv3p_op_forward_f64(P.data_ptr(), X.data_ptr(), W.data_ptr(), i3(*dims), i3(*stride), voxel, B, N,
                                       Cin, Cout, cap, out.data_ptr(), ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    pb = L.conv3p_plan_bytes(geom)
    return ws[pb:pb + B * N * cells * 4].view(torch.int32).view(B, N, cells).cpu().numpy()


@pytest.mark.parametrize("name", CASES)
def test_double_matches_reference_vectors(name):
    from pointwise_b200 import conv3p, conv3p_grad
    g = np.load(os.path.join(GOLD, name + ".npz"))
    stride, voxel = [int(s) for s in g["stride"]], float(g["voxel"])
    P, X, W, G = dev(g["points"]), dev(g["input"]).requires_grad_(), dev(g["filter"]).requires_grad_(), dev(g["grad_out"])
    y = conv3p(P, X, W, stride, [voxel])
    assert y.dtype == torch.float64
    y.backward(G)
    close(y.detach().cpu().numpy(), g["output"], f"{name} output")
    close(X.grad.cpu().numpy(), g["grad_input"], f"{name} grad_input")
    close(W.grad.cpu().numpy(), g["grad_filter"], f"{name} grad_filter")
    gi, gf = conv3p_grad(G, P, X.detach(), W.detach(), stride, [voxel])      # the Conv3pGrad mirror, called directly
    assert torch.equal(gi, X.grad) and torch.equal(gf, W.grad)
    dims = tuple(int(d) for d in g["filter"].shape[:3])
    cnt = count_table(P, dims, stride, voxel, int(g["filter"].shape[3]), int(g["filter"].shape[4]),
                      int(g["count_table"].sum()) + 64)
    assert np.array_equal(cnt, g["count_table"]), "count table differs from the reference's double neighbor_count"


def test_double_predicate_is_not_the_float_one():
    """On the sub-float lattice the float operator (same points rounded to float) must see different neighbours: the
    double path really evaluates the predicate on the double bits."""
    from pointwise_b200 import NeighborPlan
    g = np.load(os.path.join(GOLD, "f64_333_subfloat.npz"))
    stride, voxel = [int(s) for s in g["stride"]], float(g["voxel"])
    P32 = dev(g["points"].astype(np.float32))
    plan = NeighborPlan(P32, stride, [voxel])
    c32 = plan.count_table.cpu().numpy().reshape(g["count_table"].shape)
    assert not np.array_equal(c32, g["count_table"])
    c64 = count_table(dev(g["points"]), (3, 3, 3), stride, voxel, 2, 2, int(g["count_table"].sum()) + 64)
    assert np.array_equal(c64, g["count_table"])


def test_double_against_reference_object_code():
    """Larger random cases straight against the reference's double kernels (wherever oracle/_ref was built)."""
    import oracle
    if not oracle.Ref.available() or not hasattr(oracle.ref().lib, "ref_conv3p_forward_f64"):
        pytest.skip("oracle/_ref (with the double entry points) not built")
    from pointwise_b200 import conv3p
    from pointwise_b200.synth import make_problem
    R = oracle.ref()
    rng = np.random.default_rng(77)
    for dims, stride, dist, N in [((3, 3, 3), (1, 1, 1), "room", 2000), ((3, 3, 3), (2, 2, 2), "sphere", 1500),
                                  ((3, 5, 3), (1, 1, 2), "room", 900), ((4, 4, 4), (1, 1, 1), "cube", 700)]:
        P = make_problem(2, N, 6, 7, dist, seed=31)["points"].astype(np.float64) + rng.uniform(-1e-9, 1e-9, (2, N, 3))
        Xn, Wn, Gn = rng.uniform(-1, 1, (2, N, 6)), rng.uniform(-0.1, 0.1, (*dims, 6, 7)), rng.uniform(-1, 1, (2, N, 7))
        want = R.forward64(P, Xn, Wn, stride, 0.1)
        wgi, wgf = R.backward64(Gn, P, Xn, Wn, stride, 0.1)
        X, W = dev(Xn).requires_grad_(), dev(Wn).requires_grad_()
        y = conv3p(dev(P), X, W, list(stride), [0.1])
        y.backward(dev(Gn))
        close(y.detach().cpu().numpy(), want, f"{dims} {stride} output")
        close(X.grad.cpu().numpy(), wgi, f"{dims} {stride} grad_input")
        close(W.grad.cpu().numpy(), wgf, f"{dims} {stride} grad_filter")
        for b in range(2):
            cnt = count_table(dev(P[b:b + 1]), dims, list(stride), 0.1, 1, 1, 64 * N * 4)
            assert np.array_equal(cnt[0], R.neighbor_count64(P[b], stride, 0.1, dims=dims))


def test_double_call_rejects_mixed_dtypes_and_reports_overflow():
    from pointwise_b200 import _lib, conv3p
    P = torch.rand(1, 64, 3, dtype=torch.float64, device="cuda")
    with pytest.raises(TypeError, match="float64"):
        conv3p(P, torch.zeros(1, 64, 2, device="cuda"), torch.zeros(3, 3, 3, 2, 2, dtype=torch.float64, device="cuda"),
               [1, 1, 1], [0.1])
    # pair capacity too small: status OK, outputs NaN-poisoned, overflow flag in the plan header (as the float calls)
    L = _lib.lib()
    i3 = C.c_int * 3
    P = torch.zeros(1, 64, 3, dtype=torch.float64, device="cuda")          # 64 coincident points: 4096 pairs
    geom = _lib.make_geom(1, 64, [1, 1, 1], 0.1, 100)
    nbytes = L.conv3p_op_workspace_bytes_f64(geom, i3(3, 3, 3), 2, 2)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    X = torch.ones(1, 64, 2, dtype=torch.float64, device="cuda")
    W = torch.ones(3, 3, 3, 2, 2, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, 64, 2, dtype=torch.float64, device="cuda")
    _lib.check(L.conv3p_op_forward_f64(P.data_ptr(), X.data_ptr(), W.data_ptr(), i3(3, 3, 3), i3(1, 1, 1), 0.1, 1, 64,
                                       2, 2, 100, out.data_ptr(), ws.data_ptr(), nbytes, None))
    st = _lib.PlanStats()
    _lib.check(L.conv3p_plan_stats(geom, ws.data_ptr(), st, None))
    assert st.overflow and st.total_pairs == 64 * 64
    assert torch.isnan(out).any()
    y = conv3p(P, X, W, [1, 1, 1], [0.1])                                   # the Python layer grows the capacity
    assert torch.allclose(y, torch.full_like(y, 2.0))
