"""GPU tests of the drop-in consumers: the reference's two networks restated on the new
operator, checked layer-by-layer against the CPU oracle, with plan sharing and a few optimisation steps."""
import numpy as np
import pytest
import torch

from helpers import assert_close_scaled
from pointwise_b200.synth import make_points

pytestmark = pytest.mark.gpu


def selu_np(x):
    a, s = 1.6732632423543772848170429916717, 1.0507009873554804934193349852946   # selu.py:24-25
    return (s * np.where(x >= 0, x, a * np.expm1(np.minimum(x, 0)))).astype(np.float32)


def test_segmentation_net_matches_oracle_chain(port):
    """pointcnn_scene_seg_acsd.py:51-57 evaluated layer by layer with the CPU oracle and numpy SELU."""
    from pointwise_b200.nets import PointConvNetSeg
    torch.manual_seed(0)
    B, N, C, K = 2, 1024, 9, 13
    pts = make_points(B, N, "room", seed=3)
    feats = np.random.default_rng(5).uniform(-1, 1, (B, N, C)).astype(np.float32)
    net = PointConvNetSeg(K, C).cuda()
    got = net.model(torch.from_numpy(pts).cuda(), torch.from_numpy(feats).cuda()).detach().cpu().numpy()
    W = [w.detach().cpu().numpy() for w in net.filters]
    x, outs = feats, []
    for i in range(4):
        x = selu_np(port.forward(pts, x, W[i], i + 1, 0.1))
        outs.append(x)
    want = selu_np(port.forward(pts, np.concatenate(outs, axis=2), W[4], 1, 0.1))
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5)


def test_classification_net_matches_oracle_chain(port):
    """pointcnn2_acsd.py:48-75 evaluated with the CPU oracle (Conv3p layers), numpy SELU and numpy dense layers."""
    from pointwise_b200.nets import PointConvNetCls
    torch.manual_seed(0)
    B, N, C, K = 4, 1024, 3, 40
    pts = make_points(B, N, "sphere", seed=7)
    net = PointConvNetCls(K, N, C).cuda()
    got = net.model(torch.from_numpy(pts).cuda(), torch.from_numpy(pts.copy()).cuda(), is_training=False)
    got = got.detach().cpu().numpy()
    W = [w.detach().cpu().numpy() for w in net.filters]
    x, outs = pts, []
    for i in range(4):                                                  # :48-67, strides 1..4
        x = selu_np(port.forward(pts, x, W[i], i + 1, 0.1))
        outs.append(x)
    feat = np.concatenate(outs, axis=2).reshape(B, -1).astype(np.float64)      # :69-70
    w1, b1 = net.fc1.weight.detach().cpu().numpy().astype(np.float64), net.fc1.bias.detach().cpu().numpy()
    w2, b2 = net.fc2.weight.detach().cpu().numpy().astype(np.float64), net.fc2.bias.detach().cpu().numpy()
    h = selu_np(feat @ w1.T + b1).astype(np.float64)                   # :71 (no dropout at inference, :73)
    want = selu_np(h @ w2.T + b2)                                       # :75
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4)       # torch's dense layers run in TF32-free fp32


def test_selu_gradient_at_exactly_zero_follows_the_reference():
    """selu.py:25 takes the linear branch for x >= 0, so at a pre-activation of exactly 0 (all neighbour rows zero:
    padded rows, zero features) the derivative is `scale`, not scale * alpha."""
    from pointwise_b200 import conv3p
    from pointwise_b200.synth import make_problem
    scale = 1.0507009873554804934193349852946
    pr = make_problem(1, 300, 9, 9, "room", seed=19)
    P = torch.from_numpy(pr["points"]).cuda()
    x = torch.zeros(1, 300, 9, device="cuda", requires_grad=True)           # y = selu(0) = 0 everywhere
    w = torch.from_numpy(pr["filter"]).cuda().requires_grad_()
    g = torch.from_numpy(pr["grad_out"]).cuda()
    y = conv3p(P, x, w, [1, 1, 1], [0.1], activation="selu")
    assert float(y.abs().max()) == 0.0
    y.backward(g)
    x2 = torch.zeros(1, 300, 9, device="cuda", requires_grad=True)
    y2 = conv3p(P, x2, w.detach(), [1, 1, 1], [0.1])
    y2.backward(g * scale)
    assert torch.equal(x.grad, x2.grad)


def test_classification_net_trains():
    """pointcnn2_acsd.py:37-89 at the ModelNet40 shape: loss decreases, gradients reach every filter, one plan
    per stride."""
    from pointwise_b200.nets import PointConvNetCls
    from pointwise_b200 import launch_count
    torch.manual_seed(1)
    B, N, C, K = 8, 1024, 3, 40
    pts = torch.from_numpy(make_points(B, N, "sphere", seed=2)).cuda()
    net = PointConvNetCls(K, N, C).cuda()
    labels = torch.randint(0, K, (B,), device="cuda")
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)     # param.json: momentum 0.9, lr 1e-3
    losses = []
    for _ in range(6):
        opt.zero_grad()
        launch_count(reset=True)
        loss = net.loss(net.model(pts, pts.clone(), is_training=False), labels)
        loss.backward()
        n_launch = launch_count()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert losses[-1] < losses[0]
    assert n_launch > 0


def test_plan_cache_shares_plans():
    from pointwise_b200.nets import PlanCache
    pts = torch.from_numpy(make_points(2, 256, "cube", seed=1)).cuda()
    pc = PlanCache(pts, 0.1)
    a, b, c = pc.get([1, 1, 1]), pc.get(torch.tensor([1, 1, 1])), pc.get([2, 2, 2])
    assert a is b and a is not c


# ---- SELU fused into the epilogue, concat-free layout ------------------------------------------
@pytest.mark.parametrize("Cin,Cout,stride", [(9, 9, 1), (36, 13, 1), (64, 128, 1), (32, 64, 2), (40, 48, 1)])
def test_fused_selu_epilogue_matches_selu_of_plain_output(Cin, Cout, stride):
    """activation="selu" == selu(unfused output) on every forward engine (warp-per-point, tile, tensor core);
    the unfused output itself is held to the oracle in test_gpu_parity.py."""
    from pointwise_b200 import NeighborPlan, conv3p_forward
    from pointwise_b200.synth import make_problem
    pr = make_problem(3, 700, Cin, Cout, "room", seed=11)
    plan = NeighborPlan(torch.from_numpy(pr["points"]).cuda(), (stride,) * 3, 0.1)
    x, w = torch.from_numpy(pr["input"]).cuda(), torch.from_numpy(pr["filter"]).cuda()
    plain = conv3p_forward(plan, x, w)
    fused = conv3p_forward(plan, x, w, activation="selu")
    want = selu_np(plain.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(fused.cpu().numpy(), want, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("Cin,Cout,W_in,W_out,c_in,c_out", [(9, 9, 36, 36, 9, 18), (64, 128, 96, 160, 32, 16),
                                                            (9, 13, 11, 17, 1, 3)])
def test_strided_rows_in_and_out(Cin, Cout, W_in, W_out, c_in, c_out):
    """Input read from / output written into channel slices of wider buffers: same values as the dense call, and
    nothing outside the output slice is touched.  (Third case: rows not 16-byte aligned -> fp32 engines.)"""
    from pointwise_b200 import NeighborPlan, conv3p_forward
    from pointwise_b200.synth import make_problem
    B, N = 2, 600
    pr = make_problem(B, N, Cin, Cout, "room", seed=12)
    plan = NeighborPlan(torch.from_numpy(pr["points"]).cuda(), (1, 1, 1), 0.1)
    x, w = torch.from_numpy(pr["input"]).cuda(), torch.from_numpy(pr["filter"]).cuda()
    dense = conv3p_forward(plan, x, w, activation="selu")
    wide_in = torch.full((B, N, W_in), float("nan"), device="cuda")
    wide_in[:, :, c_in:c_in + Cin] = x
    wide_out = torch.full((B, N, W_out), -7.0, device="cuda")
    ret = conv3p_forward(plan, wide_in[:, :, c_in:c_in + Cin], w, activation="selu",
                         out=wide_out[:, :, c_out:c_out + Cout])
    assert ret.data_ptr() == wide_out[:, :, c_out:c_out + Cout].data_ptr()
    assert torch.equal(wide_out[:, :, c_out:c_out + Cout], dense)
    assert float((wide_out[:, :, :c_out] + 7.0).abs().max()) == 0.0
    assert float((wide_out[:, :, c_out + Cout:] + 7.0).abs().max() if c_out + Cout < W_out else 0.0) == 0.0


def test_fused_selu_gradients_match_autograd_of_unfused():
    """conv3p(..., activation="selu") and F.selu(conv3p(...)) give the same output and the same gradients."""
    import torch.nn.functional as F
    from pointwise_b200 import conv3p
    from pointwise_b200.synth import make_problem
    for Cin, Cout in [(9, 9), (64, 128)]:
        pr = make_problem(2, 800, Cin, Cout, "room", seed=13)
        P = torch.from_numpy(pr["points"]).cuda()
        g = torch.from_numpy(pr["grad_out"]).cuda()
        res = []
        for fused in (True, False):
            x = torch.from_numpy(pr["input"]).cuda().requires_grad_()
            w = torch.from_numpy(pr["filter"]).cuda().requires_grad_()
            y = conv3p(P, x, w, [1, 1, 1], [0.1], activation="selu") if fused else F.selu(conv3p(P, x, w, [1, 1, 1], [0.1]))
            y.backward(g)
            res.append((y.detach(), x.grad, w.grad))
        for a, b in zip(*res):
            scale = float(b.abs().max())
            assert float((a - b).abs().max()) <= 3e-6 * scale + 1e-7


def test_concat_free_inference_matches_model():
    """PointConvNetSeg.infer / PointConvNetCls.infer (layers write into the concat buffer, SELU in the epilogue)
    reproduce model() (separate SELU, torch.cat)."""
    from pointwise_b200.nets import PointConvNetCls, PointConvNetSeg
    torch.manual_seed(3)
    B, N = 2, 1024
    pts = torch.from_numpy(make_points(B, N, "room", seed=4)).cuda()
    feats = torch.from_numpy(np.random.default_rng(6).uniform(-1, 1, (B, N, 9)).astype(np.float32)).cuda()
    seg = PointConvNetSeg(13, 9).cuda()
    a, b = seg.model(pts, feats).detach(), seg.infer(pts, feats)
    assert float((a - b).abs().max()) <= 1e-6 * float(a.abs().max()) + 1e-7
    cls = PointConvNetCls(40, N, 3).cuda()
    a, b = cls.model(pts, pts.clone(), is_training=False).detach(), cls.infer(pts, pts.clone())
    assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-6
