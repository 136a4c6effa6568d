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
        losses.append(float(loss))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert losses[-1] < losses[0]
    assert n_launch > 0


def test_plan_cache_shares_plans():
    from pointwise_b200.nets import PlanCache
    pts = torch.from_numpy(make_points(2, 256, "cube", seed=1)).cuda()
    pc = PlanCache(pts, 0.1)
    a, b, c = pc.get([1, 1, 1]), pc.get(torch.tensor([1, 1, 1])), pc.get([2, 2, 2])
    assert a is b and a is not c
