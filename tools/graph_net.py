"""A whole network training step (forward + loss + backward of pointwise_b200.nets) as ONE CUDA graph against eager
launches: the networks at the reference's sizes launch 35-50 short kernels per step and are host-bound in eager mode.
usage: python tools/graph_net.py [seg_net|cls_net]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import NETS  # noqa: E402
from pointwise_b200 import nets, ops  # noqa: E402
from pointwise_b200.synth import make_points  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "seg_net"
per_gpu, N, cin, ncls, dist = NETS[name]
dev = torch.device("cuda")
pts = torch.from_numpy(make_points(per_gpu, N, dist, seed=0)).to(dev)
torch.manual_seed(0)
feats = pts.clone() if cin == 3 else torch.rand(per_gpu, N, cin, device=dev) * 2 - 1
if name == "seg_net":
    net = nets.PointConvNetSeg(ncls, cin).to(dev)
    labels = torch.randint(0, ncls, (per_gpu, N), device=dev)
else:
    net = nets.PointConvNetCls(ncls, N, cin).to(dev)
    labels = torch.randint(0, ncls, (per_gpu,), device=dev)
params = list(net.parameters())
for p in params:                      # static gradient buffers: the captured backward accumulates into them
    p.grad = torch.zeros_like(p)


def step():
    for p in params:
        p.grad.zero_()
    loss = net.loss(net.model(pts, feats, True), labels)
    loss.backward()
    return loss


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timeit(step)                  # also learns the plans' pair capacities (checked builds)
ref_loss = float(step())
ref_grads = [p.grad.clone() for p in params]
nets.PLAN_CHECK = False               # learned capacity, nothing read back inside the capture
ops.PREFETCH_BACKWARD = False         # single stream inside the capture
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = step()
graph = timeit(g.replay)
g.replay()
torch.cuda.synchronize()
same = abs(float(loss) - ref_loss) <= 1e-6 * abs(ref_loss) and all(
    torch.allclose(p.grad, r, rtol=1e-5, atol=1e-7) for p, r in zip(params, ref_grads))
print(json.dumps({"net": name, "points": per_gpu * N, "eager_ms_per_step": eager, "cuda_graph_ms_per_step": graph,
                  "eager_points_per_s": per_gpu * N / (eager * 1e-3), "graph_points_per_s": per_gpu * N / (graph * 1e-3),
                  "same_loss_and_gradients": bool(same)}))
