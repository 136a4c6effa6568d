"""Does a CUDA graph of the whole fwd+bwd step (plan build, forward, backward) beat eager launches?
usage: python tools/graph_step.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CONV3P_PREFETCH_BACKWARD", "0")      # single stream inside the capture
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
B, N, Cin, Cout, stride, dist = WORKLOADS[name]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
cap = int(NeighborPlan(pr["points"], stride, VOXEL, check="sync").stats.total_pairs * 1.05) + 1024


def step():
    plan = NeighborPlan(pr["points"], stride, VOXEL, check=False, capacity=cap)
    y = conv3p_forward(plan, pr["input"], pr["filter"])
    gi, gf = conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
    return y, gi, gf


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


eager = timeit(step)
ref = step()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
graph = timeit(g.replay)
same = all(torch.equal(a, b) for a, b in zip(ref, out))
print(f"{name}: eager {eager:.4f} ms/step, CUDA graph replay {graph:.4f} ms/step, identical results: {same}")
