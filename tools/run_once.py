"""Runs the headline Conv3p fwd+bwd a few times (target for ncu captures).
usage: python tools/run_once.py [iters] [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, N, Cin, Cout, stride, dist = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "headline"]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
for _ in range(iters):
    plan = NeighborPlan(pr["points"], stride, VOXEL)
    y = conv3p_forward(plan, pr["input"], pr["filter"])
    gi, gf = conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(gi.abs().mean()), float(gf.abs().mean()))
