"""Host-side vs device-side time of one bench step (tools for chasing launch-path stalls).
usage: python tools/step_breakdown.py [workload]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
B, N, Cin, Cout, stride, dist = WORKLOADS[name]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
probe = NeighborPlan(pr["points"], stride, VOXEL, check="sync")
cap = int(probe.stats.total_pairs * 1.05) + 1024


def step(sync=False):
    t = [time.perf_counter()]

    def mark():
        if sync:
            torch.cuda.synchronize()
        t.append(time.perf_counter())
    plan = NeighborPlan(pr["points"], stride, VOXEL, check=False, capacity=cap); mark()
    y = conv3p_forward(plan, pr["input"], pr["filter"]); mark()
    plan.prefetch_backward(); mark()
    gi, gf = conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"]); mark()
    return [b - a for a, b in zip(t, t[1:])]


for _ in range(5):
    step()
torch.cuda.synchronize()
for mode in (False, True):
    acc = [0.0] * 4
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        for i, v in enumerate(step(mode)):
            acc[i] += v
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{name} {'synchronised after every call' if mode else 'enqueue only (host time per call)'}: "
          f"plan {acc[0] / 20 * 1e3:.3f} ms | forward {acc[1] / 20 * 1e3:.3f} | prefetch {acc[2] / 20 * 1e3:.3f} | "
          f"backward {acc[3] / 20 * 1e3:.3f} | whole step {total:.3f} ms; allocated {torch.cuda.memory_allocated() >> 20} MiB, "
          f"reserved {torch.cuda.memory_reserved() >> 20} MiB")
