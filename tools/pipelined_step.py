"""Software pipelining across training steps: the neighbour plan of batch i+1 depends on its POINTS only, so an input
pipeline that has the next batch on the device can build it on a side stream while batch i's gradient kernels run
(the plan kernels are integer / issue-bound and fit next to the persistent tensor-core CTAs).  Measures the fwd+bwd
step with and without that overlap; every step still builds exactly one plan inside the timed region.
(The figures in profiles/r2_summary.md were taken before NeighborPlan released its buffer in the allocating stream's
order: a plan built on a foreign stream, as here, is now serialised behind its predecessor's side-stream work, and this
tool shows a loss at every size.  Kept as the record of the experiment.)
usage: python tools/pipelined_step.py [workload]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
B, N, Cin, Cout, stride, dist = WORKLOADS[name]
# two different batches, alternating: the prefetched plan really belongs to the NEXT batch
prs = [{k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=s).items()} for s in (0, 1)]
cap = max(int(NeighborPlan(p["points"], stride, VOXEL, check="sync").stats.total_pairs * 1.05) + 1024 for p in prs)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def plain(i):
    pr = prs[i & 1]
    plan = NeighborPlan(pr["points"], stride, VOXEL, check=False, capacity=cap)
    y = conv3p_forward(plan, pr["input"], pr["filter"])
    plan.prefetch_backward()
    return (y,) + tuple(conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"]))


state = {"plan": None}


def build_ahead(i):
    side.wait_stream(main)          # (the points of batch i are resident; nothing else to wait for in this tool)
    with torch.cuda.stream(side):
        plan = NeighborPlan(prs[i & 1]["points"], stride, VOXEL, check=False, capacity=cap)
    plan.buffer.record_stream(main)
    return plan


def pipelined(i):
    pr = prs[i & 1]
    plan = state["plan"]
    main.wait_event(plan._searched)
    y = conv3p_forward(plan, pr["input"], pr["filter"])
    plan.prefetch_backward()
    state["plan"] = build_ahead(i + 1)            # overlaps the gradient kernels below
    return (y,) + tuple(conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"]))


def timeit(fn, n=20):
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        out = fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


t_plain, ref = timeit(plain)
state["plan"] = build_ahead(0)
t_pipe, out = timeit(pipelined)
same = all(torch.equal(a, b) for a, b in zip(ref, out))
print(json.dumps({"workload": name, "ms_per_step": t_plain, "ms_per_step_plan_built_a_step_ahead": t_pipe,
                  "points_per_s": B * N / (t_plain * 1e-3), "points_per_s_pipelined": B * N / (t_pipe * 1e-3),
                  "identical_results": bool(same)}))
