"""Where the host-buffer pipeline's step time goes: per-stage durations (H2D, compute, D2H) measured with CUDA events
around every stage of every step of HostConv3p's three-stream pipeline, and the steady-state period.
usage: python tools/e2e_timeline.py [depth]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, N, Cin, Cout, stride, dist = WORKLOADS["headline"]
pr = make_problem(B, N, Cin, Cout, dist, seed=0)
host = {k: torch.from_numpy(v).pin_memory() for k, v in pr.items()}
dev = torch.device("cuda", 0)
cap = int(NeighborPlan(host["points"].to(dev), stride, VOXEL, check="sync").stats.total_pairs * 1.05) + 1024
f32 = torch.float32
slots = [dict(points=torch.empty((B, N, 3), dtype=f32, device=dev), input=torch.empty((B, N, Cin), dtype=f32, device=dev),
              filter=torch.empty((3, 3, 3, Cin, Cout), dtype=f32, device=dev),
              grad_out=torch.empty((B, N, Cout), dtype=f32, device=dev),
              h_out=torch.empty((B, N, Cout), dtype=f32).pin_memory(), h_gi=torch.empty((B, N, Cin), dtype=f32).pin_memory(),
              h_gf=torch.empty((3, 3, 3, Cin, Cout), dtype=f32).pin_memory(), fetched=None) for _ in range(depth)]
s_h2d, s_cmp, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
log = []
steps = 24
t_first = None
for k in range(steps):
    s = slots[k % depth]
    e = [ev() for _ in range(6)]
    if s["fetched"] is not None:
        s["fetched"].synchronize()              # the host waits for the slot's previous results (as fetch() does)
        if os.environ.get("E2E_CHECK"):
            s["keep"][0].verify(block=True)
    with torch.cuda.stream(s_h2d):
        e[0].record(s_h2d)
        for name in ("points", "input", "filter", "grad_out"):
            s[name].copy_(host[name], non_blocking=True)
        e[1].record(s_h2d)
    with torch.cuda.stream(s_cmp):
        s_cmp.wait_event(e[1])
        e[2].record(s_cmp)
        plan = NeighborPlan(s["points"], stride, VOXEL, capacity=cap, check=os.environ.get("E2E_CHECK", "") or False)
        out = conv3p_forward(plan, s["input"], s["filter"])
        plan.prefetch_backward()
        gi, gf = conv3p_backward(plan, s["grad_out"], s["input"], s["filter"])
        e[3].record(s_cmp)
        s["keep"] = (plan, out, gi, gf)
    with torch.cuda.stream(s_d2h):
        s_d2h.wait_event(e[3])
        e[4].record(s_d2h)
        s["h_out"].copy_(out, non_blocking=True)
        s["h_gi"].copy_(gi, non_blocking=True)
        s["h_gf"].copy_(gf, non_blocking=True)
        e[5].record(s_d2h)
        for t in (out, gi, gf, plan.buffer):
            t.record_stream(s_d2h)
    s["fetched"] = e[5]
    log.append(e)
torch.cuda.synchronize()
base = log[8][0]
rows = []
for k in range(8, steps):
    e = log[k]
    rows.append([base.elapsed_time(x) for x in e])
print("step:  H2D start..end | compute start..end | D2H start..end   (ms since step 8's H2D start)")
for k, r in enumerate(rows[:8]):
    print(f"{k + 8:3d}: {r[0]:7.2f}..{r[1]:7.2f} | {r[2]:7.2f}..{r[3]:7.2f} | {r[4]:7.2f}..{r[5]:7.2f}")
import statistics as st
print("mean durations: H2D %.2f ms, compute %.2f ms, D2H %.2f ms; period %.2f ms" % (
    st.mean(r[1] - r[0] for r in rows), st.mean(r[3] - r[2] for r in rows), st.mean(r[5] - r[4] for r in rows),
    (rows[-1][5] - rows[0][5]) / (len(rows) - 1)))
