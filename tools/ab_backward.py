"""A/B timing of the backward kernels: engine 0 (the grad_filter kernel reads the G store the grad_input kernel
leaves behind) against 256 (no sharing: both kernels gather).  usage: python tools/ab_backward.py [workload]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, _lib, conv3p_backward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

B, N, Cin, Cout, stride, dist = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "headline"]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
plan = NeighborPlan(pr["points"], stride, VOXEL).ensure_backward()
L = _lib.lib()
for eng, name in [(0, "shared gather (G store)"), (256, "unshared")]:
    L.conv3p_set_engine(eng)
    for _ in range(3):
        conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
    torch.cuda.synchronize()
    L.conv3p_profile_enable(1)
    for _ in range(10):
        conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
    torch.cuda.synchronize()
    buf = C.create_string_buffer(8192)
    L.conv3p_profile_read(buf, 8192)
    L.conv3p_profile_enable(0)
    print(name)
    for ln in buf.value.decode().splitlines():
        k, n, t = ln.split()
        print(f"   {k:28s} {float(t) / int(n):.4f} ms")
L.conv3p_set_engine(0)
