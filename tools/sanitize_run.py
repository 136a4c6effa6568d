"""Small fixed workload for compute-sanitizer (tools/sanitize.sh): every engine's forward / grad_input / grad_filter
plus the plan kernels on clouds small enough for the tool's ~100x slow-down.  `small` restricts it to the SIMT
kernels (racecheck does not see tcgen05/TMEM traffic and is very slow on the persistent kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward, set_engine  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"
shapes = [(2, 300, 9, 9, "simt"), (1, 260, 36, 13, "simt"), (1, 200, 40, 48, "tile")]
if not small:
    shapes += [(2, 300, 64, 128, "tc"), (1, 260, 64, 64, "tc"), (1, 200, 256, 256, "tc"), (1, 200, 32, 32, "tc"),
               (1, 260, 36, 13, "auto"), (1, 200, 100, 100, "auto")]      # zero-padded onto the tensor-core kernels
for B, N, Cin, Cout, eng in shapes:
    pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, "room", seed=1).items()}
    prev = set_engine(eng)
    for stride in (1, 2):
        plan = NeighborPlan(pr["points"], stride, 0.1)
        y = conv3p_forward(plan, pr["input"], pr["filter"], activation="selu")
        gi, gf = conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
        _, gf2 = conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"], need_input_grad=False)
    set_engine(prev)
    torch.cuda.synchronize()
    print(B, N, Cin, Cout, eng, float(y.abs().mean()), float(gi.abs().mean()), float(gf.abs().mean()), bool(torch.equal(gf, gf2)))
# the general path (other filter shapes, T = double)
from pointwise_b200 import conv3p  # noqa: E402
for dt in (torch.float32, torch.float64):
    pr = make_problem(2, 200, 3, 4, "room", seed=2)
    P, X = torch.from_numpy(pr["points"]).cuda().to(dt), torch.from_numpy(pr["input"]).cuda().to(dt).requires_grad_()
    W = (torch.rand(2, 3, 2, 3, 4, device="cuda", dtype=dt) - 0.5).requires_grad_()
    y = conv3p(P, X, W, [1, 2, 1], [0.1])
    y.sum().backward()
    torch.cuda.synchronize()
    print("general", dt, float(y.abs().mean()), float(X.grad.abs().mean()), float(W.grad.abs().mean()))
print("ok")
