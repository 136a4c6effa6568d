"""Times the T = double operator (conv3p_op_{forward,backward}_f64, general fp64 SIMT path) next to the reference's own
double CPU kernels (oracle/_ref) on the same box, and checks the two against each other on the sample.
usage: python tools/f64_timing.py > gpurun_out/f64_timing.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402  (checker / CPU baseline leg only)
from pointwise_b200 import conv3p  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

out = []
for name, B, N, Cin, Cout, dist in [("seg layer 9->9", 16, 4096, 9, 9, "room"), ("cls layer 9->9", 32, 1024, 9, 9, "sphere"),
                                     ("64->128", 4, 4096, 64, 128, "room")]:
    rng = np.random.default_rng(3)
    P = make_problem(B, N, 1, 1, dist, seed=5)["points"].astype(np.float64) + rng.uniform(-1e-9, 1e-9, (B, N, 3))
    Xn, Wn, Gn = rng.uniform(-1, 1, (B, N, Cin)), rng.uniform(-0.1, 0.1, (3, 3, 3, Cin, Cout)), rng.uniform(-1, 1, (B, N, Cout))
    Pd, Gd = torch.from_numpy(P).cuda(), torch.from_numpy(Gn).cuda()

    def step():
        X, W = torch.from_numpy(Xn).cuda().requires_grad_(), torch.from_numpy(Wn).cuda().requires_grad_()
        y = conv3p(Pd, X, W, [1, 1, 1], [0.1])
        y.backward(Gd)
        return y, X.grad, W.grad

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y, gi, gf = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rec = {"case": name, "clouds": B, "N": N, "Cin": Cin, "Cout": Cout, "gpu_ms_per_step": ms,
           "gpu_points_per_s": B * N / (ms * 1e-3), "dtype": "f64"}
    if oracle.Ref.available():
        R = oracle.ref()
        nb = min(B, 2)
        t0 = time.perf_counter()
        want = R.forward64(P[:nb], Xn[:nb], Wn, [1, 1, 1], 0.1)
        wgi, _ = R.backward64(Gn[:nb], P[:nb], Xn[:nb], Wn, [1, 1, 1], 0.1)
        dt = time.perf_counter() - t0
        rec["cpu_reference_points_per_s"] = nb * N / dt
        rec["cpu_threads"] = R.threads
        rec["cpu_sample"] = f"{nb} clouds"
        rec["output_max_abs_err"] = float(np.abs(y[:nb].detach().cpu().numpy() - want).max())
        rec["grad_input_max_abs_err"] = float(np.abs(gi[:nb].cpu().numpy() - wgi).max())
    out.append(rec)
print(json.dumps(out))
