"""Randomised parity sweep on the GPU: random batch shapes, channel counts (tensor-core shapes, zero-padded ones, tiny
ones), per-axis strides, voxel sizes and point distributions (incl. quantised coordinates and duplicated points), each
checked against the CPU oracle -- count tables bit-exact, sums inside |got - sum64| <= 1e-7 + 1e-5 * sum|terms|.
usage: python tools/fuzz_parity.py [seconds] [seed]   (test infrastructure: imports oracle/)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from helpers import assert_close_scaled  # noqa: E402
from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

CH = [1, 3, 4, 8, 9, 13, 16, 17, 20, 24, 32, 33, 36, 40, 48, 64, 70, 96, 100, 128]


def run(budget: float, seed: int):
    """-> (cases, worst error / sum|terms|); raises AssertionError on the first case out of tolerance."""
    rng = np.random.default_rng(seed)
    oracle.build()
    port = oracle.port()
    t0, cases, worst = time.time(), 0, 0.0
    while time.time() - t0 < budget:
        cases, worst = cases + 1, max(worst, one_case(rng, port))
    return cases, worst


def one_case(rng, port) -> float:
    B, N = int(rng.integers(1, 5)), int(rng.choice([1, 2, 7, 63, 128, 129, 300, 700, 1500]))
    Cin, Cout = int(rng.choice(CH)), int(rng.choice(CH))
    if Cin * Cout > 64 * 128:
        N = min(N, 300)                     # keeps the CPU checker in seconds
    stride = tuple(int(s) for s in rng.integers(1, 5, 3)) if rng.random() < 0.5 else (int(rng.integers(1, 5)),) * 3
    voxel = float(rng.choice([0.1, 0.05, 0.23]))
    dist = str(rng.choice(["room", "sphere", "cube"]))
    quant = float(rng.choice([0.05, 0.025])) if rng.random() < 0.3 else None
    pr = make_problem(B, N, Cin, Cout, dist, seed=int(rng.integers(1 << 30)), quantise=quant)
    if rng.random() < 0.2 and N > 4:        # duplicated points
        pr["points"][:, N // 2:] = pr["points"][:, :N - N // 2]
    d = {k: torch.from_numpy(v).cuda() for k, v in pr.items()}
    # check="sync": shapes repeat here with very different densities (duplicated points triple the pairs), which is
    # exactly what the learned-capacity default reports as an overflow to be rerun
    plan = NeighborPlan(d["points"], stride, voxel, check="sync")
    y = conv3p_forward(plan, d["input"], d["filter"]).cpu().numpy()
    gi, gf = conv3p_backward(plan, d["grad_out"], d["input"], d["filter"])
    cnt = plan.count_table.cpu().numpy()
    for b in range(B):
        assert np.array_equal(cnt[b], port.neighbor_count(pr["points"][b], stride, voxel)), ("count table", B, N, stride, voxel, dist)
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], stride, voxel, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, voxel, with64=True)
    tag = f"B{B} N{N} {Cin}->{Cout} s{stride} v{voxel} {dist} q{quant}"
    return max(assert_close_scaled(y, o64, oabs, 1e-5, 1e-7, "forward " + tag),
               assert_close_scaled(gi.cpu().numpy(), r[2], r[3], 1e-5, 1e-7, "grad_input " + tag),
               assert_close_scaled(gf.cpu().numpy(), r[4], r[5], 1e-5, 1e-7, "grad_filter " + tag))


if __name__ == "__main__":
    t_ = time.time()
    n_, w_ = run(float(sys.argv[1]) if len(sys.argv) > 1 else 120.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"fuzz ok: {n_} random cases in {time.time() - t_:.0f} s, worst error / sum|terms| = {w_:.2e} (bound 1e-5)")
