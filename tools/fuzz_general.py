"""Randomised sweep of the general path (filter shapes other than 3x3x3, and T = double for every shape) against the
REFERENCE's own object code (oracle/_ref, present wherever it was built): count tables bit-exact, sums within a
tolerance of the result's scale (1e-12 double; 2e-5 x 8 float, the reference sums in another order in fp32).
usage: python tools/fuzz_general.py [seconds] [seed]   (test infrastructure: imports oracle/)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
from pointwise_b200 import _lib, conv3p  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

if not oracle.Ref.available():
    print("oracle/_ref not built: nothing to compare with")
    sys.exit(0)
R = oracle.ref()
L = _lib.lib()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
i3 = C.c_int * 3
t0, cases = time.time(), {"f32": 0, "f64": 0}


def close(got, want, tol, what):
    scale = float(np.abs(want).max()) + 1e-300
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    assert err <= tol * scale + (1e-6 if got.dtype == np.float32 else 0.0), f"{what}: |err| {err:.3e} at scale {scale:.3e}"


while time.time() - t0 < budget:
    f64 = bool(rng.random() < 0.5)
    dims = tuple(int(d) for d in rng.integers(1, 6, 3))
    if not f64 and dims == (3, 3, 3):
        continue                                  # the tuned engines' case: tools/fuzz_parity.py
    B, N = int(rng.integers(1, 4)), int(rng.choice([1, 5, 64, 200, 500]))
    Cin, Cout = int(rng.integers(1, 9)), int(rng.integers(1, 9))
    stride = tuple(int(s) for s in rng.integers(1, 4, 3))
    voxel = float(rng.choice([0.1, 0.07]))
    dist = str(rng.choice(["room", "sphere", "cube"]))
    quant = 0.05 if rng.random() < 0.3 else None
    pr = make_problem(B, N, Cin, Cout, dist, seed=int(rng.integers(1 << 30)), quantise=quant)
    dt = np.float64 if f64 else np.float32
    P = pr["points"].astype(dt)
    if f64:
        P = P + rng.uniform(-1e-9, 1e-9, P.shape)
    X, G = pr["input"].astype(dt), pr["grad_out"].astype(dt)
    W = rng.uniform(-0.1, 0.1, (*dims, Cin, Cout)).astype(dt)
    if f64:
        want = R.forward64(P, X, W, stride, voxel)
        wgi, wgf = R.backward64(G, P, X, W, stride, voxel)
    else:
        want = R.forward(P, X, W, stride, np.float32(voxel))
        wgi, wgf = R.backward(G, P, X, W, stride, np.float32(voxel))
    Xt, Wt = torch.from_numpy(X).cuda().requires_grad_(), torch.from_numpy(W).cuda().requires_grad_()
    y = conv3p(torch.from_numpy(P).cuda(), Xt, Wt, list(stride), [voxel])
    y.backward(torch.from_numpy(G).cuda())
    tol = 1e-12 if f64 else 2e-5 * 8
    tag = f"{'f64' if f64 else 'f32'} {dims} s{stride} v{voxel} B{B} N{N} {Cin}->{Cout} {dist} q{quant}"
    close(y.detach().cpu().numpy(), want, tol, "output " + tag)
    close(Xt.grad.cpu().numpy(), wgi, tol, "grad_input " + tag)
    close(Wt.grad.cpu().numpy(), wgf, tol, "grad_filter " + tag)
    cases["f64" if f64 else "f32"] += 1
print(f"fuzz ok: {cases['f32']} float and {cases['f64']} double random cases of the general path in {time.time() - t0:.0f} s")
