"""Hardware probes for round-2 design questions (not tests): run on the B200 box.
 1. Does a K-major tcgen05 operand accept the SWIZZLE_128B_BASE32B layout (so one shared-memory image of the
    gathered rows can feed both the K-major grad_input MMA and the MN-major grad_filter MMA)?
 2. Where do the rows of an M = 64 accumulator land in TMEM (cta_group::1)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointwise_b200 import _lib
L = _lib.lib()
rng = np.random.default_rng(0)
N, K = 64, 64
A = rng.uniform(-1, 1, (128, K)).astype(np.float32); B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
want = A.astype(np.float64) @ B.astype(np.float64).T
a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
def run(flags):
    d = torch.full((128, N), float("nan"), device="cuda")
    st = L.conv3p_selftest_tc(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, flags, None)
    torch.cuda.synchronize()
    return st, d.cpu().numpy().astype(np.float64)
st, d = run(1)
print("baseline 3xTF32 max err", np.abs(d - want).max())
st, d = run(1 | 2)
print("probe 1: K-major + SWIZZLE_128B_BASE32B: status", st, "max err", np.abs(d - want).max(),
      "-> WORKS" if np.abs(d - want).max() < 1e-4 else "-> does NOT reproduce A*B^T")
st, d = run(1 | 4)
print("probe 2: M=64: status", st)
for lane in range(0, 128, 8):
    row = d[lane]
    match = [r for r in range(128) if np.abs(row - want[r]).max() < 1e-4]
    print(f"   TMEM lane {lane:3d} holds A row {match if match else ('zeros' if np.abs(row).max() < 1e-6 else 'other')}")
