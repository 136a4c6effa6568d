"""A/B timing of the tensor-core kernels on one workload: per-kernel CUDA-event times of forward + backward, with
the G store shared by the two gradient kernels (engine 0) and without (engine flag 256), plus -- for library variants
built with -DC3P_W2_TIMED=1 -- the phase timers of the weight-gradient producers.
usage: [CONV3P_LIB=pointwise_b200/lib/variants/<name>.so] python tools/ab_backward.py [workload] [engine flags ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import VOXEL, WORKLOADS  # noqa: E402
from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

B, N, Cin, Cout, stride, dist = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "headline"]
flags = [int(a) for a in sys.argv[2:]] or [0]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
plan = NeighborPlan(pr["points"], stride, VOXEL).ensure_backward()
L = _lib.lib()
print("library:", os.environ.get("CONV3P_LIB", "production"))
for eng in flags:
    L.conv3p_set_engine(eng)
    for _ in range(3):
        conv3p_forward(plan, pr["input"], pr["filter"])
        conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
    torch.cuda.synchronize()
    buf8 = (C.c_ulonglong * 16)()
    L.conv3p_debug_w2_cycles(buf8)
    L.conv3p_profile_enable(1)
    for _ in range(10):
        conv3p_forward(plan, pr["input"], pr["filter"])
        conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"])
    torch.cuda.synchronize()
    buf = C.create_string_buffer(8192)
    L.conv3p_profile_read(buf, 8192)
    L.conv3p_profile_enable(0)
    print(f"engine flags {eng}:", "  ".join(f"{ln.split()[0]} {float(ln.split()[2]) / int(ln.split()[1]):.4f}"
                                            for ln in buf.value.decode().splitlines() if "_tc" in ln or "group_items" in ln))
    L.conv3p_debug_w2_cycles(buf8)
    if any(buf8):
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        v = [x / sms / 10 for x in buf8]
        print(f"   weight-gradient producer warp 0, cycles per CTA: item-list barrier {v[7]:.0f}, rest of the look-ahead {v[0]:.0f} | rows arrive {v[1]:.0f} | "
              f"ring-slot wait {v[2]:.0f} | split+stores+fence+arrive {v[3]:.0f} | input panels {v[4]:.0f} | "
              f"flush {v[5]:.0f} | total {v[6]:.0f}")
        print(f"   item loader: wait for a free slot {v[8]:.0f} of {v[9]:.0f} | MMA issuer: wait input panels {v[10]:.0f}, "
              f"wait G stage {v[11]:.0f}, wait flush {v[12]:.0f} of {v[13]:.0f}")
L.conv3p_set_engine(0)
