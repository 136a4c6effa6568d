"""Engine timing (tensor cores vs fp32 SIMT) and the phase timers of the gather+MMA kernel (engine flag 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import VOXEL, WORKLOADS
from pointwise_b200 import NeighborPlan, _lib, conv3p_forward, conv3p_backward
from pointwise_b200.synth import make_problem
B, N, Cin, Cout, stride, dist = WORKLOADS["headline"]
pr = {k: torch.from_numpy(v).cuda() for k, v in make_problem(B, N, Cin, Cout, dist, seed=0).items()}
plan = NeighborPlan(pr["points"], stride, VOXEL).ensure_backward()
L = _lib.lib()
for eng, name in [(0, "auto"), (1, "simt")]:
    L.conv3p_set_engine(eng)
    for fn, label in [(lambda: conv3p_forward(plan, pr["input"], pr["filter"]), "fwd"),
                      (lambda: conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"]), "bwd")]:
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:10s} {label:7s} {e0.elapsed_time(e1)/5:.3f} ms")
L.conv3p_set_engine(0)

import ctypes as C
buf = (C.c_ulonglong * 8)()
for label, fn in [("forward", lambda: conv3p_forward(plan, pr["input"], pr["filter"])),
                  ("backward", lambda: conv3p_backward(plan, pr["grad_out"], pr["input"], pr["filter"]))]:
    L.conv3p_set_engine(32)
    L.conv3p_debug_phase_cycles(buf)
    cta = (C.c_ulonglong * 2)()
    L.conv3p_debug_cta_cycles(cta)
    fn(); torch.cuda.synchronize()
    L.conv3p_debug_phase_cycles(buf)
    L.conv3p_debug_cta_cycles(cta)
    L.conv3p_set_engine(0)
    tiles = min(torch.cuda.get_device_properties(0).multi_processor_count, (B * N + 127) // 128)
    v = [x / tiles for x in buf]
    print(f"{label}: CTA total cycles mean {cta[0] / tiles:.0f}, max {cta[1]} (the launch ends with the slowest CTA)")
    print(f"{label}: k_gather_mma2 cycles per CTA (thread 0): prologue {v[0]:.0f} | producer loop {v[1]:.0f} | "
          f"wait last MMA {v[2]:.0f} | epilogue {v[3]:.0f} || in loop: item fetch {v[4]:.0f} | gather 0 {v[5]:.0f} | "
          f"ring wait {v[6]:.0f} | stores+rep 1+arrive {v[7]:.0f}")
