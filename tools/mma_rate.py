"""Issue rate of tcgen05.mma on one SM for the operand layouts the kernels use (conv3p_debug_mma_rate): cycles per
instruction for K-major / MN-major x TF32 / BF16 x N, next to the tensor-core floor 128 * N / 256 cycles and the
shared-memory floor (operand bytes / 128 B per cycle).  The TF32 tensor peak of one SM follows from the fastest case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointwise_b200 import _lib  # noqa: E402

L = _lib.lib()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
reps = 2000
rows = []
for mode, name in [(0, "K-major tf32"), (1, "MN-major tf32"), (2, "K-major bf16"), (3, "MN-major bf16")]:
    for N in (64, 128, 256):
        _lib.check(L.conv3p_debug_mma_rate(N, mode, 50, out.data_ptr(), None))
        _lib.check(L.conv3p_debug_mma_rate(N, mode, reps, out.data_ptr(), None))
        torch.cuda.synchronize()
        cyc = int(out.item()) / (reps * 8)
        k = 16 if mode & 2 else 8
        esz = 2 if mode & 2 else 4
        smem_floor = (128 + N) * k * esz / 128.0
        rows.append({"layout": name, "N": N, "cycles_per_mma": round(cyc, 1), "tensor_floor": 128 * N / 256,
                     "smem_floor": smem_floor, "flops_per_cycle_sm": round(2 * 128 * N * k / cyc, 0)})
        print(f"{name:14s} N={N:3d}: {cyc:7.1f} cycles per MMA (tensor floor {128 * N / 256:.0f}, shared-memory floor {smem_floor:.0f})")
sm = torch.cuda.get_device_properties(0)
best = max(r["flops_per_cycle_sm"] for r in rows if "tf32" in r["layout"])
mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
print(f"best TF32 rate: {best:.0f} flop/cycle/SM -> x {sm.multi_processor_count} SMs x 1.965 GHz = "
      f"{best * sm.multi_processor_count * 1.965e9 / 1e12:.0f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"rows": rows, "reps": reps, "best_tf32_flop_per_cycle_sm": best,
           "tf32_tflops_own_loop_at_1965mhz": best * sm.multi_processor_count * 1.965e9 / 1e12},
          open("gpurun_out/mma_rate.json", "w"), indent=1)
