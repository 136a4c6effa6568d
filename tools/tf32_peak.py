"""Measures the TF32 tensor-core peak the roofline's frac_F divides by ("to be measured by the
builder"): (a) cuBLAS TF32 GEMM through torch.matmul with allow_tf32 (8192^3, best of 10 and sustained), (b) this
repo's own one-CTA tcgen05 kind::tf32 loop (conv3p_selftest_tc, N=256: issue rate of a single SM, scaled by the SM
count -- an upper bound no fused kernel can exceed).  Writes profiles/tf32_peak.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n ** 3 / (best * 1e-3) / 1e12
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
it = 0
while time.perf_counter() - t0 < 3.0:
    for _ in range(10):
        a @ b
    it += 10
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = 2 * n ** 3 * it / (e0.elapsed_time(e1) * 1e-3) / 1e12
torch.backends.cuda.matmul.allow_tf32 = False
e0.record()
a @ b
e1.record()
torch.cuda.synchronize()
fp32 = 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "cublas_fp32_tflops_same_call_no_tf32": fp32,
       "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32 tensor-core GEMM): best of 10 (burst), back to "
              "back for 3 s (sustained); CUDA events",
       "gpu": torch.cuda.get_device_name(0)}
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    out["bf16_tflops_measured_peaks_json"] = json.load(open(peaks)).get("bf16_tflops")
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tf32_peak.json", "w"), indent=1)
