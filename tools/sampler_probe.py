"""Does the clock sampler disturb the timed region?  Runs the 16-cloud fwd+bwd step 300 times per setting with
nvidia-smi polling at different periods (and with / without the clocks_event_reasons fields) and reports the step-time
outliers.  usage: python tools/sampler_probe.py"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import Runner  # noqa: E402

run = Runner("headline_b16", 0, 1, torch.device("cuda", 0))
for _ in range(10):
    run.step()
torch.cuda.synchronize()
FULL = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
out = []
for label, q, period in [("none", None, 0), ("full 20 ms", FULL, 20), ("full 100 ms", FULL, 100), ("full 200 ms", FULL, 200),
                         ("clocks only 20 ms", "clocks.sm,clocks.max.sm", 20), ("none again", None, 0)]:
    proc = None
    if q:
        proc = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(period)],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.0)
    n = 300
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    marks[0].record()
    for i in range(n):
        run.step()
        marks[i + 1].record()
    torch.cuda.synchronize()
    if proc:
        proc.terminate()
        proc.wait()
    per = np.array([a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:])])
    med = float(np.median(per))
    out.append({"sampler": label, "median_ms": round(med, 4), "mean_ms": round(float(per.mean()), 4), "max_ms": round(float(per.max()), 3),
                "steps_over_1.5x_median": int((per > 1.5 * med).sum()), "extra_ms_total": round(float((per - med).clip(0).sum()), 2)})
    print(json.dumps(out[-1]), flush=True)
