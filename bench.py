#!/usr/bin/env python
"""bench.py -- Conv3p fwd+bwd throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload "headline"): the configuration the metric is quoted on -- synthetic S3DIS-like
clouds, N=4096 points, Cin=64 -> Cout=128, 3x3x3 filter, stride 1, voxel 0.1, fp32; 64 clouds per GPU
(262,144 points), fixed per GPU as N grows (weak scaling; the batch is sharded across ranks, one NCCL
all-reduce on grad_filter).

A step = one pass of the hot path over one batch: neighbour plan (voxel sort + search + backward
lists), Conv3p forward, Conv3pGrad (grad_input + grad_filter).  `value` = points of all ranks / step
time with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over ranks).
`e2e` = the same step through the public Python/C-ABI call with HOST buffers: pinned host -> device
copies of points/input/filter/grad_out and device -> host reads of output/grad_input/grad_filter inside
the timed region.  `roofline` is for the dominant kernel, timed live with CUDA events on its stream.
`cpu_baseline` is the reference's own CPU op (oracle/_ref, compiled unmodified) on a bounded sample of
the same workload, on this box's host cores.

--impl reference times that CPU op alone (rank 0 only), each step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (clouds per GPU, N, Cin, Cout, stride, distribution)
    "headline": (64, 4096, 64, 128, (1, 1, 1), "room"),
    "headline_b16": (16, 4096, 64, 128, (1, 1, 1), "room"),   # BASELINE configs[4]: 128 clouds over 8 GPUs
    "s3dis_l1": (16, 4096, 9, 9, (1, 1, 1), "room"),
    "s3dis_l5": (16, 4096, 36, 13, (1, 1, 1), "room"),
    "modelnet_l2": (32, 1024, 9, 9, (2, 2, 2), "sphere"),
}
VOXEL = 0.1
METRIC = "conv3p_fwd_bwd_points_per_sec"
UNIT = "points/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback")


# ------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own op (oracle/_ref) or, if it was never built, the C port
# ------------------------------------------------------------------------------------------------------
def cpu_checker():
    import oracle
    if oracle.Ref.available():
        return oracle.ref()
    oracle.build()
    return oracle.best()


def cpu_sample(workload: str, clouds: int, seed: int = 0):
    from pointwise_b200.synth import make_problem
    _, N, Cin, Cout, stride, dist = WORKLOADS[workload]
    return make_problem(clouds, N, Cin, Cout, dist, seed=seed), stride


def cpu_step(chk, pr, stride):
    chk.forward(pr["points"], pr["input"], pr["filter"], stride, VOXEL)
    chk.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, VOXEL)


def run_cpu_baseline(workload: str):
    """The reference CPU op timed on a bounded sample of the workload, in a clean subprocess (no CUDA
    context, no torch threads): exactly `bench.py --impl reference --steps 2 --warmup 1`."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--workload", workload]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
        env.pop(k, None)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
            "sample": (out.stderr or "no output")[-200:]}


def reference_arm(args):
    """bench.py --impl reference: the reference CPU op alone, on this arm's config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chk = cpu_checker()
    cores = os.cpu_count() or 1
    threads = min(cores, chk.threads)
    per_gpu, N, Cin, Cout, stride, dist = WORKLOADS[args.workload]
    clouds = threads                                  # one cloud per host thread per step
    pr, stride = cpu_sample(args.workload, clouds)
    for _ in range(args.warmup):
        cpu_step(chk, pr, stride)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(chk, pr, stride)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = clouds * N / dt
    sample = f"{clouds} clouds x {N} points per step, {Cin}->{Cout}, fwd+bwd"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus, cpu_sample=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": chk.kind, "sample": sample,
                         "host_cpus": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(workload, gpus, **extra):
    per_gpu, N, Cin, Cout, stride, dist = WORKLOADS[workload]
    d = {"workload": workload, "clouds_per_gpu": per_gpu, "global_batch": per_gpu * gpus, "num_points": N,
         "cin": Cin, "cout": Cout, "filter": "3x3x3", "stride": list(stride), "voxel_size": VOXEL,
         "distribution": dist, "parallelism": f"batch-sharded dp{gpus}, one all-reduce(grad_filter)",
         "l2_policy": "working set per step (inputs+outputs+plan, ~0.7 GB) exceeds the 126 MB L2"}
    d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel: str, pts: int, Cin: int, Cout: int, kbar: float, kbar_b: float,
                      nbins_b: float = 0.0, shared: bool = False):
    """Algorithmic bytes per launch (DESIGN.md section 'Kernels and their roofs'), gather model G of
    SURVEY section 8d: every list entry is one index read plus one row read.  With the shared gather (`shared`:
    both gradients on the tensor-core kernels) the grad_input kernel also writes one Cout-wide row per non-empty
    (point, cell) slot (nbins_b of them per point) and the grad_filter kernel reads those rows back instead of
    walking the lists."""
    nW = 27 * Cin * Cout * 4
    per_point = {
        "k_gather_contract_fwd": kbar * (4 * Cin + 4) + 27 * 4 + 8 + 16 + 4 * Cout,
        "k_gather_contract_bwd_input": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_forward_tc": kbar * (4 * Cin + 4) + 27 * 4 + 8 + 16 + 4 * Cout,
        "k_backward_input_tc": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_backward_filter": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_backward_filter_tc": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_small_forward": kbar * (4 * Cin + 4) + 27 * 4 + 8 + 16 + 4 * Cout,
        "k_small_backward_input": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_small_backward_filter": kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin,
        "k_neighbor_search": 16 + 27 * 4 + 12 + kbar * 4 + 2.4 * kbar * 16,
        "k_backward_lists": 16 + kbar * (4 + 12 + 4) + 27 * 4 + kbar_b * 8,
        "k_cloud_sort": 12 + 16 + 4 + 4 * 16,
    }.get(kernel, 0.0)
    if shared and kernel == "k_backward_input_tc":
        per_point += nbins_b * 4 * Cout
    if shared and kernel == "k_backward_filter_tc":
        per_point = nbins_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin
    fixed = nW if kernel.startswith("k_gather") or kernel.endswith("_tc") or kernel.startswith("k_small_") else (2 * nW if kernel.startswith("k_backward_filter") else 0)
    return per_point * pts + fixed


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pointwise_b200 import NeighborPlan, _lib, conv3p_backward, conv3p_forward, launch_count
    from pointwise_b200.distributed import allreduce_grad_filter, shard_range
    from pointwise_b200.synth import make_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)
    per_gpu, N, Cin, Cout, stride, dname = WORKLOADS[args.workload]
    B_global = per_gpu * world
    lo, hi = shard_range(B_global, rank, world)
    # every rank generates only its own shard (seeded by rank) -- clouds are independent
    pr = make_problem(hi - lo, N, Cin, Cout, dname, seed=rank)
    pr["filter"] = make_problem(1, 8, Cin, Cout, dname, seed=0)["filter"]  # replicated weights
    host = {k: torch.from_numpy(v).pin_memory() for k, v in pr.items()}
    devt = {k: v.to(device) for k, v in host.items()}
    pts = (hi - lo) * N
    L = _lib.lib()

    def step_device():
        plan = NeighborPlan(devt["points"], stride, VOXEL, check=False, capacity=capacity)
        y = conv3p_forward(plan, devt["input"], devt["filter"])
        plan.prefetch_backward()         # what the autograd op does when a gradient is required
        gi, gf = conv3p_backward(plan, devt["grad_out"], devt["input"], devt["filter"])
        if world > 1:
            allreduce_grad_filter(gf)
        return plan, y, gi, gf

    # capacity: learned once (checked build), then every step runs without a host read-back
    probe = NeighborPlan(devt["points"], stride, VOXEL).ensure_backward()
    st = probe.read_stats()
    capacity = int(st.total_pairs * 1.05) + 1024
    kbar = st.total_pairs / pts
    kbar_b = st.backward_pairs / pts
    nbins = float((probe.count_table > 0).sum().item()) / pts
    nbins_b = float((probe.backward_count_table > 0).sum().item()) / pts
    # both gradients of this shape run on the tensor-core kernels with the shared gather (G store)?
    shared = bool(L.conv3p_backward_scratch_bytes(probe.geom, Cin, Cout) > L.conv3p_scratch_bytes(probe.geom, Cin, Cout))
    del probe

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.1 s to start: begin before the warm-up steps
    for _ in range(max(3, args.warmup)):
        step_device()
    launch_count(reset=True)
    L.conv3p_profile_enable(1)
    ms_step = timed(step_device, args.steps)
    launches = launch_count() // args.steps
    buf = (b"\0" * 8192)
    import ctypes as C
    cbuf = C.create_string_buffer(8192)
    L.conv3p_profile_read(cbuf, 8192)
    L.conv3p_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    kern = {}
    for ln in cbuf.value.decode().splitlines():
        name, n, total = ln.split()
        kern[name] = (int(n), float(total))
    value = B_global * N / (ms_step * 1e-3)

    # e2e (host buffers): every step copies its pinned-host inputs to the device, rebuilds the plan, runs
    # forward + backward (+ the all-reduce) and copies output / grad_input / grad_filter back to pinned host
    # memory; copies of neighbouring steps overlap the kernels (three streams, three staging slots).
    from pointwise_b200.host_api import HostConv3p
    pipe = HostConv3p(hi - lo, N, Cin, Cout, stride, VOXEL, device=device, capacity=capacity,
                      depth=int(os.environ.get("CONV3P_HOST_DEPTH", "3")))
    ar = allreduce_grad_filter if world > 1 else None

    def run_e2e(steps):
        tickets = []
        for _ in range(steps):
            tickets.append(pipe.submit(host["points"], host["input"], host["filter"], host["grad_out"], ar))
            if len(tickets) >= pipe.depth:
                pipe.fetch(tickets[-pipe.depth])
        for t in tickets[-(pipe.depth - 1):]:
            pipe.fetch(t)

    run_e2e(max(8, args.warmup))      # the first steps of a fresh process also fault in the pinned staging buffers
    e2e_steps = max(4, args.steps)
    barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / e2e_steps
    wall_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {})
    top = max(kern, key=lambda k: kern[k][1]) if kern else None
    roof = None
    kernels = {}
    for name, (n, total) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
        avg_ms = total / n
        ab = algorithmic_bytes(name, pts, Cin, Cout, kbar, kbar_b, nbins_b, shared)
        kernels[name] = {"launches_per_step": n // args.steps, "avg_ms": round(avg_ms, 4),
                         "share_of_step": round(total / args.steps / ms_step, 4),
                         "algorithmic_gbs": round(ab / (avg_ms * 1e-3) / 1e9, 1)}
    if top:
        n, total = kern[top]
        avg_s = total / n * 1e-3
        ab = algorithmic_bytes(top, pts, Cin, Cout, kbar, kbar_b, nbins_b, shared)
        flops = 2.0 * 27 * Cin * Cout * pts
        roof = {"kernel": top, "bound": "hbm", "achieved": ab / avg_s / 1e9, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": ab / avg_s / 1e9 / pk["hbm_gbs"],
                "traffic": traffic.get(top, {}).get("dram_bytes_per_launch"),
                "algorithmic_bytes_per_launch": ab,
                "peak_source": pk["source"] + " (MEASURED_PEAKS.json hbm_gbs)",
                "model": "gather model G (SURVEY 8d): list entries x (index + row bytes) + per-point I/O"
                         + (" + G-store rows" if shared else ""),
                "dense_tflops": flops / avg_s / 1e12,
                "frac_fp32_simt_peak_74.4": flops / avg_s / 1e12 / 74.4}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, world, mean_neighbours=round(kbar, 2),
                              mean_backward_pairs=round(kbar_b, 2), mean_nonempty_cells=round(nbins, 2),
                              mean_nonempty_backward_cells=round(nbins_b, 2), shared_backward_gather=shared),
        "clocks": clocks,
        "e2e": {"value": B_global * N / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "wall_ms_per_step": wall_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "pointwise_b200.host_api.HostConv3p (pinned host in/out, copies overlapped with compute)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = run_cpu_baseline(args.workload)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it can
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    gpu_arm(args)


if __name__ == "__main__":
    main()
