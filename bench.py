#!/usr/bin/env python
"""bench.py -- Conv3p fwd+bwd throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

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

Outside the timed regions the same process also
  * compares three clouds of the very batch it timed with the CPU checker (`parity_check`),
  * under --gpus N checks the all-reduced grad_filter: bit-identical on every rank and equal to the sum of the
    per-rank gradients (`scale_check`),
  * times BASELINE configs[4]'s per-GPU share, 16 clouds per GPU, with and without the collective (`b16`),
  * on one GPU, runs the 9-point configs[3] sweep N in {1k,4k,16k} x C in {9,64,256} (`sweep`, --no-sweep skips).

--impl reference times the reference CPU op alone (rank 0 only), each step a bounded sample of the workload.
--workload seg_net / cls_net time a whole PointConvNet step (BASELINE configs[2] / configs[1]) through conv3p().
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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
# BASELINE configs[3]: N in {1k, 4k, 16k} x C in {9, 64, 256}, C -> C, B = 2^18 / N so every point has 262,144 points
SWEEP = []
for _n, _nn in (("1k", 1024), ("4k", 4096), ("16k", 16384)):
    for _c in (9, 64, 256):
        WORKLOADS[f"sweep_n{_n}_c{_c}"] = ((1 << 18) // _nn, _nn, _c, _c, (1, 1, 1), "room")
        SWEEP.append(f"sweep_n{_n}_c{_c}")
NETS = {
    # name: (clouds per GPU, N, input channels, classes, distribution) -- BASELINE configs[2] / configs[1]
    "seg_net": (16, 4096, 9, 13, "room"),
    "cls_net": (32, 1024, 3, 40, "sphere"),
}
VOXEL = 0.1
METRIC = "conv3p_fwd_bwd_points_per_sec"
UNIT = "points/s"
RTOL, ATOL = 1e-5, 1e-7       # |got - sum64| <= ATOL + RTOL * sum|terms|  (tests/test_gpu_parity.py)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    pk = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")
    if os.path.exists(path):
        p = json.load(open(path))
        pk = dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    # TF32 tensor peak measured with this repo's own tcgen05 loop and with cuBLAS (tools/tf32_peak.py)
    tpath = os.path.join(ROOT, "profiles", "tf32_peak.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        pk["tf32_tflops"] = float(t["tf32_tflops"])
        pk["tf32_source"] = t.get("how", "profiles/tf32_peak.json")
    else:
        pk["tf32_tflops"] = pk["bf16_tflops"] / 2
        pk["tf32_source"] = "assumed half of the BF16 peak (profiles/tf32_peak.json absent)"
    return pk


# ------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode (one sample per 100 ms: the recipe's own line uses 200 ms; the GPU is under load from the
    warm-up steps on, so a timed region shorter than the period is still represented) writing to a temporary FILE, read
    back after the run: a reader thread in this process
    wakes up on every sample and contends for the GIL with the thread that launches the kernels -- with the 5 ms switch
    interval that showed up as single 8-20 ms steps in the short layer workloads."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.file = None
        self.lines = []

    def start(self):
        import tempfile
        try:
            self.file = tempfile.NamedTemporaryFile(prefix="conv3p_clocks_", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def ready(self, timeout=3.0):
        """Block until the first sample arrived: nvidia-smi attaching to the GPU stalls launches for tens of ms, which
        must not fall into the timed region (a 16-cloud step showed 5.2 ms instead of 1.2 ms when it did)."""
        t0 = time.perf_counter()
        while (self.proc and os.path.getsize(self.file.name) == 0 and time.perf_counter() - t0 < timeout
               and self.proc.poll() is None):
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        try:
            self.file.close()
            with open(self.file.name) as fh:
                self.lines = [ln.strip() for ln in fh]
            os.unlink(self.file.name)
        except OSError:
            pass
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


def bind_to_gpu_numa(index: int):
    """Best effort: pin this process (and with it the first-touch placement of its pinned host buffers) to the
    CPUs of the GPU's NUMA node, so the host-buffer path of every rank uses its own memory controller and PCIe
    root.  Returns what was done (recorded in config)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False, "why": "single node / not reported"}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if not ids:
            return {"numa_node": node, "bound": False, "why": "node cpus outside the allowed set"}
        os.sched_setaffinity(0, ids)
        return {"numa_node": node, "bound": True, "cpus": len(ids)}
    except Exception as e:  # pragma: no cover - depends on the box
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}


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


def cpu_net_step(chk, name, clouds, seed=0):
    """The Conv3p layers of the reference network (forward + Conv3pGrad of every layer) on the CPU op.  The dense
    layers / SELU / loss of the network are not on the hot path and are left out of the CPU arm."""
    from pointwise_b200.synth import make_points
    _, N, cin, ncls, dist = NETS[name]
    rng = np.random.default_rng(seed)
    pts = make_points(clouds, N, dist, seed=seed)
    layers = [(cin, 9, 1), (9, 9, 2), (9, 9, 3), (9, 9, 4)] + ([(36, ncls, 1)] if name == "seg_net" else [])
    t0 = time.perf_counter()
    for ci, co, s in layers:
        x = rng.uniform(-1, 1, (clouds, N, ci)).astype(np.float32)
        w = rng.uniform(-0.1, 0.1, (3, 3, 3, ci, co)).astype(np.float32)
        g = rng.uniform(-1, 1, (clouds, N, co)).astype(np.float32)
        chk.forward(pts, x, w, (s, s, s), VOXEL)
        chk.backward(g, pts, x, w, (s, s, s), VOXEL)
    return time.perf_counter() - t0


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
    clouds = threads                                  # one cloud per host thread per step
    if args.workload in NETS:
        _, N, cin, ncls, dist = NETS[args.workload]
        for _ in range(args.warmup):
            cpu_net_step(chk, args.workload, clouds)
        dt = sum(cpu_net_step(chk, args.workload, clouds) for _ in range(args.steps)) / max(1, args.steps)
        sample = f"{clouds} clouds x {N} points per step, the network's Conv3p layers fwd+bwd (no dense layers)"
    else:
        _, N, Cin, Cout, stride, dist = WORKLOADS[args.workload]
        pr, stride = cpu_sample(args.workload, clouds)
        for _ in range(args.warmup):
            cpu_step(chk, pr, stride)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step(chk, pr, stride)
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        sample = f"{clouds} clouds x {N} points per step, {Cin}->{Cout}, fwd+bwd"
    value = clouds * N / dt
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
    if workload in NETS:
        per_gpu, N, cin, ncls, dist = NETS[workload]
        d = {"workload": workload, "clouds_per_gpu": per_gpu, "global_batch": per_gpu * gpus, "num_points": N,
             "input_channels": cin, "classes": ncls, "voxel_size": VOXEL, "distribution": dist,
             "layers": "Conv3p 3x3x3: cin->9 s1, 9->9 s2, 9->9 s3, 9->9 s4" +
                       (", 36->classes s1" if workload == "seg_net" else ", FC 512, FC classes"),
             "parallelism": f"dp{gpus}"}
    else:
        per_gpu, N, Cin, Cout, stride, dist = WORKLOADS[workload]
        d = {"workload": workload, "clouds_per_gpu": per_gpu, "global_batch": per_gpu * gpus, "num_points": N,
             "cin": Cin, "cout": Cout, "filter": "3x3x3", "stride": list(stride), "voxel_size": VOXEL,
             "distribution": dist, "parallelism": f"batch-sharded dp{gpus}, one all-reduce(grad_filter)",
             "l2_policy": "working set per step (inputs+outputs+plan, ~0.7 GB at the headline) exceeds the 126 MB L2"}
    d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------------
# byte / flop models
# ------------------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel: str, pts: int, Cin: int, Cout: int, kbar: float, kbar_b: float,
                      nbins_b: float = 0.0, shared: bool = False):
    """Algorithmic bytes per launch, gather model G of SURVEY section 8d: every list entry is one index read plus
    one row read.  With the shared gather (`shared`: both gradients on the tensor-core kernels) the grad_input
    kernel also writes one Cout-wide row per non-empty (point, cell) slot (nbins_b of them per point) and the
    grad_filter kernel reads those rows back instead of walking the lists."""
    nW = 27 * Cin * Cout * 4
    fwd = kbar * (4 * Cin + 4) + 27 * 4 + 8 + 16 + 4 * Cout
    bwd = kbar_b * (4 * Cout + 8) + 27 * 4 + 8 + 16 + 4 * Cin
    per_point = {
        "k_gather_contract_fwd": fwd, "k_forward_tc": fwd, "k_small_forward": fwd,
        "k_gather_contract_bwd_input": bwd, "k_backward_input_tc": bwd, "k_small_backward_input": bwd,
        "k_backward_filter": bwd, "k_backward_filter_tc": bwd, "k_small_backward_filter": bwd,
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


def compulsory_bytes(kernel: str, pts: int, Cin: int, Cout: int):
    """Model A of SURVEY 8d: every input row read once, every output row written once, weights once per launch."""
    nW = 27 * Cin * Cout * 4
    if "forward" in kernel or kernel.endswith("_fwd"):
        return 4 * (3 + Cin + Cout) * pts + nW
    if "backward_input" in kernel or "bwd_input" in kernel:
        return 4 * (3 + Cin + Cout) * pts + nW
    if "backward_filter" in kernel:
        return 4 * (3 + Cin + Cout) * pts + nW
    return 0.0


def is_contraction(kernel: str) -> bool:
    return any(s in kernel for s in ("forward", "backward_input", "backward_filter", "gather_contract"))


def roofline_block(kern, pts, Cin, Cout, kbar, kbar_b, nbins, nbins_b, shared, pk, traffic):
    """`roofline` of the dominant kernel: three fractions, each against a measured peak.
       frac_G  gather model G bytes / time / HBM peak -- what north_star calls the per-point HBM-read roofline.  The
               gathers are served by L2 (a batch of feature rows is 67-134 MB), so this is a model of the list walk,
               not DRAM traffic; `frac` repeats it because north_star asks for it.
       frac_A  MEASURED DRAM bytes of the kernel (ncu dram__bytes_read+write, profiles/kernel_traffic.json) / time /
               HBM peak: how busy HBM really is.
       frac_F  useful dense flops (2 * 27 * Cin * Cout per point per contraction) / time / the MEASURED TF32 tensor
               peak (profiles/tf32_peak.json); the kernels issue 3 TF32 products per useful one (3xTF32).
       `bound` names the unit ncu shows busiest for this kernel (profiles/kernel_traffic.json 'binder')."""
    if not kern:
        return None
    top = max(kern, key=lambda k: kern[k][1])
    n, total = kern[top]
    avg_s = total / n * 1e-3
    ab = algorithmic_bytes(top, pts, Cin, Cout, kbar, kbar_b, nbins_b, shared)
    flops = 2.0 * 27 * Cin * Cout * pts if is_contraction(top) else 0.0
    tr = traffic.get(top, {})
    dram = tr.get("dram_bytes_per_launch")
    tensor = top.endswith("_tc")
    roof = {
        "kernel": top,
        "bound": tr.get("binder", "l1/shared-memory data pipe + producer latency (ncu, profiles/r2_summary.md)"
                        if tensor else "instruction issue (ncu, profiles/r2_summary.md)"),
        "achieved": ab / avg_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
        "frac": ab / avg_s / 1e9 / pk["hbm_gbs"],
        "frac_label": "frac == frac_G: L2-served gather model (list entries x (index + row bytes) + per-point I/O"
                      + (" + G-store rows" if shared else "") + "), against the measured HBM copy peak",
        "frac_G": ab / avg_s / 1e9 / pk["hbm_gbs"],
        "frac_A": (dram / avg_s / 1e9 / pk["hbm_gbs"]) if dram else None,
        "frac_F": (flops / avg_s / 1e12 / pk["tf32_tflops"]) if flops and tensor else None,
        "traffic": dram,
        "compulsory_bytes_per_launch": compulsory_bytes(top, pts, Cin, Cout),
        "algorithmic_bytes_per_launch": ab,
        "dense_tflops": flops / avg_s / 1e12 if flops else None,
        "peak_source": pk["source"], "tf32_peak_tflops": pk["tf32_tflops"], "tf32_peak_source": pk["tf32_source"],
    }
    return roof


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class Runner:
    """One workload resident on this rank's GPU: timed step, per-kernel timing, neighbour statistics."""

    def __init__(self, workload, rank, world, device, seed=None):
        import torch
        from pointwise_b200 import NeighborPlan, _lib
        from pointwise_b200.distributed import shard_range
        from pointwise_b200.synth import make_problem
        self.torch, self.L = torch, _lib.lib()
        self.workload, self.rank, self.world, self.device = workload, rank, world, device
        per_gpu, N, Cin, Cout, stride, dname = WORKLOADS[workload]
        self.N, self.Cin, self.Cout, self.stride = N, Cin, Cout, stride
        self.B_global = per_gpu * world
        lo, hi = shard_range(self.B_global, rank, world)
        self.B = hi - lo
        # every rank generates only its own shard (seeded by rank) -- clouds are independent
        self.pr = make_problem(self.B, N, Cin, Cout, dname, seed=rank if seed is None else seed)
        self.pr["filter"] = make_problem(1, 8, Cin, Cout, dname, seed=0)["filter"]  # replicated weights
        self.host = {k: torch.from_numpy(v).pin_memory() for k, v in self.pr.items()}
        self.devt = {k: v.to(device) for k, v in self.host.items()}
        self.pts = self.B * N
        # capacity: learned once (checked build), then every step runs without a host read-back
        probe = NeighborPlan(self.devt["points"], stride, VOXEL, check="sync").ensure_backward()
        st = probe.read_stats()
        self.capacity = int(st.total_pairs * 1.05) + 1024
        self.kbar = st.total_pairs / self.pts
        self.kbar_b = st.backward_pairs / self.pts
        self.nbins = float((probe.count_table > 0).sum().item()) / self.pts
        self.nbins_b = float((probe.backward_count_table > 0).sum().item()) / self.pts
        # both gradients of this shape run on the tensor-core kernels with the shared gather (G store)?
        self.shared = bool(self.L.conv3p_backward_scratch_bytes(probe.geom, Cin, Cout) >
                           self.L.conv3p_scratch_bytes(probe.geom, Cin, Cout))
        del probe
        self.pending_reduce = None

    def step(self, collective=True):
        from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
        from pointwise_b200.distributed import allreduce_grad_filter_overlapped
        d = self.devt
        plan = NeighborPlan(d["points"], self.stride, VOXEL, check=False, capacity=self.capacity)
        y = conv3p_forward(plan, d["input"], d["filter"])
        plan.prefetch_backward()         # what the autograd op does when a gradient is required
        gi, gf = conv3p_backward(plan, d["grad_out"], d["input"], d["filter"])
        if collective and self.world > 1:
            # the single collective of the path, on a side stream: it overlaps the next step's sort + search
            self.join()
            self.pending_reduce = allreduce_grad_filter_overlapped(gf)
        return plan, y, gi, gf

    def join(self):
        """Makes the current stream wait for the all-reduce still in flight (the consumer of grad_filter would)."""
        if self.pending_reduce is not None:
            self.torch.cuda.current_stream().wait_event(self.pending_reduce)
            self.pending_reduce = None


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world, device, finish=None):
    """ms per step: barrier + synchronize on both sides, CUDA events on the current stream, max over ranks."""
    import gc
    import torch
    import torch.distributed as dist
    # no cyclic garbage collection inside the timed region: a generation-2 pass over the interpreter's heap (torch,
    # numpy, the oracle bindings) is a 10-30 ms host pause, which showed up as single slow steps early in the region
    gc.collect()
    gc.disable()
    try:
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]     # one record per step: median and best
        e0.record()
        for i in range(steps):
            fn()
            marks[i].record()
        if finish:
            finish()
        e1.record()
        barrier(world)
    finally:
        gc.enable()
    ms = e0.elapsed_time(e1)
    per = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    timed.last = {"median_ms": float(np.median(per)), "best_ms": float(min(per)), "worst_ms": float(max(per)),
                  "worst_step_index": int(np.argmax(per))}
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def profile_kernels(L, fn, steps, world, device, finish=None):
    """-> (ms per step, launches per step, {kernel: (launches, total ms)}) with per-kernel CUDA events."""
    import ctypes as C
    from pointwise_b200 import launch_count
    launch_count(reset=True)
    L.conv3p_profile_enable(1)
    ms = timed(fn, steps, world, device, finish)
    launches = launch_count() // steps
    cbuf = C.create_string_buffer(16384)
    L.conv3p_profile_read(cbuf, 16384)
    L.conv3p_profile_enable(0)
    kern = {}
    for ln in cbuf.value.decode().splitlines():
        name, n, total = ln.split()
        kern[name] = (int(n), float(total))
    return ms, launches, kern


def parity_check(run: Runner, y, gi, gf, clouds=3):
    """Three clouds of the batch that was just timed, against the CPU checker (outside every timed region):
    output and grad_input rows of the full-batch call, grad_filter of a call on exactly those clouds, plus the
    additivity of the full-batch grad_filter over ALL its clouds (per-cloud B=1 calls, three of which are
    oracle-checked) -- tf_conv3p_atrous.cpp:456-504, 622-716.  Bound: |got - sum64| <= 1e-7 + 1e-5 * sum|terms|."""
    import torch
    import oracle
    from pointwise_b200 import NeighborPlan, conv3p_backward
    oracle.build()
    port = oracle.port()
    try:      # torchrun exports OMP_NUM_THREADS=1 to every rank; the checker may use this rank's share of the host
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(max(1, min(32, (os.cpu_count() or 1) // max(1, run.world))))
    except OSError:
        pass
    t0 = time.perf_counter()
    idx = sorted(set(np.linspace(0, run.B - 1, clouds).astype(int).tolist()))
    pr = {k: (v[idx] if k != "filter" else v) for k, v in run.pr.items()}
    o32, o64, oabs = port.forward(pr["points"], pr["input"], pr["filter"], run.stride, VOXEL, with64=True)
    r = port.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], run.stride, VOXEL, with64=True)
    worst = 0.0

    def ratio(got, want64, abs64):
        err = np.abs(got.astype(np.float64) - want64)
        return float((err / (ATOL + RTOL * abs64)).max())

    res = {"clouds": len(idx), "cloud_indices": idx}
    res["output_err_over_bound"] = ratio(y[idx].cpu().numpy(), o64, oabs)
    res["grad_input_err_over_bound"] = ratio(gi[idx].cpu().numpy(), r[2], r[3])
    sub = {k: torch.from_numpy(np.ascontiguousarray(v)).to(run.device) for k, v in pr.items()}
    plan = NeighborPlan(sub["points"], run.stride, VOXEL, check="sync")
    _, gf_sub = conv3p_backward(plan, sub["grad_out"], sub["input"], sub["filter"])
    res["grad_filter_err_over_bound"] = ratio(gf_sub.cpu().numpy(), r[4], r[5])
    worst = max(res["output_err_over_bound"], res["grad_input_err_over_bound"], res["grad_filter_err_over_bound"])
    # the reference's own fp32 result, when its object code is present: relative deviation (informational)
    if oracle.Ref.available():
        ref_out = oracle.ref().forward(pr["points"], pr["input"], pr["filter"], run.stride, VOXEL)
        res["output_max_rel_vs_reference_fp32"] = float(np.abs(y[idx].cpu().numpy() - ref_out).max() /
                                                        max(1e-30, np.abs(ref_out).max()))
        res["checker"] = "oracle port (float64 sums, pinned to the reference) + reference object code (fp32)"
    else:
        res["checker"] = "oracle port (float64 sums, pinned to the reference's golden vectors)"
    # additivity: grad_filter of the whole shard == sum over its clouds of single-cloud calls.  sum|terms| of the
    # whole shard is estimated from the three oracle clouds (B / 3 times theirs).
    acc = torch.zeros_like(gf, dtype=torch.float64)
    d = run.devt
    for b in range(run.B):
        p1 = NeighborPlan(d["points"][b:b + 1].contiguous(), run.stride, VOXEL, check=False,
                          capacity=run.capacity // max(1, run.B) * 4 + 4096)
        _, g1 = conv3p_backward(p1, d["grad_out"][b:b + 1].contiguous(), d["input"][b:b + 1].contiguous(), d["filter"],
                                need_input_grad=False)
        acc += g1.double()
    mag = torch.from_numpy(r[5]).to(run.device) * (run.B / len(idx))
    add_err = float(((gf.double() - acc).abs() / (ATOL + RTOL * mag)).max().item())
    res["grad_filter_additivity_err_over_bound"] = add_err
    res["max_err_over_bound"] = max(worst, add_err)
    res["bound"] = "|got - sum64| <= 1e-7 + 1e-5 * sum|terms|"
    res["ok"] = bool(np.isfinite(res["max_err_over_bound"]) and res["max_err_over_bound"] <= 1.0)
    res["cpu_seconds"] = round(time.perf_counter() - t0, 1)
    return res


def scale_check(run: Runner, gf_local_pre, gf_reduced):
    """--gpus N: the all-reduced grad_filter is bit-identical on every rank and equals the sum of the per-rank
    gradients (the multi-GPU analogue of tf_conv3p_atrous.cpp:709-716).  Rank 0's own shard is tied to the oracle by
    parity_check."""
    import torch
    import torch.distributed as dist
    world = run.world
    red = [torch.empty_like(gf_reduced) for _ in range(world)]
    pre = [torch.empty_like(gf_local_pre) for _ in range(world)]
    dist.all_gather(red, gf_reduced.contiguous())
    dist.all_gather(pre, gf_local_pre.contiguous())
    same = all(torch.equal(red[0], r) for r in red)
    tot = torch.zeros_like(gf_reduced, dtype=torch.float64)
    mag = torch.zeros_like(gf_reduced, dtype=torch.float64)
    for p in pre:
        tot += p.double()
        mag += p.double().abs()
    err = float(((gf_reduced.double() - tot).abs() / (ATOL + RTOL * mag)).max().item())
    return {"ranks": world, "bit_identical_across_ranks": bool(same), "sum_err_over_bound": err,
            "finite": bool(torch.isfinite(gf_reduced).all().item()), "ok": bool(same and err <= 1.0)}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pointwise_b200 import _lib
    from pointwise_b200.distributed import allreduce_grad_filter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"bound": False, "why": "single rank"}
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.workload in NETS:
        return net_arm(args, rank, world, device)
    L = _lib.lib()
    # nvidia-smi attaching to the GPU stalls launches for milliseconds during its first second or so (a 16-cloud run
    # showed single 3-5 ms steps when the sampler was started right before the warm-up; in steady state polling is
    # harmless at any period, tools/sampler_probe.py): start it before the inputs are generated and uploaded
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    run = Runner(args.workload, rank, world, device)
    N, Cin, Cout, pts = run.N, run.Cin, run.Cout, run.pts
    warm = max(3, args.warmup)
    for _ in range(warm):
        run.step()
    run.join()
    # Untimed rehearsal on top of the W warm-up steps.  The host enqueues many steps ahead of the GPU, and buffers that
    # were handed to a side stream (the plan buffer, for the backward-list prefetch) can only be reused once that
    # stream's event has completed -- so the caching allocator needs a deeper pool the further the host runs ahead, and
    # it grows it with cudaMalloc (tens of ms for half a GB) at whatever step that happens: single 15-45 ms steps inside
    # the timed region.  Rehearse the timed loop itself (same number of steps, no synchronisation inside) until a whole
    # rehearsal needed no new device allocation (at most four).  Without the collective: the count differs per rank.
    def device_allocs():
        return torch.cuda.memory_stats(device).get("num_device_alloc", 0)

    if world > 1:                      # one rehearsal WITH the collective, the same number of steps on every rank
        for _ in range(args.steps):
            run.step()
        run.join()
        torch.cuda.synchronize(device)
    for _ in range(4):
        before = device_allocs()
        for _ in range(args.steps):
            run.step(collective=False)
        torch.cuda.synchronize(device)
        if device_allocs() == before:
            break
    run.join()
    if rank == 0:
        sampler.ready()
    # the reported step time is taken WITHOUT the per-kernel event pairs (two cudaEventRecord per launch open small gaps
    # between dependent kernels); the per-kernel breakdown comes from a second, instrumented pass over the same steps
    allocs0 = device_allocs()
    ms_step = timed(run.step, args.steps, world, device, finish=run.join)
    step_spread = dict(timed.last)
    step_spread["device_mallocs_in_timed_region"] = device_allocs() - allocs0
    clocks = sampler.stop() if rank == 0 else None
    ms_profiled, launches, kern = profile_kernels(L, run.step, args.steps, world, device, finish=run.join)
    value = run.B_global * N / (ms_step * 1e-3)

    # ---- forward-only and backward-only, this rank's shard, no collective ---------------------------
    from pointwise_b200 import NeighborPlan, conv3p_backward, conv3p_forward
    d_ = run.devt

    def fwd_only():
        return conv3p_forward(NeighborPlan(d_["points"], run.stride, VOXEL, check=False, capacity=run.capacity),
                              d_["input"], d_["filter"])

    held = NeighborPlan(d_["points"], run.stride, VOXEL, check=False, capacity=run.capacity)

    def bwd_only():          # plan reused from forward; backward lists rebuilt every time (they are per step)
        held.has_backward = False
        return conv3p_backward(held, d_["grad_out"], d_["input"], d_["filter"])

    phases = {}
    for name, fn, what in (("forward", fwd_only, "plan build (sort + search) + forward kernel"),
                           ("backward", bwd_only, "backward lists + both gradient kernels + reduction, plan reused")):
        for _ in range(3):
            fn()
        ms_p = timed(fn, args.steps, 1, device)
        phases[name] = {"ms": ms_p, "points_per_s_per_gpu": run.pts / (ms_p * 1e-3), "includes": what}
    del held

    # ---- checks on the very batch that was timed (outside the timed regions) ----------------------------------
    plan, y, gi, gf = run.step(collective=False)
    torch.cuda.synchronize()
    checks = {}
    if world > 1:
        gf_pre = gf.clone()
        allreduce_grad_filter(gf)
        torch.cuda.synchronize()
        checks["scale_check"] = scale_check(run, gf_pre, gf)
        gf = gf_pre
    if rank == 0 and not args.no_parity:
        checks["parity_check"] = parity_check(run, y, gi, gf)
    del plan, y, gi, gf
    barrier(world)

    # ---- e2e (host buffers): every step copies its pinned-host inputs to the device, rebuilds the plan, runs
    # forward + backward (+ the all-reduce) and copies output / grad_input / grad_filter back to pinned host
    # memory; copies of neighbouring steps overlap the kernels (three streams, three staging slots).
    from pointwise_b200.host_api import HostConv3p
    pipe = HostConv3p(run.B, N, Cin, Cout, run.stride, VOXEL, device=device, capacity=run.capacity,
                      depth=int(os.environ.get("CONV3P_HOST_DEPTH", "3")))
    ar = allreduce_grad_filter if world > 1 else None
    host = run.host

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
    barrier(world)
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
    del pipe

    # ---- BASELINE configs[4]: 16 clouds per GPU, with the collective and alone on this GPU ----------------------
    b16 = None
    if args.workload == "headline" and not args.no_b16:
        r16 = Runner("headline_b16", rank, world, device)
        for _ in range(warm):
            r16.step()
        r16.join()
        ms16 = timed(r16.step, args.steps, world, device, finish=r16.join)
        ms16_alone = timed(lambda: r16.step(collective=False), args.steps, 1, device)   # no barrier, no collective
        if world > 1:
            t = torch.tensor([ms16_alone], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms16_alone = float(t.item())
        b16 = {"workload": "headline_b16 (BASELINE configs[4]: 16 clouds of 4096 points per GPU, 64->128)",
               "value": r16.B_global * N / (ms16 * 1e-3), "unit": UNIT, "ms_per_step": ms16,
               "ms_per_step_single_gpu_no_collective": ms16_alone,
               "efficiency_vs_own_n1": ms16_alone / ms16}
        del r16

    refgpu = None
    if world == 1 and args.workload == "headline" and not args.no_sweep:
        try:
            refgpu = reference_gpu_code(run)
        except Exception as e:  # the baseline must never take the benchmark down
            refgpu = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}

    # ---- BASELINE configs[3]: the 9-point sweep (one GPU) -----------------------------------------------------
    sweep = None
    if world == 1 and args.workload == "headline" and not args.no_sweep:
        sweep = run_sweep(L, device, warm=2, steps=5)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {})
    kernels = kernel_table(kern, args.steps, ms_profiled, pts, Cin, Cout, run, pk)
    roof = roofline_block(kern, pts, Cin, Cout, run.kbar, run.kbar_b, run.nbins, run.nbins_b, run.shared, pk, traffic)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, world, mean_neighbours=round(run.kbar, 2),
                              mean_backward_pairs=round(run.kbar_b, 2), mean_nonempty_cells=round(run.nbins, 2),
                              mean_nonempty_backward_cells=round(run.nbins_b, 2), shared_backward_gather=run.shared,
                              numa=numa),
        "clocks": clocks,
        "e2e": {"value": run.B_global * N / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "wall_ms_per_step": wall_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "pointwise_b200.host_api.HostConv3p (pinned host in/out, copies overlapped with compute)"},
        "gpu_launches": int(launches),
        "ms_per_step_with_kernel_timers": ms_profiled,
        "step_spread_rank0": step_spread,
        "phases": phases,
        "roofline": roof,
        "kernels": kernels,
    }
    line.update(checks)
    if b16:
        line["b16"] = b16
    if sweep:
        line["sweep"] = sweep
    if refgpu:
        line["reference_gpu_code"] = refgpu
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = run_cpu_baseline(args.workload)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_table(kern, steps, ms_step, pts, Cin, Cout, run, pk):
    kernels = {}
    for name, (n, total) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
        avg_ms = total / n
        ab = algorithmic_bytes(name, pts, Cin, Cout, run.kbar, run.kbar_b, run.nbins_b, run.shared)
        row = {"launches_per_step": n // steps, "avg_ms": round(avg_ms, 4),
               "share_of_step": round(total / steps / ms_step, 4),
               "algorithmic_gbs": round(ab / (avg_ms * 1e-3) / 1e9, 1),
               "frac_G": round(ab / (avg_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 3)}
        if is_contraction(name):
            row["dense_tflops"] = round(2.0 * 27 * Cin * Cout * pts / (avg_ms * 1e-3) / 1e12, 1)
        kernels[name] = row
    return kernels


def run_sweep(L, device, warm=2, steps=5):
    """BASELINE configs[3]: fwd+bwd throughput of the nine (N, C) points, each with its neighbour statistics, the
    engine that ran, and the three roofline fractions of its forward kernel."""
    import torch
    pk = peaks()
    out = {}
    for name in SWEEP:
        r = Runner(name, 0, 1, device, seed=0)
        for _ in range(warm):
            r.step()
        ms, launches, kern = profile_kernels(L, r.step, steps, 1, device)
        fwd = next((k for k in ("k_forward_tc", "k_small_forward", "k_gather_contract_fwd") if k in kern), None)
        row = {"points_per_s": r.B * r.N / (ms * 1e-3), "ms_per_step": round(ms, 4), "clouds": r.B, "N": r.N,
               "C": r.Cin, "kbar": round(r.kbar, 2), "nbins": round(r.nbins, 2),
               "engine": ("tcgen05 TF32 + BF16 corrections" + (", channels zero-padded" if "k_pad_channels" in kern else ""))
                         if "k_forward_tc" in kern else
                         ("fp32 warp-per-point" if "k_small_forward" in kern else "fp32 tile"),
               "kernels_ms": {k: round(v[1] / v[0], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:6]}}
        if fwd:
            t = kern[fwd][1] / kern[fwd][0] * 1e-3
            g = algorithmic_bytes(fwd, r.pts, r.Cin, r.Cout, r.kbar, r.kbar_b)
            fl = 2.0 * 27 * r.Cin * r.Cout * r.pts
            row["forward"] = {"kernel": fwd, "ms": round(t * 1e3, 4),
                              "frac_G": round(g / t / 1e9 / pk["hbm_gbs"], 3),
                              "frac_A_model": round(compulsory_bytes(fwd, r.pts, r.Cin, r.Cout) / t / 1e9 / pk["hbm_gbs"], 4),
                              "dense_tflops": round(fl / t / 1e12, 2),
                              "frac_F": round(fl / t / 1e12 / (pk["tf32_tflops"] if fwd.endswith("_tc") else 74.4), 3),
                              "frac_F_peak": "measured TF32 tensor peak" if fwd.endswith("_tc") else "fp32 SIMT 74.4 TFLOP/s"}
        out[name] = row
        del r
        torch.cuda.empty_cache()
    return out


def reference_gpu_code(run: Runner, clouds=4, steps=2):
    """The reference's own CUDA kernels (tf_conv3p_atrous.cu, compiled unmodified for sm_100a with its own flags:
    oracle/_ref/libconv3p_ref_gpu.so) on a bounded sample of the same workload, same GPU -- BASELINE.md section 2's
    optional second baseline.  A brute-force O(N^2) sweep with one thread per point and global read-modify-write per
    MAC; -use_fast_math makes it a timing baseline only (nothing is compared with it except loosely, below)."""
    import torch
    import oracle
    if not oracle.RefGpu.available():
        return None
    R = oracle.RefGpu()
    d = {k: (v[:clouds].contiguous() if k != "filter" else v) for k, v in run.devt.items()}
    stride = torch.tensor(list(run.stride), dtype=torch.int32, device=run.device)
    voxel = torch.tensor([VOXEL], dtype=torch.float32, device=run.device)
    out = torch.empty((clouds, run.N, run.Cout), device=run.device)
    gi, gf = torch.empty_like(d["input"]), torch.empty_like(d["filter"])

    def step():
        R.forward(d["points"], d["input"], d["filter"], stride, voxel, out)
        R.backward(d["grad_out"], d["points"], d["input"], d["filter"], stride, voxel, gi, gf)

    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()                       # (the op synchronises the device itself, tf_conv3p_atrous.cu:577)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    from pointwise_b200 import NeighborPlan, conv3p_forward
    mine = conv3p_forward(NeighborPlan(d["points"], run.stride, VOXEL, check="sync"), d["input"], d["filter"])
    dev = float((mine - out).abs().max() / out.abs().max())
    return {"value": clouds * run.N / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
            "sample": f"{clouds} clouds x {run.N} points per step, {run.Cin}->{run.Cout}, fwd+bwd, same GPU",
            "kind": "reference CUDA kernels compiled unmodified for sm_100a (fast-math: timing baseline, not a parity "
                    "reference)", "forward_max_rel_deviation_from_ours": dev}


def net_arm(args, rank, world, device):
    """--workload seg_net | cls_net: one full PointConvNet training step (forward, loss, backward) through the
    public conv3p() -- BASELINE configs[2] / configs[1] -- with one neighbour plan per stride shared by the layers
    (PlanCache) and, for comparison, with a plan per layer (what a literal drop-in call does)."""
    import torch
    import torch.distributed as dist
    from pointwise_b200 import launch_count, nets
    from pointwise_b200.synth import make_points
    per_gpu, N, cin, ncls, dname = NETS[args.workload]
    torch.manual_seed(0)
    pts = torch.from_numpy(make_points(per_gpu, N, dname, seed=rank)).to(device)
    feats = (pts.clone() if cin == 3 else
             torch.from_numpy(np.random.default_rng(rank).uniform(-1, 1, (per_gpu, N, cin)).astype(np.float32)).to(device))
    if args.workload == "seg_net":
        net = nets.PointConvNetSeg(ncls, cin).to(device)
        labels = torch.randint(0, ncls, (per_gpu, N), device=device)
    else:
        net = nets.PointConvNetCls(ncls, N, cin).to(device)
        labels = torch.randint(0, ncls, (per_gpu,), device=device)
    params = list(net.parameters())

    def step():
        for p in params:
            p.grad = None
        loss = net.loss(net.model(pts, feats, True), labels)
        loss.backward()
        if world > 1:
            for p in params:
                dist.all_reduce(p.grad)
        return loss

    res = {}
    for sharing in (True, False):
        nets.SHARE_PLANS = sharing
        for _ in range(max(3, args.warmup)):
            step()
        launch_count(reset=True)
        ms = timed(step, args.steps, world, device)
        res["shared_plans" if sharing else "plan_per_layer"] = {
            "ms_per_step": ms, "value": per_gpu * world * N / (ms * 1e-3), "gpu_launches": launch_count() // args.steps,
            "step_spread_rank0": dict(timed.last)}
    nets.SHARE_PLANS = True
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    best = res["shared_plans"]
    line = {"metric": "pointconvnet_train_step_points_per_sec", "value": best["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": best["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, world), "gpu_launches": best["gpu_launches"], "plans": res,
            "host_syncs_per_step": "none in steady state (deferred overflow check, NeighborPlan docstring)"}
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
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS) + sorted(NETS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[3] sweep")
    ap.add_argument("--no-b16", action="store_true", help="skip the 16-clouds-per-GPU leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU parity check of the timed batch")
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it can
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) and any other NCCL log line go
    # to stderr.  Set before torch / NCCL are loaded.
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    gpu_arm(args)


if __name__ == "__main__":
    main()
