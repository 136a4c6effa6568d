"""Host<->device copy bandwidth with 1..N ranks copying AT THE SAME TIME (one process per GPU, torchrun + gloo):
every rank moves the headline step's 205 MB up and 202 MB down on two streams, all ranks start together.  Shows
whether the host-buffer path (bench.py `e2e`) is limited by each GPU's own PCIe link or by what the ranks share
(host memory, root complex).  usage: torchrun --nproc-per-node N tools/pcie_probe_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
up, down = 205357056 // 4, 202211328 // 4
h_up = torch.empty(up, dtype=torch.float32).pin_memory()
d_up = torch.empty(up, dtype=torch.float32, device="cuda")
h_dn = torch.empty(down, dtype=torch.float32).pin_memory()
d_dn = torch.randn(down, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d_up.copy_(h_up, non_blocking=True)
    with torch.cuda.stream(s2):
        h_dn.copy_(d_dn, non_blocking=True)


res = {}
for active in sorted({1, 2, 4, 8, world} & set(range(1, world + 1))):
    for _ in range(3):
        both()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = 0.0
    if rank < active:
        t0 = time.perf_counter()
        for _ in range(10):
            both()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 10 * 1e3
    if world > 1:
        t = torch.tensor([ms])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    res[active] = ms
if rank == 0:
    base = res[min(res)]
    rows = [{"ranks_copying": k, "ms_per_step_slowest_rank": round(v, 3),
             "aggregate_GBps": round(k * (up + down) * 4 / v / 1e6, 1), "vs_one_rank": round(base / v, 2)}
            for k, v in sorted(res.items())]
    for r in rows:
        print(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"bytes_up": up * 4, "bytes_down": down * 4, "rows": rows}, open(f"gpurun_out/pcie_probe_{world}ranks.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
