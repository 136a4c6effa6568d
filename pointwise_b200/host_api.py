"""Host-buffer front end: the reference's feed/fetch calling shape on top of the CUDA operator.

The reference is driven with numpy batches through ``session.run`` (train_modelnet40_acsd.py:123-156): every
step feeds host arrays and fetches host results.  ``HostConv3p`` offers that shape for Conv3p forward +
Conv3pGrad: ``submit()`` takes pinned host tensors, ``fetch()`` returns pinned host results.  Copies and compute
run on three CUDA streams with `depth` (default 3) device staging slots, so the host->device copy of step k+1 and
the device->host copy of step k-1 overlap the kernels of step k (with only two slots the upload of step k+1 would
have to wait for the download of step k-1, which shares its slot); nothing is cached between steps -- every step
copies its own inputs, rebuilds its neighbour plan and copies its own outputs.  Keep at most `depth - 1` steps
un-fetched: submit() reuses the slot of step k - depth.
"""
from __future__ import annotations

from typing import Optional

import torch

from .ops import NeighborPlan, conv3p_backward, conv3p_forward, parse_stride, parse_voxel


class HostConv3p:
    def __init__(self, B: int, N: int, Cin: int, Cout: int, stride, voxel_size, device=None,
                 capacity: Optional[int] = None, depth: int = 3):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (B, N, Cin, Cout)
        self.stride, self.voxel = parse_stride(stride), parse_voxel(voxel_size)
        self.capacity = capacity
        self.depth = depth
        dev, f32 = self.device, torch.float32
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                points=torch.empty((B, N, 3), dtype=f32, device=dev),
                input=torch.empty((B, N, Cin), dtype=f32, device=dev),
                filter=torch.empty((3, 3, 3, Cin, Cout), dtype=f32, device=dev),
                grad_out=torch.empty((B, N, Cout), dtype=f32, device=dev),
                h_out=torch.empty((B, N, Cout), dtype=f32).pin_memory(),
                h_gi=torch.empty((B, N, Cin), dtype=f32).pin_memory(),
                h_gf=torch.empty((3, 3, 3, Cin, Cout), dtype=f32).pin_memory(),
                copied=torch.cuda.Event(), computed=torch.cuda.Event(), fetched=torch.cuda.Event(),
                keep=None))
        self.s_h2d = torch.cuda.Stream(dev)
        self.s_compute = torch.cuda.Stream(dev)
        self.s_d2h = torch.cuda.Stream(dev)
        self.step = 0

    @property
    def h2d_bytes(self) -> int:
        s = self.slots[0]
        return sum(s[k].numel() * 4 for k in ("points", "input", "filter", "grad_out"))

    @property
    def d2h_bytes(self) -> int:
        s = self.slots[0]
        return sum(s[k].numel() * 4 for k in ("h_out", "h_gi", "h_gf"))

    def submit(self, points, input, filter, grad_out, allreduce=None) -> int:
        """Enqueue one step on pinned HOST tensors; returns its ticket.  ``allreduce`` (optional) is called
        with the device grad_filter on the compute stream (multi-GPU: the single collective of the path)."""
        k = self.step
        s = self.slots[k % self.depth]
        with torch.cuda.stream(self.s_h2d):
            # the slot's previous results must have left the device before its inputs are overwritten
            self.s_h2d.wait_event(s["fetched"]) if k >= self.depth else None
            s["points"].copy_(points, non_blocking=True)
            s["input"].copy_(input, non_blocking=True)
            s["filter"].copy_(filter, non_blocking=True)
            s["grad_out"].copy_(grad_out, non_blocking=True)
            s["copied"].record(self.s_h2d)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(s["copied"])
            # overflow check deferred to fetch(): nothing synchronises while the step is enqueued
            plan = NeighborPlan(s["points"], self.stride, self.voxel, capacity=self.capacity,
                                check=True if self.capacity is None else "deferred")
            out = conv3p_forward(plan, s["input"], s["filter"])
            plan.prefetch_backward()
            gi, gf = conv3p_backward(plan, s["grad_out"], s["input"], s["filter"])
            if allreduce is not None:
                allreduce(gf)
            s["computed"].record(self.s_compute)
            s["keep"] = (plan, out, gi, gf)       # alive until fetched
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(s["computed"])
            s["h_out"].copy_(out, non_blocking=True)
            s["h_gi"].copy_(gi, non_blocking=True)
            s["h_gf"].copy_(gf, non_blocking=True)
            for t in (out, gi, gf, plan.buffer):
                t.record_stream(self.s_d2h)
            s["fetched"].record(self.s_d2h)
        self.step += 1
        return k

    def fetch(self, ticket: int):
        """Blocks until the step's results are in host memory -> (output, grad_input, grad_filter) pinned."""
        s = self.slots[ticket % self.depth]
        s["fetched"].synchronize()
        keep, s["keep"] = s["keep"], None
        if keep is not None:
            keep[0].verify(block=True)   # raises ERR_PAIR_OVERFLOW if this step's neighbour lists were incomplete
        return s["h_out"], s["h_gi"], s["h_gf"]

    def last_event(self):
        return self.slots[(self.step - 1) % self.depth]["fetched"]
