"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: batch sharding and the single
all-reduce on grad_filter.  The per-shard compute is done by the CPU oracle here (tests may use it);
on GPUs the same functions wrap the CUDA operator (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointwise_b200.distributed import allreduce_grad_filter, shard_batch, shard_range
from pointwise_b200.synth import make_problem


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, N, Cin, Cout, stride, V = 5, 300, 4, 6, (2, 2, 2), 0.1
        pr = make_problem(B, N, Cin, Cout, "room", seed=3)
        chk = oracle.port()
        mine = {k: shard_batch(torch.from_numpy(pr[k]), rank, world).numpy()
                for k in ("points", "input", "grad_out")}
        out = chk.forward(mine["points"], mine["input"], pr["filter"], stride, V)
        gi, gf = chk.backward(mine["grad_out"], mine["points"], mine["input"], pr["filter"], stride, V)
        gf_t = torch.from_numpy(gf.copy())
        allreduce_grad_filter(gf_t)
        full_out = chk.forward(pr["points"], pr["input"], pr["filter"], stride, V)
        full_gi, full_gf = chk.backward(pr["grad_out"], pr["points"], pr["input"], pr["filter"], stride, V)
        lo, hi = shard_range(B, rank, world)
        ok = (np.array_equal(out, full_out[lo:hi]) and np.array_equal(gi, full_gi[lo:hi])
              and np.allclose(gf_t.numpy(), full_gf, rtol=1e-5, atol=1e-5))
        # every rank holds the same reduced gradient
        gathered = [torch.empty_like(gf_t) for _ in range(world)]
        dist.all_gather(gathered, gf_t)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        q.put((rank, bool(ok), bool(same)))
    finally:
        dist.destroy_process_group()


def test_sharded_backward_equals_full_batch_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results), "shard results differ from the full-batch oracle"
    assert all(r[2] for r in results), "ranks disagree on the reduced grad_filter"


# ---- the same gate on real GPUs over NCCL ----------------------------------------------------
def _nccl_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pointwise_b200.distributed import conv3p_grad_sharded
        from pointwise_b200 import conv3p_grad
        B, N, Cin, Cout, stride, V = 6, 1024, 64, 128, (1, 1, 1), 0.1
        pr = make_problem(B, N, Cin, Cout, "room", seed=5)
        full = {k: torch.from_numpy(v).to(dev) for k, v in pr.items()}
        mine = {k: shard_batch(full[k], rank, world).contiguous() for k in ("points", "input", "grad_out")}
        gi, gf = conv3p_grad_sharded(mine["grad_out"], mine["points"], mine["input"], full["filter"], stride, V)
        gi_full, gf_full = conv3p_grad(full["grad_out"], full["points"], full["input"], full["filter"], stride, V)
        lo, hi = shard_range(B, rank, world)
        gathered = [torch.empty_like(gf) for _ in range(world)]
        dist.all_gather(gathered, gf)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        scale = float(gf_full.abs().max())
        ok = torch.equal(gi, gi_full[lo:hi]) and float((gf - gf_full).abs().max()) <= 2e-5 * scale
        q.put((rank, bool(ok), bool(same)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_backward_equals_full_batch_nccl():
    """Two ranks over NCCL: identical grad_filter on both, equal to the single-GPU result on the concatenated
    batch within fp32 summation tolerance; grad_input of a shard bit-identical to the full-batch rows."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in results), "shard results differ from the full-batch result"
    assert all(r[2] for r in results), "ranks disagree on the reduced grad_filter"
