"""N > 1 host logic on CPU: two gloo ranks shard a clip batch, each runs the ORACLE on its slice (the CUDA
library cannot run here), and the gathered result must equal the unsharded one bit for bit — the warp has no
cross-sample term, so batch sharding needs no collective in the op (DESIGN.md "Multi-GPU")."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from deep_video_interpolation_extrapolation_b200 import sharding


def test_batch_slice_partitions_exactly():
    for n in (0, 1, 2, 7, 16, 64):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                a, b = sharding.batch_slice(n, r, world)
                assert 0 <= a <= b <= n
                cover += list(range(a, b))
            assert cover == list(range(n))
            sizes = [sharding.batch_slice(n, r, world) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1
    with pytest.raises(ValueError):
        sharding.batch_slice(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    oracle.build()
    f0, f1 = synth.rgb(0, N, H, W), synth.rgb(1, N, H, W)
    ff, fb = synth.flow(3, N, H, W, 6.0), synth.flow(4, N, H, W, 6.0)
    mf, mb = synth.mask(2, N, H, W), synth.mask(5, N, H, W)
    go = synth.grad(6, (N, 3, H, W))
    mine = sharding.shard([f0, f1, ff, fb, mf, mb, go], rank, world)
    a, b = sharding.batch_slice(N, rank, world)
    out = oracle.forward([(mine[0], mine[1])], [mine[2], mine[3]], blends=[mine[4], mine[5]], signs=[-1, 1],
                         padding_mode="border")[0][:, 0]
    g = oracle.backward([(mine[0], mine[1])], [mine[2], mine[3]], [mine[6]], blends=[mine[4], mine[5]], signs=[-1, 1],
                        padding_mode="border")["grad_srcs"][0][0][:, 0]
    # gather the shards (test plumbing only: the op itself never communicates)
    parts = [None] * world
    dist.all_gather_object(parts, (a, b, out, g))
    # the timing rule of bench.py: the job's step time is the slowest rank's
    tmax = sharding.max_over_ranks(0.5 + rank)
    if rank == 0:
        full_out = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], 0)
        full_g = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[0])], 0)
        q.put((full_out, full_g, tmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_matches_unsharded(oracle):
    N, H, W, world = 5, 24, 40, 2  # odd batch: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    got_out, got_g, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f0, f1 = synth.rgb(0, N, H, W), synth.rgb(1, N, H, W)
    ff, fb = synth.flow(3, N, H, W, 6.0), synth.flow(4, N, H, W, 6.0)
    mf, mb = synth.mask(2, N, H, W), synth.mask(5, N, H, W)
    go = synth.grad(6, (N, 3, H, W))
    ref = oracle.forward([(f0, f1)], [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode="border")[0][:, 0]
    rg = oracle.backward([(f0, f1)], [ff, fb], [go], blends=[mf, mb], signs=[-1, 1], padding_mode="border")["grad_srcs"][0][0][:, 0]
    assert np.array_equal(got_out, ref)
    assert np.array_equal(got_g, rg)
    assert tmax == 1.5  # max over ranks of (0.5, 1.5)
    assert sharding.job_throughput(10.0, 4, 2.0, world) == 40.0


def test_chunk_slices_cover_the_batch():
    for n in range(0, 12):
        for c in (1, 2, 3, 16):
            sl = sharding.chunk_slices(n, c)
            assert [i for s in sl for i in range(s.start, s.stop)] == list(range(n))
            assert all(0 < s.stop - s.start <= c for s in sl)
    with pytest.raises(ValueError):
        sharding.chunk_slices(4, 0)


def test_numa_binding_is_best_effort_without_a_gpu():
    """bind_to_gpu_numa never raises: without NVML / a GPU it leaves the affinity alone and reports 0 cores."""
    before = os.sched_getaffinity(0)
    n = sharding.bind_to_gpu_numa(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def _sync_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    names = [f"step_{i}_loss" for i in range(12)]
    mine = {k: torch.randn((), generator=g) if i % 3 else torch.randn(3, generator=g) for i, k in enumerate(names)}
    mine["count"] = torch.tensor(float(rank + 1), dtype=torch.float64)  # a second dtype: its own flat buffer
    ref = {k: v.clone() for k, v in mine.items()}
    for t in ref.values():  # the reference's loop: one all_reduce (+ div_) per tensor (runners/InterTrainer.py:859-864)
        dist.all_reduce(t)
        t.div_(world)
    out = sharding.sync(mine, mean=True)
    assert out is mine
    same = all(torch.equal(mine[k], ref[k]) for k in mine)
    summed = sharding.sync({"x": torch.tensor([1.0, 2.0]) * (rank + 1)}, mean=False)["x"]
    ok_sum = torch.equal(summed, torch.tensor([1.0, 2.0]) * sum(range(1, world + 1)))
    if rank == 0:
        q.put((same, ok_sum))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_coalesced_sync_equals_per_tensor_all_reduce():
    """sharding.sync (one flat all-reduce per dtype) == the reference's per-scalar loop, world_size 2 on gloo."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, ok_sum = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert same and ok_sum


def test_sync_without_process_group_is_identity_mean():
    d = {"a": torch.tensor(3.0), "b": torch.tensor([1.0, 2.0])}
    sharding.sync(d, mean=True)
    assert float(d["a"]) == 3.0 and d["b"].tolist() == [1.0, 2.0]
