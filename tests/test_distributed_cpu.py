"""world_size-2 gloo test (CPU) of the multi-GPU plan of the Monte-Carlo loops: iterations are block-partitioned
with `shard_range`, every rank accumulates fp64 (sum, sum of squares) over ITS iterations only, one all-reduce
closes the loop, and mean / unbiased std match `torch.mean/std` over the full stack for any world size."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unet_research_b200 import shard_range

T, H, W, R = 23, 6, 5, 4


def sample(t: int) -> torch.Tensor:
    """Stand-in for iteration t's masked forward output: depends on the GLOBAL index only."""
    g = torch.Generator().manual_seed(1000 + t)
    return torch.rand(H, W, generator=g)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t0, t1 = shard_range(T, rank, world)
    acc = torch.zeros(2, H, W, dtype=torch.float64)
    samples = torch.zeros(R, H, W)
    for t in range(t0, t1):
        v = sample(t)
        acc[0] += v.double()
        acc[1] += v.double() ** 2
        if t < R:
            samples[t] = v
    dist.all_reduce(acc)
    dist.all_reduce(samples)
    mean = acc[0] / T
    std = ((acc[1] - acc[0] ** 2 / T) / (T - 1)).clamp_min(0).sqrt()
    if rank == 0:
        torch.save({"mean": mean, "std": std, "samples": samples}, out)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_statistics_match_full_stack(world, tmp_path):
    out = str(tmp_path / f"res{world}.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out)
    stack = torch.stack([sample(t) for t in range(T)])
    torch.testing.assert_close(res["mean"].float(), stack.mean(0), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(res["std"].float(), stack.std(0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res["samples"], stack[:R], rtol=0, atol=0)


def test_plan_segments_and_shards_cover_every_iteration_once():
    """Host logic of the Monte-Carlo loop: block sharding over ranks + runs of equal batches per rank cover [0, T) exactly
    once, rank 0 owns the first iterations (the saved samples need no exchange), at most one remainder run per rank."""
    from unet_research_b200.uncertainty import plan_segments, shard_range
    assert plan_segments(0, 125, 10) == [(10, 12, 0), (5, 1, 120)]
    assert plan_segments(7, 7, 10) == []
    assert plan_segments(3, 10, 10) == [(7, 1, 3)]
    for total, world, ib in [(1000, 8, 10), (359, 8, 10), (13, 2, 5), (5, 8, 10), (1000, 3, 16), (1, 1, 10)]:
        seen = []
        for rank in range(world):
            t0, t1 = shard_range(total, rank, world)
            segs = plan_segments(t0, t1, ib)
            assert len(segs) <= 2
            for nb, steps, start in segs:
                assert 1 <= nb <= ib and steps >= 1
                seen += list(range(start, start + nb * steps))
        assert seen == list(range(total))
        assert shard_range(total, 0, world)[0] == 0


def test_bench_elementwise_bytes_follow_the_fused_levels():
    """bench.py's algorithmic bytes of the stand-alone apply kernels: fusing a level removes its applies, never adds bytes,
    and the boolean forms equal the all / none level sets."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    none = bench.elementwise_bytes_per_step(10, fused=False)
    allf = bench.elementwise_bytes_per_step(10, fused=True)
    assert allf == bench.elementwise_bytes_per_step(10, fused=[0, 1, 2, 3, 4]) and none == bench.elementwise_bytes_per_step(10, fused=[])
    prev = none["b2u_gn_apply"]
    for lv in ([4], [3, 4], [2, 3, 4], [1, 2, 3, 4], [0, 1, 2, 3, 4]):
        cur = bench.elementwise_bytes_per_step(10, fused=lv)["b2u_gn_apply"]
        assert cur < prev
        prev = cur
    default = bench.elementwise_bytes_per_step(10, fused=[1, 2, 3, 4])
    assert allf["b2u_gn_apply"] < default["b2u_gn_apply"] < none["b2u_gn_apply"]
    assert default["b2u_gn_apply_pool"] == none["b2u_gn_apply_pool"] and default["b2u_head_fwd"] == none["b2u_head_fwd"]
