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
