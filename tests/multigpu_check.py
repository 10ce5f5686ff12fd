"""Multi-GPU checks, launched with torchrun on an N-GPU box (not a pytest file: one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

1. MC-DropBlock sharded over ranks == the single-rank result (same seed): samples bit-identical, mean/std to 1e-6.
2. Rotation ensemble sharded over ranks == single rank.
3. Data-parallel training: every rank trains on a DIFFERENT image; the all-reduced gradients equal the average of the
   per-rank gradients computed without the exchange (gathered for the check), and all ranks hold identical gradients.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import gpu_diag as D
    import unet_research_b200 as U
    from torch import nn
    from unet_research_b200 import synthetic
    from unet_research_b200 import uncertainty as UN
    h, w = 120, 116
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    ok = True

    # ---- 1. MC sharding
    m, _ = D._build_model(dev, dropblock=True)
    T = 12
    ev = U.DropBlockEval(m, num_iterations=T, return_num=5, iter_batch=2, gather_samples=True)
    torch.manual_seed(99)
    _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
    # default (no sample exchange): rank 0 owns iterations [0, return_num) whenever return_num <= its share
    ev0 = U.DropBlockEval(m, num_iterations=40, return_num=3, iter_batch=2)
    torch.manual_seed(99)
    _, (_, _, tens0) = ev0.predict_step((x, None, fov), 0)
    # single-rank reference: hide the process group from the loop
    real = UN._dist
    UN._dist = lambda: (None, 0, 1)
    try:
        ev1 = U.DropBlockEval(m, num_iterations=T, return_num=5, iter_batch=2)
        torch.manual_seed(99)
        _, (mean1, std1, tens1) = ev1.predict_step((x, None, fov), 0)
    finally:
        UN._dist = real
    e_s = float((tens - tens1).abs().max())
    e_m = float((mean - mean1).abs().max())
    e_d = float((std - std1).abs().max())
    good = e_s == 0.0 and e_m < 1e-6 and e_d < 1e-6
    if rank == 0:
        good = good and torch.equal(tens0, tens1[:3])
    ok &= good
    if rank == 0:
        print(f"[mc sharding x{world}] samples max|d| {e_s:.1e} mean {e_m:.1e} std {e_d:.1e} -> {'OK' if good else 'FAIL'}", flush=True)

    # ---- 2. rotation sharding
    m2, _ = D._build_model(dev)
    rv = U.RotationEval(m2, num_iterations=7, return_num=3, angle_batch=2, gather_samples=True)
    _, (rmean, rstd, rtens) = rv.predict_step((x, None, fov), 0)
    UN._dist = lambda: (None, 0, 1)
    try:
        rv1 = U.RotationEval(m2, num_iterations=7, return_num=3, angle_batch=2)
        _, (rmean1, rstd1, rtens1) = rv1.predict_step((x, None, fov), 0)
    finally:
        UN._dist = real
    good = float((rtens - rtens1).abs().max()) == 0.0 and float((rmean - rmean1).abs().max()) < 1e-6 and float((rstd - rstd1).abs().max()) < 1e-6
    ok &= good
    if rank == 0:
        print(f"[rotation sharding x{world}] -> {'OK' if good else 'FAIL'}", flush=True)

    # ---- 3. data-parallel gradients
    m3, _ = D._build_model(dev)
    m3.train()
    tm = U.BaseUNetTraining(m3, nn.BCELoss(), None)
    xr = synthetic.make_image(h, w, seed=100 + rank).to(dev)
    gt = synthetic.make_gt(h, w, seed=200 + rank).to(dev)
    tm.training_step((xr, gt, fov), 0).backward()
    local_g = torch.cat([p.grad.flatten() for p in m3.parameters()]).clone()
    gathered = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    want = torch.stack(gathered).mean(0)
    for p in m3.parameters():
        p.grad = None
    m3.data_parallel = True
    tm.training_step((xr, gt, fov), 0).backward()
    got = torch.cat([p.grad.flatten() for p in m3.parameters()])
    err = float((got - want).norm() / want.norm())
    same = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(same, got)
    ident = all(torch.equal(same[0], s) for s in same)
    good = err < 1e-6 and ident
    ok &= good
    if rank == 0:
        print(f"[ddp gradients x{world}] rel err vs mean of per-rank grads {err:.2e}, identical on all ranks {ident} -> {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0 and int(t.item()) == 1:
        print("MULTIGPU OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
