"""Upper-bound experiment on the CAPTURED training step (timing only; skipped kernels leave stale data):
    B2U_EXP_SKIP_CALLS=b2u_pack_conv3x3_weight_pair,... python tests/exp_train_skip.py [steps]
prints ms per step (fwd + bwd + clip + SGD, batch 1, 584x565, DropBlock, bf16, CUDA graphs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn

import unet_research_b200 as U
from unet_research_b200 import synthetic
from unet_research_b200.canonical import build_canonical

dev = torch.device("cuda")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
model, _ = build_canonical(dev, dropblock=True, compute="bf16")
model.train()
x = synthetic.make_image(584, 565, seed=1234).to(dev)
gt = synthetic.make_gt(584, 565, seed=1234).to(dev)
fov = synthetic.make_fov_mask(584, 565).to(dev)
tm = U.BaseUNetTraining(model, nn.BCELoss(), None)
opt = U.FusedSGD(model.parameters(), lr=1e-3, momentum=0.99, max_grad_norm=0.5)


def step():
    opt.zero_grad(set_to_none=True)
    loss = tm.training_step((x.clone(), gt, fov), 0)
    loss.backward()
    opt.step()


for _ in range(6):
    step()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print(f"skip=[{os.environ.get('B2U_EXP_SKIP_CALLS', '')}] lib={os.path.basename(os.environ.get('B2U_LIB', 'libb2u.so'))}: {best:.3f} ms per train step")
