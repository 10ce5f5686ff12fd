"""One eager training step (masks + forward + backward, batch 1, 584x565, DropBlock p=.15) for ncu:
    ncu --metrics ... --profile-from-start off python tests/prof_train.py      (captures the last step only)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from torch import nn

import gpu_diag as D
import unet_research_b200 as U
from unet_research_b200 import synthetic

dev = torch.device("cuda")
m, _ = D._build_model(dev, dropblock=True)
m.use_cuda_graph = False
m.train()
h, w = 584, 565
x = synthetic.make_image(h, w, seed=1234).to(dev)
gt = synthetic.make_gt(h, w).to(dev)
fov = synthetic.make_fov_mask(h, w).to(dev)
tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    for p in m.parameters():
        p.grad = None
    tm.training_step((x.clone(), gt, fov), 0).backward()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
