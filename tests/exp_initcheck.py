"""DIAGNOSTIC (not a test): the smallest eager Monte-Carlo DropBlock call (120x116, 4 iterations in batches of 2, no CUDA graph)
for `compute-sanitizer --tool initcheck` -- does any kernel of the path read device memory nobody wrote?
    PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck --log-file initcheck.log python tests/exp_initcheck.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import gpu_diag as D
import unet_research_b200 as U
from unet_research_b200 import synthetic

dev = torch.device("cuda")
m, _ = D._build_model(dev, dropblock=True)
x = synthetic.make_image(120, 116, seed=1234).to(dev)
fov = synthetic.make_fov_mask(120, 116).to(dev)
ev = U.DropBlockEval(m, num_iterations=4, return_num=2, iter_batch=2, use_cuda_graph=False)
torch.manual_seed(5)
_, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
torch.cuda.synchronize()
print("done", float(mean.sum()), float(std.sum()))
