"""On-disk formats the reference's downstream scripts read, written straight from the
device results:

  * Monte-Carlo outputs (uncertainty_tests/Dropblock_Uncertainty.py:157-165, read back by create_density_STD.py:49-66):
        <stats>/tensors/image_<id>/{mean.pt, std.pt, tensors.pt}   = torch.save of fp32 CPU tensors
        mean, std: [1,1,H,W]; tensors: [return_num,1,1,H,W]
  * Lightning-style checkpoints: {'state_dict': {...}} whose keys carry the `_model.` prefix of the LightningModule
    attribute (utils_training.py:13), so `load_from_checkpoint`-style loaders of the reference find the same 75 tensors.
"""
from __future__ import annotations

import os
from typing import Dict

import torch


def save_mc_outputs(stats_dir: str, im_id, mean: torch.Tensor, std: torch.Tensor, tensors: torch.Tensor) -> str:
    """One device->host copy per tensor, then the reference's three files.  Returns the image directory."""
    im_dir = os.path.join(stats_dir, "tensors", f"image_{im_id}")
    os.makedirs(im_dir, exist_ok=True)
    for name, t in (("mean.pt", mean), ("std.pt", std), ("tensors.pt", tensors)):
        torch.save(t.detach().to("cpu", torch.float32).contiguous(), os.path.join(im_dir, name))
    return im_dir


def load_mc_outputs(im_dir: str) -> Dict[str, torch.Tensor]:
    return {k: torch.load(os.path.join(im_dir, k + ".pt"), map_location="cpu") for k in ("mean", "std", "tensors")}


def save_checkpoint(path: str, training_module: torch.nn.Module, **extra) -> None:
    """{'state_dict': training_module.state_dict()} (+ extras): keys are `_model.<reference key>`."""
    sd = {k: v.detach().to("cpu") for k, v in training_module.state_dict().items()}
    torch.save({"state_dict": sd, **extra}, path)


def load_checkpoint_into(model: torch.nn.Module, path: str, prefix: str = "_model.") -> None:
    """Load a reference / Lightning checkpoint into a bare `UNet` (strips the `_model.` prefix)."""
    ck = torch.load(path, map_location="cpu")
    sd = ck["state_dict"] if "state_dict" in ck else ck
    model.load_state_dict({(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()})
