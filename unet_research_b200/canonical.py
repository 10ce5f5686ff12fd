"""The canonical reference configuration (base_model_tests/training.py:171-192) with synthetic weights: the model
every test, profile script, `bench.py` and `__graft_entry__.smoke()` builds."""
from __future__ import annotations

import torch
from torch import nn


def build_canonical(device, dropblock: bool = False, compute: str = "auto", init_channels: int = 1, seed: int = 1234,
                    drop_prob: float = 0.15, block_size: int = 7):
    """Canonical reference configuration (base_model_tests/training.py:171-192) with synthetic weights."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    m = U.UNet(init_channels=init_channels, filters=64, output_channels=1, model_depth=4, pool_mode='max',
               up_mode='upconv', connection='cat', same_padding=True, conv_layers_per_block=2, checkpointing=True)
    m.set_activation_function(nn.ReLU())
    if dropblock:
        m.set_dropblock(U.DropBlock2D, block_size=block_size, drop_prob=drop_prob, use_scheduler=False)
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m.create_model()
    sd = synthetic.make_state_dict(init_channels=init_channels, seed=seed)
    m.load_state_dict(sd)
    m.compute_dtype = compute
    m.to(device)
    m.eval()
    return m, {k: v.to(device) for k, v in sd.items()}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
