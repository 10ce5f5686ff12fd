"""`__graft_entry__.smoke()`: one small invocation of the hot path on a CUDA device, checked against the
oracle (this is one of the three places allowed to import `oracle/`)."""
from __future__ import annotations

import torch
from torch import nn


def build_canonical(device, dropblock: bool = False, compute: str = "bf16", init_channels: int = 1, seed: int = 1234,
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


def run(device) -> None:
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    from oracle import unet_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w = 120, 116
    x = synthetic.make_image(h, w, seed=1234).to(device)
    fov = synthetic.make_fov_mask(h, w).to(device)
    # eval forward
    model, sd = build_canonical(device)
    with torch.no_grad():
        y = model(x)
        ref = O.unet_forward(sd, x)
    e = rel_l2(y, ref)
    assert e < 1e-2, f"eval forward differs from the oracle: rel L2 {e:.3e}"
    # 2 MC-DropBlock iterations, same Philox stream as the oracle's torch.rand calls
    model, sd = build_canonical(device, dropblock=True)
    ev = U.DropBlockEval(model, num_iterations=2, return_num=2, iter_batch=2, use_cuda_graph=False)
    torch.manual_seed(1234)
    _, (mean, std, tensors) = ev.predict_step((x, None, fov), 0)
    torch.manual_seed(1234)
    rmean, rstd, rtensors = O.mc_dropblock(sd, x, fov, 2, 2, 0.15, 7)
    e1, e2 = rel_l2(tensors, rtensors), rel_l2(mean, rmean)
    assert e1 < 1e-2 and e2 < 1e-2, f"MC-DropBlock differs from the oracle: samples {e1:.3e} mean {e2:.3e}"
    assert float((std - rstd).abs().max()) < 2e-2
    torch.cuda.synchronize(device)
    print(f"smoke ok: eval rel {e:.2e}, mc samples rel {e1:.2e}, mean rel {e2:.2e}")
