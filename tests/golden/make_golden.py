"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on deterministic synthetic inputs.

Run in the build container only:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box; these small fixtures do.

Fixtures (all at the canonical model config, weights from `synthetic.make_state_dict(seed=1234)`):
  unet_eval_120x116.npz     eval forward of a 1x1x120x116 image (autopad -> 128x128) +
                            per-layer taps (mean / rms / 32 leading values of every Conv2d,
                            ConvTranspose2d, MaxPool2d output)
  unet_eval_rgb_64x80.npz   same with init_channels=3 (BASELINE config 1 wording), taps omitted
  dropblock_layer.npz       DropBlock2D(0.15, 7) on a 2x4x20x24 tensor with captured uniforms
  mc_dropblock_120x116.npz  DropBlockEval.predict_step, 3 iterations, CPU seed 1234 (mt19937 stream)
  rotation_120x116.npz      RotationEval.predict_step, angles 1..3
  train_step_120x116.npz    training_step loss + gradient norms / leading values, DropBlock p=.15 (captured seed)
  scheduler.json            LinearScheduler values as used by the reference UNet (0 -> .15 over 1500 steps)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from unet_research_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H, W = 120, 116


def tap_summary(t: torch.Tensor):
    t = t.detach().double()
    return np.array([t.mean().item(), t.pow(2).mean().sqrt().item()]), t.flatten()[:32].float().numpy()


def main():
    torch.set_num_threads(8)
    ref = ref_shims.load_reference()
    sd = synthetic.make_state_dict(seed=1234)

    # ---- eval forward with taps
    unet = ref_shims.build_reference_unet(ref)
    unet.load_state_dict(sd)
    unet.eval()
    x = synthetic.make_image(H, W, seed=1234)
    taps = {}
    hooks = []
    for name, mod in unet.named_modules():
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d, nn.MaxPool2d)):
            hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o)))
    with torch.no_grad():
        y = unet(x)
    for h in hooks:
        h.remove()
    out = {"output": y.numpy()}
    for k, v in taps.items():
        s, lead = tap_summary(v)
        out["tap_stats/" + k] = s
        out["tap_lead/" + k] = lead
    np.savez_compressed(os.path.join(OUT, "unet_eval_120x116.npz"), **out)
    print("eval", y.shape, float(y.min()), float(y.max()))

    # ---- RGB variant
    sd3 = synthetic.make_state_dict(init_channels=3, seed=1234)
    unet3 = ref_shims.build_reference_unet(ref, init_channels=3)
    unet3.load_state_dict(sd3)
    unet3.eval()
    x3 = synthetic.make_image(64, 80, channels=3, seed=7)
    with torch.no_grad():
        y3 = unet3(x3)
    np.savez_compressed(os.path.join(OUT, "unet_eval_rgb_64x80.npz"), output=y3.numpy())

    # ---- single DropBlock layer with captured uniforms
    db = ref.DropBlock2D(0.15, 7)
    db.train()
    g = torch.Generator().manual_seed(99)
    xin = torch.randn(2, 4, 20, 24, generator=g)
    captured = {}
    real_rand = torch.rand

    def rec_rand(*shape, **kw):
        u = real_rand(*shape, generator=g)
        captured["u"] = u
        return u

    torch.rand = rec_rand
    try:
        yout = db(xin)
    finally:
        torch.rand = real_rand
    np.savez_compressed(os.path.join(OUT, "dropblock_layer.npz"), x=xin.numpy(), u=captured["u"].numpy(), y=yout.numpy())

    # ---- MC DropBlock (3 iterations, CPU RNG stream)
    unet_db = ref_shims.build_reference_unet(ref, dropblock=ref.DropBlock2D, drop_prob=0.15, block_size=7)
    unet_db.load_state_dict(sd)
    ev = ref.DropBlockEval(unet_db, num_iterations=3, return_num=2, mode="save")
    ev.eval()
    mask = synthetic.make_fov_mask(H, W)
    gt = synthetic.make_gt(H, W)
    torch.manual_seed(1234)
    with torch.no_grad():
        _, (mean, std, tensors) = ev.predict_step((x, gt, mask), 0)
    np.savez_compressed(os.path.join(OUT, "mc_dropblock_120x116.npz"), mean=mean.numpy(), std=std.numpy(), tensors=tensors.numpy())
    print("mc", mean.shape, std.shape, tensors.shape, float(std.max()))

    # ---- rotation ensemble (angles 1..3)
    unet_r = ref_shims.build_reference_unet(ref)
    unet_r.load_state_dict(sd)
    rv = ref.RotationEval(unet_r, num_iterations=3, return_num=2)
    rv.eval()
    with torch.no_grad():
        _, (rmean, rstd, rtens) = rv.predict_step((x, gt, mask), 0)
    np.savez_compressed(os.path.join(OUT, "rotation_120x116.npz"), mean=rmean.numpy(), std=rstd.numpy(), tensors=rtens.numpy())
    print("rot", rmean.shape, float(rstd.max()))

    # ---- train step (DropBlock fixed at p=.15, scheduler off; CPU seed 4321)
    unet_t = ref_shims.build_reference_unet(ref, dropblock=ref.DropBlock2D, drop_prob=0.15, block_size=7)
    unet_t.load_state_dict(sd)
    tm = ref.UNetTraining(unet_t, loss_fcn=nn.BCELoss(), lr=1e-3, momentum=0.99)
    tm.train()
    torch.manual_seed(4321)
    xt = x.clone()
    loss = tm.training_step((xt, gt, mask), 0)
    loss.backward()
    gout = {"loss": np.array(loss.item())}
    for k, p in unet_t.named_parameters():
        gout["gnorm/" + k] = np.array(p.grad.double().norm().item())
        gout["glead/" + k] = p.grad.flatten()[:16].numpy()
    np.savez_compressed(os.path.join(OUT, "train_step_120x116.npz"), **gout)
    print("train loss", loss.item())

    # ---- scheduler semantics through the reference UNet (0 -> .15, 1500 steps)
    unet_s = ref_shims.build_reference_unet(ref, dropblock=ref.DropBlock2D, drop_prob=0.15, block_size=7,
                                            use_scheduler=True, start_drop_prob=0., max_drop_prob=0.15,
                                            dropblock_ls_steps=1500)
    vals = [float(v) for v in unet_s._dropblock.drop_values[[0, 1, 2, 749, 1498, 1499]]]
    json.dump({"idx": [0, 1, 2, 749, 1498, 1499], "values": vals, "len": len(unet_s._dropblock.drop_values),
               "note": "LinearScheduler is the shimmed restatement of dropblock==0.3.0 (parity unpinned)"},
              open(os.path.join(OUT, "scheduler.json"), "w"), indent=1)
    keys = list(unet.state_dict().keys())
    json.dump(keys, open(os.path.join(OUT, "state_dict_keys.json"), "w"), indent=0)
    print("keys", len(keys))


if __name__ == "__main__":
    main()
