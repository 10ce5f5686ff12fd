"""Golden vectors for `Dropblock2d_ichan` (reference utils_modules.py:86-139), generated from the UNMODIFIED reference:

    python tests/golden/make_golden_ichan.py          (build container only; needs /root/reference)

  dropblock_ichan_layer.npz   Dropblock2d_ichan(0.15, 7) on a 2x4x20x24 tensor with the captured bernoulli draw
  unet_ichan_fwd_120x116.npz  one training-mode forward of the canonical U-Net with Dropblock2d_ichan(0.15, 7) at all
                              22 sites, every bernoulli draw captured bit-packed in call order
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from unet_research_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(8)
    ref = ref_shims.load_reference()
    g = torch.Generator().manual_seed(77)
    real = torch.bernoulli
    draws = []

    def rec(p, *a, **k):
        m = real(p, generator=g)
        draws.append(m.clone())
        return m

    db = ref.Dropblock2d_ichan(0.15, 7)
    db.train()
    xin = torch.randn(2, 4, 20, 24, generator=g)
    torch.bernoulli = rec
    try:
        y = db(xin.clone())
    finally:
        torch.bernoulli = real
    np.savez_compressed(os.path.join(OUT, "dropblock_ichan_layer.npz"), x=xin.numpy(), draw=draws[0].numpy(), y=y.numpy())

    draws.clear()
    sd = synthetic.make_state_dict(seed=1234)
    unet = ref_shims.build_reference_unet(ref, dropblock=ref.Dropblock2d_ichan, drop_prob=0.15, block_size=7)
    unet.load_state_dict(sd)
    unet.eval()
    unet.apply(lambda l: setattr(l, "training", True) if type(l) == ref.Dropblock2d_ichan else None)
    x = synthetic.make_image(120, 116, seed=1234)
    torch.bernoulli = rec
    try:
        with torch.no_grad():
            out = unet(x)
    finally:
        torch.bernoulli = real
    packed = {f"draw{i:02d}": np.packbits(d.numpy().astype(np.uint8).reshape(-1)) for i, d in enumerate(draws)}
    shapes = np.array([list(d.shape) for d in draws], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "unet_ichan_fwd_120x116.npz"), output=out.numpy(), shapes=shapes, **packed)
    print("ichan layer", y.shape, "unet", out.shape, len(draws), "draws", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
