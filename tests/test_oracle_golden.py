"""CPU tests: the oracle restatement (oracle/unet_oracle.py, oracle/philox_oracle.py) against the
golden vectors generated from the UNMODIFIED reference (tests/golden/make_golden.py) and against
published known-answer vectors."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import philox_oracle as P
from oracle import unet_oracle as O
from unet_research_b200 import synthetic

H, W = 120, 116
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.fixture(scope="module")
def sd():
    return synthetic.make_state_dict(seed=1234)


def test_state_dict_keys_match_reference(golden_dir, sd):
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    assert list(sd.keys()) == keys
    assert len(keys) == 75
    assert sum(v.numel() for v in sd.values()) == 31_039_360


def test_eval_forward_matches_reference(golden_dir, sd):
    g = _load(golden_dir, "unet_eval_120x116.npz")
    x = synthetic.make_image(H, W, seed=1234)
    taps = {}
    with torch.no_grad():
        y = O.unet_forward(sd, x, taps=taps)
    np.testing.assert_allclose(y.numpy(), g["output"], rtol=0, atol=2e-6)
    # per-layer taps: reference module names -> oracle tap names
    name_map = {"down_blocks.0.0.0": "d0.c1.conv", "down_blocks.0.0.4": "d0.c2.conv",
                "down_blocks.3.0.4": "d3.c2.conv", "conn_block.4": "b.c2.conv",
                "up_blocks.0.0.0": "u0.up", "up_blocks.3.1.4": "u3.c2.conv",
                "down_blocks.1.1.0": "d1.pool", "output_conv.0": "logits"}
    for ref_name, tap in name_map.items():
        t = taps[tap].double()
        stats = np.array([t.mean().item(), t.pow(2).mean().sqrt().item()])
        np.testing.assert_allclose(stats, g["tap_stats/" + ref_name], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(t.flatten()[:32].float().numpy(), g["tap_lead/" + ref_name], rtol=1e-4, atol=1e-5)


def test_eval_forward_rgb(golden_dir):
    g = _load(golden_dir, "unet_eval_rgb_64x80.npz")
    sd3 = synthetic.make_state_dict(init_channels=3, seed=1234)
    x3 = synthetic.make_image(64, 80, channels=3, seed=7)
    with torch.no_grad():
        y = O.unet_forward(sd3, x3)
    np.testing.assert_allclose(y.numpy(), g["output"], rtol=0, atol=2e-6)


def test_dropblock_layer_matches_reference(golden_dir):
    g = _load(golden_dir, "dropblock_layer.npz")
    u = torch.from_numpy(g["u"])
    y = O.dropblock2d(torch.from_numpy(g["x"]), 0.15, 7, True, rand_fn=lambda *s, **k: u)
    np.testing.assert_array_equal(y.numpy(), g["y"])


def test_dropblock_identity_cases():
    x = torch.randn(1, 2, 16, 16)
    assert O.dropblock2d(x, 0.15, 7, training=False) is x
    assert O.dropblock2d(x, 0.0, 7, training=True) is x


def test_mc_dropblock_matches_reference(golden_dir, sd):
    g = _load(golden_dir, "mc_dropblock_120x116.npz")
    x = synthetic.make_image(H, W, seed=1234)
    mask = synthetic.make_fov_mask(H, W)
    torch.manual_seed(1234)
    mean, std, tensors = O.mc_dropblock(sd, x, mask, 3, 2, drop_prob=0.15, block_size=7)
    assert tuple(tensors.shape) == (2, 1, 1, H, W)
    np.testing.assert_allclose(tensors.numpy(), g["tensors"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(mean.numpy(), g["mean"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(std.numpy(), g["std"], rtol=0, atol=3e-6)
    assert float(mean[mask == 0].abs().max()) == 0.0 and float(std[mask == 0].abs().max()) == 0.0


def test_rotation_matches_reference(golden_dir, sd):
    g = _load(golden_dir, "rotation_120x116.npz")
    x = synthetic.make_image(H, W, seed=1234)
    mask = synthetic.make_fov_mask(H, W)
    mean, std, tensors = O.rotation_ensemble(sd, x, mask, 3, 2)
    np.testing.assert_allclose(tensors.numpy(), g["tensors"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(mean.numpy(), g["mean"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(std.numpy(), g["std"], rtol=0, atol=3e-6)


def test_rotate_restatement_vs_torchvision():
    tvf = pytest.importorskip("torchvision.transforms.functional")
    x = synthetic.make_image(37, 52, seed=3)
    for ang in (1.0, -1.0, 33.0, 90.0, 179.0, 271.0, -359.0):
        ref = tvf.rotate(x, angle=ang, interpolation=tvf.InterpolationMode.BILINEAR, fill=0)
        np.testing.assert_allclose(O.rotate_bilinear(x, ang).numpy(), ref.numpy(), rtol=0, atol=1e-6)


def test_train_step_matches_reference(golden_dir, sd):
    g = _load(golden_dir, "train_step_120x116.npz")
    x = synthetic.make_image(H, W, seed=1234)
    gt = synthetic.make_gt(H, W)
    mask = synthetic.make_fov_mask(H, W)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(4321)
    loss = O.train_step_loss(params, x, gt, mask, O.DropBlockCfg(0.15, 7, True))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-6)
    for k, p in params.items():
        gn = p.grad.double().norm().item()
        np.testing.assert_allclose(gn, float(g["gnorm/" + k]), rtol=2e-3, atol=1e-9)


def test_scheduler_values(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "scheduler.json")))
    vals = O.linear_scheduler_values(0.0, 0.15, 1500)
    assert len(vals) == g["len"]
    np.testing.assert_array_equal(vals[g["idx"]], np.array(g["values"]))


def test_site_shapes_and_uniform_count():
    sites = O.dropblock_site_shapes(592, 576)
    assert len(sites) == 22
    n_uniform = sum(c * (h - 6) * (w - 6) for c, h, w in sites)
    assert abs(n_uniform - 237e6) < 1.5e6          # ~237 M uniforms / forward
    assert abs(O.dropblock_gamma(0.15, 7, 592, 576) - 0.003125) < 5e-6
    assert abs(O.dropblock_gamma(0.15, 7, 37, 36) - 0.004384) < 5e-6


# ---------------------------------------------------------------- Philox / torch.rand stream
def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = P.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(v[0]) for v in got) == want


def test_threshold_equivalence():
    rng = np.random.default_rng(0)
    for gamma in (0.003125, 0.003191, 0.004384, 0.25, 1e-6):
        t = P.threshold_u32(gamma)
        xs = np.concatenate([rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32),
                             np.arange(max(t - 600, 0), t + 600, dtype=np.uint64).astype(np.uint32)])
        xs = xs[xs < P.threshold_hi_u32()]
        want = P.curand_uniform_from_u32(xs) < np.float32(gamma)
        np.testing.assert_array_equal(xs < t, want)
    hi = P.threshold_hi_u32()
    assert hi == 0xFFFFFF80
    assert P.curand_uniform_from_u32(np.array([hi - 1], dtype=np.uint32))[0] < 1.0


def test_offset_increment_formula():
    # B200: 148 SMs x 8 blocks of 256
    assert P.torch_rand_grid(64 * 586 * 570) == 1184
    assert P.torch_rand_offset_increment(64 * 586 * 570) == ((64 * 586 * 570 - 1) // (1184 * 256 * 4) + 1) * 4
    assert P.torch_rand_offset_increment(100) == 4


def test_dropblock_ichan_layer_matches_reference(golden_dir):
    """Dropblock2d_ichan (utils_modules.py:86-139) with the reference's own bernoulli draw fed back in."""
    g = _load(golden_dir, "dropblock_ichan_layer.npz")
    draw = torch.from_numpy(g["draw"])
    y = O.dropblock2d_ichan(torch.from_numpy(g["x"]), 0.15, 7, True, bernoulli_fn=lambda p: draw.clone())
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-6, atol=1e-7)
    # inactive in eval mode / at drop_prob 0 (:107-108)
    x = torch.from_numpy(g["x"])
    assert O.dropblock2d_ichan(x, 0.15, 7, False) is x and O.dropblock2d_ichan(x, 0.0, 7, True) is x


def test_unet_forward_ichan_matches_reference(golden_dir, sd):
    """U-Net forward with Dropblock2d_ichan at all 22 sites, the reference's 22 bernoulli draws replayed in call order."""
    g = _load(golden_dir, "unet_ichan_fwd_120x116.npz")
    shapes = g["shapes"]
    assert len(shapes) == 22
    it = iter(range(22))

    def replay(p):
        i = next(it)
        shp = tuple(int(v) for v in shapes[i])
        assert tuple(p.shape) == shp
        n = int(np.prod(shp))
        return torch.from_numpy(np.unpackbits(g[f"draw{i:02d}"])[:n].reshape(shp).astype(np.float32))

    x = synthetic.make_image(H, W, seed=1234)
    with torch.no_grad():
        y = O.unet_forward(sd, x, dropblock=O.DropBlockCfg(0.15, 7, True, mode="ichan", bernoulli_fn=replay))
    np.testing.assert_allclose(y.numpy(), g["output"], rtol=0, atol=5e-6)


def test_square_pad_resize_restatement_matches_torchvision():
    """The oracle's square_pad / resize restatement against the functions the reference actually calls
    (utils_general.square_pad via the reference import when present, torchvision TF.resize)."""
    import torchvision.transforms.functional as TF
    from oracle import ref_shims
    x = synthetic.make_image(584, 565, seed=3)
    if ref_shims.reference_available():
        ref_shims.load_reference()
        from utils import utils_general  # type: ignore
        assert torch.equal(O.square_pad(x), utils_general.square_pad(x))
    sp = O.square_pad(x)
    assert sp.shape[-2:] == (584, 584) and float(sp[..., :, :10].abs().max()) == 0.0 and float(sp[..., :, -9:].abs().max()) == 0.0
    for s in (128, 256, 584):
        np.testing.assert_allclose(O.square_pad_resize(x, s).numpy(), TF.resize(sp, size=(s, s)).numpy(), rtol=0, atol=1e-6)


def test_reference_bytecode_matches_source(golden_dir):
    """oracle/_ref (byte-code of the unmodified reference, built by oracle/build_ref.py: what the GPU box imports) is
    complete and, loaded in a fresh process, reproduces the golden eval forward bit for bit."""
    import subprocess
    import sys
    from oracle import build_ref
    if not build_ref.built():
        if not build_ref.ref_available():
            pytest.skip("oracle/_ref not built and /root/reference absent")
        assert build_ref.build_ref()
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
        "from oracle import ref_shims as R\n"
        "from unet_research_b200 import synthetic\n"
        "ref = R.load_reference(prefer='bytecode'); assert ref.kind == 'bytecode'\n"
        "u = R.build_reference_unet(ref); u.load_state_dict(synthetic.make_state_dict(seed=1234)); u.eval()\n"
        "g = np.load(%r)\n"
        "with torch.no_grad(): y = u(synthetic.make_image(120, 116, seed=1234))\n"
        "assert np.array_equal(y.numpy(), g['output']), float(np.abs(y.numpy() - g['output']).max())\n"
        "print('BYTECODE OK')\n") % (ROOT, os.path.join(golden_dir, "unet_eval_120x116.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "BYTECODE OK" in r.stdout, r.stderr[-2000:]
