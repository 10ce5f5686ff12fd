"""CPU tests of the drop-in boundary: libb2u.so loads without a GPU, exports every symbol that
include/b2u.h declares, host-only entry points validate their arguments; Python host logic."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch
from torch import nn

import unet_research_b200 as U
from unet_research_b200 import _lib, engine, synthetic
from oracle import philox_oracle as P
from oracle import unet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2u.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2u_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b2u.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == syms, "ctypes binding and header disagree"
    assert lib.b2u_version() == 1


def test_struct_sizes_match_header_layout():
    assert C.sizeof(_lib.ConvDesc) == 12 * 4
    assert C.sizeof(_lib.ApplyDesc) == 14 * 4 + 8
    assert C.sizeof(_lib.HeadDesc) == 12 * 4
    assert C.sizeof(_lib.DropblockCall) == 3 * 8 + 12 * 4


def test_argument_validation_without_gpu():
    d = _lib.ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = 1, 16, 16, 60, 64, 0, 32, 60
    rows, sgs = C.c_int(0), C.c_int(0)
    with pytest.raises(_lib.B2uError, match="cin=60"):
        _lib.call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    d.cin = d.x_cstride = 64
    _lib.call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    assert rows.value == 2 and sgs.value == 2          # 16x16 image = two 16x8 tiles; 64 ch / 32 groups
    d.h, d.w, d.cout = 592, 576, 64
    _lib.call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    assert rows.value == 2664                          # M = 340 992 = 2664 x 128
    d.h, d.w, d.cin, d.cout, d.x_cstride = 37, 36, 1024, 512, 1024
    _lib.call("b2u_convT2x2_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    assert rows.value % 4 == 0 and sgs.value == 16
    with pytest.raises(_lib.B2uError, match="even"):
        _lib.call("b2u_pool_stat_layout", 37, 36, 64, 32, C.byref(rows), C.byref(sgs))
    with pytest.raises(_lib.B2uError):
        _lib.call("b2u_gn_apply", None, None, None, None, None, None, None, None)


def test_dropblock_plan_flat_dilate_grid():
    """b2u_dropblock_plan (host only): prefix sum of the dilate blocks of every call = planes x ceil(items / 4) with
    items = ceil(h / 37) bands x ceil(w / 32) words; the canonical U-Net needs 1792 blocks per iteration."""
    sites = engine.site_shapes(592, 576, 64, 4)
    calls = (_lib.DropblockCall * len(sites))()
    expect = []
    for d, (c, h, w) in zip(calls, sites):
        d.n_img, d.c, d.h, d.w, d.block_size = 1, c, h, w, 7
        expect.append((c // 32) * ((-(-h // 37) * -(-w // 32) + 3) // 4))
    total = C.c_longlong(0)
    _lib.call("b2u_dropblock_plan", calls, len(sites), C.byref(total))
    assert [d.dilate_first_block for d in calls] == [sum(expect[:i]) for i in range(len(sites))]
    assert total.value == sum(expect) == 1792
    calls[3].block_size = 8
    with pytest.raises(_lib.B2uError, match="odd"):
        _lib.call("b2u_dropblock_plan", calls, len(sites), C.byref(total))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B2uError, match="no CPU or PyTorch fallback"):
        _lib.load()


# ----------------------------------------------------------------------------- host logic
def canonical(dropblock=None, **kw):
    m = U.UNet(init_channels=1, filters=64, output_channels=1, model_depth=4)
    m.set_activation_function(nn.ReLU())
    if dropblock:
        m.set_dropblock(U.DropBlock2D, **kw)
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m.create_model()
    return m


def test_state_dict_layout_matches_reference():
    m = canonical()
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    assert list(m.state_dict().keys()) == keys
    sd = synthetic.make_state_dict()
    m.load_state_dict(sd)                       # strict: same shapes too
    wrapped = U.BaseUNetTraining(m, None, None)
    assert all(k.startswith("_model.") for k in wrapped.state_dict().keys())
    assert len(wrapped.state_dict()) == 75


def test_cpu_input_raises_no_fallback():
    m = canonical()
    with pytest.raises(_lib.B2uError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32))


def test_unsupported_configurations_raise():
    m = U.UNet(pool_mode="avg")
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    with pytest.raises(NotImplementedError, match="pool_mode"):
        m.create_model()
    with pytest.raises(ValueError):
        U.UNet(connection="bogus")
    with pytest.raises(ValueError):
        U.UNet(conv_layers_per_block=1)


def test_scheduler_semantics(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "scheduler.json")))
    m = canonical(dropblock=True, block_size=7, drop_prob=0.15, use_scheduler=True, start_drop_prob=0.,
                  max_drop_prob=0.15, dropblock_ls_steps=1500)
    sch = m._dropblock
    assert type(sch) is U.LinearScheduler and type(sch.dropblock) is U.DropBlock2D
    assert len(sch.drop_values) == g["len"]
    np.testing.assert_array_equal(sch.drop_values[g["idx"]], np.array(g["values"]))
    sch.step()
    assert sch.dropblock.drop_prob == 0.0 and sch.i == 1          # step 0 -> identity (utils_modules.py:42-43)
    for _ in range(2000):
        sch.step()
    assert sch.dropblock.drop_prob == pytest.approx(0.15) and sch.i == 2001


def test_set_dropblock_on_only_touches_dropblock():
    m = canonical(dropblock=True, block_size=7, drop_prob=0.15, use_scheduler=False)
    m.eval()
    assert m._dropblock_state()[0] is False
    m.apply(U.set_dropblock_on)
    assert m._dropblock.training and not m.training and not m.down_blocks[0][0][1].training
    assert m._dropblock_state() == (True, 0.15, 7)
    x = torch.randn(1, 2, 9, 9)
    m._dropblock.eval()
    assert m._dropblock(x) is x                                    # identity when not training


def test_philox_host_logic_matches_oracle():
    for gamma in (0.003125, 0.003191, 0.003329, 0.003634, 0.004384, 0.5):
        lo, hi = engine.philox_thresholds(gamma)
        assert lo == P.threshold_u32(gamma) and hi == P.threshold_hi_u32()
    for numel in (100, 64 * 586 * 570, 1024 * 31 * 30, 303104 * 4 + 1):
        assert engine.rand_grid(numel, 148, 2048) == P.torch_rand_grid(numel)
        assert engine.rand_offset_increment(numel, 148, 2048) == P.torch_rand_offset_increment(numel)
    assert engine.site_shapes(592, 576, 64, 4) == O.dropblock_site_shapes(592, 576)
    for (c, h, w) in engine.site_shapes(592, 576, 64, 4):
        assert engine.dropblock_gamma(0.15, 7, h, w) == O.dropblock_gamma(0.15, 7, h, w)


def test_shard_range_partitions():
    for total, world in ((1000, 1), (1000, 8), (359, 8), (7, 8), (25, 4)):
        spans = [U.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [U.shard_range(359, r, 8)[1] - U.shard_range(359, r, 8)[0] for r in range(8)].count(45) == 7


def test_io_formats_roundtrip(tmp_path):
    """mean.pt / std.pt / tensors.pt layout (Dropblock_Uncertainty.py:157-165) and `_model.`-prefixed checkpoints."""
    import torch
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import io as bio
    mean, std, tens = torch.rand(1, 1, 8, 9), torch.rand(1, 1, 8, 9), torch.rand(3, 1, 1, 8, 9)
    d = bio.save_mc_outputs(str(tmp_path), 7, mean, std, tens)
    assert d.endswith("tensors/image_7") and sorted(os.listdir(d)) == ["mean.pt", "std.pt", "tensors.pt"]
    back = bio.load_mc_outputs(d)
    assert torch.equal(back["mean"], mean) and torch.equal(back["tensors"], tens) and back["std"].dtype == torch.float32
    m = U.UNet(init_channels=1, filters=64, output_channels=1, model_depth=4)
    m.set_activation_function(nn.ReLU())
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m.create_model()
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    p = str(tmp_path / "last.ckpt")
    bio.save_checkpoint(p, tm, epoch=3)
    ck = torch.load(p)
    assert len(ck["state_dict"]) == 75 and all(k.startswith("_model.") for k in ck["state_dict"]) and ck["epoch"] == 3
    m2 = U.UNet(init_channels=1, filters=64, output_channels=1, model_depth=4)
    m2.set_activation_function(nn.ReLU())
    m2.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m2.create_model()
    bio.load_checkpoint_into(m2, p)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_pack_batched_plan_without_gpu():
    """b2u_pack_batched_plan is host-only: flat grid prefix (one 32 x 32 tile per block) and argument validation."""
    ents = (_lib.PackEntry * 3)()
    for e, (kind, cout, cin) in zip(ents, [(0, 128, 64), (1, 64, 128), (0, 1024, 512)]):
        e.w, e.out0, e.out1, e.kind, e.cout, e.cin = 0x1000, 0x2000, None, kind, cout, cin
    blocks = C.c_int(0)
    _lib.call("b2u_pack_batched_plan", C.byref(ents), 3, C.byref(blocks))
    assert [e.first_block for e in ents] == [0, 8, 16] and blocks.value == 16 + 32 * 16
    assert C.sizeof(_lib.PackEntry) == 3 * 8 + 4 * 4
    ents[1].cin = 100
    with pytest.raises(_lib.B2uError, match="multiples of 32"):
        _lib.call("b2u_pack_batched_plan", C.byref(ents), 3, C.byref(blocks))
    ents[1].cin, ents[1].kind = 128, 7
    with pytest.raises(_lib.B2uError, match="kind"):
        _lib.call("b2u_pack_batched_plan", C.byref(ents), 3, C.byref(blocks))


def test_head_plan_without_gpu(monkeypatch):
    """b2u_head_plan is host-only: the launch shape of the c = 64 head covers every pixel octet, the block size is the one
    that fills the last trip better, the cp.async ring fits twice per SM, and the diagnostic overrides are validated."""
    for k in ("B2U_HEAD_ASYNC", "B2U_HEAD_THREADS", "B2U_HEAD_STAGES"):
        monkeypatch.delenv(k, raising=False)
    plan = (C.c_int * 6)()
    for h0, w0 in [(584, 565), (146, 141), (292, 283), (128, 128), (256, 256), (584, 584), (1, 1), (7, 3), (2048, 2048)]:
        _lib.call("b2u_head_plan", h0, w0, 148, plan)
        variant, nt, grid, trips, stages, ring = list(plan)
        gtrips = (h0 * w0 + 7) // 8
        assert variant == 1 and nt in (224, 192) and stages == 2
        assert ring == nt * stages * (8 * 16 + 8) and 2 * (ring + 1024) <= 227 * 1024
        assert 1 <= grid <= 2 * 148 and grid * (nt // 8) * trips >= gtrips          # every octet has an owner
        assert grid * (nt // 8) * (trips - 1) < gtrips                              # ... and no trip is all empty
        fill = {c: gtrips / (-(-gtrips // (296 * c // 8)) * (296 * c // 8)) for c in (224, 192)}
        if gtrips >= 296 * 224 // 8:
            assert fill[nt] >= max(fill.values()) - 1e-9
    _lib.call("b2u_head_plan", 584, 565, 148, plan)
    assert list(plan)[:4] == [1, 224, 296, 5]                                       # 41245 octets over 8288 resident groups: 99.5 % full
    monkeypatch.setenv("B2U_HEAD_ASYNC", "0")
    _lib.call("b2u_head_plan", 584, 565, 148, plan)
    assert list(plan) == [0, 256, 296, 5, 2, 0]                                     # register kernel: 9472 groups, fifth trip 35 % full
    monkeypatch.setenv("B2U_HEAD_ASYNC", "1")
    monkeypatch.setenv("B2U_HEAD_THREADS", "224")
    monkeypatch.setenv("B2U_HEAD_STAGES", "4")
    with pytest.raises(_lib.B2uError, match="does not fit twice"):
        _lib.call("b2u_head_plan", 584, 565, 148, plan)
    monkeypatch.setenv("B2U_HEAD_THREADS", "100")
    with pytest.raises(_lib.B2uError, match="multiple of 32"):
        _lib.call("b2u_head_plan", 584, 565, 148, plan)
    with pytest.raises(_lib.B2uError, match="bad arguments"):
        _lib.call("b2u_head_plan", 0, 565, 148, plan)
