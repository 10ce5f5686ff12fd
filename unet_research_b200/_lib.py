"""ctypes binding of the C-ABI library `csrc/libb2u.so` (declared in `include/b2u.h`).

There is NO CPU fallback: every compute entry point of this package goes through this
library, and `load()` raises loudly when it has not been built (`python __graft_entry__.py`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2U_LIB") or os.path.join(_HERE, "csrc", "libb2u.so")     # B2U_LIB: A/B runs of two builds

BF16, F32, F16 = 0, 1, 2


class B2uError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("dtype", C.c_int32), ("num_groups", C.c_int32), ("x_cstride", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class PackEntry(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out0", C.c_void_p), ("out1", C.c_void_p), ("kind", C.c_int32), ("cout", C.c_int32),
                ("cin", C.c_int32), ("first_block", C.c_int32)]


class ApplyDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("dtype", C.c_int32),
                ("relu", C.c_int32), ("out_cstride", C.c_int32), ("out_coffset", C.c_int32),
                ("mask2_cstride", C.c_int32), ("mask2_coffset", C.c_int32), ("images_per_call2", C.c_int32),
                ("reserved", C.c_int32 * 3), ("numel_per_call2", C.c_double)]


class HeadDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("h0", C.c_int32),
                ("w0", C.c_int32), ("dtype", C.c_int32), ("return_num", C.c_int32), ("fov_per_image", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class UnitBwdDesc(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n", "h", "w", "c", "dtype", "relu", "num_groups", "s2d", "images_per_call1", "a_cstride",
                                          "a_coffset", "mask2_cstride", "mask2_coffset", "images_per_call2", "h0", "w0")] +
                [("numel_per_call1", C.c_double), ("numel_per_call2", C.c_double)] +
                [(n, C.c_void_p) for n in ("y", "coef", "mean_rstd", "gamma", "mask1", "keep_counts1", "grad_a", "mask2",
                                           "keep_counts2", "grad_pool", "argmax", "grad_out", "out", "w_head")])


class WgradDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n", "h", "w", "cg", "cx", "x_cstride", "taps", "layout", "dtype")] + [("reserved", C.c_int32 * 3)]


class DropblockCall(C.Structure):
    _fields_ = [("philox_offset", C.c_uint64), ("center_word_off", C.c_uint64), ("mask_word_off", C.c_uint64),
                ("numel", C.c_uint32), ("grid", C.c_uint32), ("thresh_lo", C.c_uint32), ("thresh_hi", C.c_uint32),
                ("n_img", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("block_size", C.c_int32), ("count_index", C.c_int32), ("dilate_first_block", C.c_int32),
                ("reserved", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_D = C.c_double
_F = C.c_float

# name -> (restype, argtypes); every symbol include/b2u.h declares
SIGNATURES = {
    "b2u_last_error": (C.c_char_p, []),
    "b2u_version": (_I, []),
    "b2u_device_info": (_I, [C.POINTER(_I), C.POINTER(_I)]),
    "b2u_pack_conv3x3_weight": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b2u_pack_conv3x3_weight_pair": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "b2u_pack_batched_plan": (_I, [_P, _I, _P]),
    "b2u_pack_batched": (_I, [_P, _I, _I, _I, _P]),
    "b2u_pack_convT2x2_weight": (_I, [_P, _P, _I, _I, _I, _P]),
    "b2u_conv3x3_stat_layout": (_I, [C.POINTER(ConvDesc), C.POINTER(_I), C.POINTER(_I)]),
    "b2u_convT2x2_stat_layout": (_I, [C.POINTER(ConvDesc), C.POINTER(_I), C.POINTER(_I)]),
    "b2u_conv3x3_fwd": (_I, [_P, _P, _P, _P, C.POINTER(ConvDesc), _P]),
    "b2u_conv3x3_pro_fwd": (_I, [_P, _P, _P, _P, _P, _P, C.POINTER(ConvDesc), _I, _I, _P]),
    "b2u_convT2x2_fwd": (_I, [_P, _P, _P, _P, C.POINTER(ConvDesc), _P]),
    "b2u_conv_first_stat_layout": (_I, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "b2u_conv_first_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b2u_gn_finalize": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _I, _D, _F, _P, _I, _D, _P, _P]),
    "b2u_gn_finalize_ex": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _I, _D, _F, _P, _I, _D, _P, _I, _P]),
    "b2u_unit_bwd_rows": (_I, [_I, _I, _I, C.POINTER(_I)]),
    "b2u_unit_bwd_stats": (_I, [C.POINTER(UnitBwdDesc), _P, _P]),
    "b2u_unit_bwd_finalize": (_I, [_P, _I, _I, _I, _I, _P, _D, _P, _P, _P, _P, _P]),
    "b2u_unit_bwd_apply": (_I, [C.POINTER(UnitBwdDesc), _P, _P, _P]),
    "b2u_wgrad_workspace_floats": (_I, [C.POINTER(WgradDesc), C.POINTER(_LL)]),
    "b2u_wgrad": (_I, [_P, _P, _P, _P, C.POINTER(WgradDesc), _P]),
    "b2u_wgrad_first": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b2u_wgrad_first_rows": (_I, []),
    "b2u_gemm1x1_fwd": (_I, [_P, _P, _P, C.POINTER(ConvDesc), _P]),
    "b2u_pack_convT2x2_dgrad_weight": (_I, [_P, _P, _I, _I, _I, _P]),
    "b2u_gn_apply": (_I, [_P, _P, _P, _P, _P, _P, C.POINTER(ApplyDesc), _P]),
    "b2u_pool_stat_layout": (_I, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "b2u_gn_apply_pool": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, C.POINTER(ApplyDesc), _P]),
    "b2u_head_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(HeadDesc), _P]),
    "b2u_head_plan": (_I, [_I, _I, _I, C.POINTER(_I)]),
    "b2u_mc_finalize": (_I, [_P, _P, _P, _LL, _LL, _P]),
    "b2u_mc_accumulate": (_I, [_P, _P, _P, _P, _P, _I, _LL, _I, _P]),
    "b2u_advance_counter": (_I, [_P, _LL, _P]),
    "b2u_dropblock_centers": (_I, [_P, _I, C.c_uint64, _P, _P, _P]),
    "b2u_dropblock_centers_ex": (_I, [_P, _I, C.POINTER(DropblockCall), C.c_uint64, _P, _P, _P]),
    "b2u_dropblock_centers_ichan": (_I, [_P, _I, C.POINTER(DropblockCall), C.c_uint64, _P, _P, _LL, _P]),
    "b2u_dropblock_plan": (_I, [C.POINTER(DropblockCall), _I, C.POINTER(_LL)]),
    "b2u_dropblock_dilate": (_I, [_P, _I, C.POINTER(DropblockCall), _P, _P, _P, _P]),
    "b2u_dropblock_dilate_v2": (_I, [_P, _I, C.POINTER(DropblockCall), _P, _P, _LL, _P, _P, _P]),
    "b2u_dropblock_centers_from_uniform": (_I, [_P, _P, _LL, _F, _P]),
    "b2u_rotate_bilinear": (_I, [_P, _P, _I, _I, _I, _I, C.POINTER(_D), _I, _P]),
    "b2u_masked_bce_blocks": (_I, []),
    "b2u_masked_bce_fwd": (_I, [_P, _P, _P, _LL, _P, _P, _P, _P, _P]),
    "b2u_masked_bce_bwd": (_I, [_P, _P, _P, _P, _LL, _P]),
    "b2u_rotation_table": (_I, [C.POINTER(_D), _I, _I, _I, C.POINTER(_F)]),
    "b2u_rotate_in_table": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "b2u_rotate_back_accumulate": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "b2u_square_pad_resize": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b2u_square_pad_resize_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b2u_confusion_counts": (_I, [_P, _P, _P, _LL, _P, _P]),
    "b2u_sgd_step": (_I, [_P, _P, _I, _P, _F, _F, _F, _I, _P, _P]),
    "b2u_sgd_chunk_elems": (_LL, []),
}

_lib: Optional[C.CDLL] = None
launch_count = 0          # kernels launched through this binding (bench.py reports it as gpu_launches)
_LAUNCHERS = {"b2u_conv3x3_fwd": 1, "b2u_conv3x3_pro_fwd": 1, "b2u_convT2x2_fwd": 1, "b2u_conv_first_fwd": 1, "b2u_gn_finalize": 1, "b2u_gn_finalize_ex": 1,
              "b2u_gn_apply": 1, "b2u_gn_apply_pool": 1, "b2u_head_fwd": 1, "b2u_mc_finalize": 1,
              "b2u_mc_accumulate": 1, "b2u_advance_counter": 1, "b2u_dropblock_centers": 1, "b2u_dropblock_centers_ex": 1,
              "b2u_dropblock_dilate": 1, "b2u_dropblock_dilate_v2": 2, "b2u_dropblock_centers_ichan": 1, "b2u_dropblock_centers_from_uniform": 1, "b2u_rotate_bilinear": 1, "b2u_masked_bce_fwd": 2, "b2u_masked_bce_bwd": 1, "b2u_rotate_in_table": 1, "b2u_rotate_back_accumulate": 1,
              "b2u_pack_conv3x3_weight": 1, "b2u_pack_conv3x3_weight_pair": 1, "b2u_pack_batched": 1, "b2u_pack_convT2x2_weight": 1, "b2u_unit_bwd_stats": 1, "b2u_unit_bwd_finalize": 2,
              "b2u_unit_bwd_apply": 1, "b2u_wgrad": 2, "b2u_wgrad_first": 2, "b2u_gemm1x1_fwd": 1,
              "b2u_pack_convT2x2_dgrad_weight": 1, "b2u_sgd_step": 2, "b2u_confusion_counts": 1, "b2u_square_pad_resize": 1, "b2u_square_pad_resize_bwd": 1}


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2uError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                       "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.b2u_version() != 1:
        raise B2uError(f"libb2u.so version {lib.b2u_version()} does not match the binding (1)")
    _lib = lib
    return lib


# TIMING DIAGNOSTIC ONLY (upper-bound experiments on captured steps, tests/exp_train_skip.py): entry points named in
# B2U_EXP_SKIP_CALLS are not launched -- their outputs are stale, results are garbage, only the step time means anything.
_SKIP = frozenset(t for t in os.environ.get("B2U_EXP_SKIP_CALLS", "").split(",") if t)


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise B2uError with the library's message on failure."""
    global launch_count
    lib = load()
    if _SKIP and name in _SKIP:
        return
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B2uError(f"{name} failed (code {rc}): {lib.b2u_last_error().decode(errors='replace')}")
    launch_count += _LAUNCHERS.get(name, 0)


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
