"""ctypes binding of libunetk.so (the C ABI declared in include/unetk.h).

There is deliberately NO fallback: if the shared object is missing, or a call returns an error, this
module raises.  A silent eager-PyTorch path would void every parity and performance claim.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
# UNETK_LIB: another build of the same ABI (A/B timing of kernel variants on one GPU box); default = the in-tree build
LIB_PATH = Path(os.environ["UNETK_LIB"]).resolve() if os.environ.get("UNETK_LIB") else PKG / "libunetk.so"
HEADER = PKG.parent / "include" / "unetk.h"

_lib = None

_vp, _i, _i64, _sz, _fp, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_void_p, C.c_float

# name -> (restype, argtypes); must mirror include/unetk.h exactly (tests/test_abi.py checks the set).
SIGNATURES = {
    "unetk_abi_version": (_i, []),
    "unetk_last_error": (C.c_char_p, []),
    "unetk_launch_count": (_i64, []),
    "unetk_set_sm_limit": (_i, [_i]),
    "unetk_device_sms": (_i, []),
    "unetk_pack_weight": (_i, [_fp, _vp, _vp, _i, _i, _i, _vp]),
    "unetk_pack_tiles": (_i64, [_i, _i]),
    "unetk_pack_weights": (_i, [_vp, _i, _i64, _vp]),
    "unetk_conv3x3_fwd": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv_stats_partial_floats": (_sz, [_i]),
    "unetk_conv3x3_fwd_bnstats": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv3x3_dgrad": (_i, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv3x3_dgrad_cols": (_i, [_vp, _i64, _vp, _i, _i, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv3x3_dgrad_colsum": (_i, [_vp, _i64, _vp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_sums_to_f32": (_i, [_vp, _i, _fp, _i, _vp]),
    "unetk_conv3x3s2_fwd": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv3x3s2_dgrad": (_i, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv3x3s2_wgrad": (_i, [_vp, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_conv_wgrad_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "unetk_conv3x3_wgrad": (_i, [_vp, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_conv1x1_fwd": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv1x1_fwd_bnstats": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv1x1_dgrad": (_i, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_conv1x1_wgrad": (_i, [_vp, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_pack_upconv_weight": (_i, [_fp, _vp, _vp, _i, _i, _vp]),
    "unetk_upconv3x3_fwd": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_upconv3x3_fwd_affine": (_i, [_vp, _i64, _vp, _fp, _fp, _i, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_upconv3x3_dgrad": (_i, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_upconv_wgrad_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "unetk_upconv3x3_wgrad": (_i, [_vp, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_convT2x2_fwd": (_i, [_vp, _i64, _vp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_convT2x2_dgrad": (_i, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_convT2x2_wgrad": (_i, [_vp, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_stem_conv3x3_fwd": (_i, [_fp, _i64, _i64, _i64, _i64, _fp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_stem_stats_partial_floats": (_sz, [_i, _i, _i, _i]),
    "unetk_stem_conv3x3_fwd_bnstats": (_i, [_fp, _i64, _i64, _i64, _i64, _fp, _fp, _vp, _i64, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_stem_wgrad_workspace": (_sz, [_i, _i, _i, _i]),
    "unetk_stem_conv3x3_wgrad": (_i, [_fp, _i64, _i64, _i64, _i64, _vp, _i64, _fp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "unetk_chan_partial_floats": (_sz, [_i64, _i]),
    "unetk_bn_stats": (_i, [_vp, _i64, _i64, _i, _fp, _vp, _vp]),
    "unetk_bn_finalize": (_i, [_vp, _i, C.c_double, _fp, _fp, _f, _f, _fp, _fp, _vp, _fp, _fp, _fp, _fp, _vp]),
    "unetk_bn_eval_fold": (_i, [_i, _fp, _fp, _f, _fp, _fp, _fp, _fp, _fp, _fp, _vp]),
    "unetk_bn_eval_fold_bias": (_i, [_i, _fp, _fp, _f, _fp, _fp, _fp, _fp, _fp, _vp]),
    "unetk_conv3x3_fwd_affine": (_i, [_vp, _i64, _vp, _fp, _fp, _i, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_stem_conv3x3_fwd_affine": (_i, [_fp, _i64, _i64, _i64, _i64, _fp, _fp, _fp, _i, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_apply": (_i, [_vp, _i64, _fp, _fp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_apply_copies": (_i, [_vp, _i64, _fp, _fp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_bwd_reduce": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_bwd_apply": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _vp, C.c_double, _fp, _fp, _i,
                                _fp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_bwd_apply_res": (_i, [_vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _vp, C.c_double, _fp, _fp, _i, _fp, _fp, _vp, _i64, _i,
                               _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_bn_bwd_coef": (_i, [_vp, _i, C.c_double, _fp, _fp, _fp, _fp, _fp, _i, _fp, _fp, _vp]),
    "unetk_maxpool2x2_fwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _vp]),
    "unetk_maxpool2x2_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_maxpool2x2_fwd_codes": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _vp]),
    "unetk_max_unpool2x2": (_i, [_vp, _i64, _vp, _i, _vp, _i64, _i, _i, _i, _i, _vp]),
    "unetk_max_unpool2x2_bwd": (_i, [_vp, _i64, _vp, _i, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_shift_copy": (_i, [_vp, _i64, _i, _i, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "unetk_colsum": (_i, [_vp, _i64, _i64, _i, _fp, _fp, _i, _vp]),
    "unetk_head_partial_floats": (_sz, [_i64, _i]),
    "unetk_head_fwd": (_i, [_vp, _i64, _fp, _fp, _fp, _fp, _i, _i64, _i, _fp, _vp, _vp]),
    "unetk_loss_finalize": (_i, [_vp, C.c_double, _fp, _vp]),
    "unetk_bn_head_partial_floats": (_sz, [_i64, _i]),
    "unetk_bn_head_fwd": (_i, [_vp, _i64, _fp, _fp, _i, _fp, _fp, _fp, _fp, _i, _i64, _i, _fp, _vp, _vp]),
    "unetk_bn_head_bwd_reduce": (_i, [_vp, _i64, _fp, _fp, _fp, _i, _fp, _fp, _fp, _fp, _fp, _f, _i, _fp, _fp, _fp, _i,
                                      _vp, _i64, _i, _fp, _vp]),
    "unetk_bn_head_bwd_apply": (_i, [_vp, _i64, _fp, _fp, _i, _fp, _fp, _fp, _vp, _i64, _i64, _i, _vp]),
    "unetk_head_bwd": (_i, [_vp, _i64, _fp, _fp, _fp, _fp, _fp, _f, _i, _vp, _i64, _fp, _fp, _i, _i64, _i, _fp, _vp]),
    "unetk_sqnorm_partial_floats": (_sz, [_i64]),
    "unetk_grad_clip_coef": (_i, [_fp, _i64, _f, _f, _fp, _fp, _vp]),
    "unetk_rmsprop_step": (_i, [_fp, _fp, _fp, _fp, _i64, _f, _f, _f, _f, _f, _fp, _vp]),
    "unetk_rmsprop_step_dev": (_i, [_fp, _fp, _fp, _fp, _i64, _fp, _fp, _vp]),
    "unetk_add_n": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "unetk_upsample_nearest2x_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp]),
    "unetk_upsample_nearest2x_bwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_upsample_bilinear2x_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp]),
    "unetk_upsample_bilinear2x_bwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_gather_patches": (_i, [_fp, _i64, _i64, _i64, _i64, _fp, _i64, _i64, _i64, _vp, _i, _i, _i, _i, _i, _fp, _fp, _vp]),
    "unetk_tile_accumulate": (_i, [_fp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "unetk_tile_finalize": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "unetk_copy_f32_strided": (_i, [_fp, _i64, _fp, _i64, _i64, _i, _vp]),
    "unetk_gate_partial_floats": (_sz, [_i64, _i]),
    "unetk_gate_fwd": (_i, [_vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _i64, _i, _vp]),
    "unetk_gate_apply": (_i, [_vp, _i64, _fp, _fp, _fp, _vp, _i64, _i64, _i, _vp]),
    "unetk_gate_bwd_psi": (_i, [_vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _vp, _i64, _i, _fp, _fp, _vp, _i64, _i, _vp]),
    "unetk_gate_bwd_reduce": (_i, [_vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                   _vp, _vp, _fp, _fp, _i, _i64, _i, _vp]),
    "unetk_gate_bwd_apply": (_i, [_vp, _i64, _vp, _i64, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp,
                                  _i64, _vp, _i64, _i64, _i, _vp]),
    "unetk_head_multi_partial_floats": (_sz, [_i64, _i, _i]),
    "unetk_head_multi_fwd": (_i, [_vp, _i64, _fp, _fp, _fp, _i, _i64, _i, _i, _vp]),
    "unetk_head_multi_bwd": (_i, [_vp, _i64, _fp, _fp, _f, _vp, _i64, _fp, _fp, _i, _i, _i64, _i, _i, _fp, _vp]),
    "unetk_dice_partial_floats": (_sz, [_i64, _i64]),
    "unetk_dice_sums": (_i, [_fp, _fp, _i64, _i64, _f, _f, _fp, _vp, _vp]),
    "unetk_dice_bwd": (_i, [_fp, _fp, _fp, _fp, _i64, _i64, _f, _f, _fp, _vp]),
    "unetk_f32_pack_split3": (_i, [_fp, _vp, _i64, _i64, _i64, _i, _i, _i, _vp, _i, _vp]),
    "unetk_f32_stem_conv3x3": (_i, [_fp, _i64, _i64, _i64, _i64, _fp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_f32_conv3x3": (_i, [_vp, _i64, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_f32_convT2x2": (_i, [_vp, _i64, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_f32_stats_partial_doubles": (_sz, [_i64, _i]),
    "unetk_f32_stats": (_i, [_fp, _i64, _i64, _i, _vp, _vp, _vp]),
    "unetk_f32_bn_split": (_i, [_fp, _i64, _fp, _fp, _vp, _i64, _fp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "unetk_f32_head": (_i, [_fp, _i64, _fp, _fp, _fp, _i64, _i, _vp]),
}


def header_symbols() -> list[str]:
    """Every function name declared in include/unetk.h."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(unetk_[a-zA-Z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m jcfszxc_unet_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / eager fallback for the U-Net hot path."
        )
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.unetk_abi_version()
    if got != 3:
        raise RuntimeError(f"libunetk.so ABI version {got}, expected 3")
    _lib = lib
    return lib


class UnetkError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().unetk_last_error()
        raise UnetkError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


_profile: list | None = None  # when a list: (name, args, start_event, end_event) per C-ABI call


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on error."""
    if _profile is None:
        check(getattr(load(), name)(*args), name)
        return
    import torch

    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    check(getattr(load(), name)(*args), name)
    end.record()
    _profile.append((name, args, start, end))


class profile_calls:
    """Context manager: time every C-ABI call with CUDA events on the launching (current) stream."""

    def __enter__(self):
        global _profile
        self.records = []
        _profile = self.records
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None
        return False

    def summary(self):
        """name -> (calls, total_ms); call after torch.cuda.synchronize()."""
        out: dict[str, list] = {}
        for name, _args, s, e in self.records:
            slot = out.setdefault(name, [0, 0.0])
            slot[0] += 1
            slot[1] += s.elapsed_time(e)
        return out
