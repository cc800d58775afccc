"""Dice coefficient / loss on the GPU — drop-in for the reference's `utils/dice_score.py` (same import path
`from utils.dice_score import dice_coeff, dice_loss`, same signatures: dice_score.py:13-59), computed by the
hand-written reduction kernels of libunetk.so (`unetk_dice_sums`, `unetk_dice_bwd`, csrc/multiclass.cu).

Semantics restated from the reference:
  * dice_coeff clamps the prediction to [0, 1], sums over (H, W) per leading index — or over everything when
    `reduce_batch_first` (3-D input) / 2-D input — forms (2*I + eps) / (S + eps) with eps FORCED to 1e-5 whatever the
    argument says (dice_score.py:32), replaces S by 2*I when S < eps (empty mask, :35) and returns the mean;
  * multiclass_dice_coeff flattens (batch, class) into one axis first (:47-49);
  * dice_loss clamps the prediction to [1e-7, 1 - 1e-7] and returns 1 - dice with reduce_batch_first=True (:53-59).
Differentiable w.r.t. the prediction (torch.clamp passes the gradient on the closed interval).  CUDA tensors only:
there is no CPU fallback on this path.
"""
from __future__ import annotations

import torch
from torch import Tensor

from jcfszxc_unet_b200 import _lib

_EPS = 1e-5   # dice_score.py:32


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _DiceMean(torch.autograd.Function):
    """mean over groups of (2*sum(p*t) + eps) / (sum(p) + sum(t) + eps), p = clamp(inp, lo, hi); inp, target [G, n]."""

    @staticmethod
    def forward(ctx, inp: Tensor, target: Tensor, lo: float, hi: float):
        groups, n = inp.shape
        lib = _lib.load()
        partial = torch.empty(max(lib.unetk_dice_partial_floats(groups, n), 1), dtype=torch.float32, device=inp.device)
        sums = torch.empty((groups, 3), dtype=torch.float64, device=inp.device)
        _lib.call("unetk_dice_sums", inp.data_ptr(), target.data_ptr(), groups, n, lo, hi, partial.data_ptr(),
                  sums.data_ptr(), _stream())
        inter = 2.0 * sums[:, 0]
        sets = sums[:, 1] + sums[:, 2]
        empty = sets < _EPS
        sets = torch.where(empty, inter, sets)
        dice = (inter + _EPS) / (sets + _EPS)
        # d dice_g / d p_i = 2 t_i / (S+eps) - (I+eps) / (S+eps)^2 ; zero on the empty branch (dice == 1 there)
        ca = torch.where(empty, torch.zeros_like(sets), 2.0 / (sets + _EPS)) / groups
        cb = torch.where(empty, torch.zeros_like(sets), -(inter + _EPS) / (sets + _EPS) ** 2) / groups
        ctx.save_for_backward(inp, target, torch.stack([ca, cb], dim=1).float().contiguous())
        ctx.bounds = (lo, hi)
        return dice.mean().float()

    @staticmethod
    def backward(ctx, gout: Tensor):
        inp, target, coef = ctx.saved_tensors
        lo, hi = ctx.bounds
        groups, n = inp.shape
        dp = torch.empty_like(inp)
        g = gout.detach().float().reshape(1).contiguous()
        _lib.call("unetk_dice_bwd", inp.data_ptr(), target.data_ptr(), coef.data_ptr(), g.data_ptr(), groups, n, lo, hi,
                  dp.data_ptr(), _stream())
        return dp, None, None, None


def _dice(inp: Tensor, target: Tensor, reduce_batch_first: bool, lo: float, hi: float) -> Tensor:
    if inp.size() != target.size():
        raise AssertionError("dice_coeff: prediction and target must have the same size")
    if not (inp.dim() == 3 or not reduce_batch_first):
        raise AssertionError("dice_coeff: reduce_batch_first needs a 3-D input")
    if not inp.is_cuda:
        raise RuntimeError("utils.dice_score runs on the B200-native path (libunetk.so); it has no CPU fallback")
    if inp.dim() < 2:
        raise ValueError("dice_coeff: expected at least 2 dimensions")
    whole = inp.dim() == 2 or reduce_batch_first
    n = inp.numel() if whole else inp.shape[-1] * inp.shape[-2]
    groups = 1 if whole else inp.numel() // max(n, 1)
    dtype = inp.dtype
    p = inp.float().contiguous().view(groups, n)
    t = target.float().contiguous().view(groups, n)
    out = _DiceMean.apply(p, t, lo, hi)
    return out.to(dtype) if dtype.is_floating_point else out


def dice_coeff(input: Tensor, target: Tensor, reduce_batch_first: bool = False, epsilon: float = 1e-6):
    return _dice(input, target, reduce_batch_first, 0.0, 1.0)


def multiclass_dice_coeff(input: Tensor, target: Tensor, reduce_batch_first: bool = False, epsilon: float = 1e-5):
    return _dice(input.flatten(0, 1), target.flatten(0, 1), reduce_batch_first, 0.0, 1.0)


def dice_loss(input: Tensor, target: Tensor, multiclass: bool = False):
    # clamp(1e-7, 1-1e-7) followed by dice_coeff's clamp(0, 1) is the single clamp to [1e-7, 1-1e-7]
    if multiclass:
        input, target = input.flatten(0, 1), target.flatten(0, 1)
    return 1 - _dice(input, target, True, 1e-7, 1.0 - 1e-7)
