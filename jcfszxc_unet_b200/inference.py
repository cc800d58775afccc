"""Sliding-window inference with overlap averaging on the device.

Drop-in for the reference's `predict_full_image(model, device, image, patch_size, overlap, batch_size)`
(evaluate.py:28-96): same argument meaning, same patch grid (step = int(patch_size * (1 - overlap)), positions
`range(0, h - patch_size + 1, step)`), same accumulation order, same `np.divide(..., where=count != 0)` ending, same
return value (float64 numpy [1, h, w]).  What changes is where it runs: the image is uploaded once, patches are cut by
`unetk_gather_patches`, the model's eval forward runs on the fused plan, and probabilities are accumulated into
double-precision maps in HBM by `unetk_tile_accumulate`; one D2H copy of the final mask replaces one per batch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def patch_positions(h: int, w: int, patch_size: int, overlap: float):
    step = int(patch_size * (1 - overlap))
    if step <= 0:
        raise ValueError("overlap must leave a positive step")
    return [(y, x) for y in range(0, h - patch_size + 1, step) for x in range(0, w - patch_size + 1, step)]


def predict_full_image(model, device, image, patch_size=256, overlap=0.5, batch_size=4):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("predict_full_image runs on the B200-native path (libunetk.so); it has no CPU fallback")
    lib = _lib.load()
    image = np.asarray(image)
    if image.ndim == 3:
        hwc = image                                   # HxWxC as the reference expects (evaluate.py:44-46)
    else:
        raise ValueError("expected an HxWxC image")
    h, w, c = hwc.shape
    if patch_size % 2:
        raise ValueError("patch_size must be even")
    positions = patch_positions(h, w, patch_size, overlap)
    acc = torch.zeros((h, w), dtype=torch.float64, device=device)
    cnt = torch.zeros((h, w), dtype=torch.float64, device=device)
    out = torch.zeros((1, h, w), dtype=torch.float64, device=device)
    if positions:
        pool = torch.from_numpy(np.ascontiguousarray(hwc)).to(device, torch.float32).contiguous().view(1, h, w, c)
        s = pool.stride()
        half = patch_size // 2
        stream = lambda: torch.cuda.current_stream().cuda_stream
        model.eval()
        with torch.no_grad():
            for i in range(0, len(positions), batch_size):
                chunk = positions[i:i + batch_size]
                b = len(chunk)
                centers = torch.tensor([[0, y + half, x + half] for y, x in chunk], dtype=torch.int32).to(device)
                pos = torch.tensor(chunk, dtype=torch.int32).to(device)
                batch = torch.empty((b, patch_size, patch_size, c), dtype=torch.float32, device=device)
                _lib.call("unetk_gather_patches", pool.data_ptr(), s[0], s[3], s[1], s[2], None, 0, 0, 0, centers.data_ptr(),
                          b, c, patch_size, h, w, batch.data_ptr(), None, stream())
                logits = model(batch.permute(0, 3, 1, 2))                  # [b,1,P,P]; channels_last view, no copy
                if logits.shape != (b, 1, patch_size, patch_size):
                    raise ValueError(f"model returned {tuple(logits.shape)}; the sliding window expects one class")
                logits = logits.float().contiguous()
                _lib.call("unetk_tile_accumulate", logits.data_ptr(), pos.data_ptr(), b, patch_size, h, w, 1,
                          acc.data_ptr(), cnt.data_ptr(), stream())
        _lib.call("unetk_tile_finalize", acc.data_ptr(), cnt.data_ptr(), h * w, out.data_ptr(), stream())
    return out.cpu().numpy()
