"""On-device training-batch assembly.

The reference builds every batch on the host: `np.random.randint` into the list of vessel pixels, a Python loop of
numpy slices, `np.stack`, `torch.from_numpy(...).to(device, channels_last)` — a synchronous H2D copy of the whole
batch per step (train.py:126-155, 200-253).  Here the image / label pools live in HBM and one kernel
(`unetk_gather_patches`) cuts the crops straight into the channels_last batch; only 12 bytes per sample (image index
and centre) cross PCIe.  The random draw is the reference's own (`np.random.randint(0, n_valid, batch_size)` on the
same filtered sample map), so a seeded run yields bit-identical batches.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def filtered_sample_map(masks: np.ndarray, patch_size: int):
    """train.py:136-152: coordinates of the mask pixels whose patch fits inside the image."""
    half = patch_size // 2
    _, width, height = masks.shape
    sm = np.where(masks != 0)
    ok = (sm[1] >= half) & (sm[1] < width - half) & (sm[2] >= half) & (sm[2] < height - half)
    return sm[0][ok], sm[1][ok], sm[2][ok]


class PatchSampler:
    def __init__(self, images, masks, labels, patch_size: int, device="cuda:0"):
        """images: [N, H, W, C] (the layout of the reference's HDF5 pool before its transpose(0,3,1,2), train.py:131);
        masks, labels: [N, H, W].  Arrays or tensors; they are moved to `device` once as fp32."""
        if patch_size % 2:
            raise ValueError("patch_size must be even (the reference cuts [c - P//2, c + P//2))")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PatchSampler is the on-device path (libunetk.so); it has no CPU fallback")
        _lib.load()
        masks_np = masks.cpu().numpy() if isinstance(masks, torch.Tensor) else np.asarray(masks)
        self.sample_map = filtered_sample_map(masks_np, patch_size)
        if len(self.sample_map[0]) == 0:
            raise ValueError("no mask pixel leaves room for a patch of this size")
        to_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)))
        self.images = to_t(images).to(self.device, torch.float32).contiguous()      # NHWC memory
        self.labels = to_t(labels).to(self.device, torch.float32).contiguous()      # [N, H, W]
        self.n, self.h, self.w, self.c = self.images.shape
        assert tuple(self.labels.shape) == (self.n, self.h, self.w) and masks_np.shape == (self.n, self.h, self.w)
        self.patch = patch_size

    def draw(self, batch_size: int, rng=np.random):
        """The reference's draw (train.py:203-211): indices into the filtered sample map."""
        r = rng.randint(0, len(self.sample_map[0]), batch_size)
        return np.stack([self.sample_map[0][r], self.sample_map[1][r], self.sample_map[2][r]], axis=1).astype(np.int32)

    def gather(self, centers: np.ndarray):
        """centers int32 [B, 3] = (image, x, y) -> (images fp32 [B,C,P,P] channels_last, labels fp32 [B,1,P,P])."""
        centers = np.ascontiguousarray(centers, dtype=np.int32)
        b, p, half = centers.shape[0], self.patch, self.patch // 2
        if (centers[:, 0].min() < 0 or centers[:, 0].max() >= self.n or centers[:, 1].min() < half
                or centers[:, 1].max() > self.h - half or centers[:, 2].min() < half or centers[:, 2].max() > self.w - half):
            raise ValueError("patch centre outside the valid range")
        dev_centers = torch.from_numpy(centers).to(self.device, non_blocking=True)
        out_i = torch.empty((b, p, p, self.c), dtype=torch.float32, device=self.device)
        out_l = torch.empty((b, 1, p, p), dtype=torch.float32, device=self.device)
        si = self.images.stride()   # (n, h, w, c) memory -> logical [N, C, H, W] strides
        sl = self.labels.stride()
        _lib.call("unetk_gather_patches", self.images.data_ptr(), si[0], si[3], si[1], si[2], self.labels.data_ptr(),
                  sl[0], sl[1], sl[2], dev_centers.data_ptr(), b, self.c, p, self.h, self.w, out_i.data_ptr(),
                  out_l.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out_i.permute(0, 3, 1, 2), out_l     # [B,C,P,P] with channels_last strides, as train.py:248-252

    def sample(self, batch_size: int, rng=np.random):
        return self.gather(self.draw(batch_size, rng))
