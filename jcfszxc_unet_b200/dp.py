"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box,
gloo in the CPU tests).  The reference is single-device, so the semantics to preserve
are the ones that fall out of the maths of its single-device step (train.py:255-301):

  * the batch is sharded, every rank holds a full replica, no activation ever crosses ranks;
  * the loss is ONE mean BCE and ONE dice ratio over the WHOLE batch (utils/dice_score.py:13-38 called
    with a 3-D input), so the four loss sums are all-reduced before the backward and each rank
    back-propagates the global loss restricted to its pixels -> gradients are SUM-reduced, not averaged;
  * clip_grad_norm_ (train.py:299) sees the reduced gradient, hence the same norm on every rank;
  * BatchNorm statistics are per-rank by default (DDP convention); sync_bn=True all-reduces the
    per-channel (sum, sum-of-squares) pairs, which reproduces the single-device statistics exactly.

Everything here works on plain torch tensors and is device-agnostic so the world_size=2 gloo tests
exercise exactly this code.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, group=None, sync_bn: bool = False, sync_loss: bool = True, enabled: bool | None = None):
        """enabled=False: a single-process instance inside an initialised process group (reference runs in tests)."""
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if enabled is not None:
            self.enabled = self.enabled and enabled
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.sync_bn = sync_bn and self.enabled
        self.sync_loss = sync_loss and self.enabled

    # ---- sharding -------------------------------------------------------------------------------
    def shard(self, global_batch: int) -> tuple[int, int]:
        """Rank r owns images [lo, hi) of the global batch (even split required)."""
        if global_batch % self.world != 0:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {self.world}")
        per = global_batch // self.world
        return self.rank * per, (self.rank + 1) * per

    # ---- collectives ----------------------------------------------------------------------------
    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def reduce_loss_sums(self, sums: torch.Tensor, local_pixels: int) -> int:
        """sums = [sum bce, sum p*y, sum p, sum y] of this rank -> global sums (in place); returns the
        global pixel count the mean BCE divides by."""
        if self.sync_loss:
            self.all_reduce_sum_(sums)
            return local_pixels * self.world
        return local_pixels

    def reduce_bn_sums(self, sums: torch.Tensor, local_count: int) -> int:
        """sums = per-channel [2, C] statistics of this rank -> global (in place) when sync_bn."""
        if self.sync_bn:
            self.all_reduce_sum_(sums)
            return local_count * self.world
        return local_count

    def reduce_grads(self, flat_grad: torch.Tensor) -> float:
        """Sum-reduce the flat gradient; returns the scale the optimizer must apply to it:
        1 when the loss sums were global (each rank back-propagated the global loss), else 1/world
        (per-rank losses -> DDP-style average)."""
        self.all_reduce_sum_(flat_grad)
        return 1.0 if (self.sync_loss or not self.enabled) else 1.0 / self.world

    def reduce_grads_async(self, bucket: torch.Tensor):
        """Start the SUM all-reduce of one gradient bucket on the communication stream and return its handle (None
        when single-process).  The caller keeps computing (the rest of the backward) and waits on the handle before
        the optimizer: NCCL over NVLink runs beside the backward kernels instead of after them."""
        if not self.enabled:
            return None
        return dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    @staticmethod
    def wait(handle):
        if handle is not None:
            handle.wait()

    def broadcast_(self, t: torch.Tensor, src: int = 0):
        if self.enabled:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def max_over_ranks(self, value: float, device) -> float:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())
