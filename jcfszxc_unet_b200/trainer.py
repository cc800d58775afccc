"""Fused training step — the restatement of the reference's hot loop body (train.py:255-301):

    forward -> 0.5*BCEWithLogits + 0.5*dice_loss -> backward -> clip_grad_norm_(1.0) -> RMSprop

as CUDA-graph segments with the data-parallel collectives between them:

    [pack weights, forward, head+loss sums]  --all-reduce(loss sums)-->
    [loss finalize, backward of the decoder] --async all-reduce(decoder grads) on the NCCL stream ...
    [backward of the encoder]                --all-reduce(encoder grads), wait for both-->
    [global-norm clip coefficient, RMSprop]

The gradient all-reduce is bucketed in backward order (the decoder's gradients are the contiguous tail of the
flat buffer and are complete half-way through the backward), so NCCL runs beside the encoder's backward kernels.

Parameters, gradients and RMSprop state live in four flat fp32 buffers (the model's nn.Parameters are
re-homed as views, so state_dict()/checkpoints keep working).  No host synchronisation happens inside
step(); the loss is returned as a device scalar.  bf16 needs no GradScaler (the reference's
GradScaler/fp16 autocast is replaced by bf16 as BASELINE.json asks).
"""
from __future__ import annotations

import os

import torch

from . import engine, ops
from .dp import DataParallel


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-6, weight_decay: float = 1e-8,
                 momentum: float = 0.999, alpha: float = 0.99, eps: float = 1e-8, max_norm: float = 1.0,
                 builder=None, use_cuda_graph: bool = True, dp: DataParallel | None = None):
        self.model = model
        self.lr, self.wd, self.momentum, self.alpha, self.eps, self.max_norm = lr, weight_decay, momentum, alpha, eps, max_norm
        self.builder = builder or engine.build_unet_plan
        self.dp = dp or DataParallel()
        self.use_graph = use_cuda_graph and not self.dp.sync_bn  # SyncBN puts collectives between kernels
        params = [p for p in model.parameters()]
        if not params or not params[0].is_cuda:
            raise RuntimeError("Trainer: move the model to a CUDA device first (no CPU fallback on this path)")
        dev = params[0].device
        self.device = dev
        total = sum(p.numel() for p in params)
        # 16-byte aligned slices so vector kernels can run on any per-tensor view
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.sq = torch.zeros_like(self.flat_p)
        self.buf = torch.zeros_like(self.flat_p)
        self.grad_views = {}
        with torch.no_grad():
            for p, o in zip(params, offs):
                v = self.flat_p[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.grad_views[id(p)] = self.flat_g[o:o + p.numel()].view(p.shape)
        self.n_params = total
        if self.dp.enabled:
            self.dp.broadcast_(self.flat_p, 0)  # replicas start identical
        self.clip = torch.zeros(2, dtype=torch.float32, device=dev)
        self.sq_partial = None
        self.plan = None
        self.graphs = None
        self._single_graph = False
        self.images = self.labels = None
        self.steps_done = 0
        # input pipeline (prefetch / step()): two staging batches in HBM filled by a copy stream
        self._copy_stream = None
        self._stage = None          # [(images, labels)] x 2
        self._stage_ready = self._stage_free = None
        self._staged = []           # slots holding a batch not yet consumed, in order
        self._stage_next = 0

    # ------------------------------------------------------------------------------------------------
    def _build(self, n, c, h, w):
        dev = self.device
        self.plan = self.builder(self.model, n, h, w, dev, True, self.grad_views, True)
        self.images = torch.zeros((n, c, h, w), dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last)
        self.labels = torch.zeros((n, 1, h, w), dtype=torch.float32, device=dev)
        head = self.plan.head
        from .engine import Head
        if not isinstance(head, Head):
            raise NotImplementedError("Trainer fuses the reference's binary loss (train.py:264-278: BCE + dice on ONE logit map); "
                                      "with n_classes > 1 use model(x), your loss and loss.backward()")
        head.labels = self.labels
        head.dlogits = None
        head.auto_finalize = False
        if self.dp.sync_bn:
            self.plan.sync_sums = self.dp.reduce_bn_sums
        from . import _lib
        self.sq_partial = torch.empty(_lib.load().unetk_sqnorm_partial_floats(self.flat_g.numel()),
                                      dtype=torch.float32, device=dev)
        self._gscale = 1.0
        self._split_op, self._split_off = self._find_bucket_split()

    def _find_bucket_split(self):
        """(k, o): the ops[k:] — run FIRST in the backward — own exactly the parameters stored at flat offsets >= o,
        with about half of the gradient behind o.  (0, 0) when no such cut exists (single bucket)."""
        base = self.flat_g.data_ptr()
        ranges = []
        for op in self.plan.ops:
            lo, hi = None, None
            for q in self.plan.op_params(op):
                g = self.grad_views.get(id(q))
                if g is None:
                    continue
                o = (g.data_ptr() - base) // 4
                lo = o if lo is None else min(lo, o)
                hi = o + g.numel() if hi is None else max(hi, o + g.numel())
            ranges.append((lo, hi))
        total = self.flat_g.numel()
        n = len(ranges)
        prefix_hi = [0] * (n + 1)            # max offset end owned by ops[:k]
        for k in range(n):
            prefix_hi[k + 1] = max(prefix_hi[k], ranges[k][1] or 0)
        suffix_lo = [total] * (n + 1)        # min offset start owned by ops[k:]
        for k in range(n - 1, -1, -1):
            suffix_lo[k] = min(suffix_lo[k + 1], ranges[k][0] if ranges[k][0] is not None else total)
        best = (0, 0)
        for k in range(1, n):
            if prefix_hi[k] <= suffix_lo[k] < total and suffix_lo[k] > 0:
                if best == (0, 0) or abs(suffix_lo[k] - total // 2) < abs(best[1] - total // 2):
                    best = (k, int(suffix_lo[k]))
        return best

    # segments ---------------------------------------------------------------------------------------
    def _seg_forward(self):
        self.plan.forward(self.images)

    def _seg_backward(self):
        self.plan.head.finalize_loss(self._npix_total)
        self.plan.backward(self._split_op, None)

    def _seg_backward_tail(self):
        self.plan.backward(0, self._split_op)

    def _seg_optim(self):
        ops.grad_clip_coef(self.flat_g, self._gscale, self.max_norm, self.sq_partial, self.clip)
        ops.rmsprop_step(self.flat_p, self.flat_g, self.sq, self.buf, self.lr, self.alpha, self.eps, self.wd,
                         self.momentum, self.clip)

    def _seg_all(self):
        """The whole iteration as one launch sequence: single process, nothing to exchange between the segments."""
        self.plan.refresh_weights(force=True)
        self._seg_forward()
        self.plan.head.finalize_loss(self._npix_total)
        self.plan.backward(0, None)
        self._seg_optim()

    def _run_segments(self):
        head = self.plan.head
        if self.graphs is not None and self._single_graph:
            # one CUDA graph per step: the weight-gradient side stream is joined once, in front of the optimizer,
            # instead of at the end of each backward segment
            self.graphs[0].replay()
            return
        if self.graphs is not None:
            self.graphs[0].replay()
        else:
            self.plan.refresh_weights(force=True)
            self._seg_forward()
        if self.dp.sync_loss:
            self.dp.reduce_loss_sums(head.loss_sums, head.npix)
        if self.graphs is not None:
            self.graphs[1].replay()
        else:
            self._seg_backward()
        # decoder gradients (flat tail) are final: reduce them while the encoder's backward runs
        h_tail = self.dp.reduce_grads_async(self.flat_g[self._split_off:]) if self._split_op > 0 else None
        if self._split_op > 0:
            if self.graphs is not None:
                self.graphs[2].replay()
            else:
                self._seg_backward_tail()
        h_head = self.dp.reduce_grads_async(self.flat_g[:self._split_off] if self._split_op > 0 else self.flat_g)
        self.dp.wait(h_tail)
        self.dp.wait(h_head)
        if self.graphs is not None:
            self.graphs[3].replay()
        else:
            self._seg_optim()

    def _capture(self):
        # constants baked into the graphs
        graphs = []
        pool = None
        # captured on a HIGH-priority stream: the plan's weight-gradient side stream is low priority, so whenever a
        # kernel of the critical path and a weight gradient are both ready, the critical path gets the SMs first
        hp = torch.cuda.Stream(device=self.device, priority=-1)
        self._single_graph = not self.dp.enabled and os.environ.get("UNETK_SINGLE_GRAPH", "1") != "0"
        segments = ((self._seg_all,) if self._single_graph else
                    (self._seg_forward_packed, self._seg_backward, self._seg_backward_tail, self._seg_optim))
        for seg in segments:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=hp):
                seg()
            pool = g.pool()
            graphs.append(g)
        self.graphs = graphs

    def _seg_forward_packed(self):
        self.plan.refresh_weights(force=True)  # the optimizer updates weights behind torch's version counter
        self._seg_forward()

    # ------------------------------------------------------------------------------------------------
    def step(self, images: torch.Tensor | None = None, labels: torch.Tensor | None = None) -> torch.Tensor:
        """One training iteration on this rank's shard.  images [N,C,H,W], labels [N,1,H,W] (host or device); with no
        arguments: the batch handed to prefetch().  Returns the loss as a 0-dim device tensor (no sync)."""
        if images is None:
            self._take_staged()
            if self.use_graph and self.graphs is None and self.steps_done >= 1:
                torch.cuda.synchronize()
                self._capture()
            self._run_segments()
            self.steps_done += 1
            return self.plan.head.fin[0]
        n, c, h, w = images.shape
        if self.plan is None:
            self._build(n, c, h, w)
            head = self.plan.head
            self._npix_total = head.npix * (self.dp.world if self.dp.sync_loss else 1)
            self._gscale = 1.0 if (self.dp.sync_loss or not self.dp.enabled) else 1.0 / self.dp.world
        elif tuple(self.images.shape) != (n, c, h, w):
            raise ValueError(f"Trainer was built for input {tuple(self.images.shape)}, got {(n, c, h, w)}")
        self.images.copy_(images, non_blocking=True)
        self.labels.copy_(labels.reshape(self.labels.shape), non_blocking=True)
        if self.use_graph and self.graphs is None and self.steps_done >= 1:
            # step 0 ran eagerly (lazy one-time initialisation inside the library); capture now
            torch.cuda.synchronize()
            self._capture()
        self._run_segments()
        self.steps_done += 1
        return self.plan.head.fin[0]

    # ------------------------------------------------------------------------------------------------
    def prefetch(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Start the host->device copy of a FUTURE batch on a side stream (pinned host tensors make it truly
        asynchronous); the next `step()` without arguments trains on it.  Issue it right after launching the current
        step: the 64 MB copy of a 16x3x512x512 batch then hides behind the ~24 ms of compute instead of preceding it
        (the reference copies synchronously before every step, train.py:244-253)."""
        n, c, h, w = images.shape
        if self.plan is None:
            raise RuntimeError("prefetch: run one step(images, labels) first (it builds the plan for this input shape)")
        if tuple(self.images.shape) != (n, c, h, w):
            raise ValueError(f"Trainer was built for input {tuple(self.images.shape)}, got {(n, c, h, w)}")
        if self._stage is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [(torch.empty_like(self.images), torch.empty_like(self.labels)) for _ in range(2)]
            self._stage_ready = [torch.cuda.Event() for _ in range(2)]
            self._stage_free = [torch.cuda.Event() for _ in range(2)]
            for e in self._stage_free:
                e.record(torch.cuda.current_stream(self.device))
        if len(self._staged) == 2:
            raise RuntimeError("prefetch: both staging batches are in flight; call step() first")
        k = self._stage_next
        cs = self._copy_stream
        cs.wait_event(self._stage_free[k])            # the step that last read this slot has copied it out
        with torch.cuda.stream(cs):
            self._stage[k][0].copy_(images, non_blocking=True)
            self._stage[k][1].copy_(labels.reshape(self.labels.shape), non_blocking=True)
            self._stage_ready[k].record(cs)
        self._staged.append(k)
        self._stage_next ^= 1

    def _take_staged(self):
        if not self._staged:
            raise RuntimeError("step() without arguments needs a batch from prefetch()")
        k = self._staged.pop(0)
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._stage_ready[k])
        self.images.copy_(self._stage[k][0], non_blocking=True)      # device-to-device, ~0.05 ms
        self.labels.copy_(self._stage[k][1], non_blocking=True)
        self._stage_free[k].record(main)

    def loss_terms(self):
        """(loss, bce, dice) of the last step as device scalars."""
        f = self.plan.head.fin
        return f[0], f[1], f[2]

    def grad_norm(self):
        return self.clip[0]
