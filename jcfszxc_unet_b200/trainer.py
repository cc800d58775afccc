"""Fused training step — the restatement of the reference's hot loop body (train.py:255-301):

    forward -> 0.5*BCEWithLogits + 0.5*dice_loss -> backward -> clip_grad_norm_(1.0) -> RMSprop

as ONE CUDA graph per rank.  In data-parallel runs the collectives are part of that graph:

    [pack weights, forward, head + loss sums] -> all-reduce(4 loss sums) -> [loss finalize, backward ...
      ... after the op that completes a gradient bucket: all-reduce(bucket) on the NCCL stream, beside the rest of
          the backward ...] -> wait -> [global-norm clip coefficient, RMSprop]

The flat gradient is cut into ~8 buckets by byte count in the order the backward completes it (decoder first, stem
last), so what is still un-reduced when the backward ends is the last, small bucket (the first layers' parameters).
When the NCCL calls cannot be captured (older stacks) the same program runs as one graph per kernel segment with the
collectives launched between them; `use_cuda_graph=False` runs it eagerly.  All three orders are the same launch
sequence and give the same bits.

Parameters, gradients and RMSprop state live in four flat fp32 buffers (the model's nn.Parameters are re-homed as
views, so state_dict()/checkpoints keep working).  The learning rate and the other RMSprop hyper-parameters are read
by the kernel from device memory (`set_lr`), so a schedule (the reference's ReduceLROnPlateau, train.py:114-122,355)
takes effect inside captured graphs too.  No host synchronisation happens inside step(); the loss is returned as a
device scalar.  bf16 needs no GradScaler (the reference's GradScaler/fp16 autocast is replaced by bf16 as
BASELINE.json asks).
"""
from __future__ import annotations

import functools
import os
import sys

import torch

from . import engine, ops
from .dp import DataParallel


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-6, weight_decay: float = 1e-8,
                 momentum: float = 0.999, alpha: float = 0.99, eps: float = 1e-8, max_norm: float = 1.0,
                 builder=None, use_cuda_graph: bool = True, dp: DataParallel | None = None,
                 grad_buckets: int | None = None):
        self.model = model
        self._max_norm = max_norm
        self.builder = builder or engine.build_unet_plan
        self.dp = dp or DataParallel()
        self.use_graph = use_cuda_graph and not self.dp.sync_bn  # SyncBN puts host-side collectives between kernels
        self.grad_buckets = grad_buckets or int(os.environ.get("UNETK_GRAD_BUCKETS", "8"))
        # SMs left to NCCL while gradient buckets are in flight (see _launch_bucket); pair it with NCCL_MAX_CTAS
        self.sm_reserve = int(os.environ.get("UNETK_DP_SM_RESERVE", "4")) if self.dp.enabled else 0
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
        # The fused optimizer updates the parameters through raw pointers (also inside graph replays), behind torch's
        # version counter: every parameter carries this shared generation, bumped once per step, so that derived caches
        # of OTHER plans of the model (the eval plan behind model(x) in validation, predict_full_image) re-pack their
        # bf16 weights after training steps (engine.weight_stamp).
        self._gen = [0]
        with torch.no_grad():
            for p, o in zip(params, offs):
                v = self.flat_p[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p._unetk_gen = self._gen
                self.grad_views[id(p)] = self.flat_g[o:o + p.numel()].view(p.shape)
        self.n_params = total
        if self.dp.enabled:
            self.dp.broadcast_(self.flat_p, 0)  # replicas start identical: parameters ...
            for b in model.buffers():           # ... and BatchNorm running statistics (a loaded checkpoint on rank 0)
                self.dp.broadcast_(b, 0)
        # {lr, alpha, eps, weight_decay, momentum}: read from device memory by the RMSprop kernel
        self._hyper_host = [float(lr), float(alpha), float(eps), float(weight_decay), float(momentum)]
        self.hyper = torch.tensor(self._hyper_host, dtype=torch.float32, device=dev)
        self.clip = torch.zeros(2, dtype=torch.float32, device=dev)
        self.sq_partial = None
        self.plan = None
        self.graphs = None
        self.graph_mode = "eager"   # "single" (one graph, collectives inside) | "segments" | "eager"
        self._program = None
        self._handles = []
        self.images = self.labels = None
        self.steps_done = 0
        # input pipeline (prefetch / step()): two staging batches in HBM filled by a copy stream
        self._copy_stream = None
        self._stage = None          # [(images, labels)] x 2
        self._stage_ready = self._stage_free = None
        self._staged = []           # slots holding a batch not yet consumed, in order
        self._stage_next = 0

    # ---- hyper-parameters ----------------------------------------------------------------------------
    def _set_hyper(self, i: int, v: float):
        self._hyper_host[i] = float(v)
        self.hyper[i:i + 1].fill_(float(v))     # a tiny fill on the current stream: ordered before the next step

    lr = property(lambda self: self._hyper_host[0], lambda self, v: self._set_hyper(0, v))
    alpha = property(lambda self: self._hyper_host[1], lambda self, v: self._set_hyper(1, v))
    eps = property(lambda self: self._hyper_host[2], lambda self, v: self._set_hyper(2, v))
    wd = property(lambda self: self._hyper_host[3], lambda self, v: self._set_hyper(3, v))
    momentum = property(lambda self: self._hyper_host[4], lambda self, v: self._set_hyper(4, v))

    def set_lr(self, lr: float) -> None:
        """What `optimizer.param_groups[0]["lr"] = lr` (ReduceLROnPlateau.step, train.py:355) does in the reference;
        effective from the next step on, in eager and in graph mode alike."""
        self.lr = lr

    @property
    def max_norm(self):
        return self._max_norm

    @max_norm.setter
    def max_norm(self, v):
        if float(v) != self._max_norm:
            self._max_norm = float(v)
            self.graphs = None      # a launch argument of the clip kernel: re-capture

    # ------------------------------------------------------------------------------------------------
    def _build(self, n, c, h, w):
        dev = self.device
        self.plan = self.builder(self.model, n, h, w, dev, True, self.grad_views, True)
        self.images = torch.zeros((n, c, h, w), dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last)
        self.labels = torch.zeros((n, 1, h, w), dtype=torch.float32, device=dev)
        from .bridge import plan_heads
        from .engine import Head
        self.heads = plan_heads(self.plan)
        if not all(isinstance(h, Head) for h in self.heads):
            raise NotImplementedError("Trainer fuses the reference's binary loss (train.py:264-278: BCE + dice on ONE logit map); "
                                      "with n_classes > 1 use model(x), your loss and loss.backward()")
        # UNet++ with deep supervision (UNetPP.py:93-102) returns four maps; the loss is the MEAN over the heads of the
        # reference recipe (0.5 BCE + 0.5 dice) applied to each map: every head back-propagates 1/4 of it.
        # The heads' loss sums share one buffer so that data-parallel ranks all-reduce them in one call.
        nh = len(self.heads)
        self.loss_sums = torch.zeros((nh, 4), dtype=torch.float64, device=dev)
        self.loss_mean = torch.zeros((), dtype=torch.float32, device=dev)
        for k, head in enumerate(self.heads):
            head.labels = self.labels
            head.dlogits = None
            head.auto_finalize = False
            head.gscale = 1.0 / nh
            head.loss_sums = self.loss_sums[k]
        if self.dp.sync_bn:
            self.plan.sync_sums = self.dp.reduce_bn_sums
        from . import _lib
        self.sq_partial = torch.empty(_lib.load().unetk_sqnorm_partial_floats(self.flat_g.numel()),
                                      dtype=torch.float32, device=dev)
        self._gscale = 1.0
        self._cuts = self._plan_buckets(self.grad_buckets) if self.dp.enabled else []
        self._program = self._make_program()

    def _plan_buckets(self, nbuckets: int):
        """[(k, [(lo, hi), ...])] in backward order: once the backward of ops[k:] has run, the flat-gradient ranges
        [lo, hi) (floats) are final and are all-reduced while ops[:k] still run.  A parameter is final after the
        backward of the FIRST op (in forward order) that uses it; ranges are emitted whenever about total/nbuckets
        bytes have become final, the rest (the first layers) after the last op."""
        base, total = self.flat_g.data_ptr(), self.flat_g.numel()
        final_at = {}
        for i, op in enumerate(self.plan.ops):
            for q in self.plan.op_params(op):
                g = self.grad_views.get(id(q))
                if g is not None:
                    o = (g.data_ptr() - base) // 4
                    prev = final_at.get(o)
                    final_at[o] = (i if prev is None else min(prev[0], i), o + (g.numel() + 3) // 4 * 4)
        covered = sorted((o, hi) for o, (_, hi) in final_at.items())
        # parameters no op of the plan touches keep a zero gradient: they ride in the last bucket
        pos, rest = 0, []
        for o, hi in covered:
            if o > pos:
                rest.append((0, pos, o))
            pos = max(pos, hi)
        if pos < total:
            rest.append((0, pos, total))
        items = sorted([(k, o, hi) for o, (k, hi) in final_at.items()] + rest, key=lambda t: (-t[0], t[1]))
        thr = max(1, total // max(1, nbuckets))
        cuts, ready, ready_n, done = [], [], 0, 0
        for j, (k, lo, hi) in enumerate(items):
            ready.append((lo, hi))
            ready_n += hi - lo
            done += hi - lo
            last_of_op = j + 1 == len(items) or items[j + 1][0] != k
            # towards the end of the backward the buckets shrink geometrically (a bucket goes out as soon as it is at
            # least as large as everything still to come), so what is reduced AFTER the last op is small
            big_enough = ready_n >= thr or (nbuckets > 1 and ready_n >= max(1 << 16, total - done))
            if last_of_op and (big_enough or j + 1 == len(items)):
                cuts.append((k if j + 1 < len(items) else 0, self._coalesce(ready)))
                ready, ready_n = [], 0
        return cuts

    @staticmethod
    def _coalesce(ranges):
        out = []
        for lo, hi in sorted(ranges):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(lo, hi) for lo, hi in out]

    def bucket_report(self):
        """What bench.py prints as `grad_buckets`: MB per all-reduce call in launch order, and how the step is run."""
        if not self.dp.enabled or self.plan is None:
            return None
        return {"mode": self.graph_mode, "sm_reserve": self.sm_reserve, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "calls_mb": [[round((hi - lo) * 4 / 2**20, 2) for lo, hi in rs] for _, rs in self._cuts],
                "cut_after_op": [k for k, _ in self._cuts], "ops": len(self.plan.ops)}

    # program ------------------------------------------------------------------------------------------
    def _make_program(self):
        """The step as a list of ("k", fn) kernel-launch segments and ("c", fn) collectives, in launch order."""
        P, head = self.plan, self.plan.head
        prog = [("k", self._seg_forward_packed)]
        if self.dp.sync_loss:
            prog.append(("c", lambda: self.dp.reduce_loss_sums(self.loss_sums, head.npix)))
        hi, first = len(P.ops), True
        for k, ranges in self._cuts:
            if k < hi or first:
                prog.append(("k", functools.partial(self._seg_backward, k, hi, first)))
                hi, first = k, False
            prog.append(("c", functools.partial(self._launch_bucket, ranges)))
        if hi > 0 or first:
            prog.append(("k", functools.partial(self._seg_backward, 0, hi, first)))
        if self._cuts:
            prog.append(("c", self._wait_buckets))
        prog.append(("k", self._seg_optim))
        return prog

    def _seg_forward_packed(self):
        self.plan.refresh_weights(force=True)  # the optimizer updates weights behind torch's version counter
        self.plan.forward(self.images)

    def _seg_backward(self, lo, hi, first):
        if first:
            for h in self.heads:
                h.finalize_loss(self._npix_total)
        # the weight-gradient side stream is joined where a captured segment ends, and before the optimizer
        self.plan.backward(lo, hi, join=(self.graph_mode == "segments" or lo == 0))

    def _launch_bucket(self, ranges):
        """SUM all-reduce of gradient ranges that are final, issued behind BOTH the main stream (BatchNorm / bias
        gradients) and the weight-gradient side stream, without making the main stream wait for either."""
        with self.plan.behind_both_streams():
            for lo, hi in ranges:
                self._handles.append(self.dp.reduce_grads_async(self.flat_g[lo:hi]))
        if self.sm_reserve > 0:
            # From here to _wait_buckets NCCL kernels run beside the backward.  The tensor-core kernels are persistent
            # grids of one CTA per SM with a static tile order: with NCCL holding k SMs, k of their CTAs start only when
            # something else ends and the kernel takes twice as long — the all-reduce time is then added to the step
            # instead of hidden (measured: +0.43 ms at 2 GPUs, +1.05 ms at 8).  Launch them a few SMs narrower instead.
            from . import _lib
            lib = _lib.load()
            lib.unetk_set_sm_limit(max(1, lib.unetk_device_sms() - self.sm_reserve))

    def _wait_buckets(self):
        for h in self._handles:
            self.dp.wait(h)
        self._handles = []
        if self.sm_reserve > 0:
            from . import _lib
            _lib.load().unetk_set_sm_limit(0)

    def _seg_optim(self):
        ops.grad_clip_coef(self.flat_g, self._gscale, self._max_norm, self.sq_partial, self.clip)
        ops.rmsprop_step_dev(self.flat_p, self.flat_g, self.sq, self.buf, self.hyper, self.clip)

    def _run_segments(self):
        if self.graphs is not None and self.graph_mode == "single":
            self.graphs[0].replay()
        elif self.graphs is not None:
            it = iter(self.graphs)
            for kind, fn in self._program:
                if kind == "k":
                    next(it).replay()
                else:
                    fn()
        else:
            for _, fn in self._program:
                fn()
        self._gen[0] += 1

    def _loss(self):
        """The step's loss as a device scalar: the head's, or the mean over the deep-supervision heads."""
        if len(self.heads) == 1:
            return self.heads[0].fin[0]
        torch.mean(torch.stack([h.fin[0] for h in self.heads]), dim=0, out=self.loss_mean)
        return self.loss_mean

    def _capture(self):
        """Capture the step.  On a HIGH-priority stream: the plan's weight-gradient side stream is low priority, so
        whenever a kernel of the critical path and a weight gradient are both ready, the critical path gets the SMs."""
        hp = torch.cuda.Stream(device=self.device, priority=-1)
        want_single = os.environ.get("UNETK_SINGLE_GRAPH", "1") != "0" and (
            not self.dp.enabled or os.environ.get("UNETK_DP_GRAPH", "1") != "0")
        if want_single:
            self.graph_mode = "single"
            try:
                g = torch.cuda.CUDAGraph()
                # thread_local: NCCL's watchdog thread may query events while this thread captures
                with torch.cuda.graph(g, stream=hp, capture_error_mode="thread_local"):
                    for _, fn in self._program:
                        fn()
                self.graphs = [g]
                return
            except Exception as e:  # NCCL calls that cannot be captured on this stack: one graph per kernel segment
                if not self.dp.enabled:
                    raise
                print(f"[unetk] single-graph capture with NCCL failed ({type(e).__name__}: {e}); using per-segment graphs",
                      file=sys.stderr)
                self._handles = []
                torch.cuda.synchronize()
        self.graph_mode = "segments"
        graphs, pool = [], None
        for kind, fn in self._program:
            if kind != "k":
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=hp):
                fn()
            pool = g.pool()
            graphs.append(g)
        self.graphs = graphs

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
            return self._loss()
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
        return self._loss()

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

    def close(self):
        """Release the captured CUDA graphs (and the plan).  In data-parallel runs the graph holds captured NCCL kernels:
        call this BEFORE torch.distributed.destroy_process_group(), which otherwise may wait forever for the communicator
        to become idle (observed with 2 ranks on B200: every check passed, the process never exited)."""
        torch.cuda.synchronize(self.device)
        self.graphs = None
        self._handles = []
        self._program = None
        self.plan = None
        import gc

        gc.collect()
        torch.cuda.synchronize(self.device)

    def loss_terms(self):
        """(loss, bce, dice) of the last step as device scalars (deep supervision: means over the four heads)."""
        if len(self.heads) == 1:
            f = self.heads[0].fin
            return f[0], f[1], f[2]
        f = torch.stack([h.fin[:3] for h in self.heads]).mean(dim=0)
        return f[0], f[1], f[2]

    def grad_norm(self):
        return self.clip[0]
