"""Static-shape execution plans: the host side of the hot path.

A Plan is an ordered list of fused ops over pre-allocated NHWC bf16 buffers.  forward() runs the list,
backward() runs it in reverse; every op is a handful of C-ABI calls (jcfszxc_unet_b200/_lib.py) on the
current CUDA stream, so a whole step can be captured into one CUDA graph (see trainer.py).  There is no
tracing and no torch op on the path: PyTorch only owns the memory and the stream.

Buffer scheme (vanilla U-Net, reference UNetFamily/UNet.py:39-55):
  * per conv unit: `raw` (conv output, saved for the BN backward) and `out` (post BN+ReLU activation);
  * the skip connection of level i and the ConvTranspose output of the matching Up block are the two
    channel halves of ONE buffer cat[i] = [skip | up]  — torch.cat (unet_parts.py:69) never runs;
  * gradients mirror the activations, so the dgrad of the decoder conv writes d(skip) and d(up) in place.
"""
from __future__ import annotations

import torch

from . import _lib, ops

BF16 = torch.bfloat16


class Act:
    """An NHWC bf16 activation (a channel slice of `buf`) and, in training plans, its gradient."""

    __slots__ = ("t", "g", "N", "H", "W", "C")

    def __init__(self, t: torch.Tensor, g: torch.Tensor | None):
        self.t, self.g = t, g
        self.N, self.H, self.W, self.C = t.shape

    def slice(self, c0: int, c: int) -> "Act":
        return Act(self.t[..., c0:c0 + c], None if self.g is None else self.g[..., c0:c0 + c])


class Image:
    """The fp32 network input [N,C,H,W] in whatever strides the caller uses (leaf, no gradient)."""

    def __init__(self):
        self.x: torch.Tensor | None = None


class Plan:
    def __init__(self, device, N: int, H: int, W: int, training: bool, with_grad: bool | None = None):
        """training: BatchNorm uses batch statistics (model.train()); with_grad: allocate gradient buffers."""
        self.device, self.N, self.H, self.W, self.training = device, N, H, W, training
        self.with_grad = training if with_grad is None else with_grad
        if self.with_grad and not training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not on this path")
        self.ops: list = []
        self.image = Image()
        self.generation = 0
        self._need_partial = 1024
        self._need_ws = 1 << 20
        self._cmax = 8
        self.partial = self.sums = self.coef = self.ws = None
        self.grad_of = {}          # id(param) -> fp32 gradient tensor (same shape as the parameter)
        self.params = []           # parameters in registration order
        self.sync_sums = None      # optional hook(sums_view) -> None: all-reduce BN statistics (SyncBN)

    # ---- allocation -----------------------------------------------------------------------------
    def act(self, H, W, C, grad=None) -> Act:
        grad = self.with_grad if grad is None else grad
        t = torch.empty((self.N, H, W, C), dtype=BF16, device=self.device)
        g = torch.empty_like(t) if grad else None
        return Act(t, g)

    def vec(self, C, n=1):
        return torch.zeros((n, C), dtype=torch.float32, device=self.device)

    def need(self, partial_floats=0, ws_bytes=0, channels=0):
        self._need_partial = max(self._need_partial, int(partial_floats))
        self._need_ws = max(self._need_ws, int(ws_bytes))
        self._cmax = max(self._cmax, int(channels))

    def register_param(self, p: torch.nn.Parameter):
        if id(p) not in self.grad_of:
            self.params.append(p)
            self.grad_of[id(p)] = None

    def finalize(self, grad_views: dict | None = None):
        """Allocate shared scratch and gradient storage (or adopt views of a caller-owned flat buffer)."""
        dev = self.device
        self.partial = torch.empty(self._need_partial, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2 * self._cmax, dtype=torch.float64, device=dev)
        self.coef = torch.zeros(2 * self._cmax, dtype=torch.float32, device=dev)
        self.ws = torch.empty(self._need_ws, dtype=torch.uint8, device=dev)
        if self.with_grad:
            for p in self.params:
                if grad_views is not None:
                    self.grad_of[id(p)] = grad_views[id(p)]
                else:
                    self.grad_of[id(p)] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        for op in self.ops:
            op.bind(self)
        return self

    # ---- execution ------------------------------------------------------------------------------
    def refresh_weights(self, force=False):
        for op in self.ops:
            op.refresh(force)

    def forward(self, x: torch.Tensor | None = None):
        if x is not None:
            if x.dtype != torch.float32:
                x = x.float()
            self.image.x = x
        self.generation += 1
        self.refresh_weights()
        for op in self.ops:
            op.fwd()

    def backward(self):
        for op in reversed(self.ops):
            op.bwd()

    def grads(self):
        return [self.grad_of[id(p)] for p in self.params]


def _s():
    return torch.cuda.current_stream().cuda_stream


class ConvBNReLU:
    """conv3x3 (tcgen05 tap-GEMM, or the direct stem kernel when the input is the image) -> BatchNorm
    (batch statistics in training, running statistics in eval) -> ReLU, optionally fused with the 2x2
    max-pool that follows it in `Down`.  Reference: unet_parts.py:24-26 / 27-29 (+ :43)."""

    def __init__(self, plan: Plan, x, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d, out: Act,
                 pooled: Act | None = None, relu: bool = True):
        assert conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride == (1, 1)
        self.plan, self.x, self.conv, self.bn, self.out, self.pooled, self.relu = plan, x, conv, bn, out, pooled, relu
        self.stem = isinstance(x, Image)
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self.raw = plan.act(out.H, out.W, self.cout)
        self.stat = plan.vec(self.cout, 4)  # scale, shift, mean, invstd
        self.wpack = self.wpack_t = None
        self._wver = -1
        N, H, W = out.N, out.H, out.W
        lib = _lib.load()
        units = N * H * W
        plan.need(max(lib.unetk_chan_partial_floats(units, self.cout), lib.unetk_conv_stats_partial_floats(self.cout)),
                  0, self.cout)
        if plan.with_grad:
            if self.stem:
                plan.need(0, lib.unetk_stem_wgrad_workspace(N, H, W, self.cin))
            else:
                plan.need(0, lib.unetk_conv_wgrad_workspace(N, H, W, self.cin, self.cout, 9))
        for p in (conv.weight, conv.bias, bn.weight, bn.bias):
            if p is not None:
                plan.register_param(p)
        plan.ops.append(self)

    def bind(self, plan):
        g = plan.grad_of
        self.dw = g.get(id(self.conv.weight))
        self.dbias = g.get(id(self.conv.bias)) if self.conv.bias is not None else None
        self.dgamma = g.get(id(self.bn.weight)) if self.bn.weight is not None else None
        self.dbeta = g.get(id(self.bn.bias)) if self.bn.bias is not None else None

    def refresh(self, force=False):
        if self.stem:
            return
        w = self.conv.weight
        if force or w._version != self._wver or self.wpack is None:
            if self.wpack is None:
                self.wpack = torch.empty((9, self.cout, self.cin), dtype=BF16, device=w.device)
                self.wpack_t = torch.empty((9, self.cin, self.cout), dtype=BF16, device=w.device)
            _lib.call("unetk_pack_weight", w.data_ptr(), self.wpack.data_ptr(), self.wpack_t.data_ptr(), self.cout,
                      self.cin, 9, _s())
            self._wver = w._version

    def fwd(self):
        P, bn = self.plan, self.bn
        bias = self.conv.bias
        batch_stats = P.training or not bn.track_running_stats
        fused_stats = batch_stats and not self.stem
        if self.stem:
            ops.stem_fwd(P.image.x, self.conv.weight, bias.detach() if bias is not None else None, self.raw.t)
        elif fused_stats:
            # conv epilogue also produces the per-channel (sum, sum of squares) of its bf16 output
            xp, xld = ops.nhwc(self.x.t)
            yp, yld = ops.nhwc(self.raw.t)
            _lib.call("unetk_conv3x3_fwd_bnstats", xp, xld, self.wpack.data_ptr(),
                      bias.detach().data_ptr() if bias is not None else None, yp, yld, P.partial.data_ptr(),
                      P.sums.data_ptr(), self.raw.N, self.raw.H, self.raw.W, self.cin, self.cout, _s())
        else:
            ops.conv_fwd(self.x.t, self.wpack, bias.detach() if bias is not None else None, self.raw.t, 3)
        sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
        gamma = bn.weight.detach() if bn.weight is not None else None
        beta = bn.bias.detach() if bn.bias is not None else None
        if batch_stats:
            if not fused_stats:
                ops.bn_stats(self.raw.t, P.partial, P.sums)
            count = self.raw.N * self.raw.H * self.raw.W
            if P.sync_sums is not None:
                count = P.sync_sums(P.sums[: 2 * self.cout], count)
            track = bn.track_running_stats and P.training
            ops.bn_finalize(P.sums, count, gamma, beta, bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                            bn.running_mean if track else None, bn.running_var if track else None,
                            bn.num_batches_tracked if track else None, sc, sh, mu, iv)
        else:
            ops.bn_eval_fold(gamma, beta, bn.eps, bn.running_mean, bn.running_var, sc, sh, mu, iv)
        ops.bn_apply(self.raw.t, sc, sh, self.out.t, self.pooled.t if self.pooled is not None else None, self.relu)

    def bwd(self):
        P = self.plan
        sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
        g1 = self.out.g
        gp = self.pooled.g if self.pooled is not None else None
        ops.bn_bwd_reduce(self.raw.t, g1, gp, sc, sh, mu, iv, P.partial, P.sums, self.relu)
        count = self.raw.N * self.raw.H * self.raw.W
        if P.sync_sums is not None:
            count = P.sync_sums(P.sums[: 2 * self.cout], count)
        ops.bn_bwd_apply(self.raw.t, g1, gp, sc, sh, mu, iv, P.sums, count, self.dgamma, self.dbeta, P.coef,
                         self.raw.g, self.relu)
        if self.stem:
            n, cin, h, w = P.image.x.shape
            xp, sn, sc_, sh_, sw = ops._img(P.image.x)
            dyp, dyld = ops.nhwc(self.raw.g)
            _lib.call("unetk_stem_conv3x3_wgrad", xp, sn, sc_, sh_, sw, dyp, dyld, self.dw.data_ptr(), 0, n, h, w,
                      cin, self.cout, P.ws.data_ptr(), P.ws.numel(), _s())
        else:
            xp, xld = ops.nhwc(self.x.t)
            dyp, dyld = ops.nhwc(self.raw.g)
            _lib.call("unetk_conv3x3_wgrad", xp, xld, dyp, dyld, self.dw.data_ptr(), 0, self.raw.N, self.raw.H,
                      self.raw.W, self.cin, self.cout, P.ws.data_ptr(), P.ws.numel(), _s())
        if self.dbias is not None:
            ops.colsum(self.raw.g, P.partial, self.dbias)
        if not self.stem and self.x.g is not None:
            ops.conv_dgrad(self.raw.g, self.wpack_t, self.x.g, 3)


class ConvT2x2:
    """ConvTranspose2d(k=2, s=2) writing straight into the upper channel half of the concat buffer.
    Reference: Up.up, unet_parts.py:56-58,62."""

    def __init__(self, plan: Plan, x: Act, mod: torch.nn.ConvTranspose2d, out: Act):
        assert mod.kernel_size == (2, 2) and mod.stride == (2, 2) and mod.padding == (0, 0)
        self.plan, self.x, self.mod, self.out = plan, x, mod, out
        self.cin, self.cout = mod.in_channels, mod.out_channels
        assert out.H == 2 * x.H and out.W == 2 * x.W and out.C == self.cout
        self.w_fwd = self.w_dgrad = None
        self._wver = -1
        lib = _lib.load()
        if plan.with_grad:
            plan.need(lib.unetk_chan_partial_floats(out.N * out.H * out.W, self.cout),
                      lib.unetk_conv_wgrad_workspace(x.N, x.H, x.W, self.cin, self.cout, 4), self.cout)
        plan.register_param(mod.weight)
        if mod.bias is not None:
            plan.register_param(mod.bias)
        plan.ops.append(self)

    def bind(self, plan):
        self.dw = plan.grad_of.get(id(self.mod.weight))
        self.db = plan.grad_of.get(id(self.mod.bias)) if self.mod.bias is not None else None

    def refresh(self, force=False):
        w = self.mod.weight
        if force or w._version != self._wver or self.w_fwd is None:
            if self.w_fwd is None:
                self.w_dgrad = torch.empty((4, self.cin, self.cout), dtype=BF16, device=w.device)
                self.w_fwd = torch.empty((4, self.cout, self.cin), dtype=BF16, device=w.device)
            _lib.call("unetk_pack_weight", w.data_ptr(), self.w_dgrad.data_ptr(), self.w_fwd.data_ptr(), self.cin,
                      self.cout, 4, _s())
            self._wver = w._version

    def fwd(self):
        b = self.mod.bias
        ops.convT_fwd(self.x.t, self.w_fwd, b.detach() if b is not None else None, self.out.t)

    def bwd(self):
        P = self.plan
        dy = self.out.g
        xp, xld = ops.nhwc(self.x.t)
        dyp, dyld = ops.nhwc(dy)
        _lib.call("unetk_convT2x2_wgrad", xp, xld, dyp, dyld, self.dw.data_ptr(), 0, self.x.N, self.x.H, self.x.W,
                  self.cin, self.cout, P.ws.data_ptr(), P.ws.numel(), _s())
        if self.db is not None:
            ops.colsum(dy, P.partial, self.db)
        if self.x.g is not None:
            ops.convT_dgrad(dy, self.w_dgrad, self.x.g)


class MaxPool2x2:
    """Stand-alone nn.MaxPool2d(2) (unet_parts.py:43) for blocks used outside a fused model plan."""

    def __init__(self, plan: Plan, x: Act, out: Act):
        assert out.H == x.H // 2 and out.W == x.W // 2 and out.C == x.C
        self.plan, self.x, self.out = plan, x, out
        plan.ops.append(self)

    def bind(self, plan):
        pass

    def refresh(self, force=False):
        pass

    def fwd(self):
        ops.maxpool_fwd(self.x.t, self.out.t)

    def bwd(self):
        if self.x.g is not None:
            ops.maxpool_bwd(self.x.t, self.out.g, self.x.g)


class Head:
    """OutConv (1x1, C -> 1) fused with sigmoid + BCE-with-logits + dice sums when labels are attached.
    Reference: unet_parts.py:73-79; train.py:264-278; utils/dice_score.py:13-59."""

    def __init__(self, plan: Plan, x: Act, conv: torch.nn.Conv2d):
        assert conv.kernel_size == (1, 1)
        if conv.out_channels != 1:
            raise NotImplementedError("the fused head supports n_classes == 1 (every BASELINE.json config)")
        self.plan, self.x, self.conv = plan, x, conv
        self.C = conv.in_channels
        self.npix = x.N * x.H * x.W
        dev = plan.device
        self.logits = torch.empty((x.N, 1, x.H, x.W), dtype=torch.float32, device=dev)
        self.loss_sums = torch.zeros(4, dtype=torch.float64, device=dev)
        self.fin = torch.zeros(8, dtype=torch.float32, device=dev)
        self.labels: torch.Tensor | None = None     # fp32 [N,1,H,W] (contiguous) for the fused loss
        self.dlogits: torch.Tensor | None = None    # set instead of labels when autograd supplies dL/dlogits
        self.sync_loss = None                       # optional hook(sums, npix) -> global pixel count (data parallel)
        self.auto_finalize = True                   # trainer.py finalizes itself (collective between graph segments)
        self.gscale = 1.0
        plan.need(_lib.load().unetk_head_partial_floats(self.npix, self.C), 0, self.C)
        plan.register_param(conv.weight)
        if conv.bias is not None:
            plan.register_param(conv.bias)
        plan.ops.append(self)

    def bind(self, plan):
        self.dw = plan.grad_of.get(id(self.conv.weight))
        self.db = plan.grad_of.get(id(self.conv.bias)) if self.conv.bias is not None else None

    def refresh(self, force=False):
        pass

    def fwd(self):
        P = self.plan
        w = self.conv.weight.detach().view(-1)
        b = self.conv.bias.detach() if self.conv.bias is not None else None
        ops.head_fwd(self.x.t, w, b, self.labels, self.logits, P.partial, self.loss_sums if self.labels is not None else None)
        if self.labels is not None and self.auto_finalize:
            npix = self.npix
            if self.sync_loss is not None:
                npix = self.sync_loss(self.loss_sums, npix)
            self.finalize_loss(npix)

    def finalize_loss(self, npix_total: int):
        """fin <- {loss, bce, dice, 1/npix, cA, cB} from the (possibly all-reduced) loss sums."""
        ops.loss_finalize(self.loss_sums, npix_total, self.fin)

    def bwd(self):
        P = self.plan
        w = self.conv.weight.detach().view(-1)
        ops.head_bwd(self.x.t, w, self.labels, self.logits, self.fin, self.dlogits, self.gscale, self.x.g,
                     self.dw.view(-1) if self.dw is not None else None, self.db, P.partial)


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)


def build_unet_plan(model, N: int, H: int, W: int, device, training: bool, grad_views=None,
                    with_grad: bool | None = None) -> Plan:
    """Wire the vanilla U-Net (reference UNetFamily/UNet.py:14-55) into a Plan."""
    _require(H % 16 == 0 and W % 16 == 0 and H >= 16 and W >= 16,
             f"UNet plan needs H, W divisible by 16 (got {H}x{W}); F.pad of odd sizes is not on this path")
    P = Plan(device, N, H, W, training, with_grad)
    dcs = [model.inc.double_conv] + [getattr(model, f"down{i}").maxpool_conv[1].double_conv for i in range(1, 5)]
    C = [dc[3].out_channels for dc in dcs]
    # cat[i] = [skip_i | up_i]; gradients of both halves live in one buffer as well
    cats = [P.act(H >> i, W >> i, 2 * C[i]) for i in range(4)]
    x = P.image
    for i, dc in enumerate(dcs):
        h, w = H >> i, W >> i
        mid = P.act(h, w, dc[0].out_channels)
        ConvBNReLU(P, x, dc[0], dc[1], mid)
        if i < 4:
            out = cats[i].slice(0, C[i])
            pooled = P.act(h >> 1, w >> 1, C[i])
        else:
            out, pooled = P.act(h, w, C[i]), None
        ConvBNReLU(P, mid, dc[3], dc[4], out, pooled)
        x = pooled if pooled is not None else out
    y = x
    for j, i in enumerate((3, 2, 1, 0)):
        up = getattr(model, f"up{j + 1}")
        ConvT2x2(P, y, up.up, cats[i].slice(C[i], C[i]))
        dc = up.conv.double_conv
        h, w = H >> i, W >> i
        mid = P.act(h, w, dc[0].out_channels)
        ConvBNReLU(P, cats[i], dc[0], dc[1], mid)
        y = P.act(h, w, dc[3].out_channels)
        ConvBNReLU(P, mid, dc[3], dc[4], y)
    P.head = Head(P, y, model.outc.conv)
    return P.finalize(grad_views)
