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

import contextlib
import os

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
        self.packs = {}            # id(weight) -> WeightPack shared by every op that applies that weight
        self._pack_table = self._pack_ptrs = None
        self._pack_tiles = 0
        # Weight gradients run on a low-priority side stream: nothing on the backward's critical path consumes them, so
        # they fill the tensor cores while the HBM-bound BatchNorm passes of the next layer down run on the main stream.
        self.overlap_wgrad = os.environ.get("UNETK_WGRAD_STREAM", "1") != "0"
        self._side = None
        self._side_dirty = False
        self._g_init = {}          # gradient buffer -> set of channels already written during backward
        self._g_writers = {}       # gradient buffer -> {channel: number of ops that write it during backward}
        self._p_init = set()       # parameters whose gradient was already written during backward
        self._p_writers = {}       # id(parameter) -> number of ops that write its gradient during backward

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
        # Static backward schedule: walking the ops in backward order, the first op that produces (part of) a
        # gradient buffer writes it, every later one accumulates.  Tensors with several consumers (recurrent /
        # residual / dense-skip / gated variants) and shared weights (Recurrent_block) need no zero-fill pass.
        self._g_init.clear()
        self._g_writers.clear()
        self._p_init.clear()
        self._p_writers.clear()
        for op in reversed(self.ops):
            op.plan_bwd(self)
        if self.with_grad and os.environ.get("UNETK_SHARE_BIAS_SUMS", "1") != "0":
            self._share_bias_sums()
        if self.with_grad:
            for op in self.ops:
                if isinstance(op, ConvT2x2):
                    op.fuse_bias_grad(self)
        return self

    def _share_bias_sums(self):
        """The bias gradient of a conv WITHOUT BatchNorm is the per-channel sum of its output gradient (a two-pass
        reduction over the whole tensor).  Where another op that runs earlier in the backward already reduces the very
        same gradient tensor, copy its result instead:
          * ResidualConv (unet_parts.py:472-475): out = conv_block(x) + conv_skip(x); d(out) feeds both the last conv of
            conv_block (bias, no BN) and the BatchNorm (no ReLU) of conv_skip, whose d(beta) IS that column sum;
          * ResUNet's input stage (ResUNet.py:53-54): input_layer(x) + input_skip(x), two biased convs behind one sum.
        Measured on ResUNet (batch 8, 512^2): 1.2 ms of unetk_colsum per step."""
        def key(t):
            return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()))

        sources = {}
        for op in reversed(self.ops):
            if not isinstance(op, ConvBNReLU):
                continue
            op.bias_from = None
            if (op.bn is not None and not op.relu and not op.head_fused and op.pooled is None and op.dbeta is not None
                    and not op.acc_bn and self._p_writers.get(id(op.bn.bias), 0) == 1 and op.out.g is not None):
                sources.setdefault(key(op.out.g), op.dbeta)
            if op.bn is None and op.dbias is not None and op.raw.g is not None:
                k = key(op.raw.g)
                if k in sources:
                    op.bias_from = sources[k]
                elif not op.acc_b and self._p_writers.get(id(op.conv.bias), 0) == 1:
                    sources[k] = op.dbias

    def grad_acc(self, a: "Act") -> bool:
        """True if `a.g` already holds a gradient when the calling op's backward runs (=> accumulate)."""
        if a is None or a.g is None:
            return False
        st = a.g.untyped_storage().data_ptr()
        ld = a.g.stride(2)
        c0 = a.g.storage_offset() % ld
        chans = set(range(c0, c0 + a.C))
        seen = self._g_init.setdefault((st, ld), set())
        wr = self._g_writers.setdefault((st, ld), {})
        for c in chans:
            wr[c] = wr.get(c, 0) + 1
        hit = chans & seen
        if hit and hit != chans:
            raise RuntimeError("gradient buffer partially initialised: a builder wired overlapping channel slices")
        seen |= chans
        return bool(hit)

    def sole_writer(self, a: "Act") -> bool:
        """True if exactly one op writes each channel of `a.g` during backward (valid after finalize's walk)."""
        st, ld = a.g.untyped_storage().data_ptr(), a.g.stride(2)
        c0 = a.g.storage_offset() % ld
        wr = self._g_writers.get((st, ld), {})
        return all(wr.get(c, 0) == 1 for c in range(c0, c0 + a.C))

    def grad_written(self, a: "Act"):
        """Declare that `a.g` is filled from outside the plan (the loss gradient of a block-level plan)."""
        self.grad_acc(a)

    def param_acc(self, p) -> bool:
        if p is None:
            return False
        hit = id(p) in self._p_init
        self._p_init.add(id(p))
        self._p_writers[id(p)] = self._p_writers.get(id(p), 0) + 1
        return hit

    def pack_of(self, weight, taps: int, transposed_conv: bool = False, up: bool = False) -> "WeightPack":
        wp = self.packs.get(id(weight))
        if wp is None:
            wp = self.packs[id(weight)] = WeightPack(weight, taps, up)
        assert wp.up == up, "a weight cannot serve a plain conv and a sub-pixel up-conv at once"
        return wp

    # ---- execution ------------------------------------------------------------------------------
    def _build_pack_table(self):
        lib = _lib.load()
        rows, first = [], 0
        for wp in self.packs.values():
            rows.append([wp.w.data_ptr(), wp.ab.data_ptr(), wp.ba.data_ptr(), wp.a, wp.b, wp.taps, first, int(wp.up)])
            first += lib.unetk_pack_tiles(wp.a, wp.b)
        self._pack_ptrs = [r[0] for r in rows]
        self._pack_table = torch.tensor(rows, dtype=torch.int64).to(self.device)
        self._pack_tiles = first

    def refresh_weights(self, force=False):
        """Re-derive the bf16 kernel-layout weight caches from the fp32 masters: ONE batched launch for all of them."""
        packs = list(self.packs.values())
        if packs and (force or any(wp.stale() for wp in packs)):
            if self._pack_table is None or self._pack_ptrs != [wp.w.data_ptr() for wp in packs]:
                self._build_pack_table()   # masters were re-homed (Trainer's flat parameter buffer) or first use
            _lib.call("unetk_pack_weights", self._pack_table.data_ptr(), len(packs), self._pack_tiles, _s())
            for wp in packs:
                wp.mark_fresh()
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

    def backward(self, lo: int = 0, hi: int | None = None, join: bool = True):
        """Backward of ops[lo:hi] in reverse order (the whole plan by default).  join=False leaves the weight-gradient
        side stream un-joined (a later segment, or behind_both_streams(), picks it up)."""
        ops_ = self.ops[lo:hi]
        for op in reversed(ops_):
            op.bwd()
        if join:
            self.join_side()

    @contextlib.contextmanager
    def behind_both_streams(self):
        """Run the enclosed launches (gradient all-reduces) ordered after everything enqueued so far on the main
        stream AND on the weight-gradient side stream, without stalling the main stream: they are issued from the side
        stream once it has waited for the main stream's current position."""
        if self._side is None or not self._side_dirty:
            yield
            return
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            yield

    @contextlib.contextmanager
    def wgrad_stream(self):
        """Run the enclosed launches (weight gradients; they share the `ws` workspace and therefore one stream) on the
        side stream, ordered after everything enqueued on the current stream so far."""
        if not self.overlap_wgrad:
            yield
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=0)     # 0 = low; the trainer captures at -1
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            yield
        self._side_dirty = True

    def join_side(self):
        """The current stream waits for the side stream (end of a backward segment: gradients are complete)."""
        if self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side_dirty = False

    def op_params(self, op):
        """Parameters whose gradient `op` produces (every op keeps its modules as attributes)."""
        out = []
        for v in vars(op).values():
            if isinstance(v, torch.nn.Module):
                out += list(v.parameters(recurse=True))
            elif isinstance(v, torch.nn.Parameter):
                out.append(v)
            elif isinstance(v, (list, tuple)):
                out += [q for q in v if isinstance(q, torch.nn.Parameter)]
        return out

    def grads(self):
        return [self.grad_of[id(p)] for p in self.params]


def _s():
    return torch.cuda.current_stream().cuda_stream


def weight_stamp(w) -> tuple:
    """(torch version counter, library generation) of a master weight.  torch bumps `_version` on in-place torch ops
    (optimizer.step(), load_state_dict, init); kernels of this library that update parameters through raw pointers
    (Trainer's fused RMSprop, also inside CUDA-graph replays) cannot, so the Trainer that owns a parameter attaches a
    shared generation counter to it (`_unetk_gen`, bumped once per step) and every derived cache — bf16 weight packs
    of ANY plan of the model, fp32-mode splits, the embedded 1x1 stem kernel — compares both."""
    h = getattr(w, "_unetk_gen", None)
    return (w._version, h[0] if h is not None else 0)


def bn_momentum(bn, training: bool) -> float:
    """nn.BatchNorm2d's exponential_average_factor: `momentum`, or — momentum=None — the cumulative moving average
    1 / num_batches_tracked (counted AFTER this forward's increment, torch/nn/modules/batchnorm.py).  The cumulative
    factor needs the host value of the counter: one sync per BatchNorm, and not capturable into a CUDA graph."""
    if bn.momentum is not None:
        return float(bn.momentum)
    if not (training and bn.track_running_stats):
        return 0.0
    if torch.cuda.is_current_stream_capturing():
        raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) cannot run inside a CUDA graph")
    return 1.0 / float(int(bn.num_batches_tracked) + 1)


class WeightPack:
    """bf16 kernel-layout copies of one fp32 master weight [A,B,kh,kw]: ab = [T][A][B], ba = [T][B][A].
    A derived cache: re-packed when the master changes, shared by all ops using the weight."""

    def __init__(self, weight, taps, up: bool = False):
        a, b = weight.shape[0], weight.shape[1]
        assert weight.numel() == a * b * taps
        self.w, self.a, self.b, self.taps, self.up = weight, a, b, taps, up
        if up:
            # sub-pixel packs of an up_conv weight (csrc/pack.cu): ab = forward [4 taps][4 phases][Cout][Cin],
            # ba = dgrad [16 = phase*4 + tap][Cin][Cout]; 3x3 taps that read the same low-resolution pixel pre-summed
            assert taps == 9
            self.ab = torch.empty((4, 4, a, b), dtype=BF16, device=weight.device)
            self.ba = torch.empty((16, b, a), dtype=BF16, device=weight.device)
        else:
            self.ab = torch.empty((taps, a, b), dtype=BF16, device=weight.device)
            self.ba = torch.empty((taps, b, a), dtype=BF16, device=weight.device)
        self._stamp = None

    def stale(self):
        return weight_stamp(self.w) != self._stamp

    def mark_fresh(self):
        self._stamp = weight_stamp(self.w)


class ConvBNReLU:
    """conv -> [BatchNorm -> ReLU?] [+ residual] [+ fused 2x2 max-pool]: the unit every U-Net variant is made of.

    conv: 3x3 pad 1 (stride 1 or 2) or 1x1, on the tcgen05 tap-GEMM — or the direct stem kernel when the input is
    the image (a 1x1 stem conv runs as the 3x3 stem kernel with the weight embedded in the centre tap).
    bn=None: the conv output (with bias) IS `out`.  BatchNorm uses batch statistics in training plans and the
    running statistics in eval plans.  res: out = act(bn(conv(x))) + res.
    Reference: DoubleConv unet_parts.py:24-29 (+ :43 pool), conv_block :85-90, up_conv :104-106,
    Recurrent_block :119-128, RRCNN_block :143-146, ResidualConv :458-475, NestedUNet DoubleConv UNetPP.py:18-25."""

    def __init__(self, plan: Plan, x, conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d | None, out: Act,
                 pooled: Act | None = None, relu: bool = True, res: Act | None = None, up: bool = False):
        """up: `x` is the LOW-resolution input of an up_conv (unet_parts.py:103-104, nn.Upsample(scale_factor=2) then
        this 3x3 conv): the conv runs in sub-pixel form — four 2x2-tap convs of x, one per output phase, 2.25x fewer FLOPs —
        and the up-sampled tensor and its gradient are never materialised (unetk_upconv3x3_*)."""
        k, stride = conv.kernel_size[0], conv.stride[0]
        assert conv.kernel_size in ((3, 3), (1, 1)) and conv.padding == (k // 2, k // 2) and conv.stride == (stride, stride)
        assert stride == 1 or (stride == 2 and k == 3)
        assert conv.dilation == (1, 1) and conv.groups == 1
        self.plan, self.x, self.conv, self.bn, self.out, self.pooled, self.relu, self.res = plan, x, conv, bn, out, pooled, relu, res
        self.k, self.stride = k, stride
        self.stem = isinstance(x, Image)
        assert not (self.stem and stride != 1)
        self.up = up
        assert not up or (k == 3 and stride == 1 and not self.stem)
        if bn is None:
            assert pooled is None and res is None and not relu, "a conv without BatchNorm writes its output as is"
        self.cin, self.cout = conv.in_channels, conv.out_channels
        N, H, W = out.N, out.H, out.W
        if not self.stem:
            assert ((2 * x.H, 2 * x.W) == (H, W) if up else (x.H, x.W) == (stride * H, stride * W)) and x.C == self.cin, \
                "conv input does not match its output"
            if self.cin % 8 or self.cout % 8:
                raise ValueError(f"conv {self.cin}->{self.cout}: channel counts must be multiples of 8 on the tensor-core path")
        assert out.C == self.cout
        # Eval-mode plans fold BatchNorm (+ReLU) into the conv epilogue (north_star; evaluate.py:259-275 runs eval mode): the
        # conv writes `out` directly, `raw` and the bn_apply pass do not exist.  Not for units with a residual add, and
        # given up again by consumers that need the raw conv output (Head(fuse=...) takes over the BatchNorm itself).
        self.fold = (bn is not None and not plan.training and not plan.with_grad and bn.track_running_stats and res is None
                     and k == 3 and (not self.stem or (self.cout % 16 == 0 and self.cout <= 256))
                     and os.environ.get("UNETK_EVAL_FOLD", "1") != "0")
        self.raw = out if bn is None else (None if self.fold else plan.act(H, W, self.cout))
        self.stat = plan.vec(self.cout, 4) if bn is not None else None  # scale, shift, mean, invstd
        self.pack = None if self.stem else plan.pack_of(conv.weight, k * k, up=up)
        self.w3 = self.dw3 = None
        if self.stem and k == 1:
            self.w3 = torch.zeros((self.cout, self.cin, 3, 3), dtype=torch.float32, device=plan.device)
            self.dw3 = torch.zeros_like(self.w3) if plan.with_grad else None
        self._wstamp = None
        lib = _lib.load()
        units = N * H * W
        plan.need(max(lib.unetk_chan_partial_floats(units, self.cout), lib.unetk_conv_stats_partial_floats(self.cout),
                      lib.unetk_stem_stats_partial_floats(N, H, W, self.cout) if self.stem else 0), 0, self.cout)
        if plan.with_grad:
            if self.stem:
                plan.need(0, lib.unetk_stem_wgrad_workspace(N, H, W, self.cin))
            elif up:
                plan.need(0, lib.unetk_upconv_wgrad_workspace(N, x.H, x.W, self.cin, self.cout))
            else:
                plan.need(0, lib.unetk_conv_wgrad_workspace(N, H, W, self.cin, self.cout, k * k))
        for p in (conv.weight, conv.bias) + ((bn.weight, bn.bias) if bn is not None else ()):
            if p is not None:
                plan.register_param(p)
        self.acc_w = self.acc_b = self.acc_bn = self.acc_x = self.acc_res = False
        self.head_fused = False  # set by Head(fuse=self): the head applies this unit's BatchNorm + ReLU itself
        self.copies = []         # more destinations of `out` (AddN copies folded into the BatchNorm pass, see AddN)
        # x is a torch.cat whose members' gradients live elsewhere: [(first channel, channels, Act whose .g receives
        # those columns of the dgrad)], set by the builder (builders.build_nested_unet_plan); None: dgrad writes x.g
        self.x_parts = None
        self.part_acc = []
        self.colsum_sinks = []   # (channel offset in x, C, fp32 bias gradient, accumulate): see ConvT2x2.fuse_bias_grad
        self.bias_from = None    # fp32 vector that already holds the column sum of raw.g (Plan._share_bias_sums)
        plan.ops.append(self)

    def bind(self, plan):
        g = plan.grad_of
        self.dw = g.get(id(self.conv.weight))
        self.dbias = g.get(id(self.conv.bias)) if self.conv.bias is not None else None
        bn = self.bn
        self.dgamma = g.get(id(bn.weight)) if bn is not None and bn.weight is not None else None
        self.dbeta = g.get(id(bn.bias)) if bn is not None and bn.bias is not None else None

    def _res_aliases_out(self):
        r = self.res
        return r is None or r.g is None or r.g.data_ptr() == self.out.g.data_ptr()

    def plan_bwd(self, plan):
        if not plan.with_grad:
            return
        if self.bn is not None:
            a, b = plan.param_acc(self.bn.weight), plan.param_acc(self.bn.bias)
            self.acc_bn = a or b
            if not self._res_aliases_out():
                self.acc_res = plan.grad_acc(self.res)
        self.acc_w = plan.param_acc(self.conv.weight)
        self.acc_b = plan.param_acc(self.conv.bias)
        if not self.stem and self.x.g is not None:
            if self.x_parts is not None:
                self.part_acc = [plan.grad_acc(t) for _, _, t in self.x_parts]
            else:
                self.acc_x = plan.grad_acc(self.x)

    def refresh(self, force=False):
        if self.w3 is not None:
            w = self.conv.weight
            if force or weight_stamp(w) != self._wstamp:
                # 1x1 kernel -> centre tap (index 4 of 9) of the zero-padded 3x3 stem kernel
                ops.copy_f32_strided(self.w3, 9, w.detach(), 1, self.cout * self.cin, dst_offset=4)
                self._wstamp = weight_stamp(w)

    def unfold(self):
        """A consumer needs the raw conv output after all: give the unit its `raw` buffer back."""
        if self.fold:
            self.fold = False
            self.raw = self.plan.act(self.out.H, self.out.W, self.cout)

    def _fwd_folded(self):
        P, bn = self.plan, self.bn
        sc, sh = self.stat[0], self.stat[1]
        ops.bn_eval_fold_bias(bn.weight.detach() if bn.weight is not None else None, bn.bias.detach() if bn.bias is not None else None,
                              bn.eps, bn.running_mean, bn.running_var, self.conv.bias.detach() if self.conv.bias is not None else None,
                              sc, sh)
        if self.stem:
            ops.stem_fwd_affine(P.image.x, (self.w3 if self.w3 is not None else self.conv.weight).detach(), sc, sh, self.relu, self.out.t)
        elif self.up:
            ops.upconv_fwd_affine(self.x.t, self.pack.ab, sc, sh, self.relu, self.out.t)
        else:
            ops.conv_fwd_affine(self.x.t, self.pack.ab, sc, sh, self.relu, self.out.t, self.stride)
        if self.pooled is not None:
            ops.maxpool_fwd(self.out.t, self.pooled.t)

    def fwd(self):
        if self.fold:
            return self._fwd_folded()
        P, bn = self.plan, self.bn
        bias = self.conv.bias.detach() if self.conv.bias is not None else None
        batch_stats = bn is not None and (P.training or not bn.track_running_stats)
        if self.stem:
            w = self.w3 if self.w3 is not None else self.conv.weight
            if batch_stats:
                ops.stem_fwd_stats(P.image.x, w, bias, self.raw.t, P.partial, P.sums)
            else:
                ops.stem_fwd(P.image.x, w, bias, self.raw.t)
        elif self.up:
            ops.upconv_fwd(self.x.t, self.pack.ab, bias, self.raw.t, P.partial if batch_stats else None,
                           P.sums if batch_stats else None)
        elif batch_stats:
            # conv epilogue also produces the per-channel (sum, sum of squares) of its bf16 output
            ops.conv_fwd_stats(self.x.t, self.pack.ab, bias, self.raw.t, P.partial, P.sums, self.k, self.stride)
        elif self.stride == 2:
            ops.conv_fwd_stats(self.x.t, self.pack.ab, bias, self.raw.t, None, None, self.k, 2)
        else:
            ops.conv_fwd(self.x.t, self.pack.ab, bias, self.raw.t, self.k)
        if bn is None:
            return
        sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
        gamma = bn.weight.detach() if bn.weight is not None else None
        beta = bn.bias.detach() if bn.bias is not None else None
        if batch_stats:
            count = self.raw.N * self.raw.H * self.raw.W
            if P.sync_sums is not None:
                count = P.sync_sums(P.sums[: 2 * self.cout], count)
            track = bn.track_running_stats and P.training
            ops.bn_finalize(P.sums, count, gamma, beta, bn.eps, bn_momentum(bn, P.training),
                            bn.running_mean if track else None, bn.running_var if track else None,
                            bn.num_batches_tracked if track else None, sc, sh, mu, iv)
        else:
            ops.bn_eval_fold(gamma, beta, bn.eps, bn.running_mean, bn.running_var, sc, sh, mu, iv)
        if self.head_fused:
            return   # `out` is never materialised: Head.fwd reads `raw` and (scale, shift)
        if self.copies:
            ops.bn_apply_copies(self.raw.t, sc, sh, self.out.t, [a.t for a in self.copies],
                                self.pooled.t if self.pooled is not None else None, self.relu)
            return
        ops.bn_apply(self.raw.t, sc, sh, self.out.t, self.pooled.t if self.pooled is not None else None, self.relu,
                     self.res.t if self.res is not None else None)

    def bwd(self):
        P = self.plan
        if self.bn is not None and not self.head_fused:   # fused: Head.bwd has already written raw.g and dgamma/dbeta
            sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
            g1 = self.out.g
            gp = self.pooled.g if self.pooled is not None else None
            ops.bn_bwd_reduce(self.raw.t, g1, gp, sc, sh, mu, iv, P.partial, P.sums, self.relu)
            count = self.raw.N * self.raw.H * self.raw.W
            if P.sync_sums is not None:
                count = P.sync_sums(P.sums[: 2 * self.cout], count)
            # the bias of a conv in front of a train-mode BatchNorm has an identically zero gradient: written by the
            # BN backward's per-channel kernel instead of a column-sum pass over d(raw)
            # out = relu(bn(raw)) + res: d(res) = g1 leaves this pass too (UNETK_FUSE_RES_GRAD=0: a separate add pass)
            res_fused = (not self._res_aliases_out() and gp is None
                         and os.environ.get("UNETK_FUSE_RES_GRAD", "1") != "0")
            ops.bn_bwd_apply(self.raw.t, g1, gp, sc, sh, mu, iv, P.sums, count, self.dgamma, self.dbeta, P.coef,
                             self.raw.g, self.relu, accumulate=self.acc_bn, dconv_bias=self.dbias,
                             dres=self.res.g if res_fused else None, dres_accumulate=self.acc_res)
            if not self._res_aliases_out() and not res_fused:
                ops.add_n(self.res.g, [g1], accumulate=self.acc_res)
        dy = self.raw.g
        with P.wgrad_stream():
            if self.stem:
                n, cin, h, w = P.image.x.shape
                xp, sn, sc_, sh_, sw = ops._img(P.image.x)
                dyp, dyld = ops.nhwc(dy)
                three = self.w3 is None
                dw = self.dw if three else self.dw3
                _lib.call("unetk_stem_conv3x3_wgrad", xp, sn, sc_, sh_, sw, dyp, dyld, dw.data_ptr(),
                          int(self.acc_w and three), n, h, w, cin, self.cout, P.ws.data_ptr(), P.ws.numel(), _s())
                if not three:
                    ops.copy_f32_strided(self.dw, 1, self.dw3, 9, self.cout * self.cin, accumulate=self.acc_w, src_offset=4)
            elif self.up:
                ops.upconv_wgrad(self.x.t, dy, self.dw, self.acc_w, ws=P.ws)
            else:
                ops.conv_wgrad(self.x.t, dy, self.dw, self.k, self.acc_w, self.stride, ws=P.ws)
        if self.dbias is not None and self.bn is None:
            if self.bias_from is not None:
                ops.copy_f32_strided(self.dbias, 1, self.bias_from, 1, self.cout, accumulate=self.acc_b)
            else:
                ops.colsum(dy, P.partial, self.dbias, self.acc_b)
        if self.up:
            if self.x.g is not None:
                ops.upconv_dgrad(dy, self.pack.ba, self.x.g, self.acc_x)
        elif not self.stem and self.x.g is not None and self.x_parts is not None:
            for (c0, c, t), acc in zip(self.x_parts, self.part_acc):
                ops.conv_dgrad_cols(dy, self.pack.ba, c0, t.g, acc)
        elif not self.stem and self.x.g is not None:
            if self.colsum_sinks:
                # the dgrad epilogue also sums its output per channel: the bias gradient of the ConvTranspose2d
                # whose output is a channel slice of x (the concat buffer) comes for free
                ops.conv_dgrad_colsum(dy, self.pack.ba, self.x.g, P.partial, P.sums)
                for c0, c, db, acc in self.colsum_sinks:
                    ops.sums_to_f32(P.sums, c0, c, db, acc)
            else:
                ops.conv_dgrad(dy, self.pack.ba, self.x.g, self.k, self.acc_x, self.stride)


class ConvT2x2:
    """ConvTranspose2d(k=2, s=2) writing straight into a channel slice of the concat buffer.
    Reference: Up.up, unet_parts.py:56-58,62; Upsample :478-487."""

    def __init__(self, plan: Plan, x: Act, mod: torch.nn.ConvTranspose2d, out: Act):
        assert mod.kernel_size == (2, 2) and mod.stride == (2, 2) and mod.padding == (0, 0)
        self.plan, self.x, self.mod, self.out = plan, x, mod, out
        self.cin, self.cout = mod.in_channels, mod.out_channels
        assert out.H == 2 * x.H and out.W == 2 * x.W and out.C == self.cout
        self.pack = plan.pack_of(mod.weight, 4)   # ab = [4][Cin][Cout] (dgrad), ba = [4][Cout][Cin] (fwd)
        lib = _lib.load()
        if plan.with_grad:
            plan.need(lib.unetk_chan_partial_floats(out.N * out.H * out.W, self.cout),
                      lib.unetk_conv_wgrad_workspace(x.N, x.H, x.W, self.cin, self.cout, 4), self.cout)
        plan.register_param(mod.weight)
        if mod.bias is not None:
            plan.register_param(mod.bias)
        self.acc_w = self.acc_b = self.acc_x = False
        self.fused_db = False
        plan.ops.append(self)

    def bind(self, plan):
        self.dw = plan.grad_of.get(id(self.mod.weight))
        self.db = plan.grad_of.get(id(self.mod.bias)) if self.mod.bias is not None else None

    def plan_bwd(self, plan):
        if not plan.with_grad:
            return
        self.acc_w = plan.param_acc(self.mod.weight)
        self.acc_b = plan.param_acc(self.mod.bias)
        if self.x.g is not None:
            self.acc_x = plan.grad_acc(self.x)

    def fuse_bias_grad(self, plan):
        """If `out` is a channel slice of a buffer whose gradient is written by exactly one 3x3 stride-1 conv dgrad
        (the DoubleConv that consumes the concat, unet_parts.py:69-70), let that dgrad's epilogue produce the
        bias gradient (the per-channel sum of out.g) instead of a separate column-sum pass."""
        self.fused_db = False
        if self.db is None or self.out.g is None or not plan.sole_writer(self.out):
            return
        og = self.out.g
        st, ld = og.untyped_storage().data_ptr(), og.stride(2)
        c0 = og.storage_offset() % ld
        for op in plan.ops:
            if not isinstance(op, ConvBNReLU) or op.stem or op.up or op.k != 3 or op.stride != 1 or op.x.g is None or op.acc_x:
                continue
            if op.x.C < int(os.environ.get("UNETK_COLSUM_MIN_C", "128")):
                # measured (B200, UNet 512^2, same box): 128-channel full-resolution dgrad 0.512 -> 0.597 ms with the
                # statistics epilogue against a 0.13 ms column-sum pass; below 128 channels the pass is cheaper
                continue
            xg = op.x.g
            if xg.untyped_storage().data_ptr() != st or xg.stride(2) != ld or xg.shape[:3] != og.shape[:3]:
                continue
            x0 = xg.storage_offset() % ld
            if xg.storage_offset() - x0 != og.storage_offset() - c0:
                continue
            if x0 <= c0 and c0 + self.cout <= x0 + op.x.C:
                if (plan.partial.numel() < _lib.load().unetk_conv_stats_partial_floats(op.x.C)
                        or plan.sums.numel() < 2 * op.x.C):
                    return
                op.colsum_sinks.append((c0 - x0, self.cout, self.db, self.acc_b))
                self.fused_db = True
                return

    def refresh(self, force=False):
        pass

    def fwd(self):
        b = self.mod.bias
        ops.convT_fwd(self.x.t, self.pack.ba, b.detach() if b is not None else None, self.out.t)

    def bwd(self):
        P = self.plan
        dy = self.out.g
        xp, xld = ops.nhwc(self.x.t)
        dyp, dyld = ops.nhwc(dy)
        with P.wgrad_stream():
            _lib.call("unetk_convT2x2_wgrad", xp, xld, dyp, dyld, self.dw.data_ptr(), int(self.acc_w), self.x.N, self.x.H,
                      self.x.W, self.cin, self.cout, P.ws.data_ptr(), P.ws.numel(), _s())
        if self.db is not None and not self.fused_db:
            ops.colsum(dy, P.partial, self.db, self.acc_b)
        if self.x.g is not None:
            ops.convT_dgrad(dy, self.pack.ab, self.x.g, self.acc_x)


class MaxPool2x2:
    """Stand-alone nn.MaxPool2d(2) (unet_parts.py:43) for blocks used outside a fused model plan."""

    def __init__(self, plan: Plan, x: Act, out: Act):
        assert out.H == x.H // 2 and out.W == x.W // 2 and out.C == x.C
        self.plan, self.x, self.out = plan, x, out
        self.acc_x = False
        plan.ops.append(self)

    def bind(self, plan):
        pass

    def plan_bwd(self, plan):
        self.acc_x = plan.with_grad and self.x.g is not None and plan.grad_acc(self.x)

    def refresh(self, force=False):
        pass

    def fwd(self):
        ops.maxpool_fwd(self.x.t, self.out.t)

    def bwd(self):
        if self.x.g is not None:
            ops.maxpool_bwd(self.x.t, self.out.g, self.x.g, self.acc_x)


class Head:
    """OutConv (1x1, C -> 1) fused with sigmoid + BCE-with-logits + dice sums when labels are attached.
    Reference: unet_parts.py:73-79; train.py:264-278; utils/dice_score.py:13-59."""

    def __init__(self, plan: Plan, x: Act, conv: torch.nn.Conv2d, post_sigmoid: bool = False,
                 fuse: "ConvBNReLU | None" = None, own_grad: bool = False):
        """post_sigmoid: the model ends in nn.Sigmoid (ResUNet.py:47-50, UNetPP.py:105-106); `logits` then holds
        sigmoid(conv) — the model output — which train.py:264-278 feeds to the loss as if it were a logit.
        fuse: the ConvBNReLU unit that produces `x`, when the head is the ONLY consumer of x (UNet.py:53-54,
        AttentionUNet.py:83-84, UNetPP.py:104-105): its BatchNorm + ReLU pass and their backward are folded into the
        head's passes over the conv output, x and d(x) are never written (UNETK_FUSE_HEAD=0 keeps them apart).
        own_grad: x has other consumers (the deep-supervision heads of UNet++ on X[0][1..3], UNetPP.py:93-100): the head
        writes d(x) into a private buffer which one add pass folds into x.g."""
        assert conv.kernel_size == (1, 1)
        if conv.out_channels != 1:
            raise NotImplementedError("the fused head supports n_classes == 1 (every BASELINE.json config)")
        self.plan, self.x, self.conv, self.post_sigmoid = plan, x, conv, post_sigmoid
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
        self.acc_w = False
        self.own_grad, self.acc_x = own_grad, False
        self.xg = torch.empty((x.N, x.H, x.W, x.C), dtype=BF16, device=dev) if (own_grad and plan.with_grad) else None
        plan.need(_lib.load().unetk_head_partial_floats(self.npix, self.C), 0, self.C)
        self.prod = None
        if (fuse is not None and os.environ.get("UNETK_FUSE_HEAD", "1") != "0" and fuse.out is x and fuse.bn is not None
                and fuse.res is None and fuse.pooled is None and self.C in (32, 64)):
            self.prod = fuse
            fuse.unfold()
            fuse.head_fused = True
            plan.need(_lib.load().unetk_bn_head_partial_floats(self.npix, self.C), 0, self.C)
            self.dz = torch.empty(self.npix, dtype=torch.float32, device=dev) if plan.with_grad else None
        plan.register_param(conv.weight)
        if conv.bias is not None:
            plan.register_param(conv.bias)
        plan.ops.append(self)

    def bind(self, plan):
        self.dw = plan.grad_of.get(id(self.conv.weight))
        self.db = plan.grad_of.get(id(self.conv.bias)) if self.conv.bias is not None else None

    def plan_bwd(self, plan):
        if plan.with_grad:
            self.acc_w = plan.param_acc(self.conv.weight) | plan.param_acc(self.conv.bias)
            if self.x.g is not None:
                self.acc_x = plan.grad_acc(self.x)
                if self.acc_x and not self.own_grad:
                    raise RuntimeError("Head: the input of the output conv has other consumers; build it with own_grad=True")

    def refresh(self, force=False):
        pass

    def fwd(self):
        P = self.plan
        w = self.conv.weight.detach().view(-1)
        b = self.conv.bias.detach() if self.conv.bias is not None else None
        sums = self.loss_sums if self.labels is not None else None
        if self.prod is not None:
            u = self.prod
            ops.bn_head_fwd(u.raw.t, u.stat[0], u.stat[1], u.relu, w, b, self.labels, self.logits, P.partial, sums,
                            self.post_sigmoid)
        else:
            ops.head_fwd(self.x.t, w, b, self.labels, self.logits, P.partial, sums, self.post_sigmoid)
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
        if self.prod is not None:
            u = self.prod
            sc, sh, mu, iv = u.stat[0], u.stat[1], u.stat[2], u.stat[3]
            ops.bn_head_bwd_reduce(u.raw.t, sc, sh, mu, u.relu, w, self.labels, self.logits, self.fin, self.dlogits,
                                   self.gscale, self.dz, self.dw.view(-1) if self.dw is not None else None, self.db,
                                   P.sums, P.partial, self.acc_w, self.post_sigmoid)
            count = self.npix
            if P.sync_sums is not None:
                count = P.sync_sums(P.sums[: 2 * self.C], count)
            ops.bn_bwd_coef(P.sums, count, sc, mu, iv, u.dgamma, u.dbeta, P.coef, accumulate=u.acc_bn,
                            dconv_bias=u.dbias)
            ops.bn_head_bwd_apply(u.raw.t, sc, sh, u.relu, w, self.dz, P.coef, u.raw.g)
            return
        ops.head_bwd(self.x.t, w, self.labels, self.logits, self.fin, self.dlogits, self.gscale,
                     self.xg if self.own_grad else self.x.g,
                     self.dw.view(-1) if self.dw is not None else None, self.db, P.partial, self.acc_w,
                     self.post_sigmoid)
        if self.own_grad:
            ops.add_n(self.x.g, [self.xg], accumulate=self.acc_x)


class HeadMulti:
    """OutConv with n_classes > 1 (unet_parts.py:73-79): logits [N,K,H,W] fp32; the loss lives outside the library
    (nn.CrossEntropyLoss at train.py:124, utils.dice_score.multiclass_dice_coeff) and autograd supplies dL/dlogits."""

    def __init__(self, plan: Plan, x: Act, conv: torch.nn.Conv2d):
        assert conv.kernel_size == (1, 1)
        self.plan, self.x, self.conv = plan, x, conv
        self.C, self.K = conv.in_channels, conv.out_channels
        if not 1 <= self.K <= 8:
            raise NotImplementedError(f"the output conv supports 1..8 classes, got {self.K}")
        self.npix = x.N * x.H * x.W
        self.logits = torch.empty((x.N, self.K, x.H, x.W), dtype=torch.float32, device=plan.device)
        self.labels = None          # interface parity with Head (the fused loss is binary-only)
        self.dlogits: torch.Tensor | None = None
        self.gscale = 1.0
        self.acc_w = False
        plan.need(_lib.load().unetk_head_multi_partial_floats(self.npix, self.C, self.K), 0, self.C)
        plan.register_param(conv.weight)
        if conv.bias is not None:
            plan.register_param(conv.bias)
        plan.ops.append(self)

    def bind(self, plan):
        self.dw = plan.grad_of.get(id(self.conv.weight))
        self.db = plan.grad_of.get(id(self.conv.bias)) if self.conv.bias is not None else None

    def plan_bwd(self, plan):
        if plan.with_grad:
            self.acc_w = plan.param_acc(self.conv.weight) | plan.param_acc(self.conv.bias)
            if self.x.g is not None and plan.grad_acc(self.x):
                raise RuntimeError("HeadMulti: the input of the output conv must not have other consumers")

    def refresh(self, force=False):
        pass

    def fwd(self):
        if self.labels is not None:
            raise NotImplementedError("the fused BCE+dice loss is binary (n_classes == 1); compute the loss on the logits")
        xp, xld = ops.nhwc(self.x.t)
        w = self.conv.weight.detach().view(self.K, self.C)
        b = self.conv.bias.detach() if self.conv.bias is not None else None
        _lib.call("unetk_head_multi_fwd", xp, xld, ops._f32(w), ops._f32(b), self.logits.data_ptr(), self.x.N,
                  self.x.H * self.x.W, self.C, self.K, _s())

    def bwd(self):
        if self.dlogits is None:
            raise RuntimeError("HeadMulti.bwd needs dL/dlogits")
        xp, xld = ops.nhwc(self.x.t)
        dxp, dxld = ops.nhwc(self.x.g)
        w = self.conv.weight.detach().view(self.K, self.C)
        dl = self.dlogits
        assert dl.shape == self.logits.shape and dl.dtype == torch.float32 and dl.is_contiguous()
        _lib.call("unetk_head_multi_bwd", xp, xld, ops._f32(w), dl.data_ptr(), float(self.gscale), dxp, dxld,
                  ops._f32(self.dw.view(self.K, self.C)) if self.dw is not None else None, ops._f32(self.db),
                  int(self.acc_w), self.x.N, self.x.H * self.x.W, self.C, self.K, self.plan.partial.data_ptr(), _s())


class _Op:
    """Default no-op hooks of a plan op."""

    def bind(self, plan):
        pass

    def plan_bwd(self, plan):
        pass

    def refresh(self, force=False):
        pass


class AddN(_Op):
    """out = sum(inputs) as a chain of bf16 adds (x + x1 of RRCNN_block / ResUNet input, unet_parts.py:146,
    ResUNet.py:54); with ONE input it is the slice copy behind a torch.cat of NestedUNet (UNetPP.py:75-99).
    Backward: every input's gradient (+)= out.g; an input whose .g aliases out.g needs no kernel."""

    def __init__(self, plan: Plan, inputs: list, out: Act):
        assert 1 <= len(inputs) <= 4 and all((a.N, a.H, a.W, a.C) == (out.N, out.H, out.W, out.C) for a in inputs)
        self.plan, self.inputs, self.out = plan, inputs, out
        self.acc = [False] * len(inputs)
        # a plain copy of a tensor that a conv+BatchNorm unit has just produced (the torch.cat members of NestedUNet,
        # UNetPP.py:80-97) is written by that unit's BatchNorm pass itself; only the backward of this op remains
        self.fused_fwd = False
        self.scattered = False   # the consumers' dgrads write the input's gradient themselves (ConvBNReLU.x_parts)
        if len(inputs) == 1 and os.environ.get("UNETK_FUSE_COPIES", "1") != "0":
            for op in reversed(plan.ops[-8:]):
                if isinstance(op, ConvBNReLU) and op.out is inputs[0]:
                    if op.bn is not None and op.res is None and not op.head_fused and len(op.copies) < 3 and not op.fold:
                        op.copies.append(out)
                        self.fused_fwd = True
                    break
        plan.ops.append(self)

    def _alias(self, a):
        return a.g is None or a.g.data_ptr() == self.out.g.data_ptr()

    def plan_bwd(self, plan):
        if plan.with_grad and not self.scattered:
            self.acc = [False if self._alias(a) else plan.grad_acc(a) for a in self.inputs]

    def fwd(self):
        if not self.fused_fwd:
            ops.add_n(self.out.t, [a.t for a in self.inputs])

    def bwd(self):
        if self.scattered:
            return
        for a, acc in zip(self.inputs, self.acc):
            if not self._alias(a):
                ops.add_n(a.g, [self.out.g], accumulate=acc)


class Upsample2x(_Op):
    """nn.Upsample(scale_factor=2): mode "nearest" (up_conv, unet_parts.py:103) or "bilinear" with
    align_corners=True (NestedUNet.up, UNetPP.py:44), written straight into a concat slice."""

    def __init__(self, plan: Plan, x: Act, out: Act, mode: str = "nearest"):
        assert mode in ("nearest", "bilinear")
        assert (out.H, out.W, out.C) == (2 * x.H, 2 * x.W, x.C)
        self.plan, self.x, self.out, self.mode = plan, x, out, mode
        self.acc_x = False
        plan.ops.append(self)

    def plan_bwd(self, plan):
        self.acc_x = plan.with_grad and self.x.g is not None and plan.grad_acc(self.x)

    def fwd(self):
        ops.upsample2x(self.x.t, self.out.t, self.mode)

    def bwd(self):
        if self.x.g is not None:
            ops.upsample2x_bwd(self.out.g, self.x.g, self.mode, self.acc_x)


class PadInto(_Op):
    """F.pad of the Up block (unet_parts.py:64-67): the ConvTranspose output `x` (2h x 2w) lands at offset (oy, ox) =
    (diffY // 2, diffX // 2) inside `out`, the up half of the concat buffer (the skip's size), with a zero border.
    Only inputs whose sides are not multiples of 16 take this op (MaxPool2d floors); otherwise the ConvTranspose writes
    the concat slice directly.  Backward: the crop of out.g."""

    def __init__(self, plan: Plan, x: Act, out: Act, oy: int, ox: int):
        assert x.C == out.C and x.N == out.N and 0 <= oy <= out.H - x.H and 0 <= ox <= out.W - x.W
        self.plan, self.x, self.out, self.oy, self.ox = plan, x, out, oy, ox
        plan.ops.append(self)

    def plan_bwd(self, plan):
        if plan.with_grad and self.x.g is not None and plan.grad_acc(self.x):
            raise RuntimeError("PadInto: the padded tensor must not have other consumers")

    def fwd(self):
        ops.shift_copy(self.out.t, self.x.t, self.oy, self.ox)

    def bwd(self):
        if self.x.g is not None:
            ops.shift_copy(self.x.g, self.out.g, -self.oy, -self.ox)


class BNAct(_Op):
    """Stand-alone BatchNorm2d (+ReLU) on an existing activation: the pre-activation of ResidualConv
    (unet_parts.py:458-459).  Statistics pass + apply in forward; reduce + apply in backward, accumulating
    into x.g when x has other consumers (ResidualConv.conv_skip reads the same x, :467-470)."""

    def __init__(self, plan: Plan, x: Act, bn: torch.nn.BatchNorm2d, out: Act, relu: bool = True):
        assert (x.N, x.H, x.W, x.C) == (out.N, out.H, out.W, out.C) and bn.num_features == x.C
        self.plan, self.x, self.bn, self.out, self.relu = plan, x, bn, out, relu
        self.stat = plan.vec(x.C, 4)
        plan.need(_lib.load().unetk_chan_partial_floats(x.N * x.H * x.W, x.C), 0, x.C)
        for p in (bn.weight, bn.bias):
            if p is not None:
                plan.register_param(p)
        self.acc_bn = self.acc_x = False
        plan.ops.append(self)

    def bind(self, plan):
        bn = self.bn
        self.dgamma = plan.grad_of.get(id(bn.weight)) if bn.weight is not None else None
        self.dbeta = plan.grad_of.get(id(bn.bias)) if bn.bias is not None else None

    def plan_bwd(self, plan):
        if plan.with_grad:
            a, b = plan.param_acc(self.bn.weight), plan.param_acc(self.bn.bias)
            self.acc_bn = a or b
            if self.x.g is not None:
                self.acc_x = plan.grad_acc(self.x)

    def fwd(self):
        P, bn = self.plan, self.bn
        sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
        gamma = bn.weight.detach() if bn.weight is not None else None
        beta = bn.bias.detach() if bn.bias is not None else None
        if P.training or not bn.track_running_stats:
            ops.bn_stats(self.x.t, P.partial, P.sums)
            count = self.x.N * self.x.H * self.x.W
            if P.sync_sums is not None:
                count = P.sync_sums(P.sums[: 2 * self.x.C], count)
            track = bn.track_running_stats and P.training
            ops.bn_finalize(P.sums, count, gamma, beta, bn.eps, bn_momentum(bn, P.training),
                            bn.running_mean if track else None, bn.running_var if track else None,
                            bn.num_batches_tracked if track else None, sc, sh, mu, iv)
        else:
            ops.bn_eval_fold(gamma, beta, bn.eps, bn.running_mean, bn.running_var, sc, sh, mu, iv)
        ops.bn_apply(self.x.t, sc, sh, self.out.t, None, self.relu)

    def bwd(self):
        if self.x.g is None:
            raise RuntimeError("BNAct: the input carries no gradient buffer")
        P = self.plan
        sc, sh, mu, iv = self.stat[0], self.stat[1], self.stat[2], self.stat[3]
        ops.bn_bwd_reduce(self.x.t, self.out.g, None, sc, sh, mu, iv, P.partial, P.sums, self.relu)
        count = self.x.N * self.x.H * self.x.W
        if P.sync_sums is not None:
            count = P.sync_sums(P.sums[: 2 * self.x.C], count)
        ops.bn_bwd_apply(self.x.t, self.out.g, None, sc, sh, mu, iv, P.sums, count, self.dgamma, self.dbeta, P.coef,
                         self.x.g, self.relu, accumulate=self.acc_bn, draw_accumulate=self.acc_x)


class AttentionGate(_Op):
    """Attention_block.forward (unet_parts.py:170-176): out = x * sigmoid(BN1(psi(relu(BN(W_g g) + BN(W_x x))))).
    Two 1x1 tensor-core GEMMs with the BatchNorm statistics in their epilogue, then the fused gate kernels of
    csrc/gate.cu; `out` is the skip half of the decoder's concat buffer (AttentionUNet.py:65-66)."""

    def __init__(self, plan: Plan, g: Act, x: Act, mod, out: Act):
        self.plan, self.g, self.x, self.mod, self.out = plan, g, x, mod, out
        self.cg, self.bng = mod.W_g[0], mod.W_g[1]
        self.cx, self.bnx = mod.W_x[0], mod.W_x[1]
        self.cp, self.bn1 = mod.psi[0], mod.psi[1]
        self.F = self.cg.out_channels
        assert (g.N, g.H, g.W) == (x.N, x.H, x.W) == (out.N, out.H, out.W) and out.C == x.C
        assert g.C == self.cg.in_channels and x.C == self.cx.in_channels and self.cp.out_channels == 1
        N, H, W = x.N, x.H, x.W
        self.npix = N * H * W
        dev = plan.device
        self.rawg, self.rawx = plan.act(H, W, self.F), plan.act(H, W, self.F)
        self.statg, self.statx, self.stat1 = plan.vec(self.F, 4), plan.vec(self.F, 4), plan.vec(1, 4)
        self.s = torch.empty(self.npix, dtype=torch.float32, device=dev)
        self.dz = torch.empty(self.npix, dtype=torch.float32, device=dev) if plan.with_grad else None
        self.sums2 = torch.zeros(4 * self.F + 4, dtype=torch.float64, device=dev)
        self.coef1 = torch.zeros(2, dtype=torch.float32, device=dev)
        self.coefg = torch.zeros(2 * self.F, dtype=torch.float32, device=dev)
        self.coefx = torch.zeros(2 * self.F, dtype=torch.float32, device=dev)
        self.packg, self.packx = plan.pack_of(self.cg.weight, 1), plan.pack_of(self.cx.weight, 1)
        lib = _lib.load()
        plan.need(max(lib.unetk_gate_partial_floats(self.npix, self.F), lib.unetk_chan_partial_floats(self.npix, self.F),
                      lib.unetk_conv_stats_partial_floats(self.F)), 0, self.F)
        if plan.with_grad:
            plan.need(0, max(lib.unetk_conv_wgrad_workspace(N, H, W, g.C, self.F, 1),
                             lib.unetk_conv_wgrad_workspace(N, H, W, x.C, self.F, 1)))
        self.param_list = [self.cg.weight, self.cg.bias, self.bng.weight, self.bng.bias, self.cx.weight, self.cx.bias,
                           self.bnx.weight, self.bnx.bias, self.cp.weight, self.cp.bias, self.bn1.weight, self.bn1.bias]
        for p in self.param_list:
            if p is not None:
                plan.register_param(p)
        self.acc_p = False
        self.acc_x = self.acc_g = self.acc_xw = False
        plan.ops.append(self)

    def bind(self, plan):
        self.grads = [plan.grad_of.get(id(p)) if p is not None else None for p in self.param_list]

    def plan_bwd(self, plan):
        if not plan.with_grad:
            return
        hits = [plan.param_acc(p) for p in self.param_list if p is not None]
        if any(hits) and not all(hits):
            raise RuntimeError("AttentionGate: parameters partially shared with other ops")
        self.acc_p = any(hits)
        # order of the gradient writes in bwd(): x (gate multiply), then W_g dgrad -> g, then W_x dgrad -> x
        self.acc_x = plan.grad_acc(self.x)
        self.acc_g = plan.grad_acc(self.g)
        self.acc_xw = True

    def _bn(self, bn, stat, sums, count):
        P = self.plan
        sc, sh, mu, iv = stat[0], stat[1], stat[2], stat[3]
        gamma = bn.weight.detach() if bn.weight is not None else None
        beta = bn.bias.detach() if bn.bias is not None else None
        if P.training or not bn.track_running_stats:
            if P.sync_sums is not None:
                count = P.sync_sums(sums, count)
            track = bn.track_running_stats and P.training
            ops.bn_finalize(sums, count, gamma, beta, bn.eps, bn_momentum(bn, P.training),
                            bn.running_mean if track else None, bn.running_var if track else None,
                            bn.num_batches_tracked if track else None, sc, sh, mu, iv)
        else:
            ops.bn_eval_fold(gamma, beta, bn.eps, bn.running_mean, bn.running_var, sc, sh, mu, iv)

    def fwd(self):
        P, F = self.plan, self.F
        for conv, bn, pack, src, raw, stat in ((self.cg, self.bng, self.packg, self.g, self.rawg, self.statg),
                                               (self.cx, self.bnx, self.packx, self.x, self.rawx, self.statx)):
            bias = conv.bias.detach() if conv.bias is not None else None
            if F <= 32 and (P.training or not bn.track_running_stats):
                # 1x1 conv with <= 32 output channels (the full-resolution gate, 64 -> 32 @512^2): the GEMM is HBM-bound and
                # its one-chunk epilogue cannot hide the fused statistics (measured 0.346 ms fused vs 0.156 + 0.071 ms for the
                # conv and a separate statistics pass over its 268 MB output; from 64 output channels on, fused wins)
                ops.conv_fwd(src.t, pack.ab, bias, raw.t, 1)
                ops.bn_stats(raw.t, P.partial, P.sums)
            else:
                ops.conv_fwd_stats(src.t, pack.ab, bias, raw.t, P.partial, P.sums, 1, 1)
            self._bn(bn, stat, P.sums[: 2 * F], self.npix)
        gp, gld = ops.nhwc(self.rawg.t)
        xp, xld = ops.nhwc(self.rawx.t)
        f = ops._f32
        wpsi = self.cp.weight.detach().view(-1)
        bpsi = self.cp.bias.detach() if self.cp.bias is not None else None
        _lib.call("unetk_gate_fwd", gp, gld, xp, xld, f(self.statg[0]), f(self.statg[1]), f(self.statx[0]),
                  f(self.statx[1]), f(wpsi), f(bpsi), f(self.s), P.partial.data_ptr(), self.sums2.data_ptr(), self.npix,
                  F, _s())
        self._bn(self.bn1, self.stat1, self.sums2[:2], self.npix)
        ip, ild = ops.nhwc(self.x.t)
        op, old = ops.nhwc(self.out.t)
        _lib.call("unetk_gate_apply", ip, ild, f(self.s), f(self.stat1[0]), f(self.stat1[1]), op, old, self.npix,
                  self.x.C, _s())

    def bwd(self):
        P, F, f = self.plan, self.F, ops._f32
        (dwg, dbg, dgam_g, dbet_g, dwx, dbx, dgam_x, dbet_x, dwp, dbp, dgam_1, dbet_1) = self.grads
        acc = self.acc_p
        st1 = self.stat1
        dop, dold = ops.nhwc(self.out.g)
        ip, ild = ops.nhwc(self.x.t)
        dxp, dxld = ops.nhwc(self.x.g)
        s1 = self.sums2[:2]
        _lib.call("unetk_gate_bwd_psi", dop, dold, ip, ild, f(self.s), f(st1[0]), f(st1[1]), f(st1[2]), dxp, dxld,
                  int(self.acc_x), f(self.dz), P.partial.data_ptr(), s1.data_ptr(), self.npix, self.x.C, _s())
        count = self.npix
        if P.sync_sums is not None:
            count = P.sync_sums(s1, count)
        ops.bn_bwd_coef(s1, count, st1[0], st1[2], st1[3], dgam_1, dbet_1, self.coef1, acc, dconv_bias=dbp)
        gp, gld = ops.nhwc(self.rawg.t)
        xp, xld = ops.nhwc(self.rawx.t)
        sg, sx = self.statg, self.statx
        wpsi = self.cp.weight.detach().view(-1)
        sums_g, sums_x = self.sums2[4:4 + 2 * F], self.sums2[4 + 2 * F:4 + 4 * F]
        _lib.call("unetk_gate_bwd_reduce", gp, gld, xp, xld, f(sg[0]), f(sg[1]), f(sg[2]), f(sx[0]), f(sx[1]), f(sx[2]),
                  f(wpsi), f(self.s), f(self.dz), f(st1[0]), f(self.coef1), P.partial.data_ptr(), sums_g.data_ptr(),
                  sums_x.data_ptr(), f(dwp.view(-1)) if dwp is not None else None, None, int(acc), self.npix, F, _s())
        cg_count = cx_count = self.npix
        if P.sync_sums is not None:
            cg_count = P.sync_sums(sums_g, self.npix)
            cx_count = P.sync_sums(sums_x, self.npix)
        ops.bn_bwd_coef(sums_g, cg_count, sg[0], sg[2], sg[3], dgam_g, dbet_g, self.coefg, acc, dconv_bias=dbg)
        ops.bn_bwd_coef(sums_x, cx_count, sx[0], sx[2], sx[3], dgam_x, dbet_x, self.coefx, acc, dconv_bias=dbx)
        dgp, dgld = ops.nhwc(self.rawg.g)
        dxrp, dxrld = ops.nhwc(self.rawx.g)
        _lib.call("unetk_gate_bwd_apply", gp, gld, xp, xld, f(sg[0]), f(sg[1]), f(sx[0]), f(sx[1]), f(wpsi), f(self.s),
                  f(self.dz), f(st1[0]), f(self.coef1), f(self.coefg), f(self.coefx), dgp, dgld, dxrp, dxrld, self.npix,
                  F, _s())
        for src, raw, pack, dw, db, a_in in ((self.g, self.rawg, self.packg, dwg, dbg, self.acc_g),
                                             (self.x, self.rawx, self.packx, dwx, dbx, self.acc_xw)):
            with P.wgrad_stream():
                ops.conv_wgrad(src.t, raw.g, dw, 1, acc, 1, ws=P.ws)
            ops.conv_dgrad(raw.g, pack.ba, src.g, 1, a_in, 1)


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)


def build_unet_plan(model, N: int, H: int, W: int, device, training: bool, grad_views=None,
                    with_grad: bool | None = None) -> Plan:
    """Wire the vanilla U-Net (reference UNetFamily/UNet.py:14-55) into a Plan.  Any H, W >= 16: where a level's size
    is odd, MaxPool2d floors (the fused pool is replaced by the stand-alone op) and the Up block pads the ConvTranspose
    output to the skip's size (unet_parts.py:64-67, PadInto) — exactly what the reference does."""
    _require(H >= 16 and W >= 16, f"UNet plan needs H, W >= 16 (got {H}x{W}): four 2x2 max-pools")
    P = Plan(device, N, H, W, training, with_grad)
    dcs = [model.inc.double_conv] + [getattr(model, f"down{i}").maxpool_conv[1].double_conv for i in range(1, 5)]
    C = [dc[3].out_channels for dc in dcs]
    hs, ws = [H], [W]
    for _ in range(4):
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    # cat[i] = [skip_i | up_i]; gradients of both halves live in one buffer as well
    cats = [P.act(hs[i], ws[i], 2 * C[i]) for i in range(4)]
    x = P.image
    for i, dc in enumerate(dcs):
        h, w = hs[i], ws[i]
        mid = P.act(h, w, dc[0].out_channels)
        ConvBNReLU(P, x, dc[0], dc[1], mid)
        if i < 4:
            out = cats[i].slice(0, C[i])
            pooled = P.act(hs[i + 1], ws[i + 1], C[i])
        else:
            out, pooled = P.act(h, w, C[i]), None
        if pooled is not None and (h % 2 or w % 2):
            ConvBNReLU(P, mid, dc[3], dc[4], out)
            MaxPool2x2(P, out, pooled)          # odd size: the last row / column belongs to no window
        else:
            ConvBNReLU(P, mid, dc[3], dc[4], out, pooled)
        x = pooled if pooled is not None else out
    y = x
    for j, i in enumerate((3, 2, 1, 0)):
        up = getattr(model, f"up{j + 1}")
        dst = cats[i].slice(C[i], C[i])
        dy_, dx_ = hs[i] - 2 * y.H, ws[i] - 2 * y.W
        if dy_ or dx_:
            tmp = P.act(2 * y.H, 2 * y.W, C[i])
            ConvT2x2(P, y, up.up, tmp)
            PadInto(P, tmp, dst, dy_ // 2, dx_ // 2)
        else:
            ConvT2x2(P, y, up.up, dst)
        dc = up.conv.double_conv
        h, w = hs[i], ws[i]
        mid = P.act(h, w, dc[0].out_channels)
        ConvBNReLU(P, cats[i], dc[0], dc[1], mid)
        y = P.act(h, w, dc[3].out_channels)
        last = ConvBNReLU(P, mid, dc[3], dc[4], y)
    P.head = (Head(P, y, model.outc.conv, fuse=last) if model.outc.conv.out_channels == 1
              else HeadMulti(P, y, model.outc.conv))
    return P.finalize(grad_views)
