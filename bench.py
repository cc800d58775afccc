#!/usr/bin/env python
"""Headline benchmark: vanilla UNet bf16 TRAINING images/sec at 3x512x512 (BASELINE.json `metric`,
configs[1]: batch 16 per B200) on N GPUs of one node, plus the roofline of the dominant kernel and the
reference's CPU path timed beside it.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm (N=1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W     # our arm, data parallel
    python bench.py --impl reference --steps K --warmup W          # the reference's CPU implementation

One step = forward + 0.5*BCE+0.5*dice + backward + clip_grad_norm_(1.0) + RMSprop over one synthetic batch
(reference train.py:255-301).  Rank 0 prints exactly ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "UNet train images/sec @512^2 bf16"
UNIT = "images/s"
SEED = 42            # reference default, train.py:476
LABEL_DENSITY = 0.12  # vessel-like sparsity of the synthetic masks


def unet_train_gflop_per_image(size: int) -> float:
    """Algorithmic FLOPs (2*MAC) of one training image: fwd + dgrad + wgrad of every Conv2d /
    ConvTranspose2d, minus the dgrad of the stem (1155.1 GFLOP at 512^2)."""
    w = (64, 128, 256, 512, 1024)
    fwd = 0.0
    stem = 2.0 * size * size * 3 * w[0] * 9
    fwd += stem + 2.0 * size * size * w[0] * w[0] * 9
    for i in range(1, 5):
        px = (size >> i) ** 2
        fwd += 2.0 * px * w[i - 1] * w[i] * 9 + 2.0 * px * w[i] * w[i] * 9
    for i in (3, 2, 1, 0):
        px = (size >> i) ** 2
        fwd += 2.0 * (px / 4) * w[i + 1] * w[i] * 4          # ConvTranspose 2x2/s2
        fwd += 2.0 * px * (2 * w[i]) * w[i] * 9 + 2.0 * px * w[i] * w[i] * 9
    fwd += 2.0 * size * size * w[0] * 1
    return (3.0 * fwd - stem) / 1e9


# model name -> (module, class, plan builder, train GFLOP per image at 512^2 measured on the reference)
MODELS = {
    "UNet": ("UNetFamily.UNet", "UNet", "engine.build_unet_plan", 1155.1),
    "AttentionUNet": ("UNetFamily.AttentionUNet", "AttentionUNet", "builders.build_attention_unet_plan", 1593.3),
    "R2UNet": ("UNetFamily.R2UNet", "R2UNet", "builders.build_r2unet_plan", 3659.6),
    "ResUNet": ("UNetFamily.ResUNet", "ResUNet", "builders.build_resunet_plan", 1704.5),
    "NestedUNet": ("UNetFamily.UNetPP", "NestedUNet", "builders.build_nested_unet_plan", 3308.8 / 4.0),
}


def make_model(name):
    import importlib

    from jcfszxc_unet_b200 import builders, engine

    mod, cls, builder, _ = MODELS[name]
    ns = {"engine": engine, "builders": builders}
    bmod, bfn = builder.split(".")
    return getattr(importlib.import_module(mod), cls)(), getattr(ns[bmod], bfn)


def train_gflop_per_image(name: str, size: int) -> float:
    if name == "UNet":
        return unet_train_gflop_per_image(size)
    return MODELS[name][3] * (size / 512.0) ** 2   # every layer is a convolution: FLOPs scale with the pixel count


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1388.9), "bf16_tflops": d.get("bf16_tflops", 1643.8),
                "hbm_gbs": d.get("hbm_gbs", 6532.2), "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = f"/tmp/unetk_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------
# The reference's own implementation: `--impl reference` (host cores) and `--impl torch_cuda` (cuDNN on the same B200)
# ----------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
REF_CLASSES = {"UNet": ("UNet", "UNet"), "AttentionUNet": ("AttentionUNet", "AttentionUNet"), "R2UNet": ("R2UNet", "R2UNet"),
               "ResUNet": ("ResUNet", "ResUNet"), "NestedUNet": ("UNetPP", "NestedUNet")}


def _import_reference():
    """The UNMODIFIED reference modules from baseline/_ref (baseline/install_ref.py), or None when it is not installed.
    The reference's package is called `UNetFamily`, like this repository's drop-in: a process uses ONE of them, so the
    repository root leaves sys.path here (the reference arms never touch our package or our kernels)."""
    import importlib
    import types

    import torch

    if not os.path.isdir(os.path.join(REF_DIR, "UNetFamily")):
        return None
    if "UNetFamily" in sys.modules and not getattr(sys.modules["UNetFamily"], "__path__", [""])[0].startswith(REF_DIR):
        raise RuntimeError("the reference arm must not share a process with this repository's UNetFamily package")
    sys.dont_write_bytecode = True
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    _t, _l = types.ModuleType("timm"), types.ModuleType("timm.layers")     # unet_parts.py:14
    _l.trunc_normal_ = torch.nn.init.trunc_normal_
    _t.layers = _l
    sys.modules.setdefault("timm", _t)
    sys.modules.setdefault("timm.layers", _l)
    sys.path.insert(0, REF_DIR)
    mods = {name: importlib.import_module(f"UNetFamily.{mod}") for name, (mod, _) in REF_CLASSES.items()}
    dice = importlib.import_module("utils.dice_score")
    return mods, dice


def make_reference_step(model_name: str, device, size: int, batch: int):
    """One iteration of train.py:255-301 around the reference's own modules (bf16 autocast as BASELINE.json asks;
    GradScaler is a no-op for bf16; the NaN probes and loss.item() of the reference loop — host syncs — are left out,
    which only helps the baseline).  Returns (kind, step_fn): kind "reference" = the unmodified modules from
    baseline/_ref, "port" = oracle/unet_oracle.py's pinned functional restatement when baseline/_ref is absent."""
    import torch

    ref = _import_reference()
    torch.manual_seed(SEED)
    g = torch.Generator().manual_seed(SEED)
    images = torch.rand(batch, 3, size, size, generator=g)
    labels = (torch.rand(batch, 1, size, size, generator=g) < LABEL_DENSITY).float()
    images = images.to(device=device, dtype=torch.float32, memory_format=torch.channels_last)   # train.py:248-253
    labels = labels.to(device=device, dtype=torch.float32)
    dt = torch.device(device).type
    if ref is not None:
        mods, dice = ref
        mod, cls = REF_CLASSES[model_name]
        model = getattr(mods[model_name], cls)().to(device=device, memory_format=torch.channels_last).train()  # train.py:523-525
        opt = torch.optim.RMSprop(model.parameters(), lr=1e-6, weight_decay=1e-8, momentum=0.999)             # train.py:107-112
        criterion = torch.nn.BCEWithLogitsLoss()                                                             # train.py:124

        def step():
            with torch.autocast(dt, dtype=torch.bfloat16):
                masks_pred = model(images)
                masks_pred_sigmoid = torch.sigmoid(masks_pred)
                bce_loss = criterion(masks_pred, labels)
                d = dice.dice_loss(masks_pred_sigmoid.squeeze(1), labels.squeeze(1), multiclass=False)
                loss = 0.5 * bce_loss + 0.5 * d
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return loss

        return "reference", step
    from oracle import unet_oracle as O

    model, _ = make_model(model_name)
    sd = {k: v.detach().clone().to(device) for k, v in model.state_dict().items()}
    names = O.param_names(sd)
    for k in names:
        if sd[k].dim() == 4:
            sd[k] = sd[k].contiguous(memory_format=torch.channels_last)
        sd[k].requires_grad_(True)
    params = [sd[k] for k in names]
    opt = torch.optim.RMSprop(params, lr=1e-6, weight_decay=1e-8, momentum=0.999)

    def step():
        _, loss, _, _ = O.forward_loss(sd, images, labels, bf16=True, training=True, model=model_name)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    return "port", step


def run_reference_arm(args):
    """The reference's CPU implementation of the step on the box's host cores, all threads, a bounded sample per step."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, step = make_reference_step(args.model, "cpu", args.size, batch)
    warm = max(1, args.warmup) if args.budget is None else 1
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        step()
        done += 1
        if args.budget is not None and time.perf_counter() - t0 > args.budget:
            break
    dt = time.perf_counter() - t0
    ips, ms = batch * done / dt, 1e3 * dt / done
    threads = torch.get_num_threads()
    what = ("UNMODIFIED reference modules (baseline/_ref) in the loop body of train.py:255-301" if kind == "reference" else
            "oracle port of train.py:255-301 (the reference modules' arithmetic via torch.nn.functional)")
    sample = (f"{what}, bf16 autocast, {args.model}, batch {batch} of 3x{args.size}x{args.size} per step, "
              f"{warm} warm-up + {done} timed steps, {threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{_workload(args.model, args.size)}; CPU arm runs a bounded sample: batch {batch} per step",
                   "sample_batch": batch},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


def run_torch_cuda_arm(args):
    """Stock PyTorch on the SAME B200: the reference's real GPU path (train.py:248-256,525: channels_last + autocast, cuDNN
    underneath), same model / batch / seeds / step definition as our arm; timed with cudnn.benchmark off + deterministic
    (what the reference sets, utils/utils.py:30-31) and on (the strongest stock setting).  `value` is the faster one."""
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "torch_cuda", "unavailable": "no CUDA device"}), file=_OUT, flush=True)
        return 0
    dev = torch.device("cuda", args.device_index)
    torch.cuda.set_device(dev)
    B, S, K, W = args.batch, args.size, args.steps, max(3, args.warmup)
    out = {}
    kind = None
    settings = (("cudnn_benchmark_on", True),) if args.cudnn_on_only else (("cudnn_reference_settings", False), ("cudnn_benchmark_on", True))
    for label, bench_on in settings:
        torch.backends.cudnn.benchmark = bench_on
        torch.backends.cudnn.deterministic = not bench_on
        kind, step = make_reference_step(args.model, dev, S, B)
        for _ in range(W):
            step()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(K):
            loss = step()
        end.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end) / K
        out[label] = {"images_per_s": B * 1e3 / ms, "ms_per_step": ms, "loss": float(loss),
                      "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
        del step
        torch.cuda.empty_cache()
    best = max(out.values(), key=lambda v: v["images_per_s"])
    line = {
        "impl": "torch_cuda", "metric": METRIC, "value": best["images_per_s"], "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": _workload(args.model, S), "per_gpu_batch": B, "image": f"3x{S}x{S}"},
        "kind": ("UNMODIFIED reference modules (baseline/_ref)" if kind == "reference" else "oracle port (torch.nn.functional)")
                + f", stock torch {torch.__version__} + cuDNN {torch.backends.cudnn.version()}, channels_last, bf16 autocast, "
                  "torch.optim.RMSprop + clip_grad_norm_, no host sync inside a step",
        "settings": out,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


def _workload(model: str, size: int) -> str:
    if model == "UNet":
        return f"vanilla UNet(n_channels=3, n_classes=1) bf16 training step (BCE+dice, clip 1.0, RMSprop), 3x{size}x{size} synthetic"
    return f"{model} bf16 training step (BCE+dice, clip 1.0, RMSprop), 3x{size}x{size} synthetic"


def _child_json(argv, timeout):
    """Run another arm of this file in a child process (the reference arms must not share a process with our package)
    and parse its single JSON line; a failure is reported, never hidden."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), *argv], capture_output=True, text=True, timeout=timeout,
                           cwd=ROOT, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout}s"}
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    if r.returncode != 0 or not lines:
        return {"unavailable": (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    return json.loads(lines[-1])


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
CONV_FLOPS = {
    # name -> (index of N in args, taps, family)
    "unetk_conv3x3_fwd": (6, 9, "tap_gemm"), "unetk_conv3x3_dgrad": (6, 9, "tap_gemm"),
    "unetk_conv3x3_dgrad_colsum": (7, 9, "tap_gemm"),
    "unetk_conv3x3_fwd_bnstats": (8, 9, "tap_gemm"), "unetk_conv1x1_fwd_bnstats": (8, 1, "tap_gemm"),
    "unetk_conv3x3s2_fwd": (8, 9, "tap_gemm"), "unetk_conv3x3s2_dgrad": (6, 9, "tap_gemm"),
    "unetk_conv1x1_fwd": (6, 1, "tap_gemm"), "unetk_conv1x1_dgrad": (6, 1, "tap_gemm"),
    "unetk_convT2x2_fwd": (6, 4, "tap_gemm"), "unetk_convT2x2_dgrad": (6, 4, "tap_gemm"),
    "unetk_conv3x3_wgrad": (6, 9, "wgrad"), "unetk_conv1x1_wgrad": (6, 1, "wgrad"), "unetk_convT2x2_wgrad": (6, 4, "wgrad"),
    "unetk_conv3x3s2_wgrad": (6, 9, "wgrad"),
    # sub-pixel up-conv (H, W = the low-resolution size): 16 taps EXECUTED per low-resolution pixel; the reference's
    # formulation of the same result (conv3x3 on the 2x up-sampled tensor) is 36 — the kernel tables count what ran
    "unetk_upconv3x3_fwd": (8, 16, "tap_gemm"), "unetk_upconv3x3_dgrad": (6, 16, "tap_gemm"),
    "unetk_upconv3x3_wgrad": (6, 16, "wgrad"),
}


def ncu_traffic_per_launch(family, args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, from the committed ncu
    capture of this exact workload (profiles/r02_ncu_traffic.json, made by tools/ncu_traffic.py); None for any other
    workload (a number taken under a profiler is never measured live)."""
    if args.model != "UNet" or args.batch != 16 or args.size != 512:
        return None
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):     # the latest capture of this workload
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)[family]["bytes_per_launch"]
        except Exception:
            continue
    return None


def kernel_breakdown(records):
    """Group the per-call CUDA-event timings of one eager step into kernel families."""
    fam = {}
    for name, a, s, e in records:
        ms = s.elapsed_time(e)
        if name in CONV_FLOPS:
            i, taps, family = CONV_FLOPS[name]
            n, h, w, cin, cout = a[i:i + 5]
            flops = 2.0 * n * h * w * cin * cout * taps
        else:
            family, flops = name.replace("unetk_", ""), 0.0
        slot = fam.setdefault(family, {"calls": 0, "ms": 0.0, "flops": 0.0})
        slot["calls"] += 1
        slot["ms"] += ms
        slot["flops"] += flops
    return fam


def _free_cuda():
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()


def measure_ours(model_name, B, S, K, W, dev, dp, world, rank, local_rank, graph=True, detail=True, e2e=True):
    """Build the model + Trainer, warm up, and time K steps of our arm.  Returns a dict of raw measurements (every rank)."""
    import torch
    import torch.distributed as dist

    from jcfszxc_unet_b200 import _lib
    from jcfszxc_unet_b200.trainer import Trainer

    lib = _lib.load()
    torch.manual_seed(SEED)
    model, builder = make_model(model_name)
    model = model.to(dev).train()
    tr = Trainer(model, lr=1e-6, use_cuda_graph=graph, dp=dp, builder=builder)   # lr: reference default train.py:434
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    images = torch.rand(B, 3, S, S, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    labels = (torch.rand(B, 1, S, S, device=dev, generator=g) < LABEL_DENSITY).float()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (step 0 eager, graphs captured at step 1) -----------------------------------------
    for _ in range(W):
        tr.step(images, labels)
    torch.cuda.synchronize()
    out = {"trainer": tr}

    # ---- per-kernel timing + launch count: one EAGER step, every C-ABI call bracketed by CUDA events --
    graphs, use_graph, tr.graphs, tr.use_graph = tr.graphs, tr.use_graph, None, False
    overlap, tr.plan.overlap_wgrad = tr.plan.overlap_wgrad, False   # per-kernel durations: no side-stream co-scheduling
    with _lib.profile_calls() as prof:
        c0 = lib.unetk_launch_count()
        tr.step(images, labels)
        out["launches_per_step"] = lib.unetk_launch_count() - c0
    torch.cuda.synchronize()
    out["families"] = kernel_breakdown(prof.records) if detail else None
    tr.plan.overlap_wgrad = overlap
    tr.graphs, tr.use_graph = graphs, use_graph
    tr.step(images, labels)
    torch.cuda.synchronize()

    # ---- headline: K steps, inputs resident in HBM ---------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(K):
        tr.step(images, labels)
    end.record()
    barrier()
    out["ms_total"] = dp.max_over_ranks(start.elapsed_time(end), dev)
    out["clocks"] = sampler.stop() if sampler else None
    out["loss"] = float(tr.loss_terms()[0])
    if not e2e:
        return out

    # ---- end to end: host (pinned) inputs copied in every step, loss read back every step --------------
    h_images = torch.empty((B, S, S, 3), dtype=torch.float32, pin_memory=True).permute(0, 3, 1, 2)  # channels_last, pinned
    h_images.copy_(images.cpu())
    h_labels = torch.empty((B, 1, S, S), dtype=torch.float32, pin_memory=True)
    h_labels.copy_(labels.cpu())
    float(tr.step(h_images, h_labels))
    for _ in range(3):                                     # warm the pipeline itself (copy stream, staging batches)
        tr.prefetch(h_images, h_labels)
        float(tr.step())
    e2e_steps = max(3, min(K, 20))
    # Input pipeline of the public API: the H2D copy of step i+1 (pinned host memory -> staging batch, side stream) is
    # issued right after step i is launched, so it overlaps step i's compute; EVERY step still copies its own 64 MB
    # from the host inside the timed region and its loss is read back to the host before the next step is launched.
    barrier()
    start.record()
    tr.prefetch(h_images, h_labels)
    for i in range(e2e_steps):
        loss_dev = tr.step()                               # trains on the staged batch
        if i + 1 < e2e_steps:
            tr.prefetch(h_images, h_labels)                # next step's inputs: host -> device while this step runs
        loss_host = float(loss_dev)                        # D2H read of the step's result: synchronises every step
    end.record()
    barrier()
    out["e2e_ms"] = dp.max_over_ranks(start.elapsed_time(end), dev)
    out["e2e_steps"] = e2e_steps
    out["h2d"] = h_images.numel() * 4 + h_labels.numel() * 4
    out["loss_e2e"] = loss_host
    return out


def replicas_in_sync(tr, dp, dev):
    """Data parallel: every rank's parameters must be bit-identical after the timed steps (same reduced gradients, same
    optimizer arithmetic).  Compares an integer checksum of the flat fp32 parameter buffer across ranks."""
    import torch
    import torch.distributed as dist

    bits = tr.flat_p.view(torch.int32).to(torch.int64)
    mine = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 65521 + 1)).sum()])
    if not dp.enabled:
        return True
    all_ = [torch.zeros_like(mine) for _ in range(dp.world)]
    dist.all_gather(all_, mine)
    return all(bool(torch.equal(a, all_[0])) for a in all_)


# BASELINE.json configs[2..4] at their per-GPU batches on the 8-GPU box they are quoted on
VARIANTS = (("AttentionUNet", 16, 512), ("R2UNet", 8, 512), ("ResUNet", 8, 512), ("NestedUNet", 4, 1024))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jcfszxc_unet_b200 import _lib
    from jcfszxc_unet_b200.dp import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the U-Net hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the all-reduce kernels get the SMs the trainer leaves them (Trainer.sm_reserve) and no more
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("UNETK_DP_SM_RESERVE", "4"))
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    S = args.size
    strong = args.scaling == "strong"
    if strong:
        gb = args.global_batch or args.batch
        if gb % world:
            raise ValueError(f"--global-batch {gb} is not divisible by {world} ranks")
        B = gb // world
    else:
        B = args.batch
    W, K = max(3, args.warmup), args.steps
    dp = DataParallel()
    m = measure_ours(args.model, B, S, K, W, dev, dp, world, rank, local_rank, graph=not args.no_graph)
    tr = m.pop("trainer")
    in_sync = replicas_in_sync(tr, dp, dev) if world > 1 else None
    buckets = getattr(tr, "bucket_report", lambda: None)()
    tr.close()          # the captured graph holds NCCL kernels: release it before the process group is destroyed
    del tr
    _free_cuda()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fam = m["families"]
    ms_total, e2e_ms, e2e_steps = m["ms_total"], m["e2e_ms"], m["e2e_steps"]
    launches_per_step = m["launches_per_step"]
    ms_per_step = ms_total / K
    value = world * B * K / (ms_total / 1e3)
    e2e_value = world * B * e2e_steps / (e2e_ms / 1e3)
    peaks = measured_peaks()
    gflop_img = train_gflop_per_image(args.model, S)
    # dominant kernel family (by time inside the step) and its live roofline
    conv_fams = {k: v for k, v in fam.items() if v["flops"] > 0}
    dom_name = max(conv_fams, key=lambda k: conv_fams[k]["ms"])
    dom = conv_fams[dom_name]
    achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
    step_ms_eager = sum(v["ms"] for v in fam.values())
    roofline = {
        "bound": "tensor", "kernel": {"tap_gemm": "conv_gemm_kernel / conv3x3_halo_kernel / conv3x3_rows_kernel (conv3x3 and 1x1 fwd/dgrad, ConvTranspose fwd/dgrad; BN = 256 and the 128-column halo layers as CTA pairs, tcgen05 cta_group::2)",
                                      "wgrad": "wgrad3x3_kernel / wgrad_kernel (+ordered reduce)"}[dom_name],
        "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": ncu_traffic_per_launch(dom_name, args),
        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
        "launches": dom["calls"], "avg_launch_ms": dom["ms"] / dom["calls"],
        "share_of_step": dom["ms"] / step_ms_eager,
        # per GPU: images/s of ONE rank x algorithmic GFLOP per training image, against one GPU's peak
        "whole_step": {"achieved": value / world * gflop_img / 1e3,
                       "frac": value / world * gflop_img / 1e3 / peaks["bf16_tflops_sustained"],
                       "gflop_per_image": gflop_img, "per": "gpu"},
        "families": {k: {"calls": v["calls"], "ms": round(v["ms"], 3),
                         **({"tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)} if v["flops"] else {})}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
    }
    solo = world == 1
    cpu = None
    if solo and not args.no_cpu_baseline and args.model == "UNet":
        # the reference's CPU implementation on this box's host cores: a child process (it imports the reference's own
        # `UNetFamily` package, which cannot share an interpreter with ours), bounded to ~25 s of CPU work
        d = _child_json(["--impl", "reference", "--steps", "3", "--warmup", "1", "--budget", "25", "--size", str(S)], 400)
        cpu = d.get("cpu_baseline") or {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": d.get("unavailable")}
    gpu_base = None
    if solo and not args.no_gpu_baseline:
        # stock PyTorch (cuDNN, channels_last, bf16 autocast) running the reference's modules on this same B200
        d = _child_json(["--impl", "torch_cuda", "--model", args.model, "--batch", str(B), "--size", str(S), "--steps", "10",
                         "--warmup", "3", "--device-index", str(local_rank)], 600)
        gpu_base = _gpu_baseline_block(d, value)
    variants = None
    if solo and args.variants and args.model == "UNet" and not strong:
        variants = {}
        for name, vb, vs in VARIANTS:
            try:
                vm = measure_ours(name, vb, vs, 5, 3, dev, dp, 1, 0, local_rank, graph=True, detail=False, e2e=False)
                vm.pop("trainer").close()
                _free_cuda()
                ips = vb * 5 / (vm["ms_total"] / 1e3)
                gf = train_gflop_per_image(name, vs)
                slot = {"per_gpu_batch": vb, "image": f"3x{vs}x{vs}", "value": ips, "unit": UNIT, "ms_per_step": vm["ms_total"] / 5,
                        "steps": 5, "warmup": 3, "whole_step_frac": ips * gf / 1e3 / peaks["bf16_tflops_sustained"],
                        "gflop_per_image": gf, "launches_per_step": int(vm["launches_per_step"]), "loss": vm["loss"]}
                if not args.no_gpu_baseline:
                    d = _child_json(["--impl", "torch_cuda", "--model", name, "--batch", str(vb), "--size", str(vs), "--steps", "5",
                                     "--warmup", "3", "--cudnn-on-only", "--device-index", str(local_rank)], 600)
                    slot["gpu_baseline"] = _gpu_baseline_block(d, ips)
                variants[name] = slot
            except Exception as e:  # a variant that fails is reported as such, the headline line still prints
                variants[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                _free_cuda()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": _workload(args.model, S) + f", batch {B} per GPU" + (" (BASELINE.json configs[1])" if args.model == "UNet" and B == 16 and S == 512 else ""),
                   "per_gpu_batch": B, "global_batch": B * world, "image": f"3x{S}x{S}", "parallelism": f"dp{world}",
                   "batchnorm": "per-rank batch statistics" if world > 1 else "batch statistics",
                   "cuda_graph": not args.no_graph,
                   "l2": "no explicit flush: every step streams ~17 GB of activations/gradients, far beyond the 126 MB L2"},
        "clocks": m["clocks"],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": int(launches_per_step * K),
        "launches_per_step": int(launches_per_step),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gpu_baseline": gpu_base,
        "variants": variants,
        "replicas_in_sync": in_sync,
        "grad_buckets": buckets,
        "loss": m["loss"], "loss_e2e": m["loss_e2e"],
    }
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _gpu_baseline_block(d, ours_value):
    if "unavailable" in d or "value" not in d:
        return {"value": None, "unit": UNIT, "unavailable": d.get("unavailable", "no result")}
    return {"value": d["value"], "unit": UNIT, "ms_per_step": d["ms_per_step"], "kind": d["kind"], "settings": d["settings"],
            "steps": d["steps"], "warmup": d["warmup"], "ours_over_stock": ours_value / d["value"]}


_OUT = sys.stdout


def _only_json_on_stdout():
    """The contract is ONE JSON line on stdout: native libraries print there too (NCCL's version banner when the box
    sets NCCL_DEBUG), so fd 1 is pointed at stderr for the run and the line goes to a duplicate of the real stdout."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _only_json_on_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (BASELINE.json configs[1]: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--model", default="UNet", choices=sorted(MODELS), help="default: the headline config (vanilla UNet)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-CUDA (cuDNN) leg")
    ap.add_argument("--no-variants", dest="variants", action="store_false",
                    help="skip the short runs of BASELINE.json configs[2..4] (AttentionUNet, R2UNet, ResUNet, NestedUNet)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --global-batch is split over the ranks (BASELINE.json configs[2]: AttentionUNet, 16)")
    ap.add_argument("--global-batch", type=int, default=None)
    ap.add_argument("--ref-batch", type=int, default=2, help="reference CPU arm: images per step of its bounded sample")
    ap.add_argument("--budget", type=float, default=None, help="reference CPU arm: stop after this many seconds")
    ap.add_argument("--device-index", type=int, default=0)
    ap.add_argument("--cudnn-on-only", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "torch_cuda":
        return run_torch_cuda_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
