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
# CPU baseline (the oracle port of the reference step) — also the `--impl reference` arm
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(size: int, batch: int, steps: int, warmup: int, budget_s: float | None = None):
    """Times oracle.train_step (train.py:255-301 restated, bf16 autocast) on the host cores.
    Returns (images_per_s, ms_per_step, steps_done, cores)."""
    import torch

    from oracle import unet_oracle as O
    from UNetFamily.UNet import UNet

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    model = UNet(3, 1)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    names = O.param_names(sd)
    opt_state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in names}
    g = torch.Generator().manual_seed(SEED)
    images = torch.rand(batch, 3, size, size, generator=g).contiguous(memory_format=torch.channels_last)
    labels = (torch.rand(batch, 1, size, size, generator=g) < LABEL_DENSITY).float()
    for _ in range(warmup):
        O.train_step(sd, opt_state, images, labels, 1e-6, bf16=True)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        O.train_step(sd, opt_state, images, labels, 1e-6, bf16=True)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, 1e3 * dt / done, done, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = 2
    ips, ms, done, cores = cpu_reference_step_rate(args.size, batch, args.steps, max(1, args.warmup))
    sample = (f"oracle port of train.py:255-301 (bf16 autocast, reference modules' arithmetic via torch.nn.functional), "
              f"batch {batch} of 3x{args.size}x{args.size} per step, {done} timed steps, {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"vanilla UNet(3,1) bf16 training step (BCE+dice, clip 1.0, RMSprop), 3x{args.size}x{args.size} synthetic; "
                               f"CPU arm runs a bounded sample: batch {batch} per step", "sample_batch": batch},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


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
}


def ncu_traffic_per_launch(family, args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, from the committed ncu
    capture of this exact workload (profiles/r01_ncu_traffic.json, made by tools/ncu_traffic.py); None for any other
    workload (a number taken under a profiler is never measured live)."""
    if args.model != "UNet" or args.batch != 16 or args.size != 512:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            return json.load(f)[family]["bytes_per_launch"]
    except Exception:
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


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jcfszxc_unet_b200 import _lib
    from jcfszxc_unet_b200.dp import DataParallel
    from jcfszxc_unet_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the U-Net hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, S = args.batch, args.size
    W, K = max(3, args.warmup), args.steps

    torch.manual_seed(SEED)
    model, builder = make_model(args.model)
    model = model.to(dev).train()
    dp = DataParallel()
    tr = Trainer(model, lr=1e-6, use_cuda_graph=not args.no_graph, dp=dp, builder=builder)   # lr: reference default train.py:434
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
    launches_before = lib.unetk_launch_count()

    # ---- per-kernel timing + launch count: one EAGER step, every C-ABI call bracketed by CUDA events --
    graphs, use_graph, tr.graphs, tr.use_graph = tr.graphs, tr.use_graph, None, False
    overlap, tr.plan.overlap_wgrad = tr.plan.overlap_wgrad, False   # per-kernel durations: no side-stream co-scheduling
    with _lib.profile_calls() as prof:
        c0 = lib.unetk_launch_count()
        tr.step(images, labels)
        launches_per_step = lib.unetk_launch_count() - c0
    torch.cuda.synchronize()
    fam = kernel_breakdown(prof.records)
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
    ms_total = dp.max_over_ranks(start.elapsed_time(end), dev)
    clocks = sampler.stop() if sampler else None
    loss_final = float(tr.loss_terms()[0])

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
    e2e_ms = dp.max_over_ranks(start.elapsed_time(end), dev)
    h2d = h_images.numel() * 4 + h_labels.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

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
        "bound": "tensor", "kernel": {"tap_gemm": "conv_gemm_kernel / conv3x3_halo_kernel (conv3x3 and 1x1 fwd/dgrad, ConvTranspose fwd/dgrad)",
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
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.model == "UNet":
        ips, ms, done, cores = cpu_reference_step_rate(S, 2, 3, 1, budget_s=25.0)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle port of train.py:255-301 (bf16 autocast) on the host, batch 2 of 3x{S}x{S}, 1 warm-up + {done} timed steps"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"vanilla UNet(n_channels=3, n_classes=1) bf16 training step (BCE+dice, clip 1.0, RMSprop), "
                                f"batch {B} per GPU, 3x{S}x{S} synthetic (BASELINE.json configs[1])") if args.model == "UNet" else
                               (f"{args.model} bf16 training step (BCE+dice, clip 1.0, RMSprop), batch {B} per GPU, "
                                f"3x{S}x{S} synthetic"),
                   "per_gpu_batch": B, "global_batch": B * world, "image": f"3x{S}x{S}", "parallelism": f"dp{world}",
                   "batchnorm": "per-rank batch statistics" if world > 1 else "batch statistics",
                   "cuda_graph": not args.no_graph,
                   "l2": "no explicit flush: every step streams ~17 GB of activations/gradients, far beyond the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": int(launches_per_step * K),
        "launches_per_step": int(launches_per_step),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "loss": loss_final, "loss_e2e": loss_host,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (BASELINE.json configs[1]: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--model", default="UNet", choices=sorted(MODELS), help="default: the headline config (vanilla UNet)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
