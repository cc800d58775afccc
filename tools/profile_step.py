"""Per-call timing of ONE eager UNet training step (CUDA events around every C-ABI call).

    python tools/profile_step.py [--batch 16] [--size 512] [--warmup 2] > gpurun_out/step_profile.txt

Also the command to wrap in `ncu --metrics gpu__time_duration.sum` for the launch list under profiles/.
"""
import argparse
import os
import sys

import torch

os.environ.setdefault("UNETK_WGRAD_STREAM", "0")   # one kernel at a time: clean per-call durations
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONV_FLOPS, make_model  # noqa: E402
from jcfszxc_unet_b200 import _lib  # noqa: E402
from jcfszxc_unet_b200.trainer import Trainer  # noqa: E402
from UNetFamily.UNet import UNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--model", default="UNet")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    model, builder = make_model(a.model)
    model = model.to(dev).train()
    tr = Trainer(model, lr=1e-6, use_cuda_graph=False, builder=builder)
    g = torch.Generator(device=dev).manual_seed(42)
    images = torch.rand(a.batch, 3, a.size, a.size, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    labels = (torch.rand(a.batch, 1, a.size, a.size, device=dev, generator=g) < 0.12).float()
    for _ in range(a.warmup):
        tr.step(images, labels)
    torch.cuda.synchronize()
    with _lib.profile_calls() as prof:
        tr.step(images, labels)
    torch.cuda.synchronize()
    total = 0.0
    print(f"{'call':28s} {'N':>3s} {'H':>4s} {'W':>4s} {'Cin':>5s} {'Cout':>5s} {'ms':>8s} {'TFLOP/s':>8s}")
    for name, args, s, e in prof.records:
        ms = s.elapsed_time(e)
        total += ms
        if name in CONV_FLOPS:
            i, taps, _ = CONV_FLOPS[name]
            n, h, w, cin, cout = args[i:i + 5]
            tf = 2.0 * n * h * w * cin * cout * taps / (ms / 1e3) / 1e12
            print(f"{name[6:]:28s} {n:3d} {h:4d} {w:4d} {cin:5d} {cout:5d} {ms:8.3f} {tf:8.1f}")
        else:
            dims = [x for x in args if isinstance(x, int) and 0 < x < 5000][:5]
            print(f"{name[6:]:28s} {' '.join(f'{d:>5d}' for d in dims):28s} {ms:8.3f}")
    print(f"total {total:.3f} ms over {len(prof.records)} calls")
    for name, (calls, ms) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:30s} {calls:4d} calls {ms:9.3f} ms")


if __name__ == "__main__":
    main()
