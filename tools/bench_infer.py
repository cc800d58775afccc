"""Eval-mode forward throughput (what evaluate.py:259-275 runs): ours with BatchNorm folded into the conv epilogues, ours
without the fold (UNETK_EVAL_FOLD=0), and stock PyTorch (the UNMODIFIED reference modules from baseline/_ref, cuDNN,
channels_last, bf16 autocast, cudnn.benchmark on) on the same B200.

    python tools/bench_infer.py [--model UNet] [--batch 16] [--size 512] [--steps 20]
Each arm runs in its own process (the reference's `UNetFamily` package cannot share an interpreter with ours)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def arm(args):
    import torch

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(42)
    x = torch.rand(args.batch, 3, args.size, args.size, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    if args.arm == "stock":
        sys.path.insert(0, ROOT)
        import bench

        mods, _ = bench._import_reference()
        mod, cls = bench.REF_CLASSES[args.model]
        torch.manual_seed(42)
        m = getattr(mods[args.model], cls)().to(device=dev, memory_format=torch.channels_last).eval()
        torch.backends.cudnn.benchmark = True

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return m(x)
    else:
        sys.path.insert(0, ROOT)
        import bench
        from jcfszxc_unet_b200 import _lib

        lib = _lib.load()
        torch.manual_seed(42)
        m, _ = bench.make_model(args.model)
        m = m.to(dev).eval()

        def fwd():
            with torch.no_grad():
                return m(x)
    for _ in range(5):
        y = fwd()
    torch.cuda.synchronize()
    launches = None
    if args.arm != "stock":
        c0 = lib.unetk_launch_count()
        fwd()
        launches = lib.unetk_launch_count() - c0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        y = fwd()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    print(json.dumps({"arm": args.arm, "fold": os.environ.get("UNETK_EVAL_FOLD", "1"), "model": args.model, "batch": args.batch,
                      "size": args.size, "ms_per_forward": ms, "images_per_s": args.batch * 1e3 / ms, "launches": launches,
                      "out_mean": float((y[-1] if isinstance(y, (list, tuple)) else y).float().mean())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="UNet")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--arm", default=None)
    args = ap.parse_args()
    if args.arm:
        return arm(args)
    out = {}
    for name, a, env in (("ours_folded", "ours", {"UNETK_EVAL_FOLD": "1"}), ("ours_unfolded", "ours", {"UNETK_EVAL_FOLD": "0"}),
                         ("stock_cudnn", "stock", {})):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--arm", a, "--model", args.model, "--batch", str(args.batch),
                            "--size", str(args.size), "--steps", str(args.steps)], capture_output=True, text=True,
                           env={**os.environ, **env}, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        out[name] = json.loads(lines[-1]) if lines else {"error": r.stderr[-300:]}
    if "images_per_s" in out.get("ours_folded", {}) and "images_per_s" in out.get("stock_cudnn", {}):
        out["folded_over_stock"] = out["ours_folded"]["images_per_s"] / out["stock_cudnn"]["images_per_s"]
        out["folded_over_unfolded"] = out["ours_folded"]["images_per_s"] / out["ours_unfolded"]["images_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
