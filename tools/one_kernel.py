"""Run ONE library call a few times (for `ncu --set full -k regex:... -c 1`):
    python tools/one_kernel.py conv1x1_fwd_stats --n 16 --s 512 --cin 64 --cout 32
    python tools/one_kernel.py conv3x3_fwd_stats | conv3x3_dgrad | conv3x3_wgrad | convT_wgrad ..."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jcfszxc_unet_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--s", type=int, default=512)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(1)
    n, s, cin, cout = a.n, a.s, a.cin, a.cout
    k = 1 if "1x1" in a.what else 3
    x = torch.randn(n, s, s, cin, device=dev, generator=g).bfloat16()
    dy = torch.randn(n, s, s, cout, device=dev, generator=g).bfloat16()
    wt = torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.05
    w_ab, w_ba = ops.pack_weight(wt)
    y = torch.empty(n, s, s, cout, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(n, s, s, cin, device=dev, dtype=torch.bfloat16)
    dw = torch.empty(cout, cin, k, k, device=dev)
    partial = torch.empty(max(lib.unetk_conv_stats_partial_floats(cout), lib.unetk_chan_partial_floats(n * s * s, cout), 4096), device=dev)
    sums = torch.zeros(2 * max(cin, cout), dtype=torch.float64, device=dev)
    ws = torch.empty(lib.unetk_conv_wgrad_workspace(n, s, s, cin, cout, k * k), dtype=torch.uint8, device=dev)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # sub-pixel up-conv (cin -> cout, --s = the LOW-resolution size; y / dy are 2s x 2s)
    up = a.what.startswith("upconv")
    if up:
        w3 = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        w_up, w_up_t = ops.pack_upconv_weight(w3)
        y2 = torch.empty(n, 2 * s, 2 * s, cout, device=dev, dtype=torch.bfloat16)
        dy2 = torch.randn(n, 2 * s, 2 * s, cout, device=dev, generator=g).bfloat16()
        dw3 = torch.empty(cout, cin, 3, 3, device=dev)
        ws_up = torch.empty(lib.unetk_upconv_wgrad_workspace(n, s, s, cin, cout), dtype=torch.uint8, device=dev)
    fn = {
        "upconv_fwd": lambda: ops.upconv_fwd(x, w_up, None, y2, partial, sums),
        "upconv_dgrad": lambda: ops.upconv_dgrad(dy2, w_up_t, dx),
        "upconv_wgrad": lambda: ops.upconv_wgrad(x, dy2, dw3, ws=ws_up),
        "conv1x1_fwd_stats": lambda: ops.conv_fwd_stats(x, w_ab, None, y, partial, sums, 1, 1),
        "conv1x1_fwd": lambda: ops.conv_fwd(x, w_ab, None, y, 1),
        "conv1x1_dgrad": lambda: ops.conv_dgrad(dy, w_ba, dx, 1),
        "conv1x1_wgrad": lambda: ops.conv_wgrad(x, dy, dw, 1, ws=ws),
        "conv3x3_fwd": lambda: ops.conv_fwd(x, w_ab, None, y, 3),
        "bn_stats": lambda: ops.bn_stats(y, partial, sums),
        "conv3x3_fwd_stats": lambda: ops.conv_fwd_stats(x, w_ab, None, y, partial, sums, 3, 1),
        "conv3x3_dgrad": lambda: ops.conv_dgrad(dy, w_ba, dx, 3),
        "conv3x3_wgrad": lambda: ops.conv_wgrad(x, dy, dw, 3, ws=ws),
    }[a.what]
    fn()
    torch.cuda.synchronize()
    start.record()
    for _ in range(a.reps):
        fn()
    end.record()
    torch.cuda.synchronize()
    print(f"{a.what} n={n} s={s} {cin}->{cout}: {start.elapsed_time(end) / a.reps:.3f} ms per call")


if __name__ == "__main__":
    main()
