"""tcgen05.mma issue-rate probe on a B200: clocks per 128 x N x 16 MMA from resident shared memory.

    python tools/probe_mma_rate.py > gpurun_out/mma_rate.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import build as _lib  # noqa: E402  (tools/probe/libunetk_probe.so, not the product library)


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    iters = 2048
    print(f"{'N':>4s} {'grid':>4s} {'shift':>5s} {'2acc':>4s} {'btiles':>6s} {'clk/MMA':>9s} {'ideal':>6s} {'TFLOP/s/SM@1.9GHz':>18s}")
    for grid in (148,):
        for n in (64, 128, 192, 256):
            for shift in (0, 3):
                for two in (0, 1, 3):
                    if (two + 1) * n > 512:
                        continue
                    for bt in ((1, 4) if n == 256 else (1, 9)):
                        out = torch.zeros(grid, dtype=torch.int64, device=dev)
                        rc = lib.unetk_probe_mma_rate(n, grid, shift, two, iters, bt, out.data_ptr(), stream)
                        if rc != 0:
                            print(n, grid, shift, two, bt, "rc", rc, lib.unetk_last_error().decode())
                            continue
                        torch.cuda.synchronize()
                        mmas = iters * 4 * (1 + two)
                        clk = out.double().mean().item() / mmas
                        ideal = 128 * n / 256
                        tf = 2 * 128 * n * 16 / clk * 1.9e9 / 1e12
                        print(f"{n:4d} {grid:4d} {shift:5d} {two:4d} {bt:6d} {clk:9.1f} {ideal:6.0f} {tf:18.2f}")


if __name__ == "__main__":
    main()
