"""Run the tcgen05 descriptor-semantics probe on a B200 and print which variants are exact.

Usage (GPU box): python tools/probe_umma.py > gpurun_out/probe.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import build as _lib  # noqa: E402  (tools/probe/libunetk_probe.so, not the product library)


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    lib = _lib.load()
    rows = 144
    b = torch.randn(64, 64, device=dev).bfloat16()
    stream = torch.cuda.current_stream().cuda_stream
    for mode in (0, 1, 2):
        cols = 128 if mode == 2 else 64
        a = torch.randn(rows, cols, device=dev).bfloat16()
        for shift in (0, 1, 2, 3, 4, 7, 8, 9, 16):
            for bo in ((0, shift & 7) if mode != 1 else (0,)):
                d = torch.full((128, 64), float("nan"), device=dev, dtype=torch.float32)
                rc = lib.unetk_probe_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), mode, shift, bo, stream)
                if rc != 0:
                    print(f"mode {mode} shift {shift} bo {bo}: rc={rc} {lib.unetk_last_error().decode()}")
                    continue
                torch.cuda.synchronize()
                af, bf = a.float(), b.float()
                if mode in (0, 1):
                    ref = af[shift:shift + 128] @ bf.t()           # D[i][n] = sum_k A[i+shift][k] B[n][k]
                else:
                    ref = af[shift:shift + 64].t() @ bf.t()        # D[m][n] = sum_k A[k+shift][m] B[n][k]
                err = (d - ref).abs().max().item()
                print(f"mode {mode} shift {shift:2d} bo {bo}: max|err| = {err:.4g}  {'OK' if err < 1e-2 else 'MISMATCH'}")
                if bo == (shift & 7):
                    break


if __name__ == "__main__":
    main()
