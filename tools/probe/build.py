"""Builds tools/probe/libunetk_probe.so: the tcgen05 micro-probes (probe.cu) — measurement tools, not product code, and
not part of libunetk.so.  `python tools/probe/build.py` (nvcc cross-compiles sm_100a without a GPU)."""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "jcfszxc_unet_b200" / "csrc"
LIB = HERE / "libunetk_probe.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build() -> Path:
    srcs = [HERE / "probe.cu", HERE / "probe_capi.cu", CSRC / "host_common.cu"]
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", f"-I{CSRC}", "-o", str(LIB), *map(str, srcs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed on the probe library")
    return LIB


def load():
    import ctypes as C

    lib = C.CDLL(os.fspath(build() if not LIB.exists() else LIB))
    vp, i, fp = C.c_void_p, C.c_int, C.c_void_p
    lib.unetk_probe_umma.restype, lib.unetk_probe_umma.argtypes = i, [vp, vp, fp, i, i, i, vp]
    lib.unetk_probe_mma_rate.restype, lib.unetk_probe_mma_rate.argtypes = i, [i, i, i, i, i, i, vp, vp]
    lib.unetk_probe_last_error.restype = C.c_char_p
    lib.unetk_last_error = lib.unetk_probe_last_error
    return lib


if __name__ == "__main__":
    print(build())
