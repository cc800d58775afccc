"""B200-native U-Net convolutional hot path (hand-written sm_100a CUDA behind a C ABI)."""
from __future__ import annotations

import contextlib

__version__ = "0.1.0"

_precision = "bf16"


def get_precision() -> str:
    return _precision


@contextlib.contextmanager
def precision(mode: str):
    """Arithmetic of `model(x)` inside the block: "bf16" (default: what the reference runs under torch.autocast on
    GPU, train.py:255) or "fp32" (the reference without autocast; forward only, vanilla UNet; see f32.py)."""
    global _precision
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    prev, _precision = _precision, mode
    try:
        yield
    finally:
        _precision = prev


def clear_plans(module) -> None:
    """Free the cached execution plans (activation / gradient buffers) of `module` and its sub-modules."""
    from .bridge import clear_plans as _clear

    _clear(module)
