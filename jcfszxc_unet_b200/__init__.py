"""B200-native U-Net convolutional hot path (hand-written sm_100a CUDA behind a C ABI)."""
__version__ = "0.1.0"
