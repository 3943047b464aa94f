"""gomavatar_b200 — B200-native hot path of GoMAvatar (LBS -> per-face frame -> Gaussian mean/cov ->
tile splat rasterizer fwd/bwd -> photometric losses) behind the reference's own interfaces.

The compute path is hand-written sm_100a CUDA in ``csrc/`` behind a C ABI (``include/gom_b200.h``), loaded
with ctypes. There is no CPU fallback: importing the ops without the built library raises.
"""
__version__ = "0.1.0"
