"""stripenn_b200 — B200-native `stripenn compute` / `score` hot path.

Host layer (Python) mirroring the reference's call surface (src/stripenn/stripenn.py,
score.py, getStripe.py, cli.py) on top of the C ABI in include/stripenn_b200.h; all numerical
stages run in hand-written sm_100a CUDA kernels (stripenn_b200/csrc).  No CPU fallback.
"""
__version__ = "0.1.0"
