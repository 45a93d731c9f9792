"""filterpy_b200 — B200-native batched state-estimation engine behind filterpy's call surface.

Hot path only: banks of linear / unscented Kalman filters and particle
resampling, computed by hand-written sm_100a CUDA kernels behind a C-ABI (include/bke.h).
"""
__version__ = "0.1.0"
