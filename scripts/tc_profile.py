"""A few tensor-core predict launches (csrc/kf_tc.cu) for an ncu capture: python scripts/tc_profile.py [16|32]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = (1 << 19) if n == 16 else (1 << 17)
rng = np.random.default_rng(0)
a = rng.normal(size=(N, n, n)).astype(np.float32)
P0 = 2.0 * (a @ np.swapaxes(a, -1, -2) / n + np.eye(n, dtype=np.float32))
kf = KalmanFilter(n, 4, n_filters=N, dtype=np.float32, diagnostics=False)
kf.x, kf.P = rng.normal(size=(N, n)), P0
kf.F, kf.H, kf.Q, kf.R = np.eye(n) + 0.1 * rng.normal(size=(n, n)), rng.normal(size=(4, n)), 0.05 * np.eye(n), 0.5 * np.eye(4)
for _ in range(4):
    kf.predict(); kf._flush()
torch.cuda.synchronize()
