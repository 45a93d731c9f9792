"""A few launches of the tensor-core tile (csrc/kf_tc.cu) for an ncu capture:
python scripts/tc_profile.py [16|32] [predict|step]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fused = len(sys.argv) > 2 and sys.argv[2] == "step"
N = (1 << 19) if n == 16 else (1 << 17)
rng = np.random.default_rng(0)
a = rng.normal(size=(4096, n, n)).astype(np.float32)
P0 = np.tile(2.0 * (a @ np.swapaxes(a, -1, -2) / n + np.eye(n, dtype=np.float32)), (N // 4096, 1, 1))
kf = KalmanFilter(n, 4, n_filters=N, dtype=np.float32, diagnostics=False)
kf.x, kf.P = rng.normal(size=(N, n)), P0
kf.F, kf.H, kf.Q, kf.R = np.eye(n) + 0.1 * rng.normal(size=(n, n)), rng.normal(size=(4, n)), 0.05 * np.eye(n), 0.5 * np.eye(4)
z = torch.from_numpy(rng.normal(size=(N, 4)).astype(np.float32)).cuda()
for _ in range(4):
    kf.predict()
    if fused:
        kf.update(z)
    else:
        kf._flush()
torch.cuda.synchronize()
