"""A few IMM steps on 2^20 tracks x 3 models (for an ncu launch list)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter, IMMEstimator
from filterpy_b200.common import workloads as wl
N, M = 1 << 20, 3
wk = wl.kf_bank_cv2d(4096, seed=6)
reps = N // 4096
fs = []
for j in range(M):
    f = KalmanFilter(4, 2, n_filters=N, dtype=np.float32)
    f.x = np.tile(wk["x"], (reps, 1)); f.P = np.tile(wk["P"], (reps, 1, 1))
    f.F = np.tile(wk["F"], (reps, 1, 1)); f.H = wk["H"][0]; f.R = wk["R"][0]
    f.Q = np.tile(wk["Q"], (reps, 1, 1)) * (10.0 ** j)
    fs.append(f)
imm = IMMEstimator(fs, [0.5, 0.3, 0.2], np.array([[.9, .05, .05], [.1, .8, .1], [.05, .15, .8]]))
z = torch.randn(N, 2, device="cuda")
for _ in range(4):
    imm.predict(); imm.update(z)
torch.cuda.synchronize()
