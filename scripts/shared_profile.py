"""Launches the shared-model 4/2 fp32 step and the batch_filter kernel a few times (for ncu)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
N = 1 << 20
w = wl.kf_bank_cv2d(1 << 14, dtype=np.float32)
r = N >> 14
kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
kf.x = np.tile(w["x"], (r, 1)); kf.P = np.tile(w["P"], (r, 1, 1))
kf.F, kf.H, kf.Q, kf.R = w["F"][0], w["H"][0], w["Q"][0], w["R"][0]
z = torch.from_numpy(np.tile(w["zs"][0], (r, 1)).astype(np.float32)).cuda()
for _ in range(4):
    kf.predict(); kf.update(z)
torch.cuda.synchronize()
Nb, T = 1 << 18, 8
kb = KalmanFilter(4, 2, n_filters=Nb, dtype=np.float32, diagnostics=False)
rb = Nb >> 14
kb.x = np.tile(w["x"], (rb, 1)); kb.P = np.tile(w["P"], (rb, 1, 1))
for nm in "FHQR":
    setattr(kb, nm, np.tile(w[nm], (rb, 1, 1)))
zs = torch.randn(T, Nb, 2, device="cuda")
for _ in range(2):
    kb.batch_filter(zs)
torch.cuda.synchronize()
