import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import UnscentedKalmanFilter, MerweScaledSigmaPoints, ConstVelFx, RangeAzElHx
from filterpy_b200.common import workloads as wl
N = 1 << 18
uw = wl.ukf_bank_cv3d(N, steps=1)
for dt in (np.float64, np.float32):
    u = UnscentedKalmanFilter(6, 3, 0.1, RangeAzElHx(), ConstVelFx(), MerweScaledSigmaPoints(6, .5, 2., 0.), n_filters=N, dtype=dt, diagnostics=False)
    u.x = uw["x"]; u.P = uw["P"]; u.Q = uw["Q"]; u.R = uw["R"]
    z = torch.from_numpy(uw["zs"][0].astype(dt)).cuda()
    for _ in range(3):
        u.predict(); u.update(z)
torch.cuda.synchronize()
