"""Regression for the TMA-refill vs LDS race found in round 1: an update-only launch in which no
filter has a measurement must leave the whole 2^20-filter state bit-identical, every time."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_state_untouched_when_no_filter_has_a_measurement():
    import torch
    from filterpy_b200.kalman import KalmanFilter
    from filterpy_b200.common import workloads as wl
    N = 1 << 20
    w = wl.kf_bank_cv2d(N, seed=77, steps=1, dtype=np.float32)
    kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
    for k in "xPFHQR":
        setattr(kf, k, w[k])
    P0, x0 = kf.P.clone(), kf.x.clone()
    valid = np.zeros(N, dtype=bool)
    for trial in range(8):
        kf.update(w["zs"][0], valid=valid)
        assert torch.equal(kf.P, P0) and torch.equal(kf.x, x0), "trial %d" % trial


def test_fused_step_is_deterministic():
    import torch
    from filterpy_b200.kalman import KalmanFilter
    from filterpy_b200.common import workloads as wl
    N = 1 << 20
    w = wl.kf_bank_cv2d(N, seed=78, steps=1, dtype=np.float32)
    outs = []
    for trial in range(4):
        kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
        for k in "xPFHQR":
            setattr(kf, k, w[k])
        kf.predict(); kf.update(w["zs"][0])
        outs.append((kf.x.clone(), kf.P.clone()))
    for x, P in outs[1:]:
        assert torch.equal(x, outs[0][0]) and torch.equal(P, outs[0][1])
