"""torch.ops.bke.* (filterpy_b200/torch_ops): the C-ABI of include/bke.h as PyTorch operators.
CPU: the extension builds, loads, registers every operator with the expected schema and has no CPU backend.
GPU: each operator equals the ctypes-bound mirror / the oracle on the same inputs."""
import numpy as np
import pytest


def test_twin_loads_and_registers_every_operator():
    import torch
    from filterpy_b200 import torch_ops
    ops = torch_ops.load()
    for name, frag in [("kf_step", "Tensor z, float alpha_sq"), ("kf_predict", "Tensor Q, float alpha_sq"),
                       ("ukf_step", "int fx_model, int hx_model"), ("systematic_resample", "Tensor weights, float u"),
                       ("stratified_resample", "Tensor weights, Tensor uniforms")]:
        assert frag in str(getattr(ops, name).default._schema)
    # no CPU backend is registered: the product path fails loudly instead of falling back
    with pytest.raises(NotImplementedError):
        ops.systematic_resample(torch.ones(4, dtype=torch.float64) / 4, 0.5)
    with pytest.raises(NotImplementedError):
        ops.kf_predict(torch.zeros(2, 4), torch.eye(4).repeat(2, 1, 1), torch.eye(4), torch.eye(4))


@pytest.mark.gpu
def test_twin_kf_and_resample_equal_the_ctypes_path(golden):
    import torch
    from filterpy_b200 import torch_ops
    from filterpy_b200.kalman import KalmanFilter
    from filterpy_b200.monte_carlo import ResamplePlan
    from filterpy_b200.common import workloads as wl
    from oracle import resample as ors
    ops = torch_ops.load()
    N = 5003
    for dtype, td in [(np.float32, torch.float32), (np.float64, torch.float64)]:
        w = wl.kf_bank_cv2d(N, seed=3, steps=1, dtype=dtype)
        kf = KalmanFilter(4, 2, n_filters=N, dtype=dtype, diagnostics=False)
        for k in "xPFHQR":
            setattr(kf, k, w[k])
        kf.predict(); kf.update(w["zs"][0])
        d = {k: torch.from_numpy(np.ascontiguousarray(w[k]).astype(dtype)).cuda() for k in "xPFHQR"}
        z = torch.from_numpy(w["zs"][0].astype(dtype)).cuda()
        x, P = ops.kf_step(d["x"], d["P"], d["F"], d["H"], d["Q"], d["R"], z)
        t0 = 1e-6 if dtype is np.float32 else 1e-13
        assert x.dtype == td and torch.allclose(x, kf.x, rtol=t0, atol=t0 * float(kf.x.abs().max()))
        assert torch.allclose(P, kf.P, rtol=t0, atol=t0 * float(kf.P.abs().max()))
        # shared models (2-D tensors -> stride 0) and the predict-only operator
        xs, Ps = ops.kf_step(d["x"], d["P"], d["F"][0], d["H"][0], d["Q"][0], d["R"][0], z)
        xp, Pp = ops.kf_predict(d["x"], d["P"], d["F"][0], d["Q"][0])
        k2 = KalmanFilter(4, 2, n_filters=N, dtype=dtype, diagnostics=True)
        k2.x, k2.P = w["x"], w["P"]
        k2.F, k2.H, k2.Q, k2.R = w["F"][0], w["H"][0], w["Q"][0], w["R"][0]
        k2.predict(); k2.update(w["zs"][0])
        tol = 1e-5 if dtype is np.float32 else 1e-12
        for a, b in [(xs, k2.x), (Ps, k2.P), (xp, k2.x_prior), (Pp, k2.P_prior)]:
            assert torch.allclose(a, b, rtol=tol, atol=tol * float(b.abs().max()))
    wts = wl.resample_weights(1 << 16, "heavy", seed=4)
    wd = torch.from_numpy(wts).cuda()
    idx = ops.systematic_resample(wd, 0.37)
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), ors.systematic_resample_c(wts, 0.37))
    U = np.random.default_rng(2).random(len(wts))
    idx = ops.stratified_resample(wd, torch.from_numpy(U).cuda())
    assert np.array_equal(idx.cpu().numpy(), ors.stratified_resample_c(wts, U))
    with pytest.raises(IndexError):
        ops.systematic_resample(wd * 0.5, 0.9)          # resampling.py:145


@pytest.mark.gpu
def test_twin_ukf_equals_the_mirror():
    import torch
    from filterpy_b200 import torch_ops, _lib
    from filterpy_b200.kalman import UnscentedKalmanFilter, MerweScaledSigmaPoints, ConstVelFx, RangeAzElHx
    from filterpy_b200.common import workloads as wl
    ops = torch_ops.load()
    N = 3000
    uw = wl.ukf_bank_cv3d(N, seed=2, steps=1)
    ukf = UnscentedKalmanFilter(6, 3, 0.1, RangeAzElHx(), ConstVelFx(), MerweScaledSigmaPoints(6, .5, 2., 0.),
                                n_filters=N, device="cuda:0", diagnostics=False)
    ukf.x = uw["x"]; ukf.P = uw["P"]; ukf.Q = uw["Q"]; ukf.R = uw["R"]
    ukf.predict(); ukf.update(uw["zs"][0])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
    x, P = ops.ukf_step(t(uw["x"]), t(uw["P"]), t(uw["Q"]), t(uw["R"]), t(uw["zs"][0]), 0.1, .5, 2., 0.,
                        _lib.BKE_FX_CONST_VEL, _lib.BKE_HX_RANGE_AZ_EL)
    assert torch.allclose(x, ukf.x, rtol=1e-10, atol=1e-10 * float(ukf.x.abs().max()))
    assert torch.allclose(P, ukf.P, rtol=1e-10, atol=1e-10 * float(ukf.P.abs().max()))
