#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports rlabbe/filterpy 1.4.5 (@ 3b51149) from /root/reference, feeds it the seeded
synthetic inputs of ``filterpy_b200.common.workloads`` and stores inputs + the reference's
outputs as small ``.npz`` files.  The reference cannot travel to the GPU box, the vectors
can: tests compare the oracle (``oracle/``) and the CUDA path against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from filterpy.kalman import KalmanFilter, UnscentedKalmanFilter, MerweScaledSigmaPoints  # noqa: E402
from filterpy.kalman import predict as kf_predict_proc, update as kf_update_proc          # noqa: E402
from filterpy.monte_carlo import systematic_resample, stratified_resample                # noqa: E402
import filterpy                                                                            # noqa: E402

from filterpy_b200.common import workloads as wl                                          # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, reference_version=filterpy.__version__, **arrs)
    print("wrote", path, os.path.getsize(path), "bytes")


# ----------------------------------------------------------------------------- KF C1
def gen_kf_c1():
    w = wl.kf_single_cv2d(T=1000, seed=0)
    kf = KalmanFilter(4, 2)
    kf.x = w["x"].copy(); kf.P = w["P"].copy()
    kf.F, kf.H, kf.Q, kf.R = w["F"], w["H"], w["Q"], w["R"]
    means, covs, means_p, covs_p = kf.batch_filter(list(w["zs"]))
    # one-step known answers quoted in SURVEY.md §8c
    k2 = KalmanFilter(4, 2)
    k2.x = np.zeros(4); k2.P = 10 * np.eye(4)
    k2.F, k2.H, k2.Q, k2.R = w["F"], w["H"], w["Q"], w["R"]
    k2.predict(); k2.update(np.array([1., 2.]))
    save("kf_c1", **w, means=means, covs=covs, means_p=means_p, covs_p=covs_p,
         one_x=k2.x, one_P=k2.P, one_S=k2.S, one_K=k2.K,
         one_loglik=k2.log_likelihood, one_maha=float(k2.mahalanobis))


# ----------------------------------------------------------------------------- KF banks
def run_kf_bank(w, steps, alpha=1.0, none_frac=0.0, seed=5, with_control=False):
    N, n = w["x"].shape
    m = w["H"].shape[-2]
    rng = np.random.default_rng(seed)
    valid = rng.random((steps, N)) >= none_frac
    du = 2 if with_control else 0
    B = rng.standard_normal((N, n, du)) if with_control else None
    us = rng.standard_normal((steps, N, du)) if with_control else None
    keys = ["x", "P", "x_prior", "P_prior", "y", "K", "S", "SI", "loglik"]
    out = {k: [] for k in keys}
    filters = []
    for f in range(N):
        kf = KalmanFilter(n, m, dim_u=du)
        kf.x = w["x"][f].copy(); kf.P = w["P"][f].copy()
        kf.F, kf.H, kf.Q, kf.R = w["F"][f], w["H"][f], w["Q"][f], w["R"][f]
        kf.alpha = alpha
        if with_control:
            kf.B = B[f]
        filters.append(kf)
    for t in range(steps):
        rec = {k: [] for k in keys}
        for f, kf in enumerate(filters):
            kf.predict(u=us[t, f] if with_control else None)
            z = w["zs"][t, f] if valid[t, f] else None
            kf.update(z)
            rec["x"].append(kf.x.copy()); rec["P"].append(kf.P.copy())
            rec["x_prior"].append(kf.x_prior.copy()); rec["P_prior"].append(kf.P_prior.copy())
            rec["y"].append(np.asarray(kf.y, float).reshape(m))
            rec["K"].append(kf.K.copy()); rec["S"].append(kf.S.copy()); rec["SI"].append(kf.SI.copy())
            rec["loglik"].append(kf.log_likelihood if z is not None else np.nan)
        for k in keys:
            out[k].append(np.array(rec[k]))
    res = {"ref_" + k: np.array(v) for k, v in out.items()}
    res["valid"] = valid
    res["alpha"] = alpha
    if with_control:
        res["B"] = B; res["us"] = us
    return res


def gen_kf_banks():
    w = wl.kf_bank_cv2d(48, seed=1234, steps=5)
    save("kf_bank_4_2", **w, **run_kf_bank(w, 5, none_frac=0.15))
    w = wl.kf_bank_ca3d(24, seed=4321, steps=3)
    save("kf_bank_9_3", **w, **run_kf_bank(w, 3, alpha=1.02))
    # odd little shapes: random well-conditioned dense models
    rng = np.random.default_rng(99)
    for (n, m) in [(1, 1), (2, 1), (3, 2), (6, 3), (5, 5)]:
        N, steps = 8, 3
        A = rng.standard_normal((N, n, n)) * 0.3
        w = dict(x=rng.standard_normal((N, n)),
                 P=np.einsum("nij,nkj->nik", A, A) + np.eye(n),
                 F=np.eye(n) + 0.1 * rng.standard_normal((N, n, n)),
                 H=rng.standard_normal((N, m, n)),
                 Q=0.01 * np.eye(n) + np.zeros((N, n, n)),
                 R=np.eye(m) * rng.uniform(0.2, 1.0, (N, 1, 1)),
                 zs=rng.standard_normal((steps, N, m)))
        save("kf_bank_%d_%d" % (n, m), **w, **run_kf_bank(w, steps, with_control=(n == 3)))


# ----------------------------------------------------------------------------- UKF
def fx_cv(x, dt):
    o = x.copy()
    o[0::2] = x[0::2] + dt * x[1::2]
    return o


def hx_rae(x):
    px, py, pz = x[0], x[2], x[4]
    return np.array([np.sqrt(px * px + py * py + pz * pz), np.arctan2(py, px),
                     np.arctan2(pz, np.sqrt(px * px + py * py))])


def gen_ukf():
    alpha, beta, kappa = 0.5, 2.0, 0.0
    pts = MerweScaledSigmaPoints(6, alpha, beta, kappa)
    P = np.array([[4, .5, 0, 0, 0, 0], [.5, 2, 0, 0, 0, 0], [0, 0, 3, .2, 0, 0],
                  [0, 0, .2, 1, 0, 0], [0, 0, 0, 0, 5, .1], [0, 0, 0, 0, .1, 1.5]], float)
    x = np.arange(6.0)
    save("ukf_sigma", x=x, P=P, alpha=alpha, beta=beta, kappa=kappa,
         sigmas=pts.sigma_points(x, P), Wm=pts.Wm, Wc=pts.Wc,
         Wm4=MerweScaledSigmaPoints(4, .5, 2, 0).Wm, Wc4=MerweScaledSigmaPoints(4, .5, 2, 0).Wc)
    for name, linear in (("ukf_bank_rae", False), ("ukf_bank_lin", True)):
        N, steps, dt = 16, 5, 0.1
        w = wl.ukf_bank_cv3d(N, seed=2468, steps=steps, dt=dt, linear_hx=linear)
        F, Hlin = w["F"], w["H"]
        fx = (lambda s, dt: F @ s) if linear else fx_cv
        hx = (lambda s: Hlin @ s) if linear else hx_rae
        rng = np.random.default_rng(3)
        valid = rng.random((steps, N)) >= (0.0 if linear else 0.1)
        keys = ["x", "P", "x_prior", "P_prior", "K", "S", "y"]
        out = {k: [] for k in keys}
        ukfs = []
        for f in range(N):
            u = UnscentedKalmanFilter(6, 3, dt, hx, fx, MerweScaledSigmaPoints(6, alpha, beta, kappa))
            u.x = w["x"][f].copy(); u.P = w["P"][f].copy(); u.Q = w["Q"][f]; u.R = w["R"][f]
            ukfs.append(u)
        for t in range(steps):
            rec = {k: [] for k in keys}
            for f, u in enumerate(ukfs):
                u.predict()
                u.update(w["zs"][t, f] if valid[t, f] else None)
                for k in keys:
                    rec[k].append(np.array(getattr(u, k), float).copy())
            for k in keys:
                out[k].append(np.array(rec[k]))
        save(name, **w, valid=valid, dt=dt, alpha=alpha, beta=beta, kappa=kappa,
             **{"ref_" + k: np.array(v) for k, v in out.items()})


# ----------------------------------------------------------------------------- resampling
def gen_resample():
    cases = {}
    meta = []
    i = 0
    for kind in ["heavy", "uniform", "random", "zeros", "degenerate", "dyadic"]:
        for N in [1, 2, 7, 64, 1000, 4097]:
            w = wl.resample_weights(N, kind, seed=97 + N)
            np.random.seed(7 + i)
            st = np.random.get_state()
            try:
                idx = systematic_resample(w)
                ok = 1
            except IndexError:
                idx = np.zeros(0, 'i'); ok = 0
            np.random.set_state(st)
            u = np.random.random()
            np.random.set_state(st)
            try:
                idx_s = stratified_resample(w)
                ok_s = 1
            except IndexError:
                idx_s = np.zeros(0, 'i'); ok_s = 0
            np.random.set_state(st)
            U = np.random.random(N)
            cases["w%d" % i] = w; cases["u%d" % i] = u; cases["U%d" % i] = U
            cases["sys%d" % i] = idx; cases["str%d" % i] = idx_s
            meta.append((i, N, ok, ok_s, 7 + i))
            i += 1
    # the SURVEY known answer: systematic_resample([.1,.2,.3,.4]) with u = 0.5 -> [1,2,3,3]
    import filterpy.monte_carlo.resampling as rs
    old = rs.random
    rs.random = lambda *a: 0.5 if not a else np.full(a[0], 0.5)
    ka = systematic_resample([.1, .2, .3, .4]); ka_s = stratified_resample([.1, .2, .3, .4])
    rs.random = old
    save("resample", meta=np.array(meta), known_sys=ka, known_str=ka_s, **cases)


if __name__ == "__main__":
    gen_kf_c1()
    gen_kf_banks()
    gen_ukf()
    gen_resample()
