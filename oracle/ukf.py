"""Oracle: unscented Kalman filter (TEST INFRASTRUCTURE).

Restates (reference @ 3b51149):

* ``MerweScaledSigmaPoints``  sigma_points.py:124-192  (rows of the UPPER Cholesky
  factor of (n+lambda) P; weights Wm/Wc)
* ``unscented_transform``     unscented_transform.py:99-128
* ``UnscentedKalmanFilter.predict / update / cross_variance / batch_filter``
  UKF.py:364-411, 413-491, 493-504, 524-632; ``rts_smoother`` UKF.py:634-739.  Note UKF.py:407: after the first
  unscented transform the sigma points are REGENERATED from the prior, and
  UKF.py:481: P = P - K S K' (no symmetrisation).

``*_single`` take Python ``fx(x, dt)`` / ``hx(x)`` callables exactly like the
reference; ``*_bank`` vectorise over N for the closed set of device-side models
(``FX_*`` / ``HX_*`` ids below — the same ids the C-ABI takes, include/bke.h).
Parity: pinned by ``tests/golden/ukf_*.npz``.
"""
import numpy as np

# model ids (must match include/bke.h)
FX_LINEAR = 0      # x' = F x           (F[n,n] shared or [N,n,n])
FX_CONST_VEL = 1   # state = (p0,v0,p1,v1,...): p_i += dt * v_i
HX_LINEAR = 0      # z = H x            (H[m,n] shared or [N,m,n])
HX_RANGE_AZ_EL = 1  # n=6 (x,vx,y,vy,z,vz) -> (range, azimuth, elevation)
HX_RANGE_BEARING = 2  # n=4 (x,vx,y,vy) -> (range, bearing)


def merwe_weights(n, alpha, beta, kappa):
    """sigma_points.py:180-192."""
    lambda_ = alpha ** 2 * (n + kappa) - n
    c = .5 / (n + lambda_)
    Wc = np.full(2 * n + 1, c)
    Wm = np.full(2 * n + 1, c)
    Wc[0] = lambda_ / (n + lambda_) + (1 - alpha ** 2 + beta)
    Wm[0] = lambda_ / (n + lambda_)
    return Wm, Wc


def _chol_upper(A):
    """scipy.linalg.cholesky(A) (upper, reads only the upper triangle): U'U = A."""
    A = np.asarray(A, float)
    Au = np.triu(A)
    As = Au + np.swapaxes(np.triu(A, 1), -1, -2)
    L = np.linalg.cholesky(As)
    return np.swapaxes(L, -1, -2)


def merwe_sigma_points(x, P, alpha, beta, kappa):
    """sigma_points.py:160-177; x[..., n], P[..., n, n] -> sigmas[..., 2n+1, n]."""
    n = x.shape[-1]
    lambda_ = alpha ** 2 * (n + kappa) - n
    U = _chol_upper((lambda_ + n) * P)
    s0 = x[..., None, :]
    return np.concatenate([s0, s0 + U, s0 - U], axis=-2)


def unscented_transform(sigmas, Wm, Wc, noise_cov=None):
    """unscented_transform.py:99-128 (default mean / residual); leading batch axes ok."""
    x = np.einsum("s,...sn->...n", Wm, sigmas)
    y = sigmas - x[..., None, :]
    P = np.einsum("...sa,s,...sb->...ab", y, Wc, y)
    if noise_cov is not None:
        P = P + noise_cov
    return x, P


# --------------------------------------------------------------------------- single filter, callables
def ukf_predict_single(x, P, Q, fx, dt, alpha, beta, kappa):
    """UKF.py:393-411 -> (x_prior, P_prior, sigmas_f regenerated from the prior)."""
    Wm, Wc = merwe_weights(x.shape[0], alpha, beta, kappa)
    sig = merwe_sigma_points(x, P, alpha, beta, kappa)
    sig_f = np.array([fx(s, dt) for s in sig])
    x, P = unscented_transform(sig_f, Wm, Wc, Q)
    sig_f = merwe_sigma_points(x, P, alpha, beta, kappa)
    return x, P, sig_f


def ukf_update_single(x, P, sig_f, z, R, hx, alpha, beta, kappa):
    """UKF.py:442-486 -> (x, P, y, K, S, SI)."""
    if z is None:
        return x.copy(), P.copy(), None, None, None, None
    Wm, Wc = merwe_weights(x.shape[0], alpha, beta, kappa)
    sig_h = np.atleast_2d([hx(s) for s in sig_f])
    zp, S = unscented_transform(sig_h, Wm, Wc, R)
    SI = np.linalg.inv(S)
    Pxz = np.einsum("s,sa,sb->ab", Wc, sig_f - x, sig_h - zp)
    K = np.dot(Pxz, SI)
    y = z - zp
    x = x + np.dot(K, y)
    P = P - np.dot(K, np.dot(S, K.T))
    return x, P, y, K, S, SI


# --------------------------------------------------------------------------- closed set of models
def fx_apply(model, s, dt, F=None):
    """s[..., n] -> fx(s)."""
    if model == FX_LINEAR:
        if F.ndim == 2:
            return s @ F.T
        return np.einsum("nab,n...b->n...a", F, s)
    if model == FX_CONST_VEL:
        o = s.copy()
        o[..., 0::2] = s[..., 0::2] + dt * s[..., 1::2]
        return o
    raise ValueError(model)


def hx_apply(model, s, H=None):
    if model == HX_LINEAR:
        if H.ndim == 2:
            return s @ H.T
        return np.einsum("nab,n...b->n...a", H, s)
    if model == HX_RANGE_AZ_EL:
        px, py, pz = s[..., 0], s[..., 2], s[..., 4]
        rho = np.sqrt(px * px + py * py)
        r = np.sqrt(px * px + py * py + pz * pz)
        return np.stack([r, np.arctan2(py, px), np.arctan2(pz, rho)], axis=-1)
    if model == HX_RANGE_BEARING:
        px, py = s[..., 0], s[..., 2]
        return np.stack([np.sqrt(px * px + py * py), np.arctan2(py, px)], axis=-1)
    raise ValueError(model)


def ukf_step_bank(x, P, z, Q, R, dt, alpha, beta, kappa,
                  fx_model=FX_LINEAR, hx_model=HX_LINEAR, F=None, H=None, valid=None):
    """One predict + update for a bank x[N,n], P[N,n,n], z[N,m]; Q/R [n,n]/[m,m] or per filter.

    Returns dict(x, P, x_prior, P_prior, y, K, S, SI)."""
    n = x.shape[-1]
    Wm, Wc = merwe_weights(n, alpha, beta, kappa)
    sig = merwe_sigma_points(x, P, alpha, beta, kappa)
    sig_f = fx_apply(fx_model, sig, dt, F)
    xp, Pp = unscented_transform(sig_f, Wm, Wc, Q)
    sig_f = merwe_sigma_points(xp, Pp, alpha, beta, kappa)
    sig_h = hx_apply(hx_model, sig_f, H)
    zp, S = unscented_transform(sig_h, Wm, Wc, R)
    SI = np.linalg.inv(S)
    Pxz = np.einsum("s,nsa,nsb->nab", Wc, sig_f - xp[:, None, :], sig_h - zp[:, None, :])
    K = Pxz @ SI
    y = z - zp
    xn = xp + (K @ y[..., None])[..., 0]
    Pn = Pp - K @ (S @ np.swapaxes(K, -1, -2))
    if valid is not None:
        v = np.asarray(valid, bool)
        xn = np.where(v[:, None], xn, xp)
        Pn = np.where(v[:, None, None], Pn, Pp)
    return dict(x=xn, P=Pn, x_prior=xp, P_prior=Pp, y=y, K=K, S=S, SI=SI)


def ukf_rts_smoother(Xs, Ps, Q, fx, dts, alpha, beta, kappa):
    """UnscentedKalmanFilter.rts_smoother, UKF.py:696-739, for ONE filter: Xs (T,n), Ps (T,n,n),
    Q the filter's own Q (the reference never reads its Qs argument, :715), fx a callable,
    dts a list of T time steps -> (xs, Ps, Ks)."""
    if len(Xs) != len(Ps):
        raise ValueError('Xs and Ps must have the same length')
    n, dim_x = Xs.shape
    Wm, Wc = merwe_weights(dim_x, alpha, beta, kappa)
    Ks = np.zeros((n, dim_x, dim_x))
    xs, ps = Xs.copy(), Ps.copy()
    for k in reversed(range(n - 1)):
        sigmas = merwe_sigma_points(xs[k], ps[k], alpha, beta, kappa)         # :711
        sigmas_f = np.array([fx(s, dts[k]) for s in sigmas])                  # :712-713
        xb, Pb = unscented_transform(sigmas_f, Wm, Wc, Q)                     # :715-717
        Pxb = 0
        for i in range(sigmas.shape[0]):                                      # :720-724
            y = sigmas_f[i] - xb
            z = sigmas[i] - Xs[k]
            Pxb = Pxb + Wc[i] * np.outer(z, y)
        K = np.dot(Pxb, np.linalg.inv(Pb))                                    # :727
        xs[k] += np.dot(K, xs[k + 1] - xb)                                    # :730
        ps[k] += np.dot(K, ps[k + 1] - Pb).dot(K.T)                           # :731
        Ks[k] = K
    return xs, ps, Ks
