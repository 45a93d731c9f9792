"""Host-side mirror of ``filterpy.kalman.IMMEstimator`` (filterpy/kalman/IMM.py:31-260) for a BANK
of tracks: ``filters`` is a list of M ``KalmanFilter`` banks (one per motion model, each with the
same ``n_filters`` tracks).  Every step of the reference's per-object Python loops — mode
probabilities (:178-184, :239-247), mixed initial conditions (:201-213), combined estimate
(:228-237) — is one CUDA launch over all tracks (csrc/mix.cu); the per-model predict / update are
the bank kernels.  Single-mode filters (``KalmanFilter(n, m)`` without ``n_filters``) give the
reference's one-track behaviour with NumPy attributes.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._dev import StepGraph, bke_dtype, ptr, stream_ptr

__all__ = ["IMMEstimator"]


def _check_bank(filters):
    f0 = filters[0]
    for f in filters:
        if (f.dim_x, f.n_filters, f._dtype, f._device, f._single) != \
                (f0.dim_x, f0.n_filters, f0._dtype, f0._device, f0._single):
            raise ValueError('All filters must have the same state dimension')     # IMM.py:143-146
        if not f.diagnostics:
            raise ValueError("the model filters must be built with diagnostics=True (likelihoods are read)")
    if len(filters) > _lib.BKE_MM_MAX_MODELS:
        raise NotImplementedError("at most %d models" % _lib.BKE_MM_MAX_MODELS)
    return f0


def _mm_args(filters, flags=0):
    f0 = filters[0]
    a = _lib.MmArgs()
    a.n_tracks, a.dim_x, a.n_models = f0.n_filters, f0.dim_x, len(filters)
    a.dtype, a.flags = bke_dtype(f0._dtype), flags
    for j, f in enumerate(filters):
        f._flush()
        a.x[j], a.P[j] = ptr(f._x), ptr(f._P)
        a.log_likelihood[j] = ptr(f._ll)
    return a


class IMMEstimator(object):
    """IMM.py:31-260.  ``mu``: (M,) initial mode probabilities shared by the tracks, or (N, M);
    ``M``: (M, M) Markov transition matrix."""

    def __init__(self, filters, mu, M):
        if len(filters) < 2:
            raise ValueError('filters must contain at least two filters')     # IMM.py:135-136
        f0 = _check_bank(filters)
        self.filters = filters
        self.N = len(filters)                                # number of models, as in the reference (:152)
        self.n_tracks = f0.n_filters
        self._single = f0._single
        self._dtype, self._device = f0._dtype, f0._device
        self._lib = _lib.load()
        nm, nt, n = self.N, self.n_tracks, f0.dim_x
        mu = np.asarray(mu, dtype=np.float64)
        mu = mu / np.sum(mu, axis=-1, keepdims=True)         # IMM.py:139
        if mu.shape == (nm,):
            mu = np.broadcast_to(mu, (nt, nm))
        if mu.shape != (nt, nm):
            raise ValueError("mu must have shape (%d,) or (%d,%d)" % (nm, nt, nm))
        kw = dict(dtype=torch.float64, device=self._device)
        self._mu = torch.from_numpy(np.array(mu, dtype=np.float64, order='C')).to(self._device)
        self._M = torch.from_numpy(np.ascontiguousarray(np.asarray(M, dtype=np.float64))).to(self._device)
        if tuple(self._M.shape) != (nm, nm):
            raise ValueError("M must have shape (%d,%d)" % (nm, nm))
        self._cbar = torch.zeros(nt, nm, **kw)
        self._omega = torch.zeros(nt, nm, nm, **kw)
        skw = dict(dtype=self._dtype, device=self._device)
        self._x = torch.zeros(nt, n, **skw)
        self._P = torch.zeros(nt, n, n, **skw)
        # spare state buffers for the mixing step (double-buffered with the filters' own)
        self._spare = [(torch.empty(nt, n, **skw), torch.empty(nt, n, n, **skw)) for _ in filters]
        self._compute_mixing_probabilities(initial=True)
        self._compute_state_estimate()
        self._x_prior = self._x.clone(); self._P_prior = self._P.clone()
        self._x_post = self._x.clone(); self._P_post = self._P.clone()

    # ------------------------------------------------------------------ outputs
    def _vec(self, t):
        if not self._single:
            return t
        v = t[0].cpu().numpy()
        return v.reshape(-1, 1) if self.filters[0]._x_col else v

    def _mat(self, t):
        return t if not self._single else t[0].cpu().numpy()

    x = property(lambda self: self._vec(self._x))
    P = property(lambda self: self._mat(self._P))
    x_prior = property(lambda self: self._vec(self._x_prior))
    P_prior = property(lambda self: self._mat(self._P_prior))
    x_post = property(lambda self: self._vec(self._x_post))
    P_post = property(lambda self: self._mat(self._P_post))
    mu = property(lambda self: self._mat(self._mu))
    M = property(lambda self: self._M.cpu().numpy())
    cbar = property(lambda self: self._mat(self._cbar))
    omega = property(lambda self: self._mat(self._omega))

    @property
    def likelihood(self):
        """per-model likelihood of the last measurement (IMM.py:154, :174-176): (N, M), or (M,) for one track."""
        if self._single:
            return np.array([f.likelihood for f in self.filters])
        return torch.stack([f.likelihood for f in self.filters], dim=-1)

    # ------------------------------------------------------------------ steps
    def _call(self, fn, a):
        with torch.cuda.device(self._device):
            _lib.check(fn(ctypes.byref(a), stream_ptr(self._device)))

    def update(self, z):
        """IMM.py:160-184: update every model, then mode probabilities and the combined estimate."""
        for f in self.filters:
            f.update(z)
        a = _mm_args(self.filters)
        a.mu, a.cbar, a.omega, a.trans = ptr(self._mu), ptr(self._cbar), ptr(self._omega), ptr(self._M)
        self._call(self._lib.bke_mm_probabilities, a)
        self._compute_state_estimate()
        self._x_post.copy_(self._x); self._P_post.copy_(self._P)

    def predict(self, u=None):
        """IMM.py:186-226: mixed initial conditions for every model, predict, combined prior."""
        a = _mm_args(self.filters)
        a.omega, a.weights_stride = ptr(self._omega), self.N * self.N
        for j, (xs, Ps) in enumerate(self._spare):
            a.x_out[j], a.P_out[j] = ptr(xs), ptr(Ps)
        self._call(self._lib.bke_mm_mix, a)
        for j, f in enumerate(self.filters):
            self._spare[j] = f._adopt_state(*self._spare[j])           # f.x = xs[i]; f.P = Ps[i] (:216-219)
            f.predict(u)
        self._compute_state_estimate()
        self._x_prior.copy_(self._x); self._P_prior.copy_(self._P)

    def _compute_state_estimate(self):
        """IMM.py:228-237."""
        a = _mm_args(self.filters)
        a.mu, a.weights_stride = ptr(self._mu), self.N
        a.x_out[0], a.P_out[0] = ptr(self._x), ptr(self._P)
        self._call(self._lib.bke_mm_estimate, a)

    def _compute_mixing_probabilities(self, initial=False):
        """IMM.py:239-247 (cbar = mu . M, omega); the likelihood step is part of update()."""
        a = _mm_args(self.filters, flags=_lib.BKE_MM_FROM_MU)
        a.mu, a.cbar, a.omega, a.trans = ptr(self._mu), ptr(self._cbar), ptr(self._omega), ptr(self._M)
        self._call(self._lib.bke_mm_probabilities, a)

    def capture(self, fn, warmup=2):
        """Capture ``fn`` — a fixed sequence of ``predict()`` / ``update(z_buffer)`` calls — into a CUDA
        graph (``.replay()``); one IMM step is ~15 small launches, so the host side dominates otherwise.
        The model filters' state buffers rotate with period 3 (state, stored posterior, spare), so
        ``fn`` must run a multiple of 3 steps for a replay to find the buffers where it left them."""
        return StepGraph(fn, self._device, warmup)

    def __repr__(self):
        return "IMMEstimator (B200): %d models x %d tracks, dim_x=%d" % (self.N, self.n_tracks, self.filters[0].dim_x)
