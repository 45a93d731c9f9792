"""Oracle: multiple-model estimators over a list of linear Kalman filters (TEST INFRASTRUCTURE).

Restates, for ONE track (reference @ 3b51149):

* ``IMMEstimator``   filterpy/kalman/IMM.py:133-158 (init), :160-184 (update), :186-226 (predict),
  :228-237 (_compute_state_estimate), :239-247 (_compute_mixing_probabilities)
* ``MMAEFilterBank`` filterpy/kalman/mmae.py:140-153 (predict), :155-206 (update) — including the
  element-wise ``zip(self.x, self.filters, self.p)`` of :197-199, which pairs COMPONENT i of the
  mixed state with filter i.

Filters are plain dicts ``{x, P, F, H, Q, R}`` advanced with ``oracle.kf``; everything fp64.
Parity: pinned by ``tests/golden/mm.npz`` (generated from the reference's own classes).
"""
import sys

import numpy as np

from . import kf as okf


def likelihood(y, S):
    """kalman_filter.py:1203-1223: exp(logpdf(y; 0, S)), floored at float min."""
    ll = okf.log_likelihood_bank(np.asarray(y, float).reshape(1, -1), np.asarray(S, float)[None])[0]
    lk = np.exp(ll)
    return lk if lk != 0 else sys.float_info.min


class Imm(object):
    def __init__(self, filters, mu, M):
        self.filters = filters
        self.mu = np.asarray(mu, float) / np.sum(mu)
        self.M = np.asarray(M, float)
        self.N = len(filters)
        self.likelihood = np.zeros(self.N)
        self.omega = np.zeros((self.N, self.N))
        self._mixing_probabilities()
        self._state_estimate()

    def _mixing_probabilities(self):
        self.cbar = np.dot(self.mu, self.M)                                    # IMM.py:244
        for i in range(self.N):
            for j in range(self.N):
                self.omega[i, j] = (self.M[i, j] * self.mu[i]) / self.cbar[j]  # :247

    def _state_estimate(self):
        self.x = np.zeros_like(self.filters[0]["x"])
        for f, mu in zip(self.filters, self.mu):
            self.x += f["x"] * mu                                              # :232-233
        self.P = np.zeros_like(self.filters[0]["P"])
        for f, mu in zip(self.filters, self.mu):
            y = f["x"] - self.x
            self.P += mu * (np.outer(y, y) + f["P"])                           # :235-237

    def update(self, z):
        for i, f in enumerate(self.filters):
            x, P, y, K, S, SI = okf.kf_update_single(f["x"], f["P"], np.asarray(z, float), f["H"], f["R"])
            f["x"], f["P"] = x, P
            self.likelihood[i] = likelihood(y, S)                              # :174-176
        self.mu = self.cbar * self.likelihood                                  # :179
        self.mu /= np.sum(self.mu)
        self._mixing_probabilities()
        self._state_estimate()

    def predict(self):
        xs, Ps = [], []
        for i, (f, w) in enumerate(zip(self.filters, self.omega.T)):           # :201
            x = np.zeros_like(self.x)
            for kf, wj in zip(self.filters, w):
                x += kf["x"] * wj
            xs.append(x)
            P = np.zeros_like(self.P)
            for kf, wj in zip(self.filters, w):
                y = kf["x"] - x
                P += wj * (np.outer(y, y) + kf["P"])
            Ps.append(P)
        for i, f in enumerate(self.filters):                                   # :215-220
            f["x"], f["P"] = okf.kf_predict_single(xs[i].copy(), Ps[i].copy(), f["F"], f["Q"])
        self._state_estimate()


class Mmae(object):
    def __init__(self, filters, p):
        self.filters = filters
        self.p = np.asarray(p, float).copy()
        self.x = filters[0]["x"].copy()
        self.P = filters[0]["P"].copy()

    def predict(self):
        for f in self.filters:
            f["x"], f["P"] = okf.kf_predict_single(f["x"], f["P"], f["F"], f["Q"])

    def update(self, z):
        for i, f in enumerate(self.filters):
            x, P, y, K, S, SI = okf.kf_update_single(f["x"], f["P"], np.asarray(z, float), f["H"], f["R"])
            f["x"], f["P"] = x, P
            self.p[i] *= likelihood(y, S)                                      # mmae.py:182
        self.p /= sum(self.p)
        self.P = np.zeros(self.filters[0]["P"].shape)
        self.x = np.zeros(self.filters[0]["x"].shape)
        for f, p in zip(self.filters, self.p):
            self.x += np.dot(f["x"], p)                                        # :191-192
        for x, f, p in zip(self.x, self.filters, self.p):                      # :197 (components of x!)
            y = f["x"] - x
            self.P += p * (np.outer(y, y) + f["P"])
