"""Host mirror of ``filterpy.kalman.MerweScaledSigmaPoints`` (filterpy/kalman/sigma_points.py:24-208).

The object carries (n, alpha, beta, kappa) and the weights; the sigma points themselves are
generated inside the fused UKF kernel (csrc/ukf.cu) from the rows of the upper Cholesky factor of
(n + lambda) P, exactly as sigma_points.py:167-175 does.
"""
import numpy as np

__all__ = ["MerweScaledSigmaPoints"]


class MerweScaledSigmaPoints(object):
    def __init__(self, n, alpha, beta, kappa, sqrt_method=None, subtract=None):
        if sqrt_method is not None or subtract is not None:
            raise NotImplementedError(
                "custom sqrt_method / subtract are Python callables; the GPU path implements the "
                "defaults only (scipy.linalg.cholesky, np.subtract) and has no CPU fallback")
        self.n = int(n)
        self.alpha = float(alpha)
        self.beta = float(beta)
        self.kappa = float(kappa)
        self._compute_weights()

    def num_sigmas(self):
        """Number of sigma points (sigma_points.py:119-121)."""
        return 2 * self.n + 1

    def _compute_weights(self):
        """sigma_points.py:180-192."""
        n = self.n
        lambda_ = self.alpha ** 2 * (n + self.kappa) - n
        c = .5 / (n + lambda_)
        self.Wc = np.full(2 * n + 1, c)
        self.Wm = np.full(2 * n + 1, c)
        self.Wc[0] = lambda_ / (n + lambda_) + (1 - self.alpha ** 2 + self.beta)
        self.Wm[0] = lambda_ / (n + lambda_)

    def __repr__(self):
        return "MerweScaledSigmaPoints(n=%d, alpha=%g, beta=%g, kappa=%g)" % (
            self.n, self.alpha, self.beta, self.kappa)
