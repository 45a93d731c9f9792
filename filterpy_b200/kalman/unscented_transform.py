"""GPU mirror of ``filterpy.kalman.unscented_transform`` (filterpy/kalman/unscented_transform.py:22-128)."""
import numpy as np
import torch

from .. import _lib
from .._dev import bke_dtype, require_cuda, stream_ptr

__all__ = ["unscented_transform"]


def unscented_transform(sigmas, Wm, Wc, noise_cov=None, mean_fn=None, residual_fn=None):
    """``sigmas`` (k, n) -> ``(x, P)`` like the reference; a bank ``[N, k, n]`` -> ``(x[N,n], P[N,n,n])``.
    NumPy in -> NumPy out, CUDA tensors in -> CUDA tensors out.  ``mean_fn`` / ``residual_fn`` are
    Python callables in the reference: anything but the default raises NotImplementedError."""
    if mean_fn is not None or (residual_fn is not None and residual_fn is not np.subtract):
        raise NotImplementedError("mean_fn / residual_fn are Python callables; the GPU path implements the "
                                  "defaults only and has no CPU fallback")
    is_t = isinstance(sigmas, torch.Tensor)
    dev = sigmas.device if (is_t and sigmas.is_cuda) else require_cuda(None)
    dt = sigmas.dtype if (is_t and sigmas.dtype in (torch.float32, torch.float64)) else torch.float64

    def dev_of(a):
        return (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.float64))).to(device=dev, dtype=dt).contiguous()
    st = dev_of(sigmas)
    single = st.dim() == 2
    if single:
        st = st[None]
    N, k, n = st.shape
    wm, wc = dev_of(Wm), dev_of(Wc)
    if wm.numel() != k or wc.numel() != k:
        raise ValueError("Wm / Wc must have one weight per sigma point")
    nz, nstride = None, 0
    if noise_cov is not None:
        nz = dev_of(noise_cov)
        if nz.dim() == 0:
            nz = torch.eye(n, dtype=dt, device=dev) * nz
        nstride = 0 if nz.dim() == 2 else n * n
    x = torch.empty(N, n, dtype=dt, device=dev)
    P = torch.empty(N, n, n, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bke_unscented_transform(N, k, n, bke_dtype(dt), st.data_ptr(), wm.data_ptr(), wc.data_ptr(),
                                                       None if nz is None else nz.data_ptr(), nstride,
                                                       x.data_ptr(), P.data_ptr(), stream_ptr(dev)))
    if single:
        x, P = x[0], P[0]
    if is_t:
        return x, P
    return x.cpu().numpy(), P.cpu().numpy()
