"""GPU mirrors of ``filterpy.monte_carlo.systematic_resample`` / ``stratified_resample`` /
``multinomial_resample`` (filterpy/monte_carlo/resampling.py:117-150 / :80-114 / :153-176) and the
particle gather that follows them (docs/monte_carlo/resampling.rst:4-8).

Same call, same result: ``indexes`` is the ``int32`` array the reference returns for the same
weights and the same uniform draw(s) — bit for bit, because the CUDA path reproduces the strictly
sequential fp64 ``np.cumsum`` exactly (csrc/resample.cu).  The uniforms are drawn here from
``numpy.random.random`` exactly as the reference does (resampling.py:24,103,139), so seeding with
``np.random.seed`` reproduces the reference's stream.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch
from numpy.random import random

from .. import _lib
from .._dev import require_cuda, stream_ptr

__all__ = ["systematic_resample", "stratified_resample", "multinomial_resample", "residual_resample",
           "gather_particles", "exact_cumsum", "ResamplePlan", "normalize_weights",
           "residual_resample_with_uniforms"]


class ResamplePlan(object):
    """Pre-allocated workspace + output for repeated resampling of ``n`` particles on one GPU
    (nothing is allocated and the host is not synchronised per call)."""

    def __init__(self, n, device=None):
        self.n = int(n)
        self.device = require_cuda(device)
        self._lib = _lib.load()
        self.ws_bytes = int(self._lib.bke_resample_workspace_bytes(self.n))
        self.workspace = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_ptr = self.workspace.data_ptr() + off
        self.indexes = torch.empty(self.n, dtype=torch.int32, device=self.device)
        self._info = torch.zeros(8, dtype=torch.int32, device=self.device)
        self.cumsum_last = torch.zeros(1, dtype=torch.float64, device=self.device)

    def _check_w(self, weights):
        if not (isinstance(weights, torch.Tensor) and weights.is_cuda and weights.dtype == torch.float64
                and weights.is_contiguous() and weights.numel() == self.n):
            raise ValueError("weights must be a contiguous float64 CUDA tensor of %d elements" % self.n)

    def systematic(self, weights, u, out=None):
        """indexes for the offset ``u`` (resampling.py:139: positions = (u + arange(N)) / N)."""
        self._check_w(weights)
        out = self.indexes if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bke_systematic_resample(
                self.n, weights.data_ptr(), float(u), out.data_ptr(), self._ws_ptr, self.ws_bytes,
                self._info.data_ptr(), self.cumsum_last.data_ptr(), stream_ptr(self.device)))
        return out

    def stratified(self, weights, uniforms, out=None):
        """indexes for per-particle uniforms (resampling.py:103: positions = (U + range(N)) / N)."""
        self._check_w(weights)
        if not (isinstance(uniforms, torch.Tensor) and uniforms.is_cuda and uniforms.dtype == torch.float64
                and uniforms.is_contiguous() and uniforms.numel() == self.n):
            raise ValueError("uniforms must be a contiguous float64 CUDA tensor of %d elements" % self.n)
        out = self.indexes if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bke_stratified_resample(
                self.n, weights.data_ptr(), uniforms.data_ptr(), out.data_ptr(), self._ws_ptr, self.ws_bytes,
                self._info.data_ptr(), self.cumsum_last.data_ptr(), stream_ptr(self.device)))
        return out

    def normalized(self, weights, u=None, uniforms=None, out=None, weights_out=None):
        """Fused normalise + resample: ``systematic_resample(weights / S)`` (``stratified_resample`` when
        ``uniforms`` is given) with ``S`` the engine's sum of the weights.  Returns (indexes, S tensor)."""
        self._check_w(weights)
        out = self.indexes if out is None else out
        total = torch.zeros(1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bke_resample_normalized(
                self.n, weights.data_ptr(), float(u if u is not None else 0.0),
                uniforms.data_ptr() if uniforms is not None else None, out.data_ptr(),
                weights_out.data_ptr() if weights_out is not None else None, total.data_ptr(),
                self._ws_ptr, self.ws_bytes, self._info.data_ptr(), self.cumsum_last.data_ptr(),
                stream_ptr(self.device)))
        return out, total

    def cumsum(self, weights, out=None, last_one=False):
        """``np.cumsum(weights)`` bit for bit (sequential fp64 order), on the device."""
        self._check_w(weights)
        out = torch.empty(self.n, dtype=torch.float64, device=self.device) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bke_cumsum_exact(self.n, weights.data_ptr(), out.data_ptr(), int(bool(last_one)),
                                                  self._ws_ptr, self.ws_bytes, self._info.data_ptr(),
                                                  stream_ptr(self.device)))
        return out

    def multinomial(self, weights, uniforms, out=None, scratch=None, lut=True):
        """resampling.py:173-176 for the given uniforms: int64 indexes.  ``lut=False`` runs the plain
        bisection without the bracket table (same result)."""
        self._check_w(weights)
        if not (isinstance(uniforms, torch.Tensor) and uniforms.is_cuda and uniforms.dtype == torch.float64
                and uniforms.is_contiguous() and uniforms.numel() == self.n):
            raise ValueError("uniforms must be a contiguous float64 CUDA tensor of %d elements" % self.n)
        out = torch.empty(self.n, dtype=torch.int64, device=self.device) if out is None else out
        scratch = torch.empty(self.n, dtype=torch.float64, device=self.device) if scratch is None else scratch
        lut_ptr = self.indexes.data_ptr() if lut else None           # the plan's int32[n] doubles as the table
        with torch.cuda.device(self.device):
            _lib.check(self._lib.bke_multinomial_resample(
                self.n, weights.data_ptr(), uniforms.data_ptr(), out.data_ptr(), scratch.data_ptr(), lut_ptr,
                self._ws_ptr, self.ws_bytes, self._info.data_ptr(), stream_ptr(self.device)))
        return out

    def info(self):
        """int32[8] on the host: [0] positions >= cumsum[-1] (the reference raises IndexError),
        [1] sequential-fallback used, [2] tiles with raw elements, [3] long runs, [4] chain flag."""
        return self._info.cpu().numpy()

    def raise_if_overflow(self):
        inf = self.info()
        if inf[0] > 0:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (self.n, self.n))


def normalize_weights(weights, plan=None):
    """weights / weights.sum() on the device (IEEE division by the engine's tree-ordered sum).
    Returns (normalised weights, sum tensor)."""
    plan = plan or ResamplePlan(weights.numel(), weights.device)
    lib = plan._lib
    total = torch.zeros(1, dtype=torch.float64, device=weights.device)
    out = torch.empty_like(weights)
    with torch.cuda.device(weights.device):
        _lib.check(lib.bke_weights_sum(plan.n, weights.data_ptr(), total.data_ptr(), plan._ws_ptr, plan.ws_bytes,
                                       stream_ptr(weights.device)))
        _lib.check(lib.bke_weights_scale(plan.n, weights.data_ptr(), total.data_ptr(), out.data_ptr(),
                                         stream_ptr(weights.device)))
    return out, total


def _weights_on_device(weights):
    is_torch = isinstance(weights, torch.Tensor)
    if is_torch and weights.is_cuda:
        w = weights.to(torch.float64).contiguous()
        dev = w.device
    else:
        dev = require_cuda(None)
        w = torch.from_numpy(np.ascontiguousarray(np.asarray(weights, dtype=np.float64))).to(dev)
    return is_torch, w, dev


def _run(weights, stratified):
    is_torch, w, dev = _weights_on_device(weights)
    n = w.numel()
    if n == 0:
        if stratified:
            random(0)
        else:
            random()
        return torch.zeros(0, dtype=torch.int32, device=dev) if is_torch else np.zeros(0, 'i')
    plan = ResamplePlan(n, dev)
    if stratified:
        U = torch.from_numpy(random(n)).to(dev)          # resampling.py:103
        idx = plan.stratified(w, U)
    else:
        idx = plan.systematic(w, random())               # resampling.py:139
    plan.raise_if_overflow()                             # resampling.py:145 (IndexError)
    return idx if is_torch else idx.cpu().numpy()


def systematic_resample(weights):
    """resampling.py:117-150 on the GPU; returns ``ndarray`` int32 (or a CUDA tensor when given one)."""
    return _run(weights, False)


def stratified_resample(weights):
    """resampling.py:80-114 on the GPU."""
    return _run(weights, True)


def multinomial_resample(weights):
    """resampling.py:153-176 on the GPU: ``searchsorted(cumsum(w) with [-1] = 1, random(N))``.
    Returns int64 indexes (what ``np.searchsorted`` returns), ndarray or CUDA tensor like the input."""
    is_torch, w, dev = _weights_on_device(weights)
    n = w.numel()
    if n == 0:
        np.cumsum(np.zeros(0))[-1:]                       # the reference fails on cumulative_sum[-1] (:174)
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    plan = ResamplePlan(n, dev)
    U = torch.from_numpy(np.atleast_1d(random(n))).to(dev)   # resampling.py:176
    idx = plan.multinomial(w, U)
    return idx if is_torch else idx.cpu().numpy()


def residual_resample_with_uniforms(weights, uniforms_fn=random, max_sweeps=10000):
    """residual_resample for uniforms drawn by ``uniforms_fn(N - k)`` (the reference: ``random``);
    returns (indexes, info) with info = dict(k, sweeps, residual_sum)."""
    is_torch, w, dev = _weights_on_device(weights)
    n = w.numel()
    if n == 0:
        # resampling.py:70-72 on empty arrays: sum([]) = 0, cumulative_sum[-1] raises IndexError
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    lib = _lib.load()
    ws_bytes = int(lib.bke_residual_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    idx = torch.zeros(n, dtype=torch.int32, device=dev)              # np.zeros(N, 'i') (:52)
    csum = torch.empty(n, dtype=torch.float64, device=dev)
    k_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    rsum = torch.zeros(1, dtype=torch.float64, device=dev)
    st = stream_ptr(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.bke_residual_prepare(n, w.data_ptr(), idx.data_ptr(), csum.data_ptr(), k_dev.data_ptr(),
                                            rsum.data_ptr(), ws.data_ptr(), ws_bytes, st))
        k = int(k_dev.item())
        if k > n:                                                     # indexes[k] = i runs off the end (:61)
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (n, n))
        m = n - k
        U = np.atleast_1d(uniforms_fn(m))                             # resampling.py:74: random(N - k)
        sweeps = 0
        if m > 0:
            keys = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float64)).to(dev)
            r = [torch.empty(m, dtype=torch.int64, device=dev), torch.empty(m, dtype=torch.int64, device=dev)]
            changed = torch.zeros(1, dtype=torch.int32, device=dev)
            tail = idx[k:]
            # the int32 copies land directly in indexes[k:N] when that view is 4-byte aligned (always)
            _lib.check(lib.bke_searchsorted_bracket_sweep(n, csum.data_ptr(), m, keys.data_ptr(), None,
                                                          r[0].data_ptr(), tail.data_ptr(), changed.data_ptr(), st))
            cur = 0
            while True:
                sweeps += 1
                changed.zero_()
                _lib.check(lib.bke_searchsorted_bracket_sweep(n, csum.data_ptr(), m, keys.data_ptr(), r[cur].data_ptr(),
                                                              r[1 - cur].data_ptr(), tail.data_ptr(),
                                                              changed.data_ptr(), st))
                cur = 1 - cur
                if int(changed.item()) == 0:
                    break
                if sweeps >= max_sweeps:
                    raise RuntimeError("residual_resample: the bracket recurrence did not settle in %d sweeps"
                                       % max_sweeps)
    info = {"k": k, "sweeps": sweeps, "residual_sum": float(rsum.item())}
    return (idx if is_torch else idx.cpu().numpy()), info


def residual_resample(weights):
    """resampling.py:27-76 on the GPU, same call, same result for the same ``np.random`` state.

    The reference's ``residual = weights - num_copies`` (:69; not ``N*weights - num_copies``) is negative
    for every particle that got a copy, so the cumulative sum it bisects (:74) is not monotone and
    ``np.searchsorted``'s answer depends on the bracket NumPy carries from key to key.  The engine
    reproduces exactly that: the order-dependent sums in the reference's order, and NumPy's bracket
    recurrence as a parallel fixed-point iteration (csrc/residual.cu).  Returns ``ndarray`` int32 (or a
    CUDA tensor when given one)."""
    return residual_resample_with_uniforms(weights, random)[0]


def exact_cumsum(weights, last_one=False):
    """``np.cumsum(weights)`` (fp64, strictly sequential order) computed on the GPU, bit for bit."""
    is_torch, w, dev = _weights_on_device(weights)
    if w.numel() == 0:
        return w.clone() if is_torch else np.zeros(0)
    out = ResamplePlan(w.numel(), dev).cumsum(w, last_one=last_one)
    return out if is_torch else out.cpu().numpy()


def gather_particles(particles, indexes, out=None, check=True):
    """``particles[indexes]`` along axis 0 on the GPU — the step after every resample
    (docs/monte_carlo/resampling.rst:4-8: ``particles[:] = particles[indexes]``).

    ``particles`` is (N, ...) of any dtype, ``indexes`` int32 or int64 (what the resamplers return);
    NumPy in -> NumPy out, CUDA tensors in -> CUDA tensor out.  Raises IndexError for an index outside
    [0, N) like NumPy does (negative indexes are not wrapped); ``check=False`` skips that test and the
    host synchronisation it costs."""
    lib = _lib.load()
    is_torch = isinstance(particles, torch.Tensor)
    if is_torch and particles.is_cuda:
        dev = particles.device
        src = particles.contiguous()
    else:
        dev = require_cuda(None)
        src = torch.from_numpy(np.ascontiguousarray(np.asarray(particles))).to(dev)
    if isinstance(indexes, torch.Tensor):
        idx = indexes.to(dev)
    else:
        idx = torch.from_numpy(np.ascontiguousarray(np.asarray(indexes))).to(dev)
    if idx.dtype not in (torch.int32, torch.int64):
        raise IndexError("arrays used as indices must be of integer type")
    idx = idx.contiguous().reshape(-1)
    if src.dim() == 0:
        raise IndexError("too many indices for array")
    n_src = src.shape[0]
    n_out = idx.numel()
    row_bytes = (src.numel() // max(n_src, 1)) * src.element_size()
    shape = (n_out,) + tuple(src.shape[1:])
    if out is None:
        out = torch.empty(shape, dtype=src.dtype, device=dev)
    elif not (isinstance(out, torch.Tensor) and out.is_cuda and out.is_contiguous() and tuple(out.shape) == shape
              and out.dtype == src.dtype):
        raise ValueError("out must be a contiguous CUDA tensor of shape %s" % (shape,))
    if n_out and row_bytes:
        if n_src == 0:
            raise IndexError("index out of bounds for axis 0 with size 0")
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.bke_gather_rows(n_out, n_src, row_bytes, src.data_ptr(), idx.data_ptr(),
                                           1 if idx.dtype == torch.int64 else 0, out.data_ptr(), err.data_ptr(),
                                           stream_ptr(dev)))
        if check and int(err.item()):
            raise IndexError("index out of bounds for axis 0 with size %d" % n_src)
    return out if (is_torch and particles.is_cuda) else out.cpu().numpy()
