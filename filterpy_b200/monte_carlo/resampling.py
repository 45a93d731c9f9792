"""GPU mirrors of ``filterpy.monte_carlo.systematic_resample`` / ``stratified_resample``
(filterpy/monte_carlo/resampling.py:117-150 / :80-114).

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

__all__ = ["systematic_resample", "stratified_resample", "ResamplePlan", "normalize_weights"]


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


def _run(weights, stratified):
    is_torch = isinstance(weights, torch.Tensor)
    if is_torch and weights.is_cuda:
        w = weights.to(torch.float64).contiguous()
        dev = w.device
    else:
        dev = require_cuda(None)
        w = torch.from_numpy(np.ascontiguousarray(np.asarray(weights, dtype=np.float64))).to(dev)
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
