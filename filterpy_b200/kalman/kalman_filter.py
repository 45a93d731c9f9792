"""Host-side mirror of ``filterpy.kalman.KalmanFilter`` for a BANK of filters on one B200.

Same names, argument meaning and error behaviour as the reference
(``filterpy/kalman/kalman_filter.py``: ``__init__`` :387-434, ``predict`` :437-482, ``update``
:485-561, ``batch_filter`` :826-993, ``rts_smoother`` :995-1074; procedural ``predict`` :1571,
``update`` :1401, ``batch_filter`` :1664, ``rts_smoother`` :1792), with one addition: a leading ``n_filters`` axis.  All arithmetic runs in
the hand-written CUDA kernels behind the C-ABI (``include/bke.h``); this file only validates
shapes, owns the device tensors and fills the argument structs.  There is no CPU fallback.

Two modes:

* ``KalmanFilter(dim_x, dim_z)``  — *single* mode, a drop-in for one reference object:
  attributes come back as NumPy arrays with the reference's shapes (``x`` is ``(dim_x, 1)``
  by default, kalman_filter.py:399), the bank has one filter.
* ``KalmanFilter(dim_x, dim_z, n_filters=N)`` — *bank* mode: attributes are device tensors with a
  leading N axis, ``x[N,n]  P[N,n,n]  z[N,m]``; a model matrix may be given un-batched
  (``(n,n)``) = shared by the bank.

``predict()`` followed by ``update(z)`` is fused into ONE kernel launch (the predict is deferred
until the next ``update`` or until somebody looks at the state).
"""
import math
import sys

import numpy as np
import torch

from .. import _lib
from .._dev import StepGraph, bke_dtype, ptr, require_cuda, resolve_dtype, stream_ptr, to_dev
from ..common.helpers import reshape_z

__all__ = ["KalmanFilter", "predict", "update", "batch_filter", "rts_smoother"]


class _Linked(np.ndarray):
    """What the single-mode attribute getters hand out: a host copy of a device array that WRITES
    BACK.  The reference's attributes are the live arrays, so the usual idioms ``kf.P[2, 2] = 100``,
    ``kf.x[0] = z``, ``kf.F[0, 1] = dt``, ``kf.P *= 10`` must reach the filter; here they re-assign
    the attribute (which uploads it).  Views derived from it (``kf.P[2]``) do not write back."""

    def __new__(cls, arr, owner, name):
        obj = np.array(arr, copy=True).view(cls)
        obj._owner, obj._name = owner, name
        return obj

    def __array_finalize__(self, obj):
        self._owner, self._name = None, None

    def _push(self):
        if self._owner is not None:
            setattr(self._owner, self._name, np.array(self, copy=True).view(np.ndarray))

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        self._push()

    def _inplace(self, op, other):
        res = op(self.view(np.ndarray), other)
        np.ndarray.__setitem__(self, Ellipsis, res)
        self._push()
        return self

    def __iadd__(self, o): return self._inplace(np.add, o)
    def __isub__(self, o): return self._inplace(np.subtract, o)
    def __imul__(self, o): return self._inplace(np.multiply, o)
    def __itruediv__(self, o): return self._inplace(np.true_divide, o)


class KalmanFilter(object):
    def __init__(self, dim_x, dim_z, dim_u=0, n_filters=None, dtype=np.float64, device=None,
                 diagnostics=True):
        if dim_x < 1:
            raise ValueError('dim_x must be 1 or greater')      # kalman_filter.py:388-393
        if dim_z < 1:
            raise ValueError('dim_z must be 1 or greater')
        if dim_u < 0:
            raise ValueError('dim_u must be 0 or greater')
        self.dim_x, self.dim_z, self.dim_u = int(dim_x), int(dim_z), int(dim_u)
        self._single = n_filters is None
        self.n_filters = 1 if self._single else int(n_filters)
        if self.n_filters < 0:
            raise ValueError('n_filters must be 0 or greater')
        self._dtype = resolve_dtype(dtype)
        self._device = require_cuda(device)
        self._lib = _lib.load()
        self.diagnostics = bool(diagnostics)
        N, n, m = self.n_filters, self.dim_x, self.dim_z
        kw = dict(dtype=self._dtype, device=self._device)
        self._x = torch.zeros(N, n, **kw)
        self._P = torch.eye(n, **kw).repeat(N, 1, 1)
        self._Q = torch.eye(n, **kw)
        self._F = torch.eye(n, **kw)
        self._H = torch.zeros(m, n, **kw)
        self._R = torch.eye(m, **kw)
        self._B = None
        self._alpha_sq = 1.0
        self._x_col = True            # single mode: x is (n,1) like the reference default
        self._pending = None          # deferred predict: dict(u,B,F,Q)
        self._z = None
        self.inv = np.linalg.inv      # kept for API parity; only the default is supported
        if self.diagnostics:
            self._x_prior = self._x.clone(); self._P_prior = self._P.clone()
            self._x_post = self._x.clone(); self._P_post = self._P.clone()
            self._K = torch.zeros(N, n, m, **kw)
            self._y = torch.zeros(N, m, **kw)
            self._S = torch.zeros(N, m, m, **kw)
            self._SI = torch.zeros(N, m, m, **kw)
            self._ll = torch.full((N,), math.log(sys.float_info.min), **kw)
            self._status = torch.zeros(N, dtype=torch.int32, device=self._device)
        self._host = {k: np.ascontiguousarray(getattr(self, "_" + k).cpu().numpy()) for k in "FQHR"}
        self._has_update = False
        self._post_alias = False      # True: x_post / P_post are the live x / P (nothing has moved them since the update)
        self._version = 0            # bumped whenever a tensor the kernels read is re-bound
        self._args_cache = {}

    # ------------------------------------------------------------------ helpers
    def _model(self, a, rows, cols, name):
        """(rows,cols) -> shared; (N,rows,cols) -> per filter.  Returns a device tensor."""
        if np.isscalar(a):
            if rows != cols:
                raise ValueError("%s: a scalar needs a square matrix" % name)
            return torch.eye(rows, dtype=self._dtype, device=self._device) * float(a)
        t = to_dev(a, self._dtype, self._device)
        if t.dim() == 2 and tuple(t.shape) == (rows, cols):
            return t
        if t.dim() == 3 and tuple(t.shape) == (self.n_filters, rows, cols):
            return t
        if t.dim() == 1 and rows == 1 and t.shape[0] == cols:
            return t.reshape(1, cols)
        raise ValueError("%s must have shape (%d,%d) or (%d,%d,%d), got %s"
                         % (name, rows, cols, self.n_filters, rows, cols, tuple(t.shape)))

    @staticmethod
    def _stride(t):
        return 0 if t.dim() == 2 else t.shape[1] * t.shape[2]

    def _out(self, t):
        """bank mode: the device tensor; single mode: NumPy with the bank axis dropped."""
        if not self._single:
            return t
        return t[0].cpu().numpy()

    # ------------------------------------------------------------------ state attributes
    @property
    def x(self):
        self._flush()
        if not self._single:
            return self._x
        v = self._x[0].cpu().numpy()
        return _Linked(v.reshape(-1, 1) if self._x_col else v, self, "x")

    @x.setter
    def x(self, v):
        self._flush()
        n = self.dim_x
        t = to_dev(v, self._dtype, self._device)
        if self._single:
            if tuple(t.shape) == (n, 1):
                self._x_col = True
            elif tuple(t.shape) == (n,):
                self._x_col = False
            else:
                raise ValueError("x must have shape (%d,1) or (%d,), got %s" % (n, n, tuple(t.shape)))
            self._snapshot_post()
            self._x = t.reshape(1, n).clone()
            self._version += 1
        else:
            if t.dim() == 3 and t.shape[-1] == 1:
                t = t[..., 0]
            if tuple(t.shape) == (n,):
                t = t.expand(self.n_filters, n)
            if tuple(t.shape) != (self.n_filters, n):
                raise ValueError("x must have shape (%d,%d), got %s" % (self.n_filters, n, tuple(t.shape)))
            self._snapshot_post()
            self._x = t.contiguous().clone()
            self._version += 1

    @property
    def P(self):
        self._flush()
        return self._P if not self._single else _Linked(self._P[0].cpu().numpy(), self, "P")

    @P.setter
    def P(self, v):
        self._flush()
        n = self.dim_x
        if np.isscalar(v):
            v = np.eye(n) * v
        t = to_dev(v, self._dtype, self._device)
        if tuple(t.shape) == (n, n):
            t = t.expand(self.n_filters, n, n)
        if tuple(t.shape) != (self.n_filters, n, n):
            raise ValueError("P must have shape (%d,%d) or (%d,%d,%d)" % (n, n, self.n_filters, n, n))
        self._snapshot_post()
        self._P = t.contiguous().clone()
        self._version += 1

    def _adopt_state(self, x_t, P_t):
        """Re-bind the state to caller-owned device tensors (no copy) and hand back a spare pair, so
        that a caller (IMMEstimator's mixing step) can double-buffer.  When x_post / P_post are
        still the live state, the outgoing buffers BECOME the stored posterior and the previous
        posterior buffers are the spare pair: a rotation instead of a copy."""
        self._flush()
        if self.diagnostics and self._post_alias:
            spare = (self._x_post, self._P_post)
            self._x_post, self._P_post = self._x, self._P
            self._post_alias = False
        else:
            spare = (self._x, self._P)
        self._x, self._P = x_t, P_t
        self._version += 1
        return spare

    def _mk_model_prop(name, rows_attr, cols_attr):  # noqa: N805
        priv = "_" + name

        def get(self):
            t = getattr(self, priv)
            if t is None:
                return None
            if self._single:
                return _Linked(t.cpu().numpy(), self, name)
            # the caller may edit the live tensor in place: a deferred predict must run with the
            # model it was issued with (the reference's predict has already happened), and the
            # host copy can no longer be trusted
            self._flush()
            if self._host.pop(name, None) is not None:
                self._version += 1
            return t

        def set_(self, v):
            self._flush()                                   # predict(); kf.F = F2; update(): the predict used the OLD F
            self._version += 1
            self._host.pop(name, None)
            if v is None:
                setattr(self, priv, None)
                return
            cols = getattr(self, cols_attr)
            if name == "B" and cols == 0 and not np.isscalar(v):
                cols = int(np.shape(v)[-1]) if np.ndim(v) >= 1 else 1     # the reference never checks B against dim_u
                if np.ndim(v) == 1:
                    v = np.asarray(v).reshape(-1, 1); cols = 1
            t = self._model(v, getattr(self, rows_attr), cols, name)
            setattr(self, priv, t)
            if t.dim() == 2 and name in "FQHR":
                # host copy of a model shared by the bank: lets the kernels carry it in their launch
                # parameters (bke_kf_args.*_host)
                self._host[name] = np.ascontiguousarray(t.cpu().numpy())
        return property(get, set_)

    F = _mk_model_prop("F", "dim_x", "dim_x")
    Q = _mk_model_prop("Q", "dim_x", "dim_x")
    H = _mk_model_prop("H", "dim_z", "dim_x")
    R = _mk_model_prop("R", "dim_z", "dim_z")
    B = _mk_model_prop("B", "dim_x", "dim_u")
    del _mk_model_prop

    @property
    def alpha(self):
        """Fading-memory setting (kalman_filter.py:1242-1266)."""
        return self._alpha_sq ** .5

    @alpha.setter
    def alpha(self, value):
        if not np.isscalar(value) or value < 1:
            raise ValueError('alpha must be a float greater than 1')
        self._flush()
        self._alpha_sq = float(value) ** 2
        self._version += 1

    def _diag(self, name):
        if not self.diagnostics:
            raise AttributeError("%s is only kept when the filter is built with diagnostics=True" % name)
        self._flush()
        return getattr(self, "_" + name)

    def _vec_out(self, t):
        """x-like vectors follow the shape of x in single mode."""
        if not self._single:
            return t
        v = t[0].cpu().numpy()
        return v.reshape(-1, 1) if self._x_col else v

    x_prior = property(lambda self: self._vec_out(self._diag("x_prior")))
    P_prior = property(lambda self: self._out(self._diag("P_prior")))
    # x_post / P_post (kalman_filter.py:560-561) equal x / P until the next predict runs: they are
    # the live tensors until then, and are snapshotted only when a predict is launched on its own
    x_post = property(lambda self: self._vec_out(self._diag("x" if self._post_alias_now() else "x_post")))
    P_post = property(lambda self: self._out(self._diag("P" if self._post_alias_now() else "P_post")))

    def _post_alias_now(self):
        if not self.diagnostics:
            raise AttributeError("x_post / P_post are only kept when the filter is built with diagnostics=True")
        self._flush()
        return self._post_alias

    def _snapshot_post(self):
        if self.diagnostics and self._post_alias:
            self._x_post.copy_(self._x); self._P_post.copy_(self._P)
        self._post_alias = False
    K = property(lambda self: self._out(self._diag("K")))
    y = property(lambda self: self._vec_out(self._diag("y")))
    S = property(lambda self: self._out(self._diag("S")))
    SI = property(lambda self: self._out(self._diag("SI")))

    @property
    def z(self):
        if self._z is None:
            return np.array([[None] * self.dim_z]).T
        return self._vec_out(self._z)

    @property
    def status(self):
        """int32[N]: 0 ok, 1 = S was singular (the reference raises LinAlgError there)."""
        return self._diag("status")

    def check(self):
        """Raise ``np.linalg.LinAlgError`` if any filter hit a singular S (kalman_filter.py:541)."""
        st = self._diag("status")
        bad = int((st != 0).sum().item())
        if bad:
            raise np.linalg.LinAlgError("Singular matrix in %d of %d filters" % (bad, self.n_filters))

    @property
    def log_likelihood(self):
        """log-likelihood of the last measurement (kalman_filter.py:1203-1210)."""
        ll = self._diag("ll")
        return float(ll[0].item()) if self._single else ll

    @property
    def likelihood(self):
        """kalman_filter.py:1213-1223 (exp of the log-likelihood, floored at float min)."""
        ll = self._diag("ll")
        lk = torch.exp(ll).clamp_min(sys.float_info.min)
        return float(lk[0].item()) if self._single else lk

    @property
    def mahalanobis(self):
        """sqrt(y' SI y) (kalman_filter.py:1226-1239)."""
        y, SI = self._diag("y"), self._diag("SI")
        d = torch.sqrt(torch.einsum("ni,nij,nj->n", y, SI, y))
        return float(d[0].item()) if self._single else d

    # ------------------------------------------------------------------ predict / update
    def predict(self, u=None, B=None, F=None, Q=None):
        """kalman_filter.py:437-482.  Deferred: fused with the next ``update``."""
        self._flush()
        self._pending = dict(u=u, B=B, F=F, Q=Q)

    def _flush(self):
        if self._pending is not None:
            pend, self._pending = self._pending, None
            self._launch(_lib.BKE_DO_PREDICT, pend, None, None, None, None)

    def update(self, z, R=None, H=None, valid=None):
        """kalman_filter.py:485-561.  ``z`` is ``(N, dim_z)`` in bank mode (``valid`` — bool[N] —
        marks the filters that have a measurement; the others behave as ``z=None``); in single
        mode anything the reference accepts.  ``z=None`` skips the update for the whole bank."""
        pend, self._pending = self._pending, None
        if z is None:                                       # :515-520
            if pend is not None:
                self._launch(_lib.BKE_DO_PREDICT, pend, None, None, None, None)
            self._z = None
            if self.diagnostics:
                self._post_alias = True
                self._y.zero_()
            return
        m = self.dim_z
        if self._single:
            if H is None:
                z = reshape_z(z, m, 2 if self._x_col else 1)        # :527-529
            zt = to_dev(np.asarray(z, dtype=np.float64).reshape(-1), self._dtype, self._device)
            if zt.numel() != m:
                raise ValueError("z (shape %s) must be convertible to shape (%d, 1)" % (np.shape(z), m))
            zt = zt.reshape(1, m)
        elif (isinstance(z, torch.Tensor) and z.device == self._device and z.dtype == self._dtype
              and z.dim() == 2 and z.shape[0] == self.n_filters and z.shape[1] == m and z.is_contiguous()):
            zt = z                                              # already where the kernel wants it
        else:
            zt = to_dev(z, self._dtype, self._device)
            if zt.dim() == 3 and zt.shape[-1] == 1:
                zt = zt[..., 0]
            if tuple(zt.shape) != (self.n_filters, m):
                raise ValueError("z must have shape (%d,%d), got %s" % (self.n_filters, m, tuple(zt.shape)))
            zt = zt.contiguous()
        vt = None
        if valid is not None:
            vt = torch.as_tensor(valid, device=self._device).to(torch.uint8).contiguous()
            if tuple(vt.shape) != (self.n_filters,):
                raise ValueError("valid must have shape (%d,)" % self.n_filters)
        flags = _lib.BKE_DO_UPDATE | (_lib.BKE_DO_PREDICT if pend is not None else 0)
        self._launch(flags, pend, zt, vt, R, H)
        self._z = zt

    def _call(self, a):
        if torch.cuda.current_device() == self._device.index:
            _lib.check(self._lib.bke_kf_step(a, stream_ptr(self._device)))
        else:
            with torch.cuda.device(self._device):
                _lib.check(self._lib.bke_kf_step(a, stream_ptr(self._device)))

    def _launch(self, flags, pend, zt, vt, R, H):
        # steady state of a filter loop: nothing but z changed since the last identical call ->
        # reuse the argument struct (the Python side of a launch drops to a few microseconds)
        plain = R is None and H is None and (pend is None or (pend.get("u") is None and pend.get("B") is None
                                                               and pend.get("F") is None and pend.get("Q") is None))
        if plain:
            hit = self._args_cache.get(flags)
            if hit is not None and hit[0] == self._version:
                a = hit[1]
                a.z = ptr(zt); a.z_valid = ptr(vt)
                if not (flags & _lib.BKE_DO_UPDATE):
                    self._snapshot_post()                   # a predict on its own is about to move x, P
                self._call(a)
                if self.diagnostics and (flags & _lib.BKE_DO_UPDATE):
                    self._post_alias = True
                    if self._single:
                        self.check()
                return
        a = _lib.KfArgs()
        N, n, m = self.n_filters, self.dim_x, self.dim_z
        a.n_filters, a.dim_x, a.dim_z, a.dim_u = N, n, m, 0
        a.dtype = bke_dtype(self._dtype)
        a.flags = flags
        a.alpha_sq = self._alpha_sq
        a.x = a.x_out = ptr(self._x)
        a.P = a.P_out = ptr(self._P)
        keep = []
        if flags & _lib.BKE_DO_PREDICT:
            F = self._F if pend.get("F") is None else self._model(pend["F"], n, n, "F")
            Qo = pend.get("Q")
            Q = self._Q if Qo is None else self._model(Qo, n, n, "Q")      # scalar Q -> Q*I (:467-468)
            B = self._B if pend.get("B") is None else self._model(pend["B"], n, self.dim_u or np.shape(pend["B"])[-1], "B")
            u = pend.get("u")
            a.F, a.F_stride = ptr(F), self._stride(F)
            a.Q, a.Q_stride = ptr(Q), self._stride(Q)
            keep += [F, Q]
            if B is not None and u is not None:                             # :472-475
                du = B.shape[-1]
                ut = to_dev(u, self._dtype, self._device).reshape(-1, du) if not np.isscalar(u) else \
                    torch.full((1, 1), float(u), dtype=self._dtype, device=self._device)
                if ut.shape[0] not in (1, N):
                    raise ValueError("u must have shape (%d,) or (%d,%d)" % (du, N, du))
                a.dim_u = du
                a.B, a.B_stride = ptr(B), self._stride(B)
                a.u, a.u_stride = ptr(ut), (0 if ut.shape[0] == 1 else du)
                keep += [B, ut]
        if flags & _lib.BKE_DO_UPDATE:
            Rm = self._R if R is None else self._model(R, m, m, "R")        # scalar R -> R*I (:524-525)
            Hm = self._H if H is None else self._model(H, m, n, "H")
            a.H, a.H_stride = ptr(Hm), self._stride(Hm)
            a.R, a.R_stride = ptr(Rm), self._stride(Rm)
            a.z = ptr(zt)
            a.z_valid = ptr(vt)
            keep += [Rm, Hm, zt, vt]
        if plain and len(self._host) == 4:                  # every model shared and known on the host
            hm = [self._host[k] for k in "FQHR"]
            a.F_host, a.Q_host, a.H_host, a.R_host = (h.ctypes.data for h in hm)
            keep += hm
        if self.diagnostics:
            if flags & _lib.BKE_DO_PREDICT:
                a.x_prior, a.P_prior = ptr(self._x_prior), ptr(self._P_prior)
            if flags & _lib.BKE_DO_UPDATE:
                a.K, a.y, a.S, a.SI = ptr(self._K), ptr(self._y), ptr(self._S), ptr(self._SI)
                a.log_likelihood = ptr(self._ll)
                a.status = ptr(self._status)
        if not (flags & _lib.BKE_DO_UPDATE):
            self._snapshot_post()                           # a predict on its own is about to move x, P
        self._call(a)
        if plain:
            self._args_cache[flags] = (self._version, a, keep)      # keep: the tensors `a` points into
        if self.diagnostics and (flags & _lib.BKE_DO_UPDATE):
            self._post_alias = True
            if self._single:
                self.check()

    def capture(self, fn, warmup=2):
        """Capture ``fn`` — a fixed sequence of ``predict()/update(z_buffer)`` calls on this bank —
        into a CUDA graph; ``.replay()`` re-runs it with a single launch (see ``StepGraph``).  The
        state is NOT rolled back after the warm-up / capture runs: set ``x`` / ``P`` afterwards."""
        self._flush()
        return StepGraph(fn, self._device, warmup)

    # ------------------------------------------------------------------ batch_filter
    def batch_filter(self, zs, Fs=None, Qs=None, Hs=None, Rs=None, Bs=None, us=None,
                     update_first=False, saver=None, valid=None):
        """kalman_filter.py:826-993.  Bank mode: ``zs[T,N,dim_z]`` (``valid[T,N]`` optional) ->
        device tensors ``means[T,N,n] covariances[T,N,n,n] means_p covariances_p``.  Single mode:
        ``zs`` as in the reference (entries may be None), NumPy outputs with its shapes.

        With time-constant models the whole T-epoch loop is ONE kernel (bke_kf_batch_filter);
        per-epoch ``Fs/Qs/Hs/Rs/Bs/us`` or a ``saver`` run one fused launch per epoch."""
        self._flush()
        N, n, m = self.n_filters, self.dim_x, self.dim_z
        # (the reference's np.size(zs, 0), kalman_filter.py:951; a list that mixes None with arrays is
        # ragged for NumPy >= 1.24, so lists are measured with len)
        T = zs.shape[0] if isinstance(zs, torch.Tensor) else (len(zs) if isinstance(zs, (list, tuple)) else np.size(zs, 0))
        if self._single:
            zarr = np.zeros((T, 1, m))
            vmask = np.ones((T, 1), dtype=bool)
            for i, z in enumerate(zs):
                if z is None:
                    vmask[i, 0] = False
                else:
                    zarr[i, 0] = np.asarray(z, dtype=np.float64).reshape(-1)[:m] if np.size(z) == m else \
                        reshape_z(z, m, 1)
            zt = to_dev(zarr, self._dtype, self._device)
            vt = None if vmask.all() else torch.from_numpy(vmask.astype(np.uint8)).to(self._device)
        else:
            zt = to_dev(zs, self._dtype, self._device)
            if tuple(zt.shape) != (T, N, m):
                raise ValueError("zs must have shape (T,%d,%d), got %s" % (N, m, tuple(zt.shape)))
            vt = None if valid is None else torch.as_tensor(valid, device=self._device).to(torch.uint8).contiguous()
        kw = dict(dtype=self._dtype, device=self._device)
        means = torch.empty(T, N, n, **kw); means_p = torch.empty(T, N, n, **kw)
        covs = torch.empty(T, N, n, n, **kw); covs_p = torch.empty(T, N, n, n, **kw)
        per_epoch = any(v is not None for v in (Fs, Qs, Hs, Rs, Bs, us)) or saver is not None
        if not per_epoch:
            b = _lib.KfBatchArgs()
            a = b.step
            a.n_filters, a.dim_x, a.dim_z, a.dim_u = N, n, m, 0
            a.dtype = bke_dtype(self._dtype)
            a.flags = _lib.BKE_DO_PREDICT | _lib.BKE_DO_UPDATE | (_lib.BKE_UPDATE_FIRST if update_first else 0)
            a.alpha_sq = self._alpha_sq
            a.x = a.x_out = ptr(self._x); a.P = a.P_out = ptr(self._P)
            a.F, a.F_stride = ptr(self._F), self._stride(self._F)
            a.Q, a.Q_stride = ptr(self._Q), self._stride(self._Q)
            a.H, a.H_stride = ptr(self._H), self._stride(self._H)
            a.R, a.R_stride = ptr(self._R), self._stride(self._R)
            if self.diagnostics:
                a.status = ptr(self._status)
            b.n_steps = T
            b.zs, b.zs_valid = ptr(zt), ptr(vt)
            b.means, b.covariances, b.means_p, b.covariances_p = ptr(means), ptr(covs), ptr(means_p), ptr(covs_p)
            with torch.cuda.device(self._device):
                _lib.check(self._lib.bke_kf_batch_filter(b, stream_ptr(self._device)))
        else:
            def at(lst, i):
                return None if lst is None else lst[i]
            for i in range(T):
                v = None if vt is None else vt[i]
                zi = zt[i]
                if update_first:
                    self._launch(_lib.BKE_DO_UPDATE, None, zi, v, at(Rs, i), at(Hs, i))
                    means[i].copy_(self._x); covs[i].copy_(self._P)
                    self._launch(_lib.BKE_DO_PREDICT, dict(u=at(us, i), B=at(Bs, i), F=at(Fs, i), Q=at(Qs, i)),
                                 None, None, None, None)
                    means_p[i].copy_(self._x); covs_p[i].copy_(self._P)
                else:
                    pend = dict(u=at(us, i), B=at(Bs, i), F=at(Fs, i), Q=at(Qs, i))
                    if self.diagnostics:
                        self._launch(_lib.BKE_DO_PREDICT | _lib.BKE_DO_UPDATE, pend, zi, v, at(Rs, i), at(Hs, i))
                        means_p[i].copy_(self._x_prior); covs_p[i].copy_(self._P_prior)
                    else:
                        self._launch(_lib.BKE_DO_PREDICT, pend, None, None, None, None)
                        means_p[i].copy_(self._x); covs_p[i].copy_(self._P)
                        self._launch(_lib.BKE_DO_UPDATE, None, zi, v, at(Rs, i), at(Hs, i))
                    means[i].copy_(self._x); covs[i].copy_(self._P)
                if saver is not None:
                    saver.save()
        if self.diagnostics:
            self._post_alias = not update_first             # update_first ends on a predict (:980-985)
            if update_first and T > 0:
                self._x_post.copy_(means[T - 1]); self._P_post.copy_(covs[T - 1])
        if not self._single:
            return means, covs, means_p, covs_p
        self.check() if self.diagnostics else None
        shp = (T, n, 1) if self._x_col else (T, n)
        return (means[:, 0].cpu().numpy().reshape(shp), covs[:, 0].cpu().numpy(),
                means_p[:, 0].cpu().numpy().reshape(shp), covs_p[:, 0].cpu().numpy())

    def rts_smoother(self, Xs, Ps, Fs=None, Qs=None, inv=None):
        """Rauch-Tung-Striebel smoother over ``batch_filter``'s output (kalman_filter.py:995-1074).

        Bank mode: ``Xs[T,N,n]``, ``Ps[T,N,n,n]`` (the ``means`` / ``covariances`` tensors
        ``batch_filter`` returns) -> ``(x, P, K, Pp)`` tensors of the same layout.  Single mode:
        NumPy ``Xs (T,n)`` or ``(T,n,1)``, ``Ps (T,n,n)`` like the reference.  ``Fs`` / ``Qs`` are
        per-epoch lists (length T; step k uses entry k+1, :1068) or None = the filter's F / Q.
        Only the default ``inv`` (np.linalg.inv) is offered on the GPU."""
        if inv is not None and inv is not np.linalg.inv:
            raise NotImplementedError("rts_smoother: only the default inv (np.linalg.inv) runs on the GPU")
        if len(Xs) != len(Ps):
            raise ValueError('length of Xs and Ps must be the same')
        self._flush()
        n, N = self.dim_x, self.n_filters
        return _rts(self, Xs, Ps, Fs, Qs, 1, self._single, N, n)

    def __repr__(self):
        return "KalmanFilter bank (B200): n_filters=%d dim_x=%d dim_z=%d dtype=%s device=%s" % (
            self.n_filters, self.dim_x, self.dim_z, self._dtype, self._device)


def _rts(kf, Xs, Ps, Fs, Qs, shift, single, N, n):
    """Shared body of the two rts_smoother forms: fills bke_rts_args and launches."""
    dtype, device = kf._dtype, kf._device
    is_np = not isinstance(Xs, torch.Tensor)
    Xt = to_dev(Xs, dtype, device)
    Pt = to_dev(Ps, dtype, device)
    T = Xt.shape[0]
    col = False
    if single:
        col = Xt.dim() == 3 and Xt.shape[-1] == 1
        Xt = Xt.reshape(T, 1, n)
        Pt = Pt.reshape(T, 1, n, n)
    if tuple(Xt.shape) != (T, N, n) or tuple(Pt.shape) != (T, N, n, n):
        raise ValueError("Xs / Ps must have shapes (T,%d,%d) / (T,%d,%d,%d), got %s / %s"
                         % (N, n, N, n, n, tuple(Xt.shape), tuple(Pt.shape)))
    Xt = Xt.contiguous(); Pt = Pt.contiguous()

    def model(lst, default, name):
        """-> (tensor, per-filter stride, per-epoch stride)"""
        if lst is None:
            return default, KalmanFilter._stride(default), 0
        if len(lst) != T:
            raise ValueError("%s must have one entry per epoch (%d), got %d" % (name, T, len(lst)))
        mats = [to_dev(m, dtype, device) for m in lst]
        if any(m.shape[-2:] != (n, n) for m in mats):
            raise ValueError("%s entries must be (%d,%d)" % (name, n, n))
        batched = [m.dim() == 3 for m in mats]
        if any(batched) and not all(batched):
            mats = [m if m.dim() == 3 else m.expand(N, n, n) for m in mats]
        t = torch.stack(mats).contiguous()               # (T,n,n) or (T,N,n,n)
        if t.dim() == 4:
            return t, n * n, N * n * n
        return t, 0, n * n

    Ft, sF, tF = model(Fs, kf._F, "Fs")
    Qt, sQ, tQ = model(Qs, kf._Q, "Qs")
    kw = dict(dtype=dtype, device=device)
    x = torch.empty(T, N, n, **kw); P = torch.empty(T, N, n, n, **kw)
    K = torch.empty(T, N, n, n, **kw); Pp = torch.empty(T, N, n, n, **kw)
    status = torch.zeros(N, dtype=torch.int32, device=device)
    a = _lib.RtsArgs()
    a.n_filters, a.n_steps, a.dim_x, a.dtype, a.model_shift = N, T, n, bke_dtype(dtype), shift
    a.Xs, a.Ps = ptr(Xt), ptr(Pt)
    a.F, a.F_stride, a.F_step_stride = ptr(Ft), sF, tF
    a.Q, a.Q_stride, a.Q_step_stride = ptr(Qt), sQ, tQ
    a.x_out, a.P_out, a.K, a.Pp = ptr(x), ptr(P), ptr(K), ptr(Pp)
    a.status = ptr(status)
    with torch.cuda.device(device):
        _lib.check(_lib.load().bke_kf_rts_smoother(a, stream_ptr(device)))
    if not single:
        return x, P, K, Pp
    if int(status[0].item()) != 0:
        raise np.linalg.LinAlgError("Singular matrix")
    shp = (T, n, 1) if col else (T, n)
    out = (x[:, 0].reshape(shp), P[:, 0], K[:, 0], Pp[:, 0])
    return tuple(o.cpu().numpy() for o in out) if is_np else out


# ---------------------------------------------------------------------- procedural form
def _as_bank(a, tail, name, dtype, device):
    """Returns (tensor with a leading batch axis or shared, was_batched)."""
    t = to_dev(a, dtype, device)
    if t.dim() == len(tail):
        return t, False
    if t.dim() == len(tail) + 1:
        return t, True
    raise ValueError("%s has a bad number of dimensions: %s" % (name, tuple(t.shape)))


def _proc_filter(x, P, dtype, device):
    xa = np.asarray(x) if not isinstance(x, torch.Tensor) else x
    col = xa.ndim >= 2 and xa.shape[-1] == 1 and (np.ndim(P) == xa.ndim)
    xt = to_dev(x, resolve_dtype(dtype), require_cuda(device))
    if col:
        xt = xt[..., 0]
    batched = xt.dim() == 2
    N = xt.shape[0] if batched else None
    return xt, col, batched, N


def _run_proc(flags, x, P, F=None, Q=None, u=None, B=None, alpha=1., z=None, R=None, H=None,
              return_all=False, dtype=None, device=None):
    is_torch = isinstance(x, torch.Tensor)
    if dtype is None:
        dtype = x.dtype if is_torch else (np.asarray(x).dtype if np.asarray(x).dtype in (np.float32, np.float64) else np.float64)
    xt, col, batched, N = _proc_filter(x, P, dtype, device)
    n = xt.shape[-1]
    m = None
    if flags & _lib.BKE_DO_UPDATE:
        Hm = np.eye(n) if H is None else H
        m = (Hm.shape[-2] if hasattr(Hm, "shape") and np.ndim(Hm) >= 2 else 1)
    kf = KalmanFilter(n, m or 1, n_filters=N, dtype=dtype, device=device, diagnostics=return_all)
    if batched:
        kf.x = xt
    else:
        kf._x = xt.reshape(1, n).clone(); kf._x_col = col
    kf.P = P
    kf._alpha_sq = float(alpha) ** 2
    if flags & _lib.BKE_DO_PREDICT:
        kf.predict(u=None if (np.isscalar(u) and u == 0) else u, B=None if np.isscalar(B) else B,
                   F=np.eye(n) * F if np.isscalar(F) else F, Q=np.eye(n) * Q if np.isscalar(Q) else Q)
    if flags & _lib.BKE_DO_UPDATE:
        if z is None:
            kf._flush()
        else:
            kf.update(z, R=R, H=Hm if np.ndim(Hm) >= 2 else np.reshape(Hm, (1, -1)))
    kf._flush()

    def conv(t, vec=False):
        if batched:
            t2 = t[..., None] if (vec and col) else t
            return t2 if is_torch else t2.cpu().numpy()
        v = t[0]
        if vec and col:
            v = v[..., None]
        return v if is_torch else v.cpu().numpy()
    out = [conv(kf._x, True), conv(kf._P)]
    if return_all:
        if z is None:
            out += [None, None, None, None]
        else:
            ll = kf._ll if batched else kf._ll[0]
            out += [conv(kf._y, True), conv(kf._K), conv(kf._S), ll if is_torch else (ll.cpu().numpy() if batched else float(ll.item()))]
    return tuple(out)


def predict(x, P, F=1, Q=0, u=0, B=1, alpha=1., dtype=None, device=None):
    """Procedural predict (kalman_filter.py:1571-1621) on the GPU; x may carry a leading bank axis."""
    return _run_proc(_lib.BKE_DO_PREDICT, x, P, F=F, Q=Q, u=u, B=B, alpha=alpha, dtype=dtype, device=device)


def update(x, P, z, R, H=None, return_all=False, dtype=None, device=None):
    """Procedural update (kalman_filter.py:1401-1508) on the GPU."""
    if z is None:
        if return_all:
            return x, P, None, None, None, None
        return x, P
    return _run_proc(_lib.BKE_DO_UPDATE, x, P, z=z, R=R, H=H, return_all=return_all, dtype=dtype, device=device)


def batch_filter(x, P, zs, Fs, Qs, Hs, Rs, Bs=None, us=None, update_first=False, saver=None,
                 dtype=np.float64, device=None):
    """Procedural batch_filter (kalman_filter.py:1664-1788) for ONE filter with per-epoch model
    lists, run on the GPU (one fused launch per epoch)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    H0 = np.atleast_2d(Hs[0])
    kf = KalmanFilter(n, H0.shape[0], dtype=dtype, device=device, diagnostics=True)
    kf.x = x; kf.P = P
    T = len(zs) if isinstance(zs, (list, tuple)) else np.size(zs, 0)
    if us is None:
        us, Bs = None, None
    return kf.batch_filter(zs, Fs=list(Fs), Qs=list(Qs), Hs=[np.atleast_2d(h) for h in Hs], Rs=list(Rs),
                           Bs=Bs, us=us, update_first=update_first, saver=saver)


def rts_smoother(Xs, Ps, Fs, Qs, dtype=np.float64, device=None):
    """Procedural RTS smoother (kalman_filter.py:1792-1858) for ONE filter, run on the GPU.
    ``Fs`` / ``Qs``: one (n,n) matrix, or a per-epoch list (step k uses entry k, :1852)."""
    if len(Xs) != len(Ps):
        raise ValueError('length of Xs and Ps must be the same')
    Xa = np.asarray(Xs, dtype=np.float64)
    n = Xa.shape[1]
    T = Xa.shape[0]
    kf = KalmanFilter(n, 1, dtype=dtype, device=device, diagnostics=False)

    def per_epoch(m):
        a = np.asarray(m, dtype=np.float64)
        return [a] * T if a.ndim == 2 else list(m)
    return _rts(kf, Xa, np.asarray(Ps, dtype=np.float64), per_epoch(Fs), per_epoch(Qs), 0, True, 1, n)
