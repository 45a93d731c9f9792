"""Host-side mirror of ``filterpy.kalman.UnscentedKalmanFilter`` for a BANK of filters on one B200
(filterpy/kalman/UKF.py: ``__init__`` :284-362, ``predict`` :364-411, ``update`` :413-491,
``batch_filter`` :524-632).

The reference calls Python ``fx(x, dt)`` / ``hx(x)`` per sigma point (UKF.py:521-522, 463-464); a
device kernel cannot call back into Python, so ``fx`` / ``hx`` name a device-side model from the
closed set of include/bke.h instead:

    fx = LinearFx(F)        x' = F x                    (F: (n,n) shared or (N,n,n))
    fx = ConstVelFx()       (p0,v0,p1,v1,...): p += dt v
    hx = LinearHx(H)        z = H x
    hx = RangeAzElHx()      n=6 (x,vx,y,vy,z,vz) -> (range, azimuth, elevation)
    hx = RangeBearingHx()   n=4 (x,vx,y,vy)      -> (range, bearing)

Passing a Python callable, or any of the hook arguments (sqrt_fn, x_mean_fn, z_mean_fn,
residual_x, residual_z, state_add, per-call UT / fx / hx), raises NotImplementedError: there is no
CPU fallback.  As for the linear filter, ``n_filters=None`` gives a single-filter drop-in whose
attributes are NumPy arrays (``x`` is 1-D, UKF.py:298).
"""
import ctypes
import math
import sys

import numpy as np
import torch

from .. import _lib
from .._dev import bke_dtype, ptr, require_cuda, resolve_dtype, stream_ptr, to_dev
from .kalman_filter import _Linked

__all__ = ["UnscentedKalmanFilter", "LinearFx", "ConstVelFx", "LinearHx", "RangeAzElHx", "RangeBearingHx",
           "DeviceFx", "DeviceHx"]


class LinearFx(object):
    model = _lib.BKE_FX_LINEAR

    def __init__(self, F):
        self.F = F


class ConstVelFx(object):
    model = _lib.BKE_FX_CONST_VEL
    F = None


class LinearHx(object):
    model = _lib.BKE_HX_LINEAR

    def __init__(self, H):
        self.H = H


class RangeAzElHx(object):
    model = _lib.BKE_HX_RANGE_AZ_EL
    H = None


class RangeBearingHx(object):
    model = _lib.BKE_HX_RANGE_BEARING
    H = None


class _DeviceModel(object):
    """A process / measurement function given as CUDA C++ source text (the reference takes Python callables:
    UKF.py:284-288, called per sigma point at :521-522 / :463-464).  The text defines, for the element
    type ``real`` (float or double, whichever the filter uses; ``BKE_DIM_X`` / ``BKE_DIM_Z`` are defined)::

        __device__ void fx(const real *x, real *x_out, real dt, const real *args)      # DeviceFx
        __device__ void hx(const real *x, real *z_out, const real *args)               # DeviceHx

    ``arg_names`` are the keyword arguments of the reference's callable (``predict(**fx_args)`` /
    ``update(z, **hx_args)``), delivered to the function as ``args[0..]`` in that order; ``defaults`` gives
    their values until a call overrides them.  A value is a scalar (whole bank) or an array ``(N,)``
    (one per filter).  The text is compiled at run time into an instance of the same kernel the built-in
    models use (NVRTC, sm_100a; ``csrc/ukf_rtc.cu``)."""
    F = None
    H = None

    def __init__(self, source, arg_names=(), **defaults):
        self.source = str(source)
        self.arg_names = tuple(arg_names)
        unknown = set(defaults) - set(self.arg_names)
        if unknown:
            raise TypeError("defaults for unknown arguments: %s" % sorted(unknown))
        self.values = dict(defaults)

    def pack(self, overrides, n_filters, dtype, device):
        """args vector(s) for the kernel: (tensor or None, stride)."""
        unknown = set(overrides) - set(self.arg_names)
        if unknown:
            raise TypeError("unexpected keyword arguments %s (declared: %s)" % (sorted(unknown), list(self.arg_names)))
        self.values.update(overrides)
        if not self.arg_names:
            return None, 0
        missing = [k for k in self.arg_names if k not in self.values]
        if missing:
            raise TypeError("missing values for the model arguments %s" % missing)
        vals = [self.values[k] for k in self.arg_names]
        if all(np.ndim(v) == 0 and not isinstance(v, torch.Tensor) for v in vals):
            return torch.tensor([float(v) for v in vals], dtype=dtype, device=device), 0
        cols = []
        for k, v in zip(self.arg_names, vals):
            t = torch.as_tensor(v, device=device).to(dtype).reshape(-1)
            if t.numel() == 1:
                t = t.expand(n_filters)
            if t.numel() != n_filters:
                raise ValueError("argument %s must be a scalar or have one value per filter (%d)" % (k, n_filters))
            cols.append(t)
        return torch.stack(cols, dim=1).contiguous(), len(cols)


class DeviceFx(_DeviceModel):
    model = _lib.BKE_FX_USER


class DeviceHx(_DeviceModel):
    model = _lib.BKE_HX_USER


_compiled_models = {}


def _compile_model(lib, dim_x, dim_z, dtype_id, fx, hx):
    """One NVRTC build per (shape, dtype, source); shared by every filter object that uses it."""
    src = "\n".join(m.source for m in (fx, hx) if isinstance(m, _DeviceModel))
    key = (dim_x, dim_z, dtype_id, fx.model, hx.model, src)
    h = _compiled_models.get(key)
    if h is None:
        out = ctypes.c_void_p()
        _lib.check(lib.bke_ukf_model_compile(dim_x, dim_z, dtype_id, fx.model, hx.model, src.encode(),
                                             _lib.kernel_include_dirs().encode(), ctypes.byref(out)))
        h = _compiled_models[key] = out
    return h


def _no_hook(name, v):
    if v is not None:
        raise NotImplementedError(
            "%s is a Python callable hook; the GPU path implements the reference defaults only and "
            "has no CPU fallback (see filterpy_b200/kalman/UKF.py)" % name)


class UnscentedKalmanFilter(object):
    def __init__(self, dim_x, dim_z, dt, hx, fx, points, sqrt_fn=None, x_mean_fn=None, z_mean_fn=None,
                 residual_x=None, residual_z=None, state_add=None,
                 n_filters=None, dtype=np.float64, device=None, diagnostics=True):
        for nm, v in (("sqrt_fn", sqrt_fn), ("x_mean_fn", x_mean_fn), ("z_mean_fn", z_mean_fn),
                      ("residual_x", residual_x), ("residual_z", residual_z), ("state_add", state_add)):
            _no_hook(nm, v)
        if not hasattr(fx, "model") or not hasattr(hx, "model"):
            raise NotImplementedError(
                "fx / hx must be device-side models (LinearFx, ConstVelFx, LinearHx, RangeAzElHx, "
                "RangeBearingHx, or DeviceFx / DeviceHx around CUDA source text): Python callables cannot "
                "run inside the CUDA kernel and there is no CPU fallback")
        if points.n != dim_x:
            raise ValueError("expected size(x) {}, but size is {}".format(points.n, dim_x))   # sigma_points.py:153
        self._dim_x, self._dim_z = int(dim_x), int(dim_z)
        self._single = n_filters is None
        self.n_filters = 1 if self._single else int(n_filters)
        self._dtype = resolve_dtype(dtype)
        self._device = require_cuda(device)
        self._lib = _lib.load()
        self.points_fn = points
        self._dt = dt
        self._num_sigmas = points.num_sigmas()
        self.fx, self.hx = fx, hx
        self.Wm, self.Wc = points.Wm, points.Wc
        self.diagnostics = bool(diagnostics)
        N, n, m = self.n_filters, self._dim_x, self._dim_z
        kw = dict(dtype=self._dtype, device=self._device)
        self._x = torch.zeros(N, n, **kw)
        self._P = torch.eye(n, **kw).repeat(N, 1, 1)
        self._Q = torch.eye(n, **kw)
        self._R = torch.eye(m, **kw)
        self._F = None if fx.F is None else self._model(fx.F, n, n, "F")
        self._H = None if hx.H is None else self._model(hx.H, m, n, "H")
        self._pending = None
        self._z = None
        self._user_model = None
        self._fx_args = self._hx_args = (None, 0)
        if isinstance(fx, _DeviceModel) or isinstance(hx, _DeviceModel):
            with torch.cuda.device(self._device):
                self._user_model = _compile_model(self._lib, n, m, bke_dtype(self._dtype), fx, hx)
            if isinstance(fx, _DeviceModel):
                self._fx_args = fx.pack({}, N, self._dtype, self._device) if all(k in fx.values for k in fx.arg_names) else (None, 0)
            if isinstance(hx, _DeviceModel):
                self._hx_args = hx.pack({}, N, self._dtype, self._device) if all(k in hx.values for k in hx.arg_names) else (None, 0)
        if self.diagnostics:
            self._x_prior = self._x.clone(); self._P_prior = self._P.clone()
            self._x_post = self._x.clone(); self._P_post = self._P.clone()
            self._K = torch.zeros(N, n, m, **kw); self._y = torch.zeros(N, m, **kw)
            self._S = torch.zeros(N, m, m, **kw); self._SI = torch.zeros(N, m, m, **kw)
            self._ll = torch.full((N,), math.log(sys.float_info.min), **kw)
            self._status = torch.zeros(N, dtype=torch.int32, device=self._device)

    # ------------------------------------------------------------------ plumbing (as KalmanFilter)
    def _model(self, a, rows, cols, name):
        if np.isscalar(a):
            return torch.eye(rows, dtype=self._dtype, device=self._device) * float(a)
        t = to_dev(a, self._dtype, self._device)
        if tuple(t.shape) == (rows, cols) or tuple(t.shape) == (self.n_filters, rows, cols):
            return t
        raise ValueError("%s must have shape (%d,%d) or (%d,%d,%d), got %s"
                         % (name, rows, cols, self.n_filters, rows, cols, tuple(t.shape)))

    @staticmethod
    def _stride(t):
        return 0 if t.dim() == 2 else t.shape[1] * t.shape[2]

    def _out(self, t):
        return t if not self._single else t[0].cpu().numpy()

    @property
    def x(self):
        self._flush()
        return self._x if not self._single else _Linked(self._x[0].cpu().numpy(), self, "x")

    @x.setter
    def x(self, v):
        self._flush()
        t = to_dev(v, self._dtype, self._device)
        if tuple(t.shape) == (self._dim_x,):
            t = t.expand(self.n_filters, self._dim_x)
        if tuple(t.shape) != (self.n_filters, self._dim_x):
            raise ValueError("x must have shape (%d,) or (%d,%d)" % (self._dim_x, self.n_filters, self._dim_x))
        self._x = t.contiguous().clone()

    @property
    def P(self):
        self._flush()
        return self._P if not self._single else _Linked(self._P[0].cpu().numpy(), self, "P")

    @P.setter
    def P(self, v):
        self._flush()
        n = self._dim_x
        if np.isscalar(v):
            v = np.eye(n) * v
        t = to_dev(v, self._dtype, self._device)
        if tuple(t.shape) == (n, n):
            t = t.expand(self.n_filters, n, n)
        if tuple(t.shape) != (self.n_filters, n, n):
            raise ValueError("P must have shape (%d,%d) or (%d,%d,%d)" % (n, n, self.n_filters, n, n))
        self._P = t.contiguous().clone()

    # A deferred predict() must run with the Q it was issued with (the reference's predict has
    # already happened): the setters, and the bank-mode getters that hand out the live tensor,
    # flush it first.  Single mode returns write-back arrays so that ``ukf.P[2, 2] = 100`` /
    # ``ukf.Q[0, 0] = q`` reach the filter as they do in the reference.
    def _get_model(self, name):
        t = getattr(self, "_" + name)
        if self._single:
            return _Linked(t.cpu().numpy(), self, name)
        self._flush()
        return t

    def _set_model(self, name, v, dim):
        self._flush()
        setattr(self, "_" + name, self._model(v, dim, dim, name))

    Q = property(lambda self: self._get_model("Q"), lambda self, v: self._set_model("Q", v, self._dim_x))
    R = property(lambda self: self._get_model("R"), lambda self, v: self._set_model("R", v, self._dim_z))

    def _diag(self, name):
        if not self.diagnostics:
            raise AttributeError("%s is only kept when the filter is built with diagnostics=True" % name)
        self._flush()
        return getattr(self, "_" + name)

    x_prior = property(lambda self: self._out(self._diag("x_prior")))
    P_prior = property(lambda self: self._out(self._diag("P_prior")))
    x_post = property(lambda self: self._out(self._diag("x_post")))
    P_post = property(lambda self: self._out(self._diag("P_post")))
    K = property(lambda self: self._out(self._diag("K")))
    y = property(lambda self: self._out(self._diag("y")))
    S = property(lambda self: self._out(self._diag("S")))
    SI = property(lambda self: self._out(self._diag("SI")))
    status = property(lambda self: self._diag("status"))

    @property
    def z(self):
        if self._z is None:
            return np.array([[None] * self._dim_z]).T
        return self._out(self._z)

    @property
    def log_likelihood(self):
        ll = self._diag("ll")
        return float(ll[0].item()) if self._single else ll

    @property
    def likelihood(self):
        lk = torch.exp(self._diag("ll")).clamp_min(sys.float_info.min)
        return float(lk[0].item()) if self._single else lk

    @property
    def mahalanobis(self):
        y, SI = self._diag("y"), self._diag("SI")
        d = torch.sqrt(torch.einsum("ni,nij,nj->n", y, SI, y))
        return float(d[0].item()) if self._single else d

    def check(self):
        """Raise LinAlgError where the reference would (non-PD P in cholesky, singular S)."""
        st = self._diag("status")
        bad = int((st != 0).sum().item())
        if bad:
            raise np.linalg.LinAlgError("%d of %d filters: matrix not positive definite / singular"
                                        % (bad, self.n_filters))

    # ------------------------------------------------------------------ predict / update
    def predict(self, dt=None, UT=None, fx=None, **fx_args):
        """UKF.py:364-411 (deferred and fused with the next ``update``)."""
        _no_hook("UT", UT); _no_hook("fx", fx)
        if fx_args:
            if not isinstance(self.fx, _DeviceModel):
                raise NotImplementedError("fx_args are arguments of a Python callback; the built-in process models take none")
        self._flush()
        if fx_args:
            self._fx_args = self.fx.pack(fx_args, self.n_filters, self._dtype, self._device)
        self._pending = self._dt if dt is None else dt

    def _flush(self):
        if self._pending is not None:
            dt, self._pending = self._pending, None
            self._launch(_lib.BKE_DO_PREDICT, dt, None, None, None)

    def update(self, z, R=None, UT=None, hx=None, valid=None, **hx_args):
        """UKF.py:413-491.  ``z`` is ``(N, dim_z)`` in bank mode; ``z=None`` skips the update.

        Difference from the reference: an ``update`` WITHOUT a preceding ``predict`` draws its sigma
        points from the current (x, P) — what ``predict`` leaves behind (UKF.py:407) — whereas the
        reference would silently reuse the stale ``self.sigmas_f`` of the last predict (zeros on a
        fresh object).  After ``predict(); update(z)`` the two agree."""
        _no_hook("UT", UT); _no_hook("hx", hx)
        if hx_args:
            if not isinstance(self.hx, _DeviceModel):
                raise NotImplementedError("hx_args are arguments of a Python callback; the built-in measurement models take none")
            self._hx_args = self.hx.pack(hx_args, self.n_filters, self._dtype, self._device)
        dt, self._pending = self._pending, None
        if z is None:                                            # UKF.py:442-446
            if dt is not None:
                self._launch(_lib.BKE_DO_PREDICT, dt, None, None, None)
            self._z = None
            if self.diagnostics:
                self._x_post.copy_(self._x); self._P_post.copy_(self._P)
            return
        m = self._dim_z
        zt = to_dev(np.asarray(z, dtype=np.float64).reshape(1, -1) if self._single else z, self._dtype, self._device)
        if tuple(zt.shape) != (self.n_filters, m):
            raise ValueError("z must have shape (%d,%d), got %s" % (self.n_filters, m, tuple(zt.shape)))
        vt = None
        if valid is not None:
            vt = torch.as_tensor(valid, device=self._device).to(torch.uint8).contiguous()
        flags = _lib.BKE_DO_UPDATE | (_lib.BKE_DO_PREDICT if dt is not None else 0)
        self._launch(flags, self._dt if dt is None else dt, zt.contiguous(), vt, R)
        self._z = zt

    def _launch(self, flags, dt, zt, vt, R):
        a = _lib.UkfArgs()
        N, n, m = self.n_filters, self._dim_x, self._dim_z
        a.n_filters, a.dim_x, a.dim_z = N, n, m
        a.dtype = bke_dtype(self._dtype)
        a.flags = flags
        a.fx_model, a.hx_model = self.fx.model, self.hx.model
        a.dt = float(dt)
        a.alpha, a.beta, a.kappa = self.points_fn.alpha, self.points_fn.beta, self.points_fn.kappa
        a.x = a.x_out = ptr(self._x)
        a.P = a.P_out = ptr(self._P)
        a.Q, a.Q_stride = ptr(self._Q), self._stride(self._Q)
        Rm = self._R if R is None else self._model(R, m, m, "R")          # scalar R -> R*I (UKF.py:456-457)
        a.R, a.R_stride = ptr(Rm), self._stride(Rm)
        if self._F is not None:
            a.F, a.F_stride = ptr(self._F), self._stride(self._F)
        if self._H is not None:
            a.H, a.H_stride = ptr(self._H), self._stride(self._H)
        a.z, a.z_valid = ptr(zt), ptr(vt)
        if self.diagnostics:
            if flags & _lib.BKE_DO_PREDICT:
                a.x_prior, a.P_prior = ptr(self._x_prior), ptr(self._P_prior)
            if flags & _lib.BKE_DO_UPDATE:
                a.K, a.y, a.S, a.SI = ptr(self._K), ptr(self._y), ptr(self._S), ptr(self._SI)
                a.log_likelihood = ptr(self._ll)
            a.status = ptr(self._status)
        with torch.cuda.device(self._device):
            if self._user_model is not None:
                for nm, mdl, (t, _) in (("fx", self.fx, self._fx_args), ("hx", self.hx, self._hx_args)):
                    if isinstance(mdl, _DeviceModel) and mdl.arg_names and t is None:
                        raise TypeError("%s needs values for its arguments %s" % (nm, list(mdl.arg_names)))
                _lib.check(self._lib.bke_ukf_step_model(a, self._user_model, ptr(self._fx_args[0]), self._fx_args[1],
                                                        ptr(self._hx_args[0]), self._hx_args[1], stream_ptr(self._device)))
            else:
                _lib.check(self._lib.bke_ukf_step(a, stream_ptr(self._device)))
        if self.diagnostics and (flags & _lib.BKE_DO_UPDATE):
            self._x_post.copy_(self._x); self._P_post.copy_(self._P)
        if self.diagnostics and self._single:
            self.check()

    def rts_smoother(self, Xs, Ps, Qs=None, dts=None, UT=None):
        """UKF.py:634-739 on the GPU.  Bank mode: ``Xs[T,N,n]``, ``Ps[T,N,n,n]`` (what
        ``batch_filter`` returns) -> ``(xs, Ps, Ks)`` tensors; single mode NumPy ``(T,n)`` /
        ``(T,n,n)``.  ``dts``: None (the filter's dt), a scalar, or one value per epoch.  ``Qs`` is
        accepted and, exactly like the reference (:715 uses ``self.Q``), not used."""
        _no_hook("UT", UT)
        if len(Xs) != len(Ps):
            raise ValueError('Xs and Ps must have the same length')
        self._flush()
        N, n = self.n_filters, self._dim_x
        is_np = not isinstance(Xs, torch.Tensor)
        Xt = to_dev(Xs, self._dtype, self._device)
        Pt = to_dev(Ps, self._dtype, self._device)
        T = Xt.shape[0]
        if self._single:
            Xt = Xt.reshape(T, 1, n); Pt = Pt.reshape(T, 1, n, n)
        if tuple(Xt.shape) != (T, N, n) or tuple(Pt.shape) != (T, N, n, n):
            raise ValueError("Xs / Ps must have shapes (T,%d,%d) / (T,%d,%d,%d)" % (N, n, N, n, n))
        Xt = Xt.contiguous(); Pt = Pt.contiguous()
        kw = dict(dtype=self._dtype, device=self._device)
        xs = torch.empty(T, N, n, **kw); Pso = torch.empty(T, N, n, n, **kw); Ks = torch.empty(T, N, n, n, **kw)
        status = torch.zeros(N, dtype=torch.int32, device=self._device)
        a = _lib.UkfRtsArgs()
        a.n_filters, a.n_steps, a.dim_x, a.dtype = N, T, n, bke_dtype(self._dtype)
        a.fx_model = self.fx.model
        a.alpha, a.beta, a.kappa = self.points_fn.alpha, self.points_fn.beta, self.points_fn.kappa
        a.dt = float(self._dt)
        dts_t = None
        if dts is not None:
            if np.isscalar(dts):
                a.dt = float(dts)
            else:
                d = np.asarray(dts, dtype=np.float64).reshape(-1)
                if d.shape[0] != T:
                    raise ValueError("dts must have one entry per epoch (%d)" % T)
                dts_t = torch.from_numpy(np.ascontiguousarray(d)).to(self._device)
                a.dts = ptr(dts_t)
        a.Xs, a.Ps = ptr(Xt), ptr(Pt)
        a.Q, a.Q_stride = ptr(self._Q), self._stride(self._Q)
        if self._F is not None:
            a.F, a.F_stride = ptr(self._F), self._stride(self._F)
        a.x_out, a.P_out, a.K = ptr(xs), ptr(Pso), ptr(Ks)
        a.status = ptr(status)
        with torch.cuda.device(self._device):
            if isinstance(self.fx, _DeviceModel):
                # the reference calls self.fx(sigma, dt) without keyword arguments here (UKF.py:712): the model's
                # current argument values stand in for the defaults of its callable
                if self.fx.arg_names and self._fx_args[0] is None:
                    raise TypeError("fx needs values for its arguments %s" % list(self.fx.arg_names))
                _lib.check(self._lib.bke_ukf_rts_smoother_model(ctypes.byref(a), self._user_model, ptr(self._fx_args[0]),
                                                                self._fx_args[1], stream_ptr(self._device)))
            else:
                _lib.check(self._lib.bke_ukf_rts_smoother(ctypes.byref(a), stream_ptr(self._device)))
        if not self._single:
            return xs, Pso, Ks
        if int(status[0].item()) != 0:
            raise np.linalg.LinAlgError("matrix not positive definite / singular")
        out = (xs[:, 0], Pso[:, 0], Ks[:, 0])
        return tuple(o.cpu().numpy() for o in out) if is_np else out

    def batch_filter(self, zs, Rs=None, dts=None, UT=None, saver=None, valid=None):
        """UKF.py:524-632: predict/update over the epochs of ``zs`` (bank: ``zs[T,N,m]``), one fused
        launch per epoch.  Returns ``(means, covariances)``."""
        _no_hook("UT", UT)
        try:
            z0 = zs[0]
        except TypeError:
            raise TypeError('zs must be list-like')                       # UKF.py:593-596
        m = self._dim_z
        if self._single:
            if m == 1:
                if not (np.isscalar(z0) or (np.ndim(z0) == 1 and len(z0) == 1)):
                    raise TypeError('zs must be a list of scalars or 1D, 1 element arrays')
            elif z0 is not None and len(z0) != m:
                raise TypeError('each element in zs must be a 1D array of length {}'.format(m))
        T = len(zs)
        N, n = self.n_filters, self._dim_x
        kw = dict(dtype=self._dtype, device=self._device)
        means = torch.empty(T, N, n, **kw); covs = torch.empty(T, N, n, n, **kw)
        for i in range(T):
            self.predict(dt=None if dts is None else dts[i])
            z = zs[i]
            v = None if valid is None else valid[i]
            self.update(z, None if Rs is None else Rs[i], valid=v)
            means[i].copy_(self._x); covs[i].copy_(self._P)
            if saver is not None:
                saver.save()
        if not self._single:
            return means, covs
        return means[:, 0].cpu().numpy(), covs[:, 0].cpu().numpy()
