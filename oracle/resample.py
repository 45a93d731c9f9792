"""Oracle: particle-filter resampling (TEST INFRASTRUCTURE).

Restates ``filterpy/monte_carlo/resampling.py`` (reference @ 3b51149):

* ``systematic_resample``  resampling.py:117-150
* ``stratified_resample``  resampling.py:80-114
* ``multinomial_resample`` resampling.py:153-176

The reference draws its uniforms from the process-global legacy RandomState
(resampling.py:24,103,139); here the uniform(s) are explicit arguments so that a
test can feed the same value to the reference, the oracle and the CUDA path.

Three equivalent forms:

``*_loop``     the literal two-pointer merge of the reference (pure Python; small N);
``*_vec``      ``searchsorted(cumsum(w), positions, side='right')`` — identical for
               non-negative weights because ``np.cumsum`` is a strictly sequential
               fp64 accumulation (resampling.py:142) and is then monotone;
``*_c``        ``oracle.c`` (sequential cumsum + merge in C; full-size arrays).

The reference raises IndexError when positions[-1] >= cumsum[-1]
(resampling.py:145, j runs off the end); the oracle raises the same.
Parity: the reference has NO tests for this module ("parity unpinned" by its own
tests); it is pinned here against the reference function itself, see
``tests/golden/resample_*.npz``.
"""
import numpy as np


def positions_systematic(N, u):
    """resampling.py:139 — (u + arange(N)) / N in fp64, exactly this op order."""
    return (u + np.arange(N)) / N


def positions_stratified(N, U):
    """resampling.py:103 — (U[N] + range(N)) / N."""
    return (np.asarray(U, float) + np.arange(N)) / N


def resample_loop(weights, positions):
    """resampling.py:141-149 (also :105-113), literal."""
    N = len(weights)
    indexes = np.zeros(N, 'i')
    cumulative_sum = np.cumsum(weights)
    i, j = 0, 0
    while i < N:
        if positions[i] < cumulative_sum[j]:   # IndexError when j == N, as in the reference
            indexes[i] = j
            i += 1
        else:
            j += 1
    return indexes


def resample_vec(weights, positions):
    w = np.asarray(weights)
    if np.any(w < 0) or not np.all(np.isfinite(w)):
        return resample_loop(w, positions)
    c = np.cumsum(w)
    idx = np.searchsorted(c, positions, side='right')
    if len(idx) and idx.max() >= len(w):
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(w), len(w)))
    return idx.astype('i')


def systematic_resample_loop(weights, u):
    return resample_loop(weights, positions_systematic(len(weights), u))


def stratified_resample_loop(weights, U):
    return resample_loop(weights, positions_stratified(len(weights), U))


def systematic_resample_vec(weights, u):
    return resample_vec(weights, positions_systematic(len(weights), u))


def stratified_resample_vec(weights, U):
    return resample_vec(weights, positions_stratified(len(weights), U))


def multinomial_resample_vec(weights, U):
    """resampling.py:173-176 with the uniforms ``random(len(weights))`` passed in."""
    cumulative_sum = np.cumsum(weights)
    cumulative_sum[-1] = 1.
    return np.searchsorted(cumulative_sum, U)


def multinomial_resample_loop(weights, U):
    """The same with the cumulative sum and the left bisection spelled out (pure Python; small N)."""
    N = len(weights)
    c = [0.0] * N
    acc = None
    for j in range(N):
        acc = float(weights[j]) if j == 0 else acc + float(weights[j])
        c[j] = acc
    c[-1] = 1.
    out = np.zeros(len(U), np.int64)
    for q, key in enumerate(U):
        lo, hi = 0, N
        while lo < hi:
            mid = (lo + hi) // 2
            if c[mid] < key:
                lo = mid + 1
            else:
                hi = mid
        out[q] = lo
    return out


# --------------------------------------------------------------------------- C port
def _clib():
    from . import cbuild
    return cbuild.load()


def systematic_resample_c(weights, u):
    """oracle.c:oracle_systematic_resample — sequential cumsum + merge, one thread."""
    import ctypes
    lib = _clib()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    N = w.shape[0]
    idx = np.empty(N, dtype=np.int32)
    rc = lib.oracle_systematic_resample(
        w.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(N), ctypes.c_double(u),
        idx.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (N, N))
    return idx


def stratified_resample_c(weights, U):
    import ctypes
    lib = _clib()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    N = w.shape[0]
    idx = np.empty(N, dtype=np.int32)
    rc = lib.oracle_stratified_resample(
        w.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(N),
        U.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (N, N))
    return idx


def multinomial_resample_c(weights, U):
    """oracle.c:oracle_multinomial_resample — sequential cumsum, last := 1, left bisection per key."""
    import ctypes
    lib = _clib()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    N = w.shape[0]
    idx = np.empty(U.shape[0], dtype=np.int64)
    scratch = np.empty(N, dtype=np.float64)
    lib.oracle_multinomial_resample(
        w.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(N), U.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int64(U.shape[0]), scratch.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p))
    return idx
