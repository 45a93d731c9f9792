"""Oracle: particle-filter resampling (TEST INFRASTRUCTURE).

Restates ``filterpy/monte_carlo/resampling.py`` (reference @ 3b51149):

* ``systematic_resample``  resampling.py:117-150
* ``stratified_resample``  resampling.py:80-114
* ``multinomial_resample`` resampling.py:153-176
* ``residual_resample``    resampling.py:27-76 (+ NumPy's ``npy_binsearch``, restated below)

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


def binsearch_left(arr, keys):
    """``np.searchsorted(arr, keys)`` (side='left') as NumPy computes it, for ANY ``arr`` (sorted or
    not): numpy/_core/src/npysort/binsearch.cpp ``binsearch<Tag, side>`` (NumPy 2.3) keeps the bracket of
    the previous key — ``[min_idx, arr_len)`` if ``last_key < key`` else ``[0, max_idx + 1)`` — and
    orders doubles with NaN last.  On a sorted array the bracket is irrelevant; on residual_resample's
    non-monotone cumulative sum (resampling.py:69-74) it decides the answer."""
    def lt(a, b):
        return a < b or (b != b and a == a)
    n = len(arr)
    out = np.empty(len(keys), np.int64)
    if len(keys) == 0:
        return out
    lo, hi = 0, n
    last = keys[0]
    for i, key in enumerate(keys):
        if lt(last, key):
            hi = n
        else:
            lo = 0
            hi = hi + 1 if hi < n else n
        last = key
        while lo < hi:
            mid = lo + ((hi - lo) >> 1)
            if lt(arr[mid], key):
                lo = mid + 1
            else:
                hi = mid
        out[i] = lo
    return out


def residual_prepare(weights):
    """resampling.py:52-72 without the uniforms: (indexes with the first k entries filled, k,
    cumulative_sum, sum(residual)); every sum in the reference's order (builtin ``sum`` :70 and
    ``np.cumsum`` :71 are both one fp64 add at a time)."""
    weights = np.asarray(weights, dtype=np.float64)
    N = len(weights)
    indexes = np.zeros(N, 'i')
    num_copies = (np.floor(N * weights)).astype(int)
    k = int(np.maximum(num_copies, 0).sum())
    if k > N:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (N, N))
    indexes[:k] = np.repeat(np.arange(N), np.maximum(num_copies, 0))
    residual = weights - num_copies
    s = 0.0
    for r in residual:                       # builtin sum(): left to right
        s = s + r
    with np.errstate(all="ignore"):
        residual = residual / s
    cumulative_sum = np.cumsum(residual)
    cumulative_sum[-1] = 1.
    return indexes, k, cumulative_sum, s


def residual_resample_vec(weights, U):
    """resampling.py:27-76 with the uniforms ``random(N - k)`` passed in; the bisection by the
    restatement above (``binsearch_left``), not by NumPy."""
    indexes, k, cumulative_sum, _ = residual_prepare(weights)
    N = len(indexes)
    U = np.atleast_1d(U)
    assert len(U) == N - k
    indexes[k:N] = binsearch_left(cumulative_sum, U)
    return indexes


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
