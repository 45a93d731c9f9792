"""Host-side shape rules shared by the mirrors."""
import numpy as np

__all__ = ["reshape_z"]


def reshape_z(z, dim_z, ndim):
    """Bring a measurement to the layout the filter state uses.

    Same contract as ``filterpy.common.reshape_z`` (filterpy/common/helpers.py:324-342): the
    result is ``(dim_z, 1)`` for a column-vector state (``ndim == 2``), ``(dim_z,)`` for a 1-D
    state, a scalar for a 0-d state; ``ValueError`` when ``z`` cannot be viewed as a
    ``dim_z``-vector."""
    arr = np.atleast_2d(z)
    if arr.shape[1] == dim_z:       # a row (or dim_z x dim_z... like the reference, transpose first)
        arr = arr.T
    if arr.shape != (dim_z, 1):
        raise ValueError("z (shape {}) must be convertible to shape ({}, 1)".format(arr.shape, dim_z))
    if ndim == 1:
        return arr[:, 0]
    if ndim == 0:
        return arr[0, 0]
    return arr
