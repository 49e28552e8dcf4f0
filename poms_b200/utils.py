"""Drop-in for the fixture helpers of /root/reference/sources/utils.py (device containers)."""
import numpy as np

from .stencil import StencilVector, StencilMatrix, StencilVectorSpace
from . import bsplines as bs

__all__ = ["populate_1d_matrix", "populate_2d_matrix", "populate_2d_vector",
           "array_to_vect_stencil", "array_to_mat_stencil"]


def populate_1d_matrix(M, diag):
    """M[i, k] = k, M[i, 0] = diag (/root/reference/sources/utils.py:7-16)."""
    p = M.pads[0]
    for k in range(-p, p + 1):
        M[:, k] = k
    M[:, 0] = diag
    M.remove_spurious_entries()


def populate_2d_matrix(M, diag):
    """M[:, :, k1, k2] = 10 - |k1| - |k2| (/root/reference/sources/utils.py:20-27)."""
    p1, p2 = M.pads
    for k1 in range(-p1, p1 + 1):
        for k2 in range(-p2, p2 + 1):
            M[:, :, k1, k2] = 10.0 - abs(k1) - abs(k2)
    M.remove_spurious_entries()


def populate_2d_vector(X):
    """X = 1 on the owned entries (/root/reference/sources/utils.py:31-39)."""
    X.data.fill_(1.0)


def array_to_vect_stencil(v_space, v_arr):
    """Global array -> stencil vector (/root/reference/sources/utils.py:92-101)."""
    return StencilVector.from_array(v_space, v_arr)


def array_to_mat_stencil(n, p, v_arr):
    """Half-bandwidth-p band of a dense (n, n) array as a 1-D StencilMatrix
    (/root/reference/sources/utils.py:105-134)."""
    V = StencilVectorSpace([n], [p], [False])
    M = StencilMatrix(V, V)
    M._data[...] = bs.dense_to_band(np.asarray(v_arr, dtype=float), p)
    return M
