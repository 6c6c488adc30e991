"""sc-PML stretch factors, computed on the device (reference: fdfdpy/pml.py).

``S_create`` keeps the reference signature and return type (four scipy diagonal matrices of the
INVERSE stretch factors, pml.py:44-89); the numbers come from the CUDA kernel ``pml_axis_kernel``.
"""
import numpy as np
import scipy.sparse as sp

from .constants import DEFAULT_MATRIX_FORMAT


def inverse_sfactors(omega, L0, N, Npml, dl):
    """(1/sx_f, 1/sx_b, 1/sy_f, 1/sy_b) 1-D arrays from the device."""
    from .core import MaxwellOperator
    op = MaxwellOperator(omega, np.ones((int(N[0]), int(N[1]))), dl, [int(Npml[0]), int(Npml[1])], 'Ez', L0)
    return op.sfactors()


def S_create(omega, L0, N, Npml, xrange, yrange=None, matrix_format=DEFAULT_MATRIX_FORMAT):
    N = np.atleast_1d(np.asarray(N))
    Npml = np.atleast_1d(np.asarray(Npml))
    if len(N) < 2:
        raise ValueError("the B200 path is 2-D: N must have two entries")
    nx, ny = int(N[0]), int(N[1])
    dl = float(np.diff(xrange)[0]) / nx
    isxf, isxb, isyf, isyb = inverse_sfactors(omega, L0, (nx, ny), Npml, dl)
    M = nx * ny

    def diag(v2d):
        return sp.spdiags(v2d.reshape(-1), 0, M, M, format=matrix_format)

    ones_x, ones_y = np.ones((nx, 1)), np.ones((1, ny))
    return (diag(isxf[:, None] * ones_y), diag(isxb[:, None] * ones_y),
            diag(ones_x * isyf[None, :]), diag(ones_x * isyb[None, :]))
