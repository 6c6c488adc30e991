"""Derivative operators (reference: fdfdpy/derivatives.py).

On the hot path the derivatives are never materialised: the stencil kernels apply them matrix
free.  ``createDws`` is kept, with the reference signature, for callers that want the unscaled
periodic difference matrices as scipy objects (a format export, built from index arithmetic).
"""
import numpy as np
import scipy.sparse as sp

from .constants import DEFAULT_MATRIX_FORMAT


def createDws(w, s, dL, N, matrix_format=DEFAULT_MATRIX_FORMAT):
    """Periodic forward ('f') / backward ('b') difference along 'x' or 'y' (derivatives.py:7-34)."""
    if w not in ('x', 'y') or s not in ('f', 'b'):
        raise ValueError("w must be 'x' or 'y' and s must be 'f' or 'b'")
    nx = int(N[0])
    ny = int(N[1]) if len(N) > 1 else 1
    d = float(dL[0]) if w == 'x' else (float(dL[1]) if len(N) > 1 else np.inf)
    n = nx if w == 'x' else ny
    idx = np.arange(n)
    nb = (idx + 1) % n if s == 'f' else (idx - 1) % n
    sign = 1.0 if s == 'f' else -1.0
    # row i: sign * (u[neighbour] - u[i]); duplicates (n == 1 or 2) are summed like scipy.diags would
    one = sp.coo_matrix((np.concatenate([np.full(n, -sign), np.full(n, sign)]),
                         (np.concatenate([idx, idx]), np.concatenate([idx, nb]))), shape=(n, n))
    if w == 'x':
        return (1 / d) * sp.kron(one, sp.eye(ny), format=matrix_format)
    return (1 / d) * sp.kron(sp.eye(nx), one, format=matrix_format)


def unpack_derivs(derivs):
    """(Dyb, Dxb, Dxf, Dyf) from the derivs dictionary (derivatives.py:37-44)."""
    return (derivs['Dyb'], derivs['Dxb'], derivs['Dxf'], derivs['Dyf'])
