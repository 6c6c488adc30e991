"""Linear algebra entry points with the reference's names and signatures (fdfdpy/linalg.py),
backed by the CUDA library.  Solver names 'pardiso' and 'scipy' are both accepted and both run
the GPU structured direct solver (there is no CPU solver in this package).
"""
from time import time

import numpy as np

from .constants import DEFAULT_MATRIX_FORMAT, DEFAULT_SOLVER
from .core import MaxwellOperator, mode_solve
from .pml import S_create
from .derivatives import createDws

DIRECT_SOLVERS = ('pardiso', 'scipy', 'direct', 'b200')
KRYLOV_SOLVERS = ('bicgstab', 'cocg')


def grid_average(center_array, w):
    """Values at cell edges: mean with the lower neighbour along 'x' or 'y' (linalg.py:14-20).
    Host-side helper for small windows (flux probes); the operator kernels average on the device."""
    xy = {'x': 0, 'y': 1}
    return (np.roll(center_array, 1, axis=xy[w]) + center_array) / 2


def dL(N, xrange, yrange=None):
    """Grid spacing from the domain extents (linalg.py:23-32)."""
    if yrange is None:
        L = np.array([np.diff(xrange)[0]])
    else:
        L = np.array([np.diff(xrange)[0], np.diff(yrange)[0]])
    return L / N


def is_equal(matrix1, matrix2):
    """True if two operators / sparse matrices hold the same entries (linalg.py:35-37)."""
    a = matrix1.to_scipy() if isinstance(matrix1, MaxwellOperator) else matrix1
    b = matrix2.to_scipy() if isinstance(matrix2, MaxwellOperator) else matrix2
    return (a != b).nnz == 0


class _LazyDerivs(dict):
    """The ``derivs`` dictionary of the reference (linalg.py:107-112), materialised on first access:
    scipy matrices S^-1 D built from the device PML factors (format export, not used by the solver)."""

    def __init__(self, op):
        super().__init__()
        self._op = op

    def _build(self):
        import scipy.sparse as sp
        op = self._op
        isxf, isxb, isyf, isyb = op.sfactors()
        nx, ny = op.nx, op.ny
        M = nx * ny
        d = [(nx * op.dl) / nx, (ny * op.dl) / ny]
        ox, oy = np.ones((nx, 1)), np.ones((1, ny))

        def diag(v):
            return sp.spdiags(v.reshape(-1), 0, M, M, format='csr')
        dict.__setitem__(self, 'Dyb', diag(ox * isyb[None, :]).dot(createDws('y', 'b', d, [nx, ny])))
        dict.__setitem__(self, 'Dxb', diag(isxb[:, None] * oy).dot(createDws('x', 'b', d, [nx, ny])))
        dict.__setitem__(self, 'Dxf', diag(isxf[:, None] * oy).dot(createDws('x', 'f', d, [nx, ny])))
        dict.__setitem__(self, 'Dyf', diag(ox * isyf[None, :]).dot(createDws('y', 'f', d, [nx, ny])))

    def __getitem__(self, k):
        if not dict.__len__(self):
            self._build()
        return dict.__getitem__(self, k)

    def keys(self):
        if not dict.__len__(self):
            self._build()
        return dict.keys(self)


def construct_A(omega, xrange, yrange, eps_r, NPML, pol, L0, averaging=True, timing=False,
                matrix_format=DEFAULT_MATRIX_FORMAT):
    """Build the Maxwell operator on the device (linalg.py:39-114).  Returns ``(A, derivs)`` where A
    is a ``MaxwellOperator`` (``A.dot``, ``A.shape``, ``A.to_scipy()``) and derivs is a lazily
    materialised dictionary of the four PML-scaled derivative matrices."""
    eps_r = np.asarray(eps_r)
    if pol not in ('Ez', 'Hz'):
        raise ValueError("something went wrong and pol is not one of Ez, Hz, instead was given {}".format(pol))
    t = time()
    dl = float(np.diff(xrange)[0]) / eps_r.shape[0]
    A = MaxwellOperator(omega, eps_r, dl, NPML, pol, L0, averaging=averaging)
    if timing:
        print('Operator assembly took {:.4f} seconds'.format(time() - t))
    return (A, _LazyDerivs(A))


def solver_eigs(A, Neigs, guess_value=0, guess_vector=None, timing=False):
    """Eigenpairs nearest ``guess_value`` (linalg.py:104-115).  ``A`` must be a ``ModeOperator``
    describing the 1-D waveguide line; the shift-invert solve runs in the CUDA mode kernel."""
    from .source.mode import ModeOperator
    if not isinstance(A, ModeOperator):
        raise TypeError("solver_eigs on the B200 path takes a fdfdpy_b200.source.mode.ModeOperator")
    t = time()
    vals, vecs = A.eigs(Neigs, guess_value)
    if timing:
        print('Elapsed time for eigs() is %.4f secs' % (time() - t))
    return (vals, vecs)


def solver_direct(A, b, timing=False, solver=DEFAULT_SOLVER):
    """Solve A x = b with the GPU direct solver (linalg.py:123-149).  ``A`` is a MaxwellOperator;
    its factorisation is cached on the operator and reused while its planes do not change."""
    b = np.asarray(b).astype(np.complex128).reshape((-1,))
    if not b.any():
        return np.zeros(b.shape)
    if not isinstance(A, MaxwellOperator):
        raise TypeError("solver_direct on the B200 path takes a MaxwellOperator (see construct_A)")
    t = time()
    s = solver.lower()
    if s in DIRECT_SOLVERS:
        x = A.solve(b)
    elif s in KRYLOV_SOLVERS:
        x, info = A.krylov(b, method=s, tol=1e-12, maxiter=200000)
        if not info['converged']:
            raise RuntimeError("{} did not converge: {}".format(s, info))
    else:
        raise ValueError('Invalid solver choice: {}, options are pardiso or scipy'.format(str(solver)))
    if timing:
        print('Linear system solve took {:.2f} seconds'.format(time() - t))
    return x.reshape(-1)


def solver_complex2real(A11, A12, b, timing=False, solver=DEFAULT_SOLVER):
    """Solve A11 x + A12 conj(x) = b (linalg.py:152-186).

    The reference expands this into a real 2N x 2N sparse LU.  Here A11 is a MaxwellOperator and
    A12 the diagonal (1-D/2-D array) of the anti-linear term; the R-linear system is solved by
    BiCGSTAB in the real inner product, right-preconditioned with A11's cached factorisation."""
    b = np.asarray(b).astype(np.complex128).reshape((-1,))
    if not b.any():
        return np.zeros(b.shape)
    if not isinstance(A11, MaxwellOperator):
        raise TypeError("solver_complex2real on the B200 path takes a MaxwellOperator for A11")
    t = time()
    c12 = np.asarray(A12.diagonal() if hasattr(A12, 'diagonal') and getattr(A12, 'ndim', 1) == 2
                     and A12.shape == A11.shape else A12).reshape(-1)
    x, info = A11.krylov(b, method='bicgstab', tol=1e-13, maxiter=200, check_every=1, precondition=True,
                         c12=c12, fused=False)
    if info['relres'] > 1e-9:
        raise RuntimeError("Jacobian solve did not converge: {}".format(info))
    if timing:
        print('Linear system solve took {:.2f} seconds'.format(time() - t))
    return x.reshape(-1)
