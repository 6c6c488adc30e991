"""Born and Newton iterations for the Kerr problem (reference: fdfdpy/nonlinear_solvers.py).

Each iteration needs a solve with A + Anl(E) (Born) or with the R-linear Jacobian (Newton).  The
reference re-factorises a sparse matrix every time (and a real 2N x 2N one for Newton); here the
factorisation of the LINEAR operator is kept on the GPU and preconditions a short BiCGSTAB run on
the perturbed operator (``Simulation.nl_strategy = 'reuse'``), with an exact re-factorisation as
the fallback / 'refactor' strategy.

By default the iteration itself runs inside the library (``fdfd_nl_solve_host``): the fields never leave the
device between iterations.  ``Simulation.nl_device = False`` selects the host-driven loops below, which call the
same solves one at a time (kept for user-supplied start fields / inspection and as the reference-shaped code path).
"""
from copy import deepcopy

import numpy as np
import numpy.linalg as la

from .constants import EPSILON_0


def _start_field(simulation, Estart):
    if Estart is not None:
        return Estart
    if simulation.fields['Ez'] is None:
        (_, _, Ez) = simulation.solve_fields()
        return Ez
    return deepcopy(simulation.fields['Ez'])


def _kerr_plane(simulation):
    """K(x) with eps_nl = K |E|^2 and d eps_nl / dE = K conj(E): the sum of the Kerr terms (nonlinearity.py:24-31)."""
    K = np.zeros(simulation.eps_r.shape, dtype=np.complex128)
    for nli in simulation.nonlinearity:
        if nli.nl_type != 'kerr':
            return None
        K = K + 3 * nli.chi * nli.nl_region * nli._weight(simulation.eps_r)
    return K


def _device_loop(simulation, Estart, method, conv_threshold, max_num_iter):
    """The whole Born / Newton iteration inside the library (fdfd_nl_solve_host): permittivity update, residual,
    Jacobian diagonal and convergence norm are kernels, one scalar per iteration comes back.  Returns
    ``(Ez, conv_array)`` or None when the device loop does not apply (switched off, non-Kerr term)."""
    import ctypes as C
    from . import _lib
    if not getattr(simulation, 'nl_device', True):
        return None
    K = _kerr_plane(simulation)
    if K is None:
        return None
    Ez = _lib.as_c128(_start_field(simulation, Estart)).copy()
    b = _lib.as_c128(np.asarray(simulation.src) * 1j * simulation.omega)
    lin = simulation._linear_factors()
    op_nl = simulation._nl_operator(np.zeros(simulation.eps_r.shape))
    work = op_nl.direct()
    conv = np.zeros(max_num_iter, dtype=np.float64)
    iters, inner = C.c_int(0), C.c_int(0)
    strategy = {'reuse': 0, 'refactor': 1}[simulation.nl_strategy]
    _lib.check(op_nl.lib.fdfd_nl_solve_host(op_nl.h, lin.h, work.h, _lib.ptr(_lib.as_c128(K)), _lib.ptr(b), _lib.ptr(Ez),
                                            {'born': 0, 'newton': 1}[method], strategy, float(conv_threshold),
                                            int(max_num_iter), _lib.ptr(conv), C.byref(iters), C.byref(inner)))
    work.factored = False                  # the work operator's planes moved on with every iteration
    simulation.last_solve = dict(nl_iterations=iters.value, krylov_iterations=inner.value)
    if conv[max(iters.value - 1, 0)] > conv_threshold:
        print("the simulation did not converge, reached {}".format(conv[max(iters.value - 1, 0)]))
    return Ez.reshape(simulation.Nx, simulation.Ny), conv.reshape(-1, 1)


def born_solve(simulation, Estart=None, conv_threshold=1e-10, max_num_iter=50, averaging=True):
    """Fixed-point iteration E <- (A + Anl(E))^-1 b (nonlinear_solvers.py:14-52)."""
    if simulation.pol != 'Ez':
        raise ValueError('Invalid polarization: {}'.format(str(simulation.pol)))
    dev = _device_loop(simulation, Estart, 'born', conv_threshold, max_num_iter)
    if dev is not None:
        Ez, conv_array = dev
        simulation.compute_nl(Ez)
        Hx, Hy = simulation._ensure_operator().derive_fields(Ez, averaging=averaging)
        return (Hx, Hy, Ez, conv_array)
    conv_array = np.zeros((max_num_iter, 1))
    Ez = _start_field(simulation, Estart)
    convergence = np.inf
    for istep in range(max_num_iter):
        Eprev = Ez
        simulation.compute_nl(Eprev)
        (Hx, Hy, Ez) = simulation.solve_fields(include_nl=True, averaging=averaging)
        convergence = la.norm(Ez - Eprev) / la.norm(Ez)
        conv_array[istep] = convergence
        if convergence < conv_threshold:
            break
    if convergence > conv_threshold:
        print("the simulation did not converge, reached {}".format(convergence))
    return (Hx, Hy, Ez, conv_array)


def nl_eq_and_jac(simulation, averaging=True, Ex=None, Ey=None, Ez=None, compute_jac=True,
                  matrix_format=None):
    """f(E) = (A + Anl(E)) E - i w src and the two Jacobian blocks (nonlinear_solvers.py:113-149).

    Returns ``(fE, Jac11, Jac12)``: Jac11 is the work MaxwellOperator holding
    A + Anl + diag(dAde E) on the device, Jac12 the diagonal conj(dAde) E of the anti-linear part."""
    if simulation.pol != 'Ez':
        raise ValueError('Invalid polarization: {}'.format(str(simulation.pol)))
    omega = simulation.omega
    k = omega ** 2 * EPSILON_0 * simulation.L0
    simulation.compute_nl(Ez)
    Anl = simulation._nl_operator(simulation.eps_nl)
    fE = Anl.dot(Ez).reshape(-1) - np.asarray(simulation.src).reshape(-1) * 1j * omega
    fE = fE.reshape(-1, 1)
    if not compute_jac:
        return fE
    # diag(dAde * E) folds into the operator diagonal as an effective eps_nl
    Jac11 = simulation._nl_operator(simulation.eps_nl + simulation.dnl_de * Ez)
    Jac12 = (np.conj(simulation.dnl_de * k) * Ez).reshape(-1)
    return (fE, Jac11, Jac12)


def newton_solve(simulation, Estart=None, conv_threshold=1e-10, max_num_iter=50, averaging=True,
                 solver=None, jac_solver='c2r', matrix_format=None):
    """Newton's method on f(E) = 0, solving J dE = f each step (nonlinear_solvers.py:55-110)."""
    if simulation.pol != 'Ez':
        raise ValueError('Invalid polarization: {}'.format(str(simulation.pol)))
    dev = _device_loop(simulation, Estart, 'newton', conv_threshold, max_num_iter)
    if dev is not None:
        Ez, conv_array = dev
        simulation.compute_nl(Ez)                     # fields of the converged permittivity, as below
        (Hx, Hy, Ez) = simulation.solve_fields(include_nl=True, averaging=averaging)
        return (Hx, Hy, Ez, conv_array)
    conv_array = np.zeros((max_num_iter, 1))
    Ez = _start_field(simulation, Estart)
    convergence = np.inf
    for istep in range(max_num_iter):
        Eprev = Ez
        (fx, Jac11, Jac12) = nl_eq_and_jac(simulation, Ez=Eprev)
        if fx.any():
            Ediff = simulation._solve_perturbed(Jac11, fx.reshape(-1), c12=Jac12)
        else:
            Ediff = np.zeros(fx.size, dtype=np.complex128)
        Ez = Eprev - Ediff.reshape(simulation.Nx, simulation.Ny)
        convergence = la.norm(Ez - Eprev) / la.norm(Ez)
        conv_array[istep] = convergence
        if convergence < conv_threshold:
            break
    # fields of the converged permittivity
    simulation.compute_nl(Ez)
    (Hx, Hy, Ez) = simulation.solve_fields(include_nl=True, averaging=averaging)
    if convergence > conv_threshold:
        print("the simulation did not converge, reached {}".format(convergence))
    return (Hx, Hy, Ez, conv_array)
