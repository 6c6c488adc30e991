"""The reference's public class, ``Simulation`` (fdfdpy/simulation.py), on the B200 path.

Same constructor, attributes and methods; the operator lives on the GPU, ``solve_fields`` runs
the structured direct solver (factorisation cached across right-hand sides and re-used as a
preconditioner for the nonlinear iterations) and hands numpy arrays back.
"""
import numbers
from copy import deepcopy
from time import time

import numpy as np

from .constants import (DEFAULT_LENGTH_SCALE, DEFAULT_MATRIX_FORMAT, DEFAULT_SOLVER, EPSILON_0, MU_0)
from .core import MaxwellOperator
from .geometry import grow, plane_slices
from .linalg import DIRECT_SOLVERS, KRYLOV_SOLVERS, _LazyDerivs, grid_average
from .nonlinearity import Nonlinearity
from .source.mode import mode

FIELD_NAMES = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')


class FdfdInputError(ValueError, AssertionError, TypeError):
    """Bad constructor argument.  The reference's tests expect ValueError (tests/test_simulation.py)
    while its code raises AssertionError / TypeError for some of them (simulation.py:258-265, :23);
    this class is all three."""


def _blank_fields():
    return {f: None for f in FIELD_NAMES}


class Simulation:

    def __init__(self, omega, eps_r, dl, NPML, pol, L0=DEFAULT_LENGTH_SCALE):
        self._check_inputs(omega, eps_r, dl, NPML, pol, L0)
        self.L0 = L0
        self.omega = float(omega)
        self.dl = float(dl)
        self.NPML = [int(n) for n in NPML]
        self.pol = pol
        (self.Nx, self.Ny) = eps_r.shape
        self.mu_r = np.ones((self.Nx, self.Ny))
        self.src = np.zeros((self.Nx, self.Ny))
        self.xrange = [0, float(self.Nx * self.dl)]
        self.yrange = [0, float(self.Ny * self.dl)]
        self.timings = {}
        self.last_solve = None  # residual / iteration record of the most recent linear solve
        self._op = None        # linear operator A(eps_r) and its cached factorisation
        self._derivs = None
        self._op_nl = None     # work operator for A + Anl and the Newton Jacobian
        self.nl_device = True        # Born / Newton loops run inside the library (False: host-driven loops)
        self.nl_strategy = 'reuse'   # 'reuse': linear factors precondition the nonlinear solves;
        #                              'refactor': factorise A + Anl every time, as the reference does
        self.eps_r = eps_r     # builds the system operator (simulation.py:38)
        self.modes = []
        self.nonlinearity = []
        self.eps_nl = np.zeros(eps_r.shape)
        self.dnl_de = np.zeros(eps_r.shape)
        self.dnl_deps = np.zeros(eps_r.shape)

    # ------------------------------------------------------------------ inputs
    @staticmethod
    def _check_inputs(omega, eps_r, dl, NPML, pol, L0):
        """Argument checks (simulation.py:256-265, tightened to what tests/test_simulation.py asks)."""
        def real_scalar(v):
            return isinstance(v, numbers.Real) and not isinstance(v, bool) or \
                (isinstance(v, np.ndarray) and v.ndim == 0 and np.isrealobj(v))
        if not real_scalar(omega) or not float(omega) > 0:
            raise FdfdInputError("omega must be a single positive number, was supplied {}".format(omega))
        if not real_scalar(dl) or not float(dl) > 0:
            raise FdfdInputError("dl must be a single positive number, was supplied {}".format(dl))
        if not real_scalar(L0) or not float(L0) > 0:
            raise FdfdInputError("L0 must be a positive number, was supplied {},".format(str(L0)))
        Simulation._check_eps(eps_r)
        try:
            n_npml = len(NPML)
        except TypeError:
            raise FdfdInputError("NPML must be a list of length 2, was supplied {}".format(NPML))
        if n_npml != 2:
            raise FdfdInputError("yrange must be a list of length 2, was supplied {}, which is of length {}"
                                 .format(str(NPML), n_npml))
        if not (int(NPML[0]) >= 0 and int(NPML[1]) >= 0):
            raise FdfdInputError("both elements of NPML must be >= 0")
        if int(NPML[0]) >= eps_r.shape[0] or int(NPML[1]) >= eps_r.shape[1]:
            raise FdfdInputError("NPML {} does not fit in a {} grid".format(list(NPML), eps_r.shape))
        if not isinstance(pol, str) or pol not in ('Ez', 'Hz'):
            raise FdfdInputError("pol must be one of 'Ez' or 'Hz'")

    @staticmethod
    def _check_eps(eps_r):
        """Shared by the constructor and the eps_r setter."""
        if not isinstance(eps_r, np.ndarray) or eps_r.ndim != 2:
            raise FdfdInputError("eps_r must be a 2-D numpy array")
        # large arrays are checked on the DEVICE right after the upload (eps_r setter): the host scan costs 17 ms at
        # 4096^2, the flag kernel 0.05 ms
        if eps_r.size <= Simulation._HOST_SCAN_MAX and np.any(np.real(eps_r) < 0):
            raise FdfdInputError("eps_r must not be negative")
        if min(eps_r.shape) < 4:
            # the reference accepts degenerate grids (e.g. Ny = 1); the structured solver's elimination tree
            # needs 4 cells per axis, and this is the place to say so (not the first solve_fields)
            raise FdfdInputError("the B200 solver needs at least 4 cells per axis, got a {} grid".format(eps_r.shape))

    _DEVICE_STATE = ('_op', '_op_nl', '_derivs')
    _HOST_SCAN_MAX = 1 << 20     # arrays up to this many cells are validated / scanned with numpy on the host

    def __deepcopy__(self, memo):
        """Everything the reference's ``deepcopy(simulation)`` preserves (fields, sources, modes, nonlinearity,
        normalisation) is copied; device handles are not: the twin owns no operator until its ``eps_r`` is
        assigned or its first solve builds one from the copied permittivity (no redundant assembly when the
        caller replaces eps_r right away, as ``mode.compute_normalization`` does)."""
        twin = Simulation.__new__(Simulation)
        memo[id(self)] = twin
        for key, val in self.__dict__.items():
            if key in self._DEVICE_STATE:
                continue
            twin.__dict__[key] = deepcopy(val, memo)
        twin._op = twin._op_nl = twin._derivs = None
        return twin

    def _ensure_operator(self):
        if self._op is None:
            self._op = MaxwellOperator(self.omega, self.__eps_r, self.dl, self.NPML, self.pol, self.L0)
            self._op_nl = None
            self._derivs = _LazyDerivs(self._op)
        return self._op

    @property
    def A(self):
        """The system operator (simulation.py:38 keeps the scipy matrix here; this is the device-resident one)."""
        return self._ensure_operator()

    @property
    def derivs(self):
        self._ensure_operator()
        return self._derivs

    # ------------------------------------------------------------------ operator
    @property
    def eps_r(self):
        return self.__eps_r

    @eps_r.setter
    def eps_r(self, new_eps):
        """Reassigning eps_r re-assembles A on the device and drops the cached factorisation
        (simulation.py:80-89)."""
        new_eps = np.asarray(new_eps)
        self._check_eps(new_eps)
        if (int(self.NPML[0]) >= new_eps.shape[0] or int(self.NPML[1]) >= new_eps.shape[1]):
            raise FdfdInputError("NPML {} does not fit in a {} grid".format(list(self.NPML), new_eps.shape))
        old_eps = getattr(self, '_Simulation__eps_r', None)
        self.__eps_r = new_eps
        (self.Nx, self.Ny) = new_eps.shape
        t = time()
        if self._op is not None and (self._op.nx, self._op.ny) == new_eps.shape:
            self._op.assemble(new_eps)
            if self._op._direct is not None:
                self._op._direct.has_factors = False
        else:
            self._op = MaxwellOperator(self.omega, new_eps, self.dl, self.NPML, self.pol, self.L0)
            self._op_nl = None
        if new_eps.size > self._HOST_SCAN_MAX and (self._op.eps_flags() & 2):
            self._op = None                 # nothing usable was assembled: the next use rebuilds from the kept eps_r
            if old_eps is not None:
                self.__eps_r = old_eps
                (self.Nx, self.Ny) = old_eps.shape
            raise FdfdInputError("eps_r must not be negative")
        self.timings['assemble'] = time() - t
        self._derivs = _LazyDerivs(self._op)
        self.fields = _blank_fields()
        self.fields_nl = _blank_fields()

    def reset_eps(self, new_eps):
        # kept for compatibility (simulation.py:91-102)
        self.eps_r = new_eps

    # ------------------------------------------------------------------ sources
    def setup_modes(self):
        for modei in self.modes:
            modei.setup_src(self)

    def add_mode(self, neff, direction_normal, center, width, scale=1, order=1):
        self.modes.append(mode(neff, direction_normal, center, width, scale=scale, order=order))

    # ------------------------------------------------------------------ nonlinearity
    def add_nl(self, chi, nl_region, nl_type='kerr', eps_scale=False, eps_max=None):
        # chi is given in SI and stored in units of L0 (simulation.py:72-75)
        self.nonlinearity.append(Nonlinearity(chi / np.square(self.L0), nl_region, nl_type, eps_scale, eps_max))

    def compute_nl(self, e, matrix_format=DEFAULT_MATRIX_FORMAT):
        """Evaluate eps_nl, d eps_nl/de and d eps_nl/d eps for the field ``e`` (simulation.py:58-70)."""
        self.eps_nl = np.zeros(self.eps_r.shape)
        self.dnl_de = np.zeros(self.eps_r.shape)
        self.dnl_deps = np.zeros(self.eps_r.shape)
        for nli in self.nonlinearity:
            self.eps_nl = self.eps_nl + nli.eps_nl(e, self.eps_r)
            self.dnl_de = self.dnl_de + nli.dnl_de(e, self.eps_r)
            self.dnl_deps = self.dnl_deps + nli.dnl_deps(e, self.eps_r)

    @property
    def Anl(self):
        """Diagonal matrix w^2 eps0 L0 eps_nl, as a scipy object for inspection (simulation.py:68-70)."""
        import scipy.sparse as sp
        n = self.Nx * self.Ny
        return sp.spdiags(self.omega ** 2 * EPSILON_0 * self.L0 * np.asarray(self.eps_nl).reshape(-1), 0, n, n,
                          format=DEFAULT_MATRIX_FORMAT)

    def _nl_operator(self, eps_nl_eff):
        """Work operator A + w^2 eps0' diag(eps_nl_eff); shares the grid with the linear one."""
        if self._op_nl is None:
            self._op_nl = MaxwellOperator(self.omega, self.eps_r, self.dl, self.NPML, self.pol, self.L0,
                                          eps_nl=eps_nl_eff)
        else:
            self._op_nl.assemble(self.eps_r, eps_nl_eff)
        return self._op_nl

    def _linear_factors(self):
        d = self._ensure_operator().direct()
        if not d.factored:
            t = time()
            d.factor()
            self.timings['factor'] = time() - t
        return d

    def compute_index_shift(self):
        """Array of the nonlinear refractive-index shift (simulation.py:104-111)."""
        _ = self.solve_fields()
        _ = self.solve_fields_nl()
        index_nl = np.sqrt(np.real(self.eps_r + self.eps_nl))
        index_lin = np.sqrt(np.real(self.eps_r))
        return np.abs(index_nl - index_lin)

    # ------------------------------------------------------------------ linear solve
    def solve_fields(self, include_nl=False, timing=False, averaging=True, solver=DEFAULT_SOLVER,
                     matrix_format=DEFAULT_MATRIX_FORMAT):
        """Solve A x = i w src on the GPU and derive the in-plane fields (simulation.py:113-178)."""
        t0 = time()
        if self.pol not in ('Ez', 'Hz'):
            raise ValueError('Invalid polarization: {}'.format(str(self.pol)))
        s = solver.lower()
        if s not in DIRECT_SOLVERS + KRYLOV_SOLVERS:
            raise ValueError('Invalid solver choice: {}, options are pardiso or scipy'.format(str(solver)))
        src = np.asarray(self.src)
        op = self._ensure_operator() if not include_nl else self._nl_operator(self.eps_nl)
        if not include_nl and s in DIRECT_SOLVERS and (src.size > self._HOST_SCAN_MAX or src.any()):
            # the hot path: ONE library call (factorisation included when eps_r changed), b = i w src formed on the
            # device; a large src is not scanned for the all-zero case here, the library returns zero fields for it
            d = op.direct()
            X, f1, f2 = d.solve_fields(src, 1j * self.omega, averaging=averaging)
            if d.last_factor_ms is not None:
                self.timings['factor'] = d.last_factor_ms * 1e-3
            self.last_solve = dict(relres=d.last_relres, refine_steps=d.last_refine_steps)
        else:
            b = src * 1j * self.omega
            if not b.any():
                X = np.zeros(b.shape, dtype=np.complex128)      # linalg.py:129-130
            elif s in KRYLOV_SOLVERS:
                X, info = op.krylov(b, method=s, tol=1e-12, maxiter=500000)
                self.last_solve = info
                if not info['converged']:
                    raise RuntimeError("{} did not converge: {}".format(s, info))
            else:
                X = self._solve_perturbed(op, b)
            X = X.reshape(self.Nx, self.Ny)
            f1, f2 = op.derive_fields(X, averaging=averaging)
        names = ('Hx', 'Hy', 'Ez') if self.pol == 'Ez' else ('Ex', 'Ey', 'Hz')
        if not include_nl:
            for k, v in zip(names, (f1, f2, X)):
                self.fields[k] = v
        self.timings['solve_fields'] = time() - t0
        if timing:
            print('Linear system solve took {:.2f} seconds'.format(time() - t0))
        return (f1, f2, X)

    def _solve_perturbed(self, op, b, c12=None, x0=None):
        """Solve with the work operator ``op`` = A + diagonal perturbation (+ anti-linear c12 term).

        'reuse': BiCGSTAB on op, right-preconditioned by the LINEAR operator's cached factorisation;
        falls back to an exact factorisation of op when that stalls.  'refactor': always exact."""
        if self.nl_strategy == 'reuse':
            op.preconditioner = self._linear_factors()
            X, info = op.krylov(b, x0=x0, method='bicgstab', tol=1e-13, maxiter=40, check_every=1,
                                precondition=True, c12=c12, fused=False)
            self.last_solve = info
            if info['relres'] <= 1e-11:
                return X
        op.preconditioner = None
        d = op.direct()
        d.factor()
        if c12 is None:
            X = d.solve(b)
            self.last_solve = dict(relres=d.last_relres, refine_steps=d.last_refine_steps)
            return X
        X, info = op.krylov(b, x0=x0, method='bicgstab', tol=1e-13, maxiter=200, check_every=1, precondition=True,
                            c12=c12, fused=False)
        self.last_solve = info
        if info['relres'] > 1e-9:
            raise RuntimeError("Jacobian solve did not converge: {}".format(info))
        return X

    # ------------------------------------------------------------------ nonlinear solve
    def solve_fields_nl(self, timing=False, averaging=True, Estart=None, solver_nl='newton',
                        conv_threshold=1e-10, max_num_iter=50, solver=DEFAULT_SOLVER,
                        matrix_format=DEFAULT_MATRIX_FORMAT):
        """Nonlinear (Kerr) solve by Born or Newton iteration (simulation.py:180-254)."""
        from .nonlinear_solvers import born_solve, newton_solve
        if self.pol not in ('Ez', 'Hz'):
            raise ValueError('Invalid polarization: {}'.format(str(self.pol)))
        allowed = {'born': born_solve, 'newton': newton_solve}
        if solver_nl not in allowed:
            raise AssertionError("solver must be one of {'born', 'newton'}")
        (f1, f2, fz, conv_array) = allowed[solver_nl](self, Estart, conv_threshold, max_num_iter,
                                                      averaging=averaging)
        names = ('Hx', 'Hy', 'Ez') if self.pol == 'Ez' else ('Ex', 'Ey', 'Hz')
        for k, v in zip(names, (f1, f2, fz)):
            self.fields_nl[k] = v
        return (f1, f2, fz, conv_array)

    # ------------------------------------------------------------------ probes
    def flux_probe(self, direction_normal, center, width, nl=False):
        """Poynting flux through a line (simulation.py:267-327).  O(width) host arithmetic on fields
        that are already on the host."""
        sx, sy = plane_slices(direction_normal, center, width)
        src = self.fields_nl if nl else self.fields
        names = ('Hx', 'Hy', 'Ez') if self.pol == 'Ez' else ('Ex', 'Ey', 'Hz')
        f1, f2, fz = (src[k] for k in names)
        window = fz[grow(sx), grow(sy)]
        # transverse field moved to the cell edges; the extra row/column only feeds the average
        fz_x = grid_average(window, 'x')[:-1, :-1]
        fz_y = grid_average(window, 'y')[:-1, :-1]
        if self.pol == 'Ez':
            if direction_normal == "x":
                return self.dl * np.sum(-1 / 2 * np.real(fz_x * np.conj(f2[sx, sy])))
            return self.dl * np.sum(1 / 2 * np.real(fz_y * np.conj(f1[sx, sy])))
        if direction_normal == "x":
            return self.dl * np.sum(1 / 2 * np.real(f2[sx, sy] * np.conj(fz_x)))
        return self.dl * np.sum(-1 / 2 * np.real(f1[sx, sy] * np.conj(fz_y)))

    def init_design_region(self, design_region, eps_m, style=''):
        """Initialise the permittivity inside ``design_region`` (simulation.py:355-383)."""
        outside = design_region == 0
        if style == 'halfway':
            eps = self.eps_r
            eps[design_region == 1] = eps_m / 2 + 1 / 2
        elif style in ('full', 'empty', 'random'):
            fill = {'full': lambda: eps_m * np.ones(self.eps_r.shape),
                    'empty': lambda: np.ones(self.eps_r.shape),
                    'random': lambda: (eps_m - 1) * np.random.random(self.eps_r.shape) + 1}[style]
            eps = fill()
            eps[outside] = self.eps_r[outside]
        else:
            return
        self.eps_r = eps

    # ------------------------------------------------------------------ plotting (optional dependency)
    def _plot(self, kind, **kw):
        from . import plot
        return getattr(plot, kind)(self, **kw)

    def plt_abs(self, **kw):
        return self._plot('plt_abs', **kw)

    def plt_re(self, **kw):
        return self._plot('plt_re', **kw)

    def plt_diff(self, **kw):
        return self._plot('plt_diff', **kw)

    def plt_eps(self, **kw):
        return self._plot('plt_eps', **kw)
