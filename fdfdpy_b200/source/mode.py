"""Modal sources (reference: fdfdpy/source/mode.py).  The eigensolve runs in the CUDA mode kernel."""
from copy import deepcopy

import numpy as np

from ..constants import DEFAULT_MATRIX_FORMAT, EPSILON_0, MU_0
from ..core import mode_solve
from ..geometry import plane_slices


class ModeOperator:
    """The 1-D waveguide operator on a source line (mode.py:78-88), held as its defining data; the
    shifted solves and the Rayleigh-Ritz iteration happen on the device."""

    def __init__(self, eps_line, omega, dl, pol, L0, averaged):
        line = np.asarray(eps_line).reshape(-1)
        # The CUDA mode kernel solves the REAL symmetric eigenproblem (for Hz it symmetrises with sqrt(eps)).
        # The reference hands a complex matrix to ARPACK, so lossy or metallic cross-sections work there; here
        # they would silently lose their imaginary part / give NaN, so they are rejected.
        if np.iscomplexobj(line) and np.any(np.abs(np.imag(line)) > 1e-12 * np.max(np.abs(line))):
            raise ValueError("modal source: the permittivity on the source line must be real (lossless); "
                             "got complex values")
        line = np.real(line)
        if line.size < 3 or not np.all(line > 0) or not np.all(np.isfinite(line)):
            raise ValueError("modal source: the permittivity on the source line must be positive and finite "
                             "on at least 3 cells")
        self.eps_line = line
        self.omega, self.dl, self.pol, self.L0, self.averaged = omega, dl, pol, L0, bool(averaged)
        n = self.eps_line.size
        self.shape = (n, n)

    def eigs(self, k, sigma):
        """k eigenpairs nearest sigma, closest first; vectors as columns like scipy's eigs."""
        neff = np.sqrt(sigma) / (self.omega * np.sqrt(MU_0 * self.L0 * EPSILON_0 * self.L0))
        vals, vecs = mode_solve(self.eps_line, self.omega, self.dl, self.pol, self.L0, neff, order=k,
                                averaged=self.averaged)
        # ARPACK's eigenvector sign is arbitrary (it depends on its random start vector), so a source built from
        # it -- and every field it drives -- is defined up to a global factor -1 in the reference.  Here the sign
        # is fixed: the largest-magnitude component is positive.
        for v in vecs:
            if v[np.argmax(np.abs(v))] < 0:
                v *= -1
        return vals.astype(np.complex128), vecs.T.astype(np.complex128)


class mode:
    """A modal source definition: same constructor and methods as the reference class (mode.py:9-27)."""

    def __init__(self, neff, direction_normal, center, width, scale, order=1):
        self.neff, self.order, self.scale = neff, order, scale
        self.direction_normal, self.center, self.width = direction_normal, center, width

    def setup_src(self, simulation, matrix_format=DEFAULT_MATRIX_FORMAT):
        self.compute_normalization(simulation, matrix_format=matrix_format)
        self.insert_mode(simulation, simulation.src, matrix_format=matrix_format)

    def _straight_guide(self, eps):
        """Permittivity of a uniform waveguide continuing the cross-section under the source plane,
        and the probe centre mirrored to the far side (mode.py:41-52)."""
        nx, ny = eps.shape
        guide = np.ones((nx, ny))
        probe = list(self.center)
        top = np.max(np.abs(eps))
        if self.direction_normal == "x":
            guide[:, eps[self.center[0], :] > 1] = top
            probe[0] = nx - probe[0]
        elif self.direction_normal == "y":
            guide[eps[:, self.center[1]] > 1, :] = top
            probe[1] = ny - probe[1]
        else:
            raise ValueError("The value of direction_normal is not x or y!")
        return guide, probe

    def compute_normalization(self, simulation, matrix_format=DEFAULT_MATRIX_FORMAT):
        """Run the source in a straight waveguide and record the injected power on ``simulation``
        as ``W_in`` and ``E2_in`` (mode.py:29-62)."""
        guide, probe = self._straight_guide(simulation.eps_r)
        twin = deepcopy(simulation)
        twin.eps_r = guide
        self.insert_mode(twin, twin.src, matrix_format=matrix_format)
        twin.solve_fields()
        simulation.W_in = twin.flux_probe(self.direction_normal, probe, self.width)
        # the reference reads fields['Ez'] here, which only exists for Ez runs (mode.py:60-61);
        # use the transverse field of whichever polarisation is being solved.
        fz = twin.fields[twin.pol]
        simulation.E2_in = np.sum(np.abs(fz) ** 2 * np.abs(twin.src))

    def insert_mode(self, simulation, destination, matrix_format=DEFAULT_MATRIX_FORMAT):
        """Solve for the mode profile on the source line and write it into ``destination`` (mode.py:64-108)."""
        from ..linalg import solver_eigs
        sx, sy = plane_slices(self.direction_normal, self.center, self.width)
        line = simulation.eps_r[sx, sy]
        # the reference edge-averages the 2-D slice along axis 0 (mode.py:82): a no-op for the (1, N)
        # slice of an x-normal plane, a real average for the (N, 1) slice of a y-normal plane
        A = ModeOperator(line, simulation.omega, simulation.dl, simulation.pol, simulation.L0,
                         averaged=(self.direction_normal == "y"))
        beta = simulation.omega * np.sqrt(MU_0 * simulation.L0 * EPSILON_0 * simulation.L0) * self.neff
        _, vecs = solver_eigs(A, self.order, guess_value=beta ** 2)
        profile = (vecs[:, self.order - 1] * self.scale).reshape(line.shape)
        destination[sx, sy] = np.abs(profile) * np.sign(np.real(profile))
