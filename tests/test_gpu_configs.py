"""Parity at the sizes BASELINE.json's configs state (the small-grid versions live in test_gpu_simulation.py):

* config 1: Hz straight waveguide, 1000 x 1000, modal source through add_mode / setup_modes
  (source/mode.py:26-108) -- mode eigensolve + normalisation run + linear solve, against the oracle;
* config 2: the reference's own Kerr workload, tests/test_nonlinear_solvers.py:14-57 at its own 500 x 350
  grid -- Born and Newton against oracle.born_solve / oracle.newton_solve.

The oracle's sparse LU needs about a minute per 10^6 unknowns, so these are the slow end of the GPU suite.
Tolerances are north_star's: relative L2 <= 1e-8 on fields, solver residual <= 1e-10, W_in to 1e-7.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import fdfd_oracle as orc

pytestmark = pytest.mark.gpu

OMEGA = 2 * np.pi * 200e12
L0 = 1e-6


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_config1_hz_waveguide_modal_source_1000():
    from fdfdpy_b200 import Simulation
    n, dl, npml = 1000, 0.01, [15, 15]
    eps = np.ones((n, n))
    eps[:, 450:550] = 12.25                                   # 1 um wide silicon-like guide along x
    sim = Simulation(OMEGA, eps, dl, npml, 'Hz')
    sim.add_mode(3.5, 'x', [20, 500], 300, scale=1)
    sim.setup_modes()                                         # mode.py:26-62: normalisation run in the straight guide
    ex, ey, hz = sim.solve_fields()
    assert sim.last_solve["relres"] <= 1e-10
    # ---- oracle: the same pipeline restated from its parts
    (ix, _), (iy0, iy1) = orc._plane_indices('x', [20, 500], 300)
    prof, beta2 = orc.mode_profile(eps[ix, iy0:iy1], OMEGA, dl, 'Hz', L0, 3.5, direction_normal='x')
    src_o = np.zeros((n, n))
    src_o[ix, iy0:iy1] = prof
    if np.vdot(src_o, sim.src).real < 0:                      # ARPACK's eigenvector sign is arbitrary
        src_o = -src_o
    assert relerr(sim.src, src_o) < 1e-8
    ref = orc.solve_fields(OMEGA, eps, dl, npml, 'Hz', L0, src_o)
    for mine, theirs, name in zip((ex, ey, hz), ref, ('Ex', 'Ey', 'Hz')):
        assert relerr(mine, theirs) < 1e-8, name
    # the device IS the straight guide, so the normalisation run (mode.py:41-58) is this same solve
    w_in = orc.flux_probe(ref, dl, 'Hz', 'x', [n - 20, 500], 300)
    assert_allclose(sim.W_in, w_in, rtol=1e-7)
    assert_allclose(sim.E2_in, np.sum(np.abs(ref[2]) ** 2 * np.abs(src_o)), rtol=1e-7)
    # transmission of a straight lossless guide is ~1 between two planes inside the domain
    t = sim.flux_probe('x', [700, 500], 300) / sim.flux_probe('x', [300, 500], 300)
    assert abs(t - 1) < 2e-3


def _kerr_workload():
    """tests/test_nonlinear_solvers.py:14-40 of the reference, verbatim sizes."""
    n0, dl, chi3 = 3.4, 0.01, 2.8e-18
    width, L, L_chi3 = 1, 5, 4
    wv, lv = int(width / dl), int(L_chi3 / dl)
    nx, ny = int(L / dl), int(3.5 * width / dl)
    eps = np.ones((nx, ny))
    eps[:, int(ny / 2 - wv / 2):int(ny / 2 + wv / 2)] = np.square(n0)
    region = np.zeros(eps.shape)
    region[int(nx / 2 - lv / 2):int(nx / 2 + lv / 2), int(ny / 2 - wv / 2):int(ny / 2 + wv / 2)] = 1
    return n0, dl, chi3, eps, region, wv, ny


@pytest.mark.parametrize("strategy", ["reuse", "refactor"])
def test_config2_kerr_born_newton_500x350(strategy):
    from fdfdpy_b200 import Simulation
    n0, dl, chi3, eps, region, wv, ny = _kerr_workload()
    npml = [15, 15]
    assert eps.shape == (500, 350)
    sim = Simulation(OMEGA, eps, dl, npml, 'Ez')
    sim.nl_strategy = strategy
    sim.add_mode(n0, 'x', [17, int(ny / 2)], wv * 3)
    sim.setup_modes()
    sim.add_nl(chi3, region, eps_scale=True, eps_max=np.max(eps))
    src0 = np.array(sim.src)

    def kerr(e):
        return orc.kerr_terms(e, eps, chi3 / L0 ** 2, region, eps_scale=True, eps_max=np.max(eps))

    # the reference's own assertion at its three source strengths: Born and Newton agree (it asks for 1e-3)
    results = {}
    for srcval in np.logspace(1, 3, 3):
        sim.src = src0 * srcval
        sim.fields = {k: None for k in sim.fields}
        hx_n, hy_n, e_newton, conv_n = sim.solve_fields_nl(solver_nl='newton')
        e_newton = np.array(e_newton)
        hx_b, hy_b, e_born, conv_b = sim.solve_fields_nl(solver_nl='born')
        assert relerr(e_newton, e_born) < 1e-8, srcval
        results[float(srcval)] = (e_newton, np.array(e_born), np.array(hy_b), conv_n, conv_b)
    if strategy == "refactor":
        return                                                # oracle parity once is enough (below)
    # oracle parity: moderate power, both iterations; highest power (eps_nl ~ 1, 35 Born / 6 Newton steps), Newton
    e_newton, e_born, hy_b, conv_n, conv_b = results[100.0]
    rhx, rhy, rez, rconv = orc.born_solve(OMEGA, eps, dl, npml, L0, src0 * 100.0, kerr)
    assert relerr(e_born, rez) < 1e-8 and relerr(hy_b, rhy) < 1e-8
    assert np.count_nonzero(conv_b) == np.count_nonzero(rconv)
    rhx, rhy, rez, rconv = orc.newton_solve(OMEGA, eps, dl, npml, L0, src0 * 100.0, kerr)
    assert relerr(e_newton, rez) < 1e-8
    assert np.count_nonzero(conv_n) == np.count_nonzero(rconv)
    e_newton, e_born, hy_b, conv_n, conv_b = results[1000.0]
    rhx, rhy, rez, rconv = orc.newton_solve(OMEGA, eps, dl, npml, L0, src0 * 1000.0, kerr)
    assert relerr(e_newton, rez) < 1e-8
    assert np.count_nonzero(conv_n) == np.count_nonzero(rconv)
    assert np.max(kerr(rez)[0]) > 0.02                        # the nonlinearity is not a perturbation here (index shift ~ 1e-2)
