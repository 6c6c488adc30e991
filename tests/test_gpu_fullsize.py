"""Full-size parity of the headline workload (4096 x 4096 Ez, bench.py's synthetic slab) through
size-independent properties: the oracle's matrix-free operator and derived-field formulas applied to
the GPU result (the oracle cannot SOLVE at this size in test time, but it can check a solution),
linearity of the cached factorisation over several right-hand sides, and reciprocity of the
row-scaled operator."""
import numpy as np
import pytest

from oracle import fdfd_oracle as orc

pytestmark = pytest.mark.gpu

N = 4096


@pytest.fixture(scope="module")
def workload():
    import bench
    from fdfdpy_b200 import Simulation
    eps = bench.synthetic_eps(N)
    sim = Simulation(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
    planes = orc.stencil_planes(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
    return bench, sim, eps, planes


def test_headline_solve_checked_by_the_oracle_operator(workload):
    bench, sim, eps, planes = workload
    src = bench.synthetic_src(N)
    sim.src = src
    hx, hy, ez = sim.solve_fields()
    assert sim.last_solve["relres"] <= 1e-10
    b = 1j * bench.OMEGA0 * src
    r = b - orc.apply_planes(planes, ez)                      # the oracle's operator on the GPU's solution
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 1e-10
    rhx, rhy = orc.derived_fields(ez, bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
    assert np.linalg.norm(hx - rhx) / np.linalg.norm(rhx) <= 1e-12
    assert np.linalg.norm(hy - rhy) / np.linalg.norm(rhy) <= 1e-12


def test_linearity_and_reciprocity_at_full_size(workload):
    bench, sim, eps, planes = workload
    d = sim._op.direct()
    if not d.factored:
        d.factor()
    i, j = (1500, 1700), (2600, 900)
    b = np.zeros((3, N, N), dtype=np.complex128)
    b[0][i] = 1.0
    b[1][j] = 1.0
    b[2] = (2 - 1j) * b[0] + 0.5j * b[1]
    x = d.solve(b).reshape(3, N, N)                           # one cached factorisation, three right-hand sides
    assert d.last_relres <= 1e-10
    lin = (2 - 1j) * x[0] + 0.5j * x[1]
    assert np.linalg.norm(x[2] - lin) / np.linalg.norm(lin) <= 1e-10
    # D A is complex symmetric, D = diag(sxf[ix] syf[iy])  =>  G(j,i) / d_i = G(i,j) / d_j
    isxf, _, isyf, _ = orc.pml_inverse_factors(bench.OMEGA0, bench.L0, (N, N), bench.NPML, bench.DL)
    di = 1.0 / (isxf[i[0]] * isyf[i[1]])
    dj = 1.0 / (isxf[j[0]] * isyf[j[1]])
    gji, gij = x[0][j], x[1][i]
    assert abs(gji / di - gij / dj) <= 1e-9 * abs(gij / dj)
