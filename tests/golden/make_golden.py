"""Generate golden input/output vectors by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py

Writes small ``.npz`` fixtures next to this file.  Requires /root/reference (build container
only); see ``_ref_import.py`` for the two missing third-party imports that are stubbed and
the upstream bare-name bugs that are patched at import time.  The fixtures are committed;
tests never read /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import_reference()
from fdfdpy import Simulation  # noqa: E402
from fdfdpy.pml import S_create  # noqa: E402
from fdfdpy.linalg import construct_A  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def csr_parts(A):
    A = A.tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return dict(data=A.data, indices=A.indices.astype(np.int32), indptr=A.indptr.astype(np.int32))


def gold_operator():
    """S-factors and the assembled A for both polarisations on a ragged small grid."""
    rng = np.random.default_rng(7)
    omega = 2 * np.pi * 200e12
    out = {}
    for tag, (nx, ny, npml, dl, L0) in {
        "a": (23, 17, [4, 3], 0.03, 1e-6),
        "b": (16, 30, [0, 5], 0.02, 1e-6),
        "c": (12, 12, [6, 6], 0.001, 1e-4),   # 2*NPML >= N: low-side branch wins everywhere
    }.items():
        eps = 1 + 11 * rng.random((nx, ny))
        out[tag + "_eps"] = eps
        out[tag + "_meta"] = np.array([omega, dl, L0, npml[0], npml[1]], dtype=np.float64)
        xr, yr = [0, float(nx * dl)], [0, float(ny * dl)]
        # S_create returns the four diagonal matrices of INVERSE stretch factors (pml.py:44-89);
        # (create_sfactor compares its kind argument with ``is``, so it is only reliable when
        # called from inside pml.py with literals -- go through S_create.)
        mats = S_create(omega, L0, np.array([nx, ny]), np.array(npml), xr, yr)
        for k, M in zip(("isxf", "isxb", "isyf", "isyb"), mats):
            d2 = M.diagonal().reshape(nx, ny)
            out[f"{tag}_{k}"] = d2[:, 0] if k[2] == "x" else d2[0, :]
            assert np.array_equal(d2, np.broadcast_to(d2[:, :1] if k[2] == "x" else d2[:1, :], d2.shape))
        for pol in ("Ez", "Hz"):
            A, derivs = construct_A(omega, xr, yr, eps, npml, pol, L0)
            for k, v in csr_parts(A).items():
                out[f"{tag}_{pol}_A_{k}"] = v
            if pol == "Ez":
                for dn, D in derivs.items():
                    for k, v in csr_parts(D).items():
                        out[f"{tag}_{dn}_{k}"] = v
    save("operator", **out)


def gold_linear():
    """solve_fields for Ez and Hz: a point dipole in a dielectric box (config 0 shrunk) and the
    full 200x200 config-0 Ez field."""
    omega = 2 * np.pi * 200e12
    out = {}
    for pol in ("Ez", "Hz"):
        nx, ny, dl, npml = 64, 48, 0.04, [8, 6]
        eps = np.ones((nx, ny))
        eps[20:44, 14:34] = 6.0
        sim = Simulation(omega, eps, dl, npml, pol)
        sim.src[30, 20] = 1
        sim.src[12, 40] = -0.5
        f = sim.solve_fields()
        out[pol + "_eps"] = eps
        out[pol + "_src"] = sim.src
        out[pol + "_meta"] = np.array([omega, dl, 1e-6, npml[0], npml[1]])
        out[pol + "_f1"], out[pol + "_f2"], out[pol + "_fz"] = f
        out[pol + "_flux_x"] = sim.flux_probe('x', [40, 24], 20)
        out[pol + "_flux_y"] = sim.flux_probe('y', [32, 36], 30)
    save("linear_small", **out)

    # config 0: notebooks/Examples.ipynb cell 2 (200x200 Ez point dipole) + dielectric box
    eps = np.ones((200, 200))
    eps[120:160, 60:140] = 4.0
    sim = Simulation(omega, eps, 0.02, [15, 15], 'Ez')
    sim.src[100, 100] = 1
    hx, hy, ez = sim.solve_fields()
    save("config0_dipole", eps=eps, meta=np.array([omega, 0.02, 1e-6, 15, 15]),
         ez=ez.astype(np.complex128), hx_probe=hx[::10, ::10], hy_probe=hy[::10, ::10])


def gold_mode():
    """Modal source insertion, normalisation and flux (tests/test_flux.py workload, coarse grid)."""
    omega = 2 * np.pi * 200e12
    out = {}
    dl = 0.02
    eps = np.ones((150, 50))
    eps[:, 20:30] = 12.25
    for pol in ("Ez", "Hz"):
        sim = Simulation(omega, eps, dl, [10, 10], pol)
        sim.add_mode(3.5, 'x', [15, 25], 30, scale=1)
        m = sim.modes[0]
        m.insert_mode(sim, sim.src)
        out[pol + "_src"] = np.array(sim.src)
        f = sim.solve_fields()
        out[pol + "_fz"] = f[2]
        out[pol + "_flux"] = sim.flux_probe('x', [75, 25], 30)
    # normal to y, order 2 (exercises the (N,1) slice and vecs[:, order-1])
    epsT = np.ones((60, 140))
    epsT[20:40, :] = 12.25
    for pol in ("Ez", "Hz"):
        sim = Simulation(omega, epsT, dl, [10, 10], pol)
        sim.add_mode(3.5, 'y', [30, 15], 44, scale=2, order=2)
        sim.modes[0].insert_mode(sim, sim.src)
        out[pol + "_srcT"] = np.array(sim.src)
    # full setup_modes (normalisation run) is only functional for Ez upstream (mode.py:60)
    sim = Simulation(omega, eps, dl, [10, 10], 'Ez')
    sim.add_mode(3.5, 'x', [15, 25], 30, scale=1)
    sim.setup_modes()
    out["Ez_W_in"] = sim.W_in
    out["Ez_E2_in"] = sim.E2_in
    out["Ez_src_setup"] = np.array(sim.src)
    out["eps"] = eps
    out["epsT"] = epsT
    out["meta"] = np.array([omega, dl, 1e-6, 10, 10])
    # tests/test_flux.py workload at full size: the upstream assertion (flux1 == flux2) does not
    # hold for the reference itself because the unit-norm mode profile is not rescaled with dl;
    # record what the reference actually returns (flux, W_in) for both resolutions.
    rows = []
    for (dl2, shape, wg, c, w, p) in [(0.01, (300, 100), (40, 60), [20, 50], 60, [150, 50]),
                                      (0.005, (600, 200), (80, 120), [20, 100], 120, [300, 100])]:
        e = np.ones(shape)
        e[:, wg[0]:wg[1]] = 12.25
        sim = Simulation(omega, e, dl2, [15, 15], 'Ez')
        sim.add_mode(3.5, 'x', c, w, scale=1)
        sim.setup_modes()
        sim.solve_fields()
        rows.append([sim.flux_probe('x', p, w), sim.W_in])
    out["flux_test"] = np.array(rows)
    save("mode_source", **out)


def gold_nonlinear():
    """Born and Newton Kerr solves (tests/test_nonlinear_solvers.py workload, coarse grid)."""
    n0 = 3.4
    omega = 2 * np.pi * 200e12
    dl = 0.04
    chi3 = 2.8e-18
    width, L, L_chi3 = 1, 5, 4
    wv, lv = int(width / dl), int(L_chi3 / dl)
    nx, ny = int(L / dl), int(3.5 * width / dl)
    eps = np.ones((nx, ny))
    eps[:, int(ny / 2 - wv / 2):int(ny / 2 + wv / 2)] = n0 ** 2
    region = np.zeros(eps.shape)
    region[int(nx / 2 - lv / 2):int(nx / 2 + lv / 2), int(ny / 2 - wv / 2):int(ny / 2 + wv / 2)] = 1
    sim = Simulation(omega, eps, dl, [10, 10], 'Ez')
    sim.add_mode(n0, 'x', [12, int(ny / 2)], wv * 3)
    sim.setup_modes()
    sim.add_nl(chi3, region, eps_scale=True, eps_max=np.max(eps))
    sim.src *= 300.0
    src = np.array(sim.src)
    hx, hy, ez_lin = sim.solve_fields()
    out = dict(eps=eps, region=region, src=src, ez_lin=ez_lin,
               meta=np.array([omega, dl, 1e-6, 10, 10, chi3, np.max(eps)]))
    hx, hy, ez, conv = sim.solve_fields_nl(solver_nl='born')
    out["born_ez"], out["born_conv"], out["born_hy"] = ez, conv, hy
    hx, hy, ez, conv = sim.solve_fields_nl(solver_nl='newton')
    out["newton_ez"], out["newton_conv"], out["newton_hy"] = ez, conv, hy
    out["eps_nl_final"] = np.array(sim.eps_nl)
    save("nonlinear", **out)


if __name__ == "__main__":
    gold_operator()
    gold_linear()
    gold_mode()
    gold_nonlinear()
