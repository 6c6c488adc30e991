"""Pin the CPU oracle against vectors produced by the unmodified reference (tests/golden)."""
import numpy as np
import scipy.sparse as sp

from oracle import fdfd_oracle as orc


def _csr(g, prefix, n):
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))


def _relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def test_sfactors_and_operator(golden):
    g = golden("operator")
    for tag in "abc":
        eps = g[tag + "_eps"]
        omega, dl, L0, npx, npy = g[tag + "_meta"]
        npml = [int(npx), int(npy)]
        nx, ny = eps.shape
        inv = orc.pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
        for mine, k in zip(inv, ("isxf", "isxb", "isyf", "isyb")):
            np.testing.assert_allclose(mine, g[f"{tag}_{k}"], rtol=1e-13)
        if npml[0] > 0:
            assert abs(inv[0][0].imag) > 1e-3      # the PML is really there
        for pol in ("Ez", "Hz"):
            ref = _csr(g, f"{tag}_{pol}_A", nx * ny)
            mine = orc.construct_A(omega, eps, dl, npml, pol, L0)
            d = (mine - ref)
            assert abs(d).max() <= 1e-13 * abs(ref).max(), (tag, pol)
            # same sparsity pattern apart from explicit zeros
            assert (abs(mine) > 0).nnz == (abs(ref) > 0).nnz
        # derivative operators with PML scaling (Ez derivs dict)
        isxf, isxb, isyf, isyb = orc.pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
        u = np.random.default_rng(0).standard_normal((nx, ny)) + 0j
        for name, fn, s, ax in (("Dxb", orc.d_back, isxb, 0), ("Dyb", orc.d_back, isyb, 1),
                                ("Dxf", orc.d_fwd, isxf, 0), ("Dyf", orc.d_fwd, isyf, 1)):
            ref = _csr(g, f"{tag}_{name}", nx * ny).dot(u.reshape(-1)).reshape(nx, ny)
            assert _relerr(fn(u, s, dl, ax), ref) < 1e-14


def test_apply_planes_matches_csr(golden):
    g = golden("operator")
    eps = g["a_eps"]
    omega, dl, L0, npx, npy = g["a_meta"]
    planes = orc.stencil_planes(omega, eps, dl, [int(npx), int(npy)], "Hz", L0)
    u = np.random.default_rng(1).standard_normal(eps.shape) + 1j
    ref = _csr(g, "a_Hz_A", eps.size).dot(u.reshape(-1)).reshape(eps.shape)
    assert _relerr(orc.apply_planes(planes, u), ref) < 1e-14


def test_linear_solve_small(golden):
    g = golden("linear_small")
    for pol in ("Ez", "Hz"):
        omega, dl, L0, npx, npy = g[pol + "_meta"]
        npml = [int(npx), int(npy)]
        f = orc.solve_fields(omega, g[pol + "_eps"], dl, npml, pol, L0, g[pol + "_src"])
        for mine, key in zip(f, ("_f1", "_f2", "_fz")):
            assert _relerr(mine, g[pol + key]) < 1e-10, (pol, key)
        np.testing.assert_allclose(orc.flux_probe(f, dl, pol, 'x', [40, 24], 20), g[pol + "_flux_x"], rtol=1e-8)
        np.testing.assert_allclose(orc.flux_probe(f, dl, pol, 'y', [32, 36], 30), g[pol + "_flux_y"], rtol=1e-8)


def test_config0_dipole(golden):
    g = golden("config0_dipole")
    omega, dl, L0, npx, npy = g["meta"]
    src = np.zeros((200, 200))
    src[100, 100] = 1
    hx, hy, ez = orc.solve_fields(omega, g["eps"], dl, [int(npx), int(npy)], "Ez", L0, src)
    assert _relerr(ez, g["ez"]) < 1e-10
    assert _relerr(hx[::10, ::10], g["hx_probe"]) < 1e-10
    assert _relerr(hy[::10, ::10], g["hy_probe"]) < 1e-10


def _up_to_sign(a, b):
    return min(_relerr(a, b), _relerr(-np.asarray(a), b))


def test_mode_source(golden):
    g = golden("mode_source")
    omega, dl, L0, npx, npy = g["meta"]
    eps, epsT = g["eps"], g["epsT"]
    for pol in ("Ez", "Hz"):
        prof, _ = orc.mode_profile(eps[15, 10:40], omega, dl, pol, L0, 3.5, direction_normal='x')
        assert _up_to_sign(prof, g[pol + "_src"][15, 10:40]) < 1e-8, pol
        profT, _ = orc.mode_profile(epsT[8:52, 15], omega, dl, pol, L0, 3.5, order=2, scale=2,
                                    direction_normal='y')
        assert _up_to_sign(profT, g[pol + "_srcT"][8:52, 15]) < 1e-8, pol
        f = orc.solve_fields(omega, eps, dl, [int(npx), int(npy)], pol, L0, g[pol + "_src"])
        assert _relerr(f[2], g[pol + "_fz"]) < 1e-9
        np.testing.assert_allclose(orc.flux_probe(f, dl, pol, 'x', [75, 25], 30), g[pol + "_flux"], rtol=1e-8)


def test_nonlinear(golden):
    g = golden("nonlinear")
    omega, dl, L0, npx, npy, chi3, eps_max = g["meta"]
    npml = [int(npx), int(npy)]
    eps, region, src = g["eps"], g["region"], g["src"]
    chi = chi3 / L0 ** 2                       # simulation.py:74 (add_nl)

    def kerr(e):
        return orc.kerr_terms(e, eps, chi, region, eps_scale=True, eps_max=eps_max)

    _, _, ez_lin = orc.solve_fields(omega, eps, dl, npml, "Ez", L0, src)
    assert _relerr(ez_lin, g["ez_lin"]) < 1e-10
    hx, hy, ez, conv = orc.born_solve(omega, eps, dl, npml, L0, src, kerr, e_start=ez_lin)
    assert _relerr(ez, g["born_ez"]) < 1e-8
    assert _relerr(hy, g["born_hy"]) < 1e-8
    nz = np.count_nonzero(g["born_conv"])
    assert np.count_nonzero(conv) == nz
    np.testing.assert_allclose(conv[:nz - 1], g["born_conv"][:nz - 1], rtol=1e-3)
    hx, hy, ez, conv = orc.newton_solve(omega, eps, dl, npml, L0, src, kerr, e_start=ez_lin)
    assert _relerr(ez, g["newton_ez"]) < 1e-8
    nz = np.count_nonzero(g["newton_conv"])
    assert np.count_nonzero(conv) == nz
    # the nonlinearity is strong enough to matter in this fixture
    assert _relerr(g["born_ez"], g["ez_lin"]) > 1e-4
