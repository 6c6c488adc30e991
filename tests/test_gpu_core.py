"""GPU parity tests of the C-ABI building blocks against the CPU oracle and the golden vectors.

Tolerances: the path is fp64; element-wise operator parity is held to 1e-13 relative, solved
fields to <= 1e-8 relative L2 (north-star bound) -- in practice they land near 1e-12.
"""
import numpy as np
import pytest

from oracle import fdfd_oracle as orc

pytestmark = pytest.mark.gpu

OMEGA = 2 * np.pi * 200e12


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def core():
    from fdfdpy_b200 import core
    return core


def test_zgemm_hook():
    from fdfdpy_b200 import _lib
    lib = _lib.load()
    _lib.require_gpu()
    rng = np.random.default_rng(0)

    def crand(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for (M, N, K, batch) in [(25, 25, 9, 7), (64, 64, 32, 3), (70, 130, 17, 2), (200, 96, 64, 1), (13, 40, 5, 11),
                             (520, 530, 70, 2), (1000, 700, 64, 1), (300, 200, 500, 3)]:
        A, C0 = crand(batch, M, K), crand(batch, M, N)
        for transb in (0, 1):
            B = crand(batch, N, K) if transb else crand(batch, K, N)
            prod = A @ (np.transpose(B, (0, 2, 1)) if transb else B)
            for mode in (0, 1):
                C = C0.copy()
                _lib.check(lib.fdfd_zgemm_batched_host(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, batch, mode,
                                                       transb, 0))
                ref = prod if mode == 0 else C0 - prod
                assert relerr(C, ref) < 1e-14, (M, N, K, batch, mode, transb)


def test_zgemm_4m_variant_matches():
    """The textbook 4M complex product stays available (fdfd_zgemm_set_variant bit 1) and agrees with the default 3M
    form to rounding; both against numpy."""
    from fdfdpy_b200 import _lib
    lib = _lib.load()
    _lib.require_gpu()
    rng = np.random.default_rng(2)
    for (M, N, K, batch) in [(150, 170, 90, 2), (1100, 900, 260, 1), (30, 28, 11, 9)]:
        A = rng.standard_normal((batch, M, K)) + 1j * rng.standard_normal((batch, M, K))
        B = rng.standard_normal((batch, N, K)) + 1j * rng.standard_normal((batch, N, K))
        ref = A @ np.transpose(B, (0, 2, 1))
        outs = []
        for variant in (0, 2, 1, 3):
            _lib.check(lib.fdfd_zgemm_set_variant(variant))
            C = np.zeros((batch, M, N), dtype=complex)
            _lib.check(lib.fdfd_zgemm_batched_host(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, batch, 0, 1, 0))
            assert relerr(C, ref) < 1e-14, (M, N, K, variant)
            outs.append(C)
        _lib.check(lib.fdfd_zgemm_set_variant(0))
        assert relerr(outs[0], outs[1]) < 1e-14
    # a purely real product keeps a clean imaginary part under 3M too (T3 - T1 - T2 cancels to rounding of |Re|)
    A = rng.standard_normal((1, 200, 300)) + 0j
    B = rng.standard_normal((1, 180, 300)) + 0j
    C = np.zeros((1, 200, 180), dtype=complex)
    _lib.check(lib.fdfd_zgemm_batched_host(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), 200, 180, 300, 1, 0, 1, 0))
    assert np.abs(C.imag).max() <= 1e-13 * np.abs(C.real).max()


def test_zgemm_lower_schur_update():
    """S -= G F^T on the lower tiles only: the lower triangle is exact, entries of tiles strictly
    above the diagonal tiles are untouched (persistent kernel for the big case, tiled for the small)."""
    from fdfdpy_b200 import _lib
    lib = _lib.load()
    _lib.require_gpu()
    rng = np.random.default_rng(1)
    for (M, K, batch) in [(24, 3, 50), (100, 40, 9), (200, 130, 5), (1300, 300, 1), (770, 64, 3)]:
        A = rng.standard_normal((batch, M, K)) + 1j * rng.standard_normal((batch, M, K))
        B = rng.standard_normal((batch, M, K)) + 1j * rng.standard_normal((batch, M, K))
        C0 = rng.standard_normal((batch, M, M)) + 1j * rng.standard_normal((batch, M, M))
        C = C0.copy()
        _lib.check(lib.fdfd_zgemm_batched_host(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, M, K, batch, 1, 1, 1))
        ref = C0 - A @ np.transpose(B, (0, 2, 1))
        assert relerr(np.tril(C), np.tril(ref)) < 1e-14, (M, K, batch)
        ts = 32 if M <= 32 else 64
        ti = np.arange(M) // ts
        above = ti[None, :] > ti[:, None]
        assert np.array_equal(C[:, above], C0[:, above]), (M, K, batch)


def test_operator_parity_golden(core, golden):
    g = golden("operator")
    for tag in "abc":
        eps = g[tag + "_eps"]
        omega, dl, L0, npx, npy = g[tag + "_meta"]
        npml = [int(npx), int(npy)]
        for pol in ("Ez", "Hz"):
            op = core.MaxwellOperator(omega, eps, dl, npml, pol, L0)
            for mine, k in zip(op.sfactors(), ("isxf", "isxb", "isyf", "isyb")):
                np.testing.assert_allclose(mine, g[f"{tag}_{k}"], rtol=1e-13)
            planes = op.planes()
            ref = orc.stencil_planes(omega, eps, dl, npml, pol, L0)
            for p, r in zip(planes, ref):
                assert np.abs(p - r).max() <= 1e-13 * np.abs(ref[0]).max(), (tag, pol)
            # against the reference's own CSR matrix
            import scipy.sparse as sp
            n = eps.size
            Aref = sp.csr_matrix((g[f"{tag}_{pol}_A_data"], g[f"{tag}_{pol}_A_indices"], g[f"{tag}_{pol}_A_indptr"]),
                                 shape=(n, n))
            d = op.to_scipy() - Aref
            assert abs(d).max() <= 1e-13 * abs(Aref).max()
            u = np.random.default_rng(3).standard_normal(eps.shape) + 1j * np.random.default_rng(4).standard_normal(eps.shape)
            ref_y = Aref.dot(u.reshape(-1)).reshape(eps.shape)
            assert relerr(op.dot(u), ref_y) < 1e-14
            assert relerr(op.dot(u, fused=True), ref_y) < 1e-13       # matrix-free kernels, both polarisations
            if pol == "Hz":                                           # Hz without edge averaging (linalg.py:74-76)
                op.assemble(eps, averaging=False)
                Ana = orc.construct_A(omega, eps, dl, npml, pol, L0, averaging=False)
                ref_na = Ana.dot(u.reshape(-1)).reshape(eps.shape)
                assert relerr(op.dot(u), ref_na) < 1e-14
                assert relerr(op.dot(u, fused=True), ref_na) < 1e-13


def test_fused_hz_stencil_layouts(core):
    """The matrix-free Hz kernel (48 B/cell: face weights rebuilt from eps_r by shuffles) on ragged shapes: widths
    that are not multiples of the warp / CTA, row counts that are not multiples of the rows a thread marches, lossy
    (complex) permittivity, several vectors per call, and the Kerr diagonal."""
    rng = np.random.default_rng(31)
    for (nx, ny), npml in [((33, 31), [4, 5]), ((70, 129), [6, 9]), ((19, 260), [0, 7]), ((128, 128), [10, 10])]:
        eps = (1 + 5 * rng.random((nx, ny))) * (1 + 0.03j * rng.random((nx, ny)))
        eps_nl = 0.1 * rng.random((nx, ny))
        U = rng.standard_normal((2, nx, ny)) + 1j * rng.standard_normal((2, nx, ny))
        for nl in (None, eps_nl):
            op = core.MaxwellOperator(OMEGA, eps, 0.03, npml, "Hz", 1e-6, eps_nl=nl)
            planes = orc.stencil_planes(OMEGA, eps, 0.03, npml, "Hz", 1e-6, eps_nl=nl)
            ref = np.stack([orc.apply_planes(planes, u) for u in U])
            assert relerr(op.dot(U), ref) < 1e-14, (nx, ny)
            assert relerr(op.dot(U, fused=True), ref) < 1e-13, (nx, ny)
            for rows in (8, 4):
                _ = op.lib.fdfd_stencil_set_variant(rows, 0)
                assert relerr(op.dot(U, fused=True), ref) < 1e-13, (nx, ny, rows)


def test_apply_multivector_and_nl(core):
    rng = np.random.default_rng(5)
    nx, ny = 70, 45
    eps = 1 + 3 * rng.random((nx, ny))
    eps_nl = 0.1 * rng.random((nx, ny))
    op = core.MaxwellOperator(OMEGA, eps, 0.03, [6, 5], "Ez", 1e-6, eps_nl=eps_nl)
    planes = orc.stencil_planes(OMEGA, eps, 0.03, [6, 5], "Ez", 1e-6, eps_nl=eps_nl)
    U = rng.standard_normal((3, nx, ny)) + 1j * rng.standard_normal((3, nx, ny))
    ref = np.stack([orc.apply_planes(planes, u) for u in U])
    assert relerr(op.dot(U), ref) < 1e-14
    assert relerr(op.dot(U, fused=True), ref) < 1e-13


@pytest.mark.parametrize("shape,npml", [((16, 16), [3, 3]), ((23, 17), [4, 3]), ((64, 48), [8, 6]),
                                        ((37, 90), [0, 7]), ((130, 75), [10, 10])])
@pytest.mark.parametrize("pol", ["Ez", "Hz"])
def test_direct_solver_vs_oracle(core, shape, npml, pol):
    rng = np.random.default_rng(11)
    nx, ny = shape
    eps = 1 + 5 * rng.random((nx, ny))
    op = core.MaxwellOperator(OMEGA, eps, 0.04, npml, pol, 1e-6)
    b = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    d = op.direct()
    x0 = d.solve(b, max_refine=0)                      # raw substitution, no refinement
    A = orc.construct_A(OMEGA, eps, 0.04, npml, pol, 1e-6)
    ref = orc.sparse_solve(A, b).reshape(nx, ny)
    assert relerr(x0, ref) < 1e-9, "unrefined"
    x = d.solve(b, max_refine=3, tol=1e-13)
    assert d.last_relres < 1e-11
    assert relerr(x, ref) < 1e-10


def test_direct_multi_rhs_reuses_factorisation(core):
    rng = np.random.default_rng(12)
    nx, ny = 48, 52
    eps = 1 + 5 * rng.random((nx, ny))
    op = core.MaxwellOperator(OMEGA, eps, 0.04, [6, 6], "Hz", 1e-6)
    d = op.direct()
    d.factor()
    B = rng.standard_normal((11, nx, ny)) + 1j * rng.standard_normal((11, nx, ny))
    X = d.solve(B)
    A = orc.construct_A(OMEGA, eps, 0.04, [6, 6], "Hz", 1e-6)
    for j in range(11):
        ref = orc.sparse_solve(A, B[j]).reshape(nx, ny)
        assert relerr(X[j], ref) < 1e-10, j
    # zero right-hand side gives zeros (linalg.py:129-130)
    assert not d.solve(np.zeros((nx, ny))).any()


def test_direct_many_rhs_tensor_substitution(core):
    """>= 8 right-hand sides per pass take the tensor-pipe substitution (mrhs.cuh: skinny GEMMs, factor blocks streamed
    once for 8 / 16 columns, K-split partial sums at the top of the tree): 37 = 16 + 16 + 4 + 1 columns, checked
    against the single-column path and the oracle; the second grid is large enough for separators cut into 512-node
    pieces and for the K-split."""
    rng = np.random.default_rng(14)
    for (nx, ny), pol, nrhs, ncheck in [((150, 131), "Hz", 37, 37), ((640, 600), "Ez", 24, 3)]:
        eps = 1 + 5 * (rng.random((nx, ny)) > 0.6)
        npml = [10, 12]
        op = core.MaxwellOperator(OMEGA, eps, 0.03, npml, pol, 1e-6)
        d = op.direct()
        B = rng.standard_normal((nrhs, nx, ny)) + 1j * rng.standard_normal((nrhs, nx, ny))
        X = d.solve(B, max_refine=0).reshape(nrhs, nx, ny)          # raw substitution: no refinement to hide behind
        for j in (0, nrhs - 1):
            xj = d.solve(B[j], max_refine=0).reshape(nx, ny)        # the matrix-vector path
            assert relerr(X[j], xj) < 1e-10, (nx, j)
        X = d.solve(B).reshape(nrhs, nx, ny)
        assert d.last_relres < 1e-10
        A = orc.construct_A(OMEGA, eps, 0.03, npml, pol, 1e-6)
        import scipy.sparse.linalg as spl
        lu = spl.splu(A.tocsc())
        for j in list(range(ncheck)) if ncheck == nrhs else [0, nrhs // 2, nrhs - 1]:
            ref = lu.solve(B[j].ravel()).reshape(nx, ny)
            assert relerr(X[j], ref) < 1e-8, (nx, j)


def test_config0_dipole_golden(core, golden):
    g = golden("config0_dipole")
    omega, dl, L0, npx, npy = g["meta"]
    op = core.MaxwellOperator(omega, g["eps"], dl, [int(npx), int(npy)], "Ez", L0)
    src = np.zeros((200, 200))
    src[100, 100] = 1
    ez = op.solve(src * 1j * omega).reshape(200, 200)
    assert relerr(ez, g["ez"]) < 1e-8
    hx, hy = op.derive_fields(ez)
    assert relerr(hx[::10, ::10], g["hx_probe"]) < 1e-8
    assert relerr(hy[::10, ::10], g["hy_probe"]) < 1e-8


def test_derived_fields_hz(core, golden):
    g = golden("linear_small")
    omega, dl, L0, npx, npy = g["Hz_meta"]
    op = core.MaxwellOperator(omega, g["Hz_eps"], dl, [int(npx), int(npy)], "Hz", L0)
    hz = op.solve(g["Hz_src"] * 1j * omega).reshape(g["Hz_eps"].shape)
    assert relerr(hz, g["Hz_fz"]) < 1e-8
    ex, ey = op.derive_fields(hz)
    assert relerr(ex, g["Hz_f1"]) < 1e-8
    assert relerr(ey, g["Hz_f2"]) < 1e-8


def test_krylov_small(core):
    rng = np.random.default_rng(13)
    nx, ny = 40, 36
    eps = 1 + 2 * rng.random((nx, ny))
    npml = [8, 8]
    op = core.MaxwellOperator(OMEGA, eps, 0.05, npml, "Ez", 1e-6)
    b = np.zeros((nx, ny), dtype=complex)
    b[20, 18] = 1j * OMEGA
    A = orc.construct_A(OMEGA, eps, 0.05, npml, "Ez", 1e-6)
    ref = orc.sparse_solve(A, b).reshape(nx, ny)
    # north_star's bar for fp64 fields: relative L2 error <= 1e-8 against the reference's solve
    x, info = op.krylov(b, method="bicgstab", tol=1e-12, maxiter=20000, check_every=20)
    assert info["relres"] < 1e-10, info
    assert relerr(x, ref) < 1e-8
    x, info = op.krylov(b, method="cocg", tol=1e-12, maxiter=20000, check_every=20)
    assert info["relres"] < 1e-10, info
    assert relerr(x, ref) < 1e-8
    # preconditioned by the cached factorisation of a PERTURBED operator: few iterations
    op2 = core.MaxwellOperator(OMEGA, eps, 0.05, npml, "Ez", 1e-6)
    op2.direct().factor()
    eps_pert = eps * (1 + 1e-3 * rng.random((nx, ny)))
    op2.assemble(eps_pert)               # new planes, old factors stay cached as the preconditioner
    assert op2.direct().has_factors and not op2.direct().factored
    x, info = op2.krylov(b, method="bicgstab", tol=1e-12, maxiter=50, check_every=1, precondition=True)
    A2 = orc.construct_A(OMEGA, eps_pert, 0.05, npml, "Ez", 1e-6)
    ref2 = orc.sparse_solve(A2, b).reshape(nx, ny)
    assert info["iters"] <= 10, info
    assert relerr(x, ref2) < 1e-9
    # restarted GMRES: unpreconditioned with a restart long enough to be full GMRES on this 1440-unknown grid (Ez and
    # Hz), with a short restart (several cycles), and right-preconditioned by the stale factors
    for pol in ("Ez", "Hz"):
        opg = op if pol == "Ez" else core.MaxwellOperator(OMEGA, eps, 0.05, npml, "Hz", 1e-6)
        refg = ref if pol == "Ez" else orc.sparse_solve(orc.construct_A(OMEGA, eps, 0.05, npml, "Hz", 1e-6), b).reshape(nx, ny)
        x, info = opg.krylov(b, method="gmres", tol=1e-12, maxiter=1500, restart=700)
        assert info["relres"] < 1e-10 and info["iters"] <= 1440, info
        assert relerr(x, refg) < 1e-8, pol
    x, info = op.krylov(b, method="gmres", tol=1e-11, maxiter=20000, restart=150)
    assert info["relres"] < 1e-10 and info["iters"] > 150, info       # went through restarts
    assert relerr(x, ref) < 1e-8
    x, info = op2.krylov(b, method="gmres", tol=1e-12, maxiter=50, restart=20, precondition=True)
    assert info["iters"] <= 10 and info["relres"] < 1e-11, info
    assert relerr(x, ref2) < 1e-9


def test_krylov_preconditioned_512(core):
    """The Krylov path at a size where unpreconditioned iteration is hopeless: 512 x 512 Ez, the cached factors of
    a NEARBY operator (permittivity changed inside a design region) precondition BiCGSTAB on the matrix-free
    stencil; parity with the oracle's sparse solve of the CHANGED operator at the fp64 bar (1e-8)."""
    rng = np.random.default_rng(17)
    n = 512
    eps = np.ones((n, n))
    eps[:, 236:276] = 12.25                                     # waveguide
    eps[180:330, 150:360] = 1 + 11 * (rng.random((150, 210)) > 0.5)
    npml = [15, 15]
    op = core.MaxwellOperator(OMEGA, eps, 0.02, npml, "Ez", 1e-6)
    op.direct().factor()
    eps2 = eps.copy()
    eps2[230:280, 200:300] += 0.05 * rng.random((50, 100))      # the design step
    op.assemble(eps2)
    b = np.zeros((n, n), dtype=complex)
    b[60, 256] = 1j * OMEGA
    x, info = op.krylov(b, method="bicgstab", tol=1e-12, maxiter=60, check_every=1, precondition=True, fused=True)
    assert info["relres"] < 1e-10 and info["iters"] <= 40, info
    ref = orc.sparse_solve(orc.construct_A(OMEGA, eps2, 0.02, npml, "Ez", 1e-6), b).reshape(n, n)
    assert relerr(x, ref) < 1e-8
    xg, infog = op.krylov(b, method="gmres", tol=1e-12, maxiter=60, restart=30, precondition=True, fused=True)
    assert infog["relres"] < 1e-10 and infog["iters"] <= 2 * info["iters"], (infog, info)
    assert relerr(xg, ref) < 1e-8


def test_residual_guard(core):
    """The residual contract of the direct solve: stale factors (operator changed after the factorisation, flag
    forced) make plain refinement stall; the call must either recover through BiCGSTAB on the factors or fail
    loudly -- never return a field whose residual is above 1e-10."""
    from fdfdpy_b200._lib import FdfdError
    rng = np.random.default_rng(23)
    nx, ny = 96, 80
    eps = 1 + 5 * rng.random((nx, ny))
    npml = [8, 8]
    op = core.MaxwellOperator(OMEGA, eps, 0.04, npml, "Ez", 1e-6)
    d = op.direct()
    d.factor()
    b = rng.standard_normal((2, nx, ny)) + 1j * rng.standard_normal((2, nx, ny))
    # (1) mild change: one refinement step is not enough, the Krylov fallback recovers
    eps1 = eps * (1 + 2e-3 * rng.random((nx, ny)))
    op.assemble(eps1)
    d.factored = True                                           # pretend the factors are current
    x = d.solve(b, max_refine=1, tol=1e-12).reshape(2, nx, ny)
    assert d.last_refine_steps >= 1000, d.last_refine_steps     # the fallback ran
    assert d.last_relres <= 1e-10
    A1 = orc.construct_A(OMEGA, eps1, 0.04, npml, "Ez", 1e-6)
    for j in range(2):
        assert relerr(x[j], orc.sparse_solve(A1, b[j]).reshape(nx, ny)) < 1e-8
    # (2) unrelated operator: nothing converges, the call fails instead of returning garbage
    op.assemble(1 + 11 * rng.random((nx, ny)))
    d.factored = True
    with pytest.raises(FdfdError, match="residual"):
        d.solve(b[0], max_refine=2, tol=1e-12)
    # (3) raw substitution (max_refine = 0) stays unguarded: it reports, it does not judge
    d.solve(b[0], max_refine=0)
    assert d.last_relres > 1e-10
    # (4) non-finite right-hand side is rejected
    bad = b[0].copy()
    bad[3, 3] = np.nan
    with pytest.raises(FdfdError, match="finite"):
        d.solve(bad)


def test_mode_solve_golden(core, golden):
    g = golden("mode_source")
    omega, dl, L0, npx, npy = g["meta"]
    eps, epsT = g["eps"], g["epsT"]

    def up_to_sign(a, b):
        return min(relerr(a, b), relerr(-a, b))

    for pol in ("Ez", "Hz"):
        vals, vecs = core.mode_solve(eps[15, 10:40], omega, dl, pol, L0, 3.5, order=1, averaged=False)
        ref, refval = orc.mode_profile(eps[15, 10:40], omega, dl, pol, L0, 3.5, direction_normal='x')
        assert abs(vals[0] - refval.real) < 1e-10 * abs(refval), pol
        assert up_to_sign(vecs[0], ref) < 1e-8, pol
        assert up_to_sign(vecs[0], g[pol + "_src"][15, 10:40]) < 1e-8, pol
        vals, vecs = core.mode_solve(epsT[8:52, 15], omega, dl, pol, L0, 3.5, order=2, averaged=True)
        assert up_to_sign(2 * vecs[1], g[pol + "_srcT"][8:52, 15]) < 1e-8, pol


@pytest.mark.parametrize("pol", ["Ez", "Hz"])
def test_single_slab_operator_matches_whole_grid(core, pol):
    """The slab code path (extended layout, halo rows, interior-only kernels, padded Krylov vectors) with one
    slab wrapping onto itself must reproduce the whole-grid operator."""
    from fdfdpy_b200.distributed import SlabOperator
    rng = np.random.default_rng(5)
    nx, ny = 40, 36
    eps = 1 + 2 * rng.random((nx, ny))
    npml = [8, 8]
    op = core.MaxwellOperator(OMEGA, eps, 0.05, npml, pol, 1e-6)
    slab = SlabOperator(OMEGA, eps, 0.05, npml, pol, 1e-6)
    x = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    ref = orc.construct_A(OMEGA, eps, 0.05, npml, pol, 1e-6).dot(x.ravel()).reshape(nx, ny)
    assert relerr(slab.dot(x), ref) < 1e-13
    assert relerr(op.dot(x), ref) < 1e-13
    assert relerr(slab.dot(x, fused=True), ref) < 1e-13
    b = np.zeros((nx, ny), dtype=complex)
    b[20, 18] = 1j * OMEGA
    sol = orc.sparse_solve(orc.construct_A(OMEGA, eps, 0.05, npml, pol, 1e-6), b).reshape(nx, ny)
    for method in ("bicgstab", "cocg"):
        xs, info = slab.krylov(b, method=method, tol=1e-12, maxiter=20000, check_every=20)
        assert info["relres"] < 1e-10, (method, info)
        assert relerr(xs, sol) < 1e-8


def test_complex64_stencil_and_krylov(core):
    """complex64 storage of the matrix-free path (fp64 arithmetic): north_star's tolerance is a relative L2
    error <= 1e-4 against the fp64 reference."""
    rng = np.random.default_rng(21)
    nx, ny = 48, 40
    eps = 1 + 2 * rng.random((nx, ny))
    npml = [8, 8]
    for pol in ("Ez", "Hz"):
        op = core.MaxwellOperator(OMEGA, eps, 0.05, npml, pol, 1e-6)
        A = orc.construct_A(OMEGA, eps, 0.05, npml, pol, 1e-6)
        x = (rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))).astype(np.complex64)
        ref = A.dot(x.astype(np.complex128).ravel()).reshape(nx, ny)
        for fused in (False, True):
            y = op.dot(x, fused=fused)
            assert y.dtype == np.complex64
            assert relerr(y, ref) < 1e-6, (pol, fused)
        # odd ny: one column per thread instead of the float4 pair kernel; ragged row count for both
        for shape in ((45, 41), (45, 38)):
            eps2 = 1 + 2 * rng.random(shape)
            op2 = core.MaxwellOperator(OMEGA, eps2, 0.05, npml, pol, 1e-6)
            x2 = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
            ref2 = orc.construct_A(OMEGA, eps2, 0.05, npml, pol, 1e-6).dot(x2.astype(np.complex128).ravel()).reshape(shape)
            assert relerr(op2.dot(x2, fused=True), ref2) < 1e-6, (pol, shape)
        b = np.zeros((nx, ny), dtype=np.complex64)
        b[24, 20] = 1j * OMEGA
        sol = orc.sparse_solve(A, b.astype(np.complex128)).reshape(nx, ny)
        for method in ("bicgstab", "cocg"):
            xs, info = op.krylov(b, method=method, tol=2e-6, maxiter=20000, check_every=20)
            assert xs.dtype == np.complex64
            assert info["relres"] < 1e-4, (pol, method, info)
            assert relerr(xs, sol) < 1e-4, (pol, method, relerr(xs, sol))
