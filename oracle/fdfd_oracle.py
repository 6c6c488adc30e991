"""CPU oracle for the fdfdpy 2-D FDFD hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the reference algorithm, used only as the
checker in ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``.  Nothing under ``fdfdpy_b200/`` imports it; the product
path is CUDA only and fails loudly when the extension is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
fixtures in ``tests/golden/*.npz`` that were produced by importing the unmodified reference
(``tests/golden/make_golden.py``; the reference has no golden vectors of its own, its tests
are self-consistency checks, which are restated in ``tests/`` too).

Third-party algorithms on the path that live outside /root/reference:
* the sparse LU: the reference defaults to MKL Pardiso through ``pyMKL`` (setup.py:18,
  unpinned; not installable here) and offers SuperLU through ``scipy.sparse.linalg.spsolve``
  (linalg.py:139).  The oracle uses scipy's SuperLU (``splu``), the reference's own
  alternative branch; both are direct solvers, so they agree to round-off*cond(A).
* the modal eigensolve: ARPACK shift-invert through ``scipy.sparse.linalg.eigs``
  (linalg.py:113, called from source/mode.py:92).  The oracle calls the same routine.

Layout conventions follow the reference: fields are (Nx, Ny) C-ordered, the unknown vector is
``field.reshape(-1)`` so the y index is fastest (derivatives.py:9-11).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

# fdfdpy/constants.py:3-6
EPSILON_0 = 8.85418782e-12
MU_0 = 1.25663706e-6
C_0 = np.sqrt(1 / EPSILON_0 / MU_0)
ETA_0 = np.sqrt(MU_0 / EPSILON_0)


# --------------------------------------------------------------------------------------
# sc-PML stretching factors (pml.py:7-41)
# --------------------------------------------------------------------------------------
def pml_sfactor(omega, L0, n, npml, dl, kind):
    """1-D complex stretch factor s[i] for an axis of ``n`` cells of size ``dl``.

    Restates pml.py:21-41 (create_sfactor) with sig_w / S (pml.py:7-18) folded in:
    polynomial grading m=4, ln R = -12.  ``kind`` is 'f' (forward / half-cell positions)
    or 'b' (backward / integer positions).  Note the reference's asymmetric bounds:
    cells ``i <= npml`` on the low side, ``i > n - npml`` on the high side.
    """
    s = np.ones(n, dtype=np.complex128)
    if npml < 1:
        return s
    m, ln_r = 4, -12.0
    thick = npml * dl
    sig_max = -(m + 1) * ln_r / (2 * ETA_0 * thick)
    shift = 0.5 if kind == 'f' else 1.0
    i = np.arange(n)
    lo = i <= npml
    hi = (~lo) & (i > n - npml)
    depth = np.zeros(n)
    depth[lo] = dl * (npml - i[lo] + shift)
    depth[hi] = dl * (i[hi] - (n - npml) - shift)
    sel = lo | hi
    sig = sig_max * (depth[sel] / thick) ** m
    s[sel] = 1 - 1j * sig / (omega * EPSILON_0 * L0)
    return s


def pml_inverse_factors(omega, L0, shape, npml, dl):
    """(1/sx_f, 1/sx_b, 1/sy_f, 1/sy_b) as 1-D arrays; the diagonals of pml.py:44-89."""
    nx, ny = shape
    return (1 / pml_sfactor(omega, L0, nx, npml[0], dl, 'f'),
            1 / pml_sfactor(omega, L0, nx, npml[0], dl, 'b'),
            1 / pml_sfactor(omega, L0, ny, npml[1], dl, 'f'),
            1 / pml_sfactor(omega, L0, ny, npml[1], dl, 'b'))


# --------------------------------------------------------------------------------------
# material averaging (linalg.py:14-20)
# --------------------------------------------------------------------------------------
def edge_average(a, axis):
    """Mean of each cell with its lower neighbour (periodic): linalg.py:14-20."""
    return (np.roll(a, 1, axis=axis) + a) / 2


# --------------------------------------------------------------------------------------
# derivative operators applied matrix-free (derivatives.py:7-34 composed with the PML
# diagonals as in linalg.py:56-59)
# --------------------------------------------------------------------------------------
def d_back(u, inv_s_b, dl, axis):
    """S_b^-1 (u[i] - u[i-1]) / dl along ``axis``, periodic wrap (derivatives.py:24-26,31-33)."""
    sh = [1, 1]
    sh[axis] = -1
    return inv_s_b.reshape(sh) * (u - np.roll(u, 1, axis=axis)) / dl


def d_fwd(u, inv_s_f, dl, axis):
    """S_f^-1 (u[i+1] - u[i]) / dl along ``axis``, periodic wrap (derivatives.py:21-23,28-30)."""
    sh = [1, 1]
    sh[axis] = -1
    return inv_s_f.reshape(sh) * (np.roll(u, -1, axis=axis) - u) / dl


# --------------------------------------------------------------------------------------
# the Maxwell operator as five stencil planes (linalg.py:39-114, construct_A)
# --------------------------------------------------------------------------------------
def stencil_planes(omega, eps_r, dl, npml, pol, L0, eps_nl=None, averaging=True):
    """Five (Nx, Ny) complex planes (c0, cxm, cxp, cym, cyp) with

        (A u)[i,j] = c0 u[i,j] + cxm u[i-1,j] + cxp u[i+1,j] + cym u[i,j-1] + cyp u[i,j+1]

    (indices periodic).  Ez: A = (Dxf Dxb + Dyf Dyb)/mu0' + w^2 eps0' eps_r   (linalg.py:50-64)
    Hz: A = Dxf ex^-1 Dxb + Dyf ey^-1 Dyb + w^2 mu0'                          (linalg.py:67-98)
    with eps0' = eps0*L0, mu0' = mu0*L0 and ex, ey the edge-averaged permittivities.
    ``eps_nl`` adds w^2 eps0' eps_nl to the diagonal (simulation.py:68-70, Anl).
    """
    eps_r = np.asarray(eps_r)
    nx, ny = eps_r.shape
    e0, m0 = EPSILON_0 * L0, MU_0 * L0
    isxf, isxb, isyf, isyb = pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
    if pol == 'Ez':
        wx = np.full((nx, ny), 1 / m0, dtype=np.complex128)   # weight on the backward x-difference
        wy = wx
        shift = omega ** 2 * e0 * eps_r.astype(np.complex128)
    elif pol == 'Hz':
        if averaging:
            wx = 1 / (e0 * edge_average(eps_r, 0)).astype(np.complex128)
            wy = 1 / (e0 * edge_average(eps_r, 1)).astype(np.complex128)
        else:
            wx = wy = 1 / (e0 * eps_r).astype(np.complex128)
        shift = np.full((nx, ny), omega ** 2 * m0, dtype=np.complex128)
    else:
        raise ValueError("pol must be 'Ez' or 'Hz', got {}".format(pol))
    # flux through the lower face of cell i is  wb[i] * (u[i]-u[i-1]) / dl
    bx = isxb[:, None] * wx / dl
    by = isyb[None, :] * wy / dl
    cxm = isxf[:, None] * bx / dl
    cxp = isxf[:, None] * np.roll(bx, -1, axis=0) / dl
    cym = isyf[None, :] * by / dl
    cyp = isyf[None, :] * np.roll(by, -1, axis=1) / dl
    c0 = shift - (cxm + cxp) - (cym + cyp)
    if eps_nl is not None:
        c0 = c0 + omega ** 2 * e0 * np.asarray(eps_nl)
    return c0, cxm, cxp, cym, cyp


def planes_to_csr(planes):
    """Assemble the five stencil planes into a scipy CSR matrix (row = i*Ny + j)."""
    c0, cxm, cxp, cym, cyp = planes
    nx, ny = c0.shape
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
    row = (ii * ny + jj).ravel()
    cols = [row,
            (((ii - 1) % nx) * ny + jj).ravel(),
            (((ii + 1) % nx) * ny + jj).ravel(),
            (ii * ny + (jj - 1) % ny).ravel(),
            (ii * ny + (jj + 1) % ny).ravel()]
    vals = [p.ravel() for p in (c0, cxm, cxp, cym, cyp)]
    A = sp.coo_matrix((np.concatenate(vals), (np.tile(row, 5), np.concatenate(cols))),
                      shape=(nx * ny, nx * ny))
    return A.tocsr()   # duplicate (wrap) entries are summed, as scipy does for the reference


def apply_planes(planes, u):
    """Matrix-free A u for a (Nx, Ny) field."""
    c0, cxm, cxp, cym, cyp = planes
    return (c0 * u + cxm * np.roll(u, 1, 0) + cxp * np.roll(u, -1, 0)
            + cym * np.roll(u, 1, 1) + cyp * np.roll(u, -1, 1))


def construct_A(omega, eps_r, dl, npml, pol, L0, averaging=True):
    """CSR system matrix; the counterpart of linalg.py:39 (construct_A)."""
    return planes_to_csr(stencil_planes(omega, eps_r, dl, npml, pol, L0, averaging=averaging))


# --------------------------------------------------------------------------------------
# linear solve and derived fields (linalg.py:123-149, simulation.py:115-178)
# --------------------------------------------------------------------------------------
def sparse_solve(A, b):
    """Direct solve; all-zero right-hand side short-circuits to zeros (linalg.py:129-130)."""
    b = np.asarray(b, dtype=np.complex128).reshape(-1)
    if not b.any():
        return np.zeros(b.shape)
    return spl.splu(sp.csc_matrix(A)).solve(b)


def derived_fields(X, omega, eps_tot, dl, npml, pol, L0, averaging=True):
    """The two in-plane fields from the solved transverse field X (simulation.py:138-176)."""
    nx, ny = X.shape
    e0, m0 = EPSILON_0 * L0, MU_0 * L0
    _, isxb, _, isyb = pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
    dxb = d_back(X, isxb, dl, 0)
    dyb = d_back(X, isyb, dl, 1)
    if pol == 'Ez':
        hx = -1 / 1j / omega / m0 * dyb
        hy = 1 / 1j / omega / m0 * dxb
        return hx, hy
    if averaging:
        ex_w = e0 * edge_average(eps_tot, 0)
        ey_w = e0 * edge_average(eps_tot, 1)
    else:
        ex_w = ey_w = e0 * eps_tot
    ex = 1 / 1j / omega * (dyb / ey_w)
    ey = -1 / 1j / omega * (dxb / ex_w)
    return ex, ey


def solve_fields(omega, eps_r, dl, npml, pol, L0, src, eps_nl=None, averaging=True):
    """Full linear solve: returns (Hx, Hy, Ez) or (Ex, Ey, Hz) (simulation.py:115-178)."""
    eps_r = np.asarray(eps_r)
    planes = stencil_planes(omega, eps_r, dl, npml, pol, L0, eps_nl=eps_nl, averaging=averaging)
    A = planes_to_csr(planes)
    X = sparse_solve(A, np.asarray(src) * 1j * omega).reshape(eps_r.shape)
    eps_tot = eps_r if eps_nl is None else eps_r + eps_nl
    f1, f2 = derived_fields(X.astype(np.complex128), omega, eps_tot, dl, npml, pol, L0, averaging)
    return f1, f2, X


# --------------------------------------------------------------------------------------
# modal source (source/mode.py:64-108) and flux probe (simulation.py:267-327)
# --------------------------------------------------------------------------------------
def _plane_indices(direction_normal, center, width):
    if direction_normal == 'x':
        return (center[0], center[0] + 1), (int(center[1] - width / 2), int(center[1] + width / 2))
    if direction_normal == 'y':
        return (int(center[0] - width / 2), int(center[0] + width / 2)), (center[1], center[1] + 1)
    raise ValueError("The value of direction_normal is neither x nor y!")


def mode_operator(eps_line, omega, dl, pol, L0, direction_normal='x'):
    """Dense 1-D waveguide operator whose eigenvalues are beta^2 (source/mode.py:78-90).

    Periodic second difference on the slice, no PML.  Ez: w^2 mu0' eps + Dxf Dxb;
    Hz: w^2 mu0' eps + eps Dxf ex^-1 Dxb.  ex is ``grid_average(slice, 'x')`` of the 2-D
    slice (mode.py:82): for a plane normal to x the slice is (1, N) so the roll along
    axis 0 is a no-op and ex = eps; for a plane normal to y the slice is (N, 1) and ex is
    the true edge average along the line.  The oracle keeps that asymmetry.
    """
    n = eps_line.size
    e0, m0 = EPSILON_0 * L0, MU_0 * L0
    eps = e0 * eps_line.astype(np.float64)
    eye = np.eye(n)
    db = (eye - np.roll(eye, -1, axis=1)) / dl      # (db u)[i] = (u[i]-u[i-1])/dl
    df = (np.roll(eye, 1, axis=1) - eye) / dl       # (df u)[i] = (u[i+1]-u[i])/dl
    if pol == 'Ez':
        return omega ** 2 * m0 * np.diag(eps) + df @ db
    eps_x = (np.roll(eps, 1) + eps) / 2 if direction_normal == 'y' else eps
    return omega ** 2 * m0 * np.diag(eps) + np.diag(eps) @ df @ np.diag(1 / eps_x) @ db


def mode_profile(eps_line, omega, dl, pol, L0, neff, order=1, scale=1, direction_normal='x'):
    """Signed-magnitude mode profile inserted as the source (source/mode.py:91-108).

    The eigenvector comes from ARPACK shift-invert around (w sqrt(mu0' eps0') neff)^2; its
    overall sign is arbitrary (ARPACK start vector), so consumers compare up to +-1.
    """
    e0, m0 = EPSILON_0 * L0, MU_0 * L0
    A = sp.csr_matrix(mode_operator(eps_line, omega, dl, pol, L0, direction_normal))
    beta = omega * np.sqrt(m0 * e0) * neff
    vals, vecs = spl.eigs(A, k=order, sigma=beta ** 2, which='LM')
    v = vecs[:, order - 1] * scale
    return np.abs(v) * np.sign(np.real(v)), vals[order - 1]


def flux_probe(fields, dl, pol, direction_normal, center, width):
    """Poynting flux through a line (simulation.py:267-327).  ``fields`` = solve_fields tuple."""
    ix, iy = _plane_indices(direction_normal, center, width)
    f1, f2, fz = fields
    win = fz[ix[0]:ix[1] + 1, iy[0]:iy[1] + 1]
    fz_x = edge_average(win, 0)[:-1, :-1]
    fz_y = edge_average(win, 1)[:-1, :-1]
    cut = (slice(ix[0], ix[1]), slice(iy[0], iy[1]))
    if pol == 'Ez':
        hx, hy = f1, f2
        if direction_normal == 'x':
            return dl * np.sum(-0.5 * np.real(fz_x * np.conj(hy[cut])))
        return dl * np.sum(0.5 * np.real(fz_y * np.conj(hx[cut])))
    ex, ey = f1, f2
    if direction_normal == 'x':
        return dl * np.sum(0.5 * np.real(ey[cut] * np.conj(fz_x)))
    return dl * np.sum(-0.5 * np.real(ex[cut] * np.conj(fz_y)))


# --------------------------------------------------------------------------------------
# Kerr nonlinearity and the Born / Newton iterations (nonlinearity.py, nonlinear_solvers.py)
# --------------------------------------------------------------------------------------
def kerr_terms(e, eps_r, chi, nl_region, eps_scale=False, eps_max=None):
    """(eps_nl, d eps_nl/d e) for the Kerr model (nonlinearity.py:19-32).  chi is already /L0^2."""
    if eps_scale:
        w = (eps_r - 1) / (eps_max - 1)
    else:
        w = 1.0
    return (3 * chi * nl_region * np.abs(e) ** 2 * w,
            3 * chi * nl_region * np.conj(e) * w)


def born_solve(omega, eps_r, dl, npml, L0, src, kerr, e_start=None, tol=1e-10, max_iter=50):
    """Fixed-point (Born) iteration for Ez (nonlinear_solvers.py:14-52).  ``kerr(e)`` -> (eps_nl, dnl_de)."""
    if e_start is None:
        _, _, ez = solve_fields(omega, eps_r, dl, npml, 'Ez', L0, src)
    else:
        ez = e_start
    conv = np.zeros((max_iter, 1))
    for it in range(max_iter):
        prev = ez
        eps_nl, _ = kerr(prev)
        hx, hy, ez = solve_fields(omega, eps_r, dl, npml, 'Ez', L0, src, eps_nl=eps_nl)
        conv[it] = np.linalg.norm(ez - prev) / np.linalg.norm(ez)
        if conv[it] < tol:
            break
    return hx, hy, ez, conv


def newton_solve(omega, eps_r, dl, npml, L0, src, kerr, e_start=None, tol=1e-10, max_iter=50):
    """Newton iteration on f(E) = (A + Anl(E)) E - i w src = 0 (nonlinear_solvers.py:55-149).

    The Jacobian acts on (dE, dE*); the reference solves it in the real 2N x 2N form
    [[Re(J11+J12), -Im(J11-J12)], [Im(J11+J12), Re(J11-J12)]] (linalg.py:152-186).
    """
    e0 = EPSILON_0 * L0
    n = eps_r.size
    if e_start is None:
        _, _, ez = solve_fields(omega, eps_r, dl, npml, 'Ez', L0, src)
    else:
        ez = e_start
    conv = np.zeros((max_iter, 1))
    for it in range(max_iter):
        prev = ez
        eps_nl, dnl = kerr(prev)
        anl = planes_to_csr(stencil_planes(omega, eps_r, dl, npml, 'Ez', L0, eps_nl=eps_nl))
        e_vec = prev.reshape(-1)
        f = anl.dot(e_vec) - np.asarray(src).reshape(-1) * 1j * omega
        dade = dnl.reshape(-1) * omega ** 2 * e0
        j11 = anl + sp.diags(dade * e_vec)
        j12 = sp.diags(np.conj(dade) * e_vec)
        big = sp.bmat([[(j11 + j12).real, -(j11 - j12).imag],
                       [(j11 + j12).imag, (j11 - j12).real]]).tocsc()
        rhs = np.concatenate([f.real, f.imag])
        if not f.any():
            d = np.zeros(n, dtype=np.complex128)
        else:
            sol = spl.splu(big).solve(rhs)
            d = sol[:n] + 1j * sol[n:]
        ez = prev - d.reshape(eps_r.shape)
        conv[it] = np.linalg.norm(ez - prev) / np.linalg.norm(ez)
        if conv[it] < tol:
            break
    eps_nl, _ = kerr(ez)
    hx, hy, ez = solve_fields(omega, eps_r, dl, npml, 'Ez', L0, src, eps_nl=eps_nl)
    return hx, hy, ez, conv
