"""numpy/scipy model of the restricted additive Schwarz preconditioner the slab path uses (csrc/krylov.cu schwarz_apply,
csrc/operator.cu schwarz_sfactor_kernel): same subdomains (slab + overlap + artificial PML, closed into a local torus),
same stretch factors, scipy SuperLU in place of the device factors.  Test infrastructure: used by the CPU test of the
algorithm and to study iteration counts (python tests/schwarz_model.py NX NY SLABS OVERLAP NPML_SUB [device|rods|vac])."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from oracle import fdfd_oracle as orc


def ez_planes(omega, eps, dl, L0, isxf, isxb, isyf, isyb):
    """The five Ez stencil planes from explicit 1-D inverse stretch factors (orc.stencil_planes computes its own)."""
    nx, ny = eps.shape
    e0, m0 = orc.EPSILON_0 * L0, orc.MU_0 * L0
    w = np.full((nx, ny), 1 / m0, dtype=complex)
    bx, by = isxb[:, None] * w / dl, isyb[None, :] * w / dl
    cxm, cxp = isxf[:, None] * bx / dl, isxf[:, None] * np.roll(bx, -1, 0) / dl
    cym, cyp = isyf[None, :] * by / dl, isyf[None, :] * np.roll(by, -1, 1) / dl
    return omega ** 2 * e0 * eps - (cxm + cxp) - (cym + cyp), cxm, cxp, cym, cyp


def artificial_sigma(depth_cells, npml_s, dl, omega, L0):
    """Im part added to s in the artificial layer: the reference's grading (pml.py:7-18), m = 4, ln R = -12."""
    thick = npml_s * dl
    smax = 5 * 12.0 / (2 * orc.ETA_0 * thick)
    return smax * (depth_cells * dl / thick) ** 4 / (omega * orc.EPSILON_0 * L0)


def subdomain_sfactors(gis, rows, npml_s, dl, omega, L0):
    nl = len(rows)
    loc = np.arange(nl)
    hi0 = nl - npml_s
    db = np.where(loc < npml_s, npml_s - loc, np.where(loc >= hi0, loc - hi0 + 1, 0)).astype(float)
    df = np.where(loc < npml_s, npml_s - loc - 0.5, np.where(loc >= hi0, loc - hi0 + 0.5, 0)).astype(float)
    sxf = 1 / gis[0][rows] - 1j * artificial_sigma(df, npml_s, dl, omega, L0)
    sxb = 1 / gis[1][rows] - 1j * artificial_sigma(db, npml_s, dl, omega, L0)
    return 1 / sxf, 1 / sxb


def build(omega, eps, dl, npml, L0, slabs, overlap, npml_s):
    """Global CSR matrix and the preconditioner M(r) -> z as a function on flat vectors."""
    nx, ny = eps.shape
    gis = orc.pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
    A = orc.planes_to_csr(ez_planes(omega, eps, dl, L0, *gis))
    ext = overlap + npml_s
    subs = []
    for r in range(slabs):
        x0, x1 = slab_rows(nx, slabs, r)
        rows = np.arange(x0 - ext, x1 + ext) % nx
        isxf, isxb = subdomain_sfactors(gis, rows, npml_s, dl, omega, L0)
        Al = orc.planes_to_csr(ez_planes(omega, eps[rows], dl, L0, isxf, isxb, gis[2], gis[3]))
        subs.append((x0, x1, rows, spl.splu(sp.csc_matrix(Al))))

    def M(rv):
        r2 = rv.reshape(nx, ny)
        out = np.zeros_like(r2, dtype=complex)
        for x0, x1, rows, lu in subs:
            nl = len(rows)
            rl = np.zeros((nl, ny), complex)
            rl[npml_s:nl - npml_s] = r2[rows[npml_s:nl - npml_s]]        # owned + overlap rows, PML rows stay zero
            out[x0:x1] = lu.solve(rl.ravel()).reshape(nl, ny)[ext:ext + (x1 - x0)]     # restricted: owned rows only
        return out.ravel()
    return A, M


def slab_rows(nx, slabs, r):
    base, extra = divmod(nx, slabs)
    x0 = r * base + min(r, extra)
    return x0, x0 + base + (1 if r < extra else 0)


def build_rank(omega, eps, dl, npml, L0, slabs, rank, overlap, npml_s):
    """What ONE rank holds (SlabOperator.setup_schwarz): its rows, the subdomain's rows and the subdomain's factors."""
    nx, ny = eps.shape
    gis = orc.pml_inverse_factors(omega, L0, (nx, ny), npml, dl)
    ext = overlap + npml_s
    x0, x1 = slab_rows(nx, slabs, rank)
    rows = np.arange(x0 - ext, x1 + ext) % nx
    isxf, isxb = subdomain_sfactors(gis, rows, npml_s, dl, omega, L0)
    Al = orc.planes_to_csr(ez_planes(omega, eps[rows], dl, L0, isxf, isxb, gis[2], gis[3]))
    return dict(x0=x0, x1=x1, rows=rows, lu=spl.splu(sp.csc_matrix(Al)), ny=ny)


def rank_apply(sub, r_owned, comm, rank, world, overlap, npml_s):
    """z = M^-1 r on this rank's rows with the message pattern of csrc/krylov.cu schwarz_apply: the first / last
    ``overlap`` owned rows go to the lower / upper neighbour (periodic in the rank index), theirs fill the overlap
    rows of the zero-extended subdomain right-hand side, only the owned rows of the local solution are kept.
    ``comm`` has sendrecv(send_array, dst, recv_shape, src) (tests/test_dist_cpu.GlooComm)."""
    ny, nown, ext = sub["ny"], sub["x1"] - sub["x0"], overlap + npml_s
    rl = np.zeros((len(sub["rows"]), ny), complex)
    rl[ext:ext + nown] = r_owned
    if overlap > 0:
        lower, upper = (rank - 1) % world, (rank + 1) % world
        if world == 1:
            hi, lo = r_owned[:overlap], r_owned[-overlap:]
        else:
            hi = comm.sendrecv(r_owned[:overlap], lower, (overlap, ny), upper)
            lo = comm.sendrecv(r_owned[-overlap:], upper, (overlap, ny), lower)
        rl[npml_s:ext] = lo
        rl[ext + nown:ext + nown + overlap] = hi
    return sub["lu"].solve(rl.ravel()).reshape(-1, ny)[ext:ext + nown]


def solve(omega, eps, dl, npml, L0, b, slabs, overlap=4, npml_s=12, tol=1e-10, maxiter=2000, method="gmres", restart=80):
    """Right-preconditioned GMRES(restart) (what the slab path uses: csrc/krylov.cu krylov_gmres) or BiCGSTAB; returns
    (x, iterations, true relative residual, A).  scipy preconditions from the left; for the iteration COUNT of this
    study that makes no difference worth modelling."""
    A, M = build(omega, eps, dl, npml, L0, slabs, overlap, npml_s)
    count = [0]

    def cb(_):
        count[0] += 1
    Mop = spl.LinearOperator(A.shape, M, dtype=complex)
    if method == "gmres":
        x, info = spl.gmres(A, b.ravel(), M=Mop, rtol=tol, atol=0.0, restart=restart, maxiter=max(1, maxiter // restart),
                            callback=cb, callback_type="pr_norm")
    else:
        x, info = spl.bicgstab(A, b.ravel(), M=Mop, rtol=tol, atol=0.0, maxiter=maxiter, callback=cb)
    relres = np.linalg.norm(A @ x - b.ravel()) / np.linalg.norm(b)
    return x.reshape(eps.shape), count[0], relres, A


def device_eps(nx, ny):
    e = np.full((nx, ny), 2.1)
    c = ny // 2
    e[:, c - 10:c + 10] = 12.0
    e[nx // 2 - nx // 8:nx // 2 + nx // 8, c + 14:c + 34] = 12.0
    return e


if __name__ == "__main__":
    a = sys.argv[1:]
    nx, ny, P, ov, ns = (int(v) for v in a[:5])
    kind = a[5] if len(a) > 5 else "device"
    omega, dl, L0, npml = 2 * np.pi * 200e12, 0.02, 1e-6, [15, 15]
    if kind == "rods":
        sys.path.insert(0, ".")
        from bench import synthetic_eps
        eps = synthetic_eps(max(nx, ny))[:nx, :ny]
    else:
        eps = device_eps(nx, ny) if kind == "device" else np.ones((nx, ny))
    b = np.zeros((nx, ny), complex)
    b[nx // 5, ny // 2] = 1j * omega
    t = time.time()
    _, its, rr, _ = solve(omega, eps, dl, npml, L0, b, P, ov, ns)
    print(f"{kind} {nx}x{ny} slabs={P} overlap={ov} npml_sub={ns}: gmres iterations={its} relres={rr:.1e} ({time.time() - t:.0f}s)")
