"""numpy model of the GPU direct solver: executes an ndplan exactly the way the CUDA kernels do
(padded batched fronts, symmetric front algebra on the row-scaled operator, blocked Gauss-Jordan
inversion of the eliminated block, level-by-level forward/backward solve).
Used to validate the plan on CPU and as a line-by-line reference for the kernels.

Row scaling: D A with D = diag(sxf[ix] * syf[iy]) is complex SYMMETRIC (the forward stretch factors
are the only asymmetry of the sc-PML operator), so only the lower triangle of a front is kept and
    E^-1,   G = F_RE E^-1,   S = F_RR - G F_RE^T
is all a level computes; the substitution uses  u_E = E^-1 f_E - G^T u_R."""
import numpy as np


def row_scale(isxf, isyf):
    """d[node] = sxf[ix] * syf[iy] = 1 / (isxf[ix] * isyf[iy])"""
    return (1.0 / (np.asarray(isxf)[:, None] * np.asarray(isyf)[None, :])).reshape(-1)


def gj_inverse(E, tile=32):
    """Blocked Gauss-Jordan inversion of every matrix in the batch (pivoting inside tiles only)."""
    F = E.copy()
    nb, k, _ = F.shape
    for j0 in range(0, k, tile):
        J = slice(j0, min(j0 + tile, k))
        P = np.linalg.inv(F[:, J, J])
        C = F[:, :, J].copy()
        C[:, J, :] = 0
        F[:, :, J] = 0
        tw = J.stop - J.start
        F[:, J, J] = np.eye(tw)
        R = P @ F[:, J, :]
        F[:, J, :] = R
        F -= C @ R
    return F


def _lower_to_full(L):
    """Symmetric matrix from its lower triangle (the upper one of the input is ignored)."""
    T = np.tril(L)
    return T + np.transpose(np.tril(L, -1), (0, 2, 1))


def factor(levels, planes, nx, ny, dscale, tile=32, comm=None):
    """``comm`` (send(array, dst) / recv(shape, src) / allreduce(array)) runs a plan produced by
    ndplan.shard_plan: the exchanges sit exactly where nd_factor / nd_solve_chunk put them."""
    c0, cxm, cxp, cym, cyp = [p.reshape(-1) for p in planes]
    store = []
    S_prev = None
    for lv in levels:
        if getattr(lv, "send_to", -1) >= 0:
            comm.send(S_prev[0], lv.send_to)
        if getattr(lv, "recv_from", -1) >= 0:
            S_prev = np.concatenate([S_prev[:1], comm.recv(S_prev[0].shape, lv.recv_from)[None]])
        if lv.nb == 0:
            store.append(None)
            S_prev = np.zeros((0, lv.mmax, lv.mmax), dtype=np.complex128)
            continue
        F = np.zeros((lv.nb, lv.nmax, lv.nmax), dtype=np.complex128)       # lower triangle only
        for b in range(lv.nb):
            c = lv.cls[b]
            for s in range(lv.k_cls[c], lv.kmax):          # identity on padded pivots
                F[b, s, s] = 1
        if lv.kind == "leaf":
            for b in range(lv.nb):
                c = lv.cls[b]
                for s in range(lv.nmax):
                    r, u = lv.slot_right[c, s], lv.slot_up[c, s]
                    if r < 0:
                        continue
                    x = (lv.x0[b] + lv.slot_lx[c, s]) % nx
                    y = (lv.y0[b] + lv.slot_ly[c, s]) % ny
                    node = x * ny + y
                    F[b, s, s] += c0[node] * dscale[node]
                    F[b, max(s, r), min(s, r)] += cxp[node] * dscale[node]
                    F[b, max(s, u), min(s, u)] += cyp[node] * dscale[node]
        else:
            mc = lv.child_mmax
            for b in range(lv.nb):
                c = lv.cls[b]
                for ch, cmap in ((lv.ch1[b], lv.c1map[c]), (lv.ch2[b], lv.c2map[c])):
                    idx = cmap[:mc]
                    ok = np.where(idx >= 0)[0]
                    for a in ok:
                        for bb in ok:
                            p, q = idx[a], idx[bb]
                            if p >= q:
                                F[b, p, q] += S_prev[ch][max(a, bb), min(a, bb)]
        k = lv.kmax
        Einv = gj_inverse(_lower_to_full(F[:, :k, :k]), tile)
        FRE = F[:, k:, :k]
        G = FRE @ Einv
        S_prev = np.tril(F[:, k:, k:] - G @ np.transpose(FRE, (0, 2, 1)))   # lower triangle only
        store.append((Einv, G))
    return store


def solve(levels, store, b, nx, ny, dscale, comm=None):
    b = np.asarray(b, dtype=np.complex128).reshape(-1) * dscale
    ring_prev = None
    ysave = []
    for lv, fac in zip(levels, store):
        if getattr(lv, "send_to", -1) >= 0:
            comm.send(ring_prev[0], lv.send_to)
        if getattr(lv, "recv_from", -1) >= 0:
            ring_prev = np.concatenate([ring_prev[:1], comm.recv(ring_prev[0].shape, lv.recv_from)[None]])
        if lv.nb == 0:
            ysave.append(None)
            ring_prev = np.zeros((0, lv.mmax), dtype=np.complex128)
            continue
        Einv, G = fac
        f = np.zeros((lv.nb, lv.nmax), dtype=np.complex128)
        if lv.kind == "leaf":
            for i in range(lv.nb):
                c = lv.cls[i]
                for s in range(lv.nmax):
                    if lv.slot_right[c, s] >= 0:
                        x = (lv.x0[i] + lv.slot_lx[c, s]) % nx
                        y = (lv.y0[i] + lv.slot_ly[c, s]) % ny
                        f[i, s] = b[x * ny + y]
        else:
            for i in range(lv.nb):
                c = lv.cls[i]
                for ch, cmap in ((lv.ch1[i], lv.c1map[c]), (lv.ch2[i], lv.c2map[c])):
                    idx = cmap[:lv.child_mmax]
                    ok = idx >= 0
                    np.add.at(f[i], idx[ok], ring_prev[ch][ok])
        k = lv.kmax
        fe = f[:, :k]
        ring_prev = f[:, k:] - np.einsum('bmk,bk->bm', G, fe)
        ysave.append(np.einsum('bkj,bj->bk', Einv, fe))
    # backward
    u_parent = None
    out = np.zeros(nx * ny, dtype=np.complex128)
    for li in range(len(levels) - 1, -1, -1):
        lv = levels[li]
        k = lv.kmax
        par = levels[li + 1] if li < len(levels) - 1 else None
        remote_child = par is not None and getattr(par, "recv_from", -1) >= 0
        u = np.zeros((lv.nb + (1 if remote_child else 0), lv.nmax), dtype=np.complex128)
        if par is not None:
            for pb in range(par.nb):
                c = par.cls[pb]
                for ch, cmap in ((par.ch1[pb], par.c1map[c]), (par.ch2[pb], par.c2map[c])):
                    idx = cmap[:par.child_mmax]
                    ok = idx >= 0
                    u[ch, k + np.where(ok)[0]] = u_parent[pb, idx[ok]]
            if remote_child:                       # the second child's ring solution goes back to its rank
                comm.send(u[1], par.recv_from)
                u = u[:1]
            if getattr(par, "send_to", -1) >= 0:
                u[0] = comm.recv(u[0].shape, par.send_to)
        if lv.nb == 0:
            u_parent = u
            continue
        G = store[li][1]
        u[:, :k] = ysave[li] - np.einsum('bmk,bm->bk', G, u[:, k:])
        u_parent = u
        if lv.kind == "leaf":
            for i in range(lv.nb):
                c = lv.cls[i]
                for s in range(lv.nmax):
                    if lv.slot_right[c, s] >= 0:
                        x = (lv.x0[i] + lv.slot_lx[c, s]) % nx
                        y = (lv.y0[i] + lv.slot_ly[c, s]) % ny
                        out[x * ny + y] = u[i, s]
    if comm is not None:
        out = comm.allreduce(out)                  # every cell is written by exactly one rank
    return out.reshape(nx, ny)
