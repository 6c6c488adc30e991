"""numpy model of the GPU direct solver: executes an ndplan exactly the way the CUDA kernels do
(padded batched fronts, blocked Gauss-Jordan sweep, level-by-level forward/backward solve).
Used to validate the plan on CPU and as a line-by-line reference for the kernels."""
import numpy as np


def sweep(F, k, tile=32):
    """In-place blocked Gauss-Jordan sweep of the leading k pivots of every front in the batch.
    Afterwards F = [[Z, Z*F_ER], [-F_RE*Z, S]] with Z = F_EE^-1, S the Schur complement."""
    nb, n, _ = F.shape
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


def factor(levels, planes, nx, ny, tile=32):
    c0, cxm, cxp, cym, cyp = [p.reshape(-1) for p in planes]
    store = []
    S_prev = None
    for lv in levels:
        F = np.zeros((lv.nb, lv.nmax, lv.nmax), dtype=np.complex128)
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
                    nr = ((x + 1) % nx) * ny + y
                    nu = x * ny + (y + 1) % ny
                    F[b, s, s] += c0[node]
                    F[b, s, r] += cxp[node]
                    F[b, r, s] += cxm[nr]
                    F[b, s, u] += cyp[node]
                    F[b, u, s] += cym[nu]
        else:
            mc = lv.child_mmax
            for b in range(lv.nb):
                c = lv.cls[b]
                for ch, cmap in ((lv.ch1[b], lv.c1map[c]), (lv.ch2[b], lv.c2map[c])):
                    idx = cmap[:mc]
                    ok = idx >= 0
                    ii = idx[ok]
                    F[b][np.ix_(ii, ii)] += S_prev[ch][np.ix_(np.where(ok)[0], np.where(ok)[0])]
        sweep(F, lv.kmax, tile)
        k = lv.kmax
        store.append((F[:, :k, :].copy(), F[:, k:, :k].copy()))      # [Z | X], -W
        S_prev = F[:, k:, k:].copy()
    return store


def solve(levels, store, b, nx, ny):
    b = np.asarray(b, dtype=np.complex128).reshape(-1)
    ring_prev = None
    ysave = []
    for lv, (EZX, RW) in zip(levels, store):
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
        ring_prev = f[:, k:] + np.einsum('bmk,bk->bm', RW, fe)
        ysave.append(np.einsum('bkj,bj->bk', EZX[:, :, :k], fe))
    # backward
    u_parent = None
    out = np.zeros(nx * ny, dtype=np.complex128)
    for li in range(len(levels) - 1, -1, -1):
        lv = levels[li]
        EZX, _ = store[li]
        k = lv.kmax
        u = np.zeros((lv.nb, lv.nmax), dtype=np.complex128)
        if li < len(levels) - 1:
            par = levels[li + 1]
            for pb in range(par.nb):
                c = par.cls[pb]
                for ch, cmap in ((par.ch1[pb], par.c1map[c]), (par.ch2[pb], par.c2map[c])):
                    idx = cmap[:par.child_mmax]
                    ok = idx >= 0
                    u[ch, k + np.where(ok)[0]] = u_parent[pb, idx[ok]]
        u[:, :k] = ysave[li] - np.einsum('bkm,bm->bk', EZX[:, :, k:], u[:, k:])
        u_parent = u
        if lv.kind == "leaf":
            for i in range(lv.nb):
                c = lv.cls[i]
                for s in range(lv.nmax):
                    if lv.slot_right[c, s] >= 0:
                        x = (lv.x0[i] + lv.slot_lx[c, s]) % nx
                        y = (lv.y0[i] + lv.slot_ly[c, s]) % ny
                        out[x * ny + y] = u[i, s]
    return out.reshape(nx, ny)
