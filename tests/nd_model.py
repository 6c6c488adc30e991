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


def gj_inverse_pingpong(E, tile=64):
    """The block Gauss-Jordan inversion as csrc/direct.cu gj_invert_batch runs it: per pivot tile p one tile inverse
    P = M_pp^-1 and one OUT-OF-PLACE update of every tile (src -> dst, buffers swapped after each step):
        dst_pp = P     dst_pj = P src_pj     dst_ip = -src_ip P     dst_ij = src_ij - src_ip (P src_pj)
    The last tile may be ragged."""
    src = np.array(E, dtype=complex)
    nb, n, _ = src.shape
    nt = (n + tile - 1) // tile
    sl = [slice(t * tile, min((t + 1) * tile, n)) for t in range(nt)]
    for p in range(nt):
        dst = np.empty_like(src)
        P = np.linalg.inv(src[:, sl[p], sl[p]])
        for i in range(nt):
            for j in range(nt):
                if i == p and j == p:
                    dst[:, sl[p], sl[p]] = P
                elif i == p:
                    dst[:, sl[p], sl[j]] = P @ src[:, sl[p], sl[j]]
                elif j == p:
                    dst[:, sl[i], sl[p]] = -src[:, sl[i], sl[p]] @ P
                else:
                    dst[:, sl[i], sl[j]] = src[:, sl[i], sl[j]] - src[:, sl[i], sl[p]] @ (P @ src[:, sl[p], sl[j]])
        src = dst
    return src


def _lower_to_full(L):
    """Symmetric matrix from its lower triangle (the upper one of the input is ignored)."""
    T = np.tril(L)
    return T + np.transpose(np.tril(L, -1), (0, 2, 1))


def factor(levels, planes, nx, ny, dscale, tile=32, comm=None):
    """``comm`` (send(array, dst) / recv(shape, src) / allreduce(array)) runs a plan produced by
    ndplan.shard_plan: the exchanges sit exactly where nd_factor / nd_solve_chunk put them.
    A plan with distributed fronts (``levels.dist``) continues in ``dist_factor`` after the local levels."""
    c0, cxm, cxp, cym, cyp = [p.reshape(-1) for p in planes]
    store = []
    S_prev = None
    dist = list(getattr(levels, "dist", ()))
    nlocal = dist[0].level0 if dist else len(levels)
    for lv in levels[:nlocal]:
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
    if dist:
        store += [None] * (len(levels) - nlocal)
        store.append(dist_factor(dist, levels, S_prev[0], comm, tile))    # one extra entry: the distributed fronts
    return store


# ------------------------------------------------------------------------------------------
# distributed fronts (ndplan.DistFront), executed the way csrc/distfront.cuh does: block rows dealt to the group,
# personalised all-to-all assembly, owner inverts the pivot block and broadcasts it, panel all-gather, every rank
# updates its own block rows; replicated substitution vectors.
# ------------------------------------------------------------------------------------------
class _DistState:
    pass


def _group_bcast(comm, df, me, arr, root, shape):
    if me == root:
        for p in range(df.gsize):
            if p != root:
                comm.send(arr, df.gbase + p)
        return arr
    return comm.recv(shape, df.gbase + root)


def dist_factor(dist, levels, S_local, comm, tile):
    """S_local: the Schur block (lower) of this rank's single front on the level below the first distributed front."""
    rank = comm.dist.get_rank()
    states = []
    child_rows, child_S = None, None          # my rows (ring indices) of the child's Schur block, and the rows themselves
    for j, df in enumerate(dist):
        me = rank - df.gbase
        g, n = df.gsize, df.n
        cidx = 1 if me >= g // 2 else 0
        inv = df.inv[cidx]
        if j == 0:
            child_rows = np.arange(S_local.shape[0])
            child_S = _lower_to_full(S_local[None])[0]
        row_of = {a: i for i, a in enumerate(child_rows)}
        rows_of = [np.concatenate([np.arange(df.bstart[b], df.bstart[b + 1]) for b in range(len(df.bowner)) if df.bowner[b] == o]
                                  or [np.zeros(0, dtype=int)]).astype(int) for o in range(g)]
        mine = rows_of[me]

        def pack(dst_rows):
            out = np.zeros((len(dst_rows), n), dtype=np.complex128)
            for lp, p in enumerate(dst_rows):
                a = inv[p]
                if a < 0:
                    continue
                for q in range(p + 1):
                    b = inv[q]
                    if b < 0:
                        continue
                    hi, lo = max(a, b), min(a, b)
                    if hi in row_of:                       # the child entry (hi, lo) lives on the rank that owns row hi
                        out[lp, q] = child_S[row_of[hi], lo]
            return out

        F = pack(mine)
        for t in range(1, g):
            d, src = (me + t) % g, (me - t) % g
            sbuf = pack(rows_of[d])
            rb = comm.sendrecv(sbuf if len(rows_of[d]) else None, df.gbase + d,
                               F.shape if len(mine) else None, df.gbase + src)
            if rb is not None:
                F += rb
        st = _DistState()
        st.df, st.me, st.cidx, st.mine = df, me, cidx, mine
        st.Einv, st.G, st.below = [], [], []
        for sidx in range(df.nsteps):
            col0, b1 = int(df.bstart[sidx]), int(df.bstart[sidx + 1])
            k, owner = b1 - col0, int(df.bowner[sidx])
            Einv = None
            if me == owner:
                loc = np.flatnonzero((mine >= col0) & (mine < b1))
                E = F[loc][:, col0:b1]
                Einv = gj_inverse(_lower_to_full(E[None]), tile)[0]
            Einv = _group_bcast(comm, df, me, Einv, owner, (k, k))
            below = np.flatnonzero(mine >= b1)
            FRE = F[below][:, col0:b1]
            G = FRE @ Einv
            st.Einv.append(Einv)
            st.G.append(G)
            st.below.append(below)
            if b1 == n:
                continue
            # all-gather of the F_RE panel in global row order
            panel = np.zeros((n - b1, k), dtype=np.complex128)
            for o in range(g):
                rows_o = rows_of[o][rows_of[o] >= b1]
                if o == me:
                    panel[rows_o - b1] = FRE
                    for p in range(g):
                        if p != me and len(rows_o):
                            comm.send(FRE, df.gbase + p)
                elif len(rows_o):
                    panel[rows_o - b1] = comm.recv((len(rows_o), k), df.gbase + o)
            upd = G @ panel.T                               # my rows x all remaining columns
            for i, li in enumerate(below):
                p = mine[li]
                F[li, b1:p + 1] -= upd[i, :p + 1 - b1]     # lower part only
        states.append(st)
        ring = np.flatnonzero(mine >= df.kfull)
        child_rows = mine[ring] - df.kfull
        child_S = F[ring][:, df.kfull:]
        # the parent needs full rows of the symmetric block, but only owns the lower part of each: entry (hi, lo) is
        # looked up on the rank that owns row hi, which is exactly what pack() does -- so the lower part suffices.
    return states


def solve(levels, store, b, nx, ny, dscale, comm=None):
    b = np.asarray(b, dtype=np.complex128).reshape(-1) * dscale
    ring_prev = None
    ysave = []
    dist = list(getattr(levels, "dist", ()))
    nlocal = dist[0].level0 if dist else len(levels)
    for lv, fac in zip(levels[:nlocal], store):
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
    top_ring = None
    if dist:
        top_ring = dist_solve(store[-1], ring_prev[0], comm)      # ring solution of my last local front
    # backward
    u_parent = None
    out = np.zeros(nx * ny, dtype=np.complex128)
    for li in range(nlocal - 1, -1, -1):
        lv = levels[li]
        k = lv.kmax
        par = levels[li + 1] if li < nlocal - 1 else None
        remote_child = par is not None and getattr(par, "recv_from", -1) >= 0
        u = np.zeros((lv.nb + (1 if remote_child else 0), lv.nmax), dtype=np.complex128)
        if par is None and top_ring is not None:
            u[0, k:k + len(top_ring)] = top_ring
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


def dist_solve(states, my_ring, comm):
    """Forward and backward substitution through the distributed fronts; returns the ring solution of this rank's
    local child front.  Front vectors are replicated inside a group, as in csrc/distfront.cuh."""
    vecs = []
    for st in states:
        df, me, g, n = st.df, st.me, st.df.gsize, st.df.n
        partner = df.gbase + (me + g // 2) % g
        mc_mine, mc_other = df.mc[st.cidx], df.mc[1 - st.cidx]
        my_ring = np.asarray(my_ring)[:mc_mine]
        if me < g // 2:
            comm.send(my_ring, partner)
            other = comm.recv((mc_other,), partner)
        else:
            other = comm.recv((mc_other,), partner)
            comm.send(my_ring, partner)
        rings = (my_ring, other) if st.cidx == 0 else (other, my_ring)
        f = np.zeros(n, dtype=np.complex128)
        for c in (0, 1):
            ok = df.inv[c] >= 0
            f[ok] += rings[c][df.inv[c][ok]]
        yE = np.zeros(df.kfull, dtype=np.complex128)
        for sidx in range(df.nsteps):
            col0, b1 = int(df.bstart[sidx]), int(df.bstart[sidx + 1])
            fE = _group_bcast(comm, df, me, f[col0:b1].copy(), int(df.bowner[sidx]), (b1 - col0,))
            f[col0:b1] = fE
            yE[col0:b1] = st.Einv[sidx] @ fE
            rows = st.mine[st.below[sidx]]
            f[rows] -= st.G[sidx] @ fE
        # replicate the ring: every block from its owner
        for b in range(df.nsteps, len(df.bowner)):
            sl = slice(int(df.bstart[b]), int(df.bstart[b + 1]))
            f[sl] = _group_bcast(comm, df, me, f[sl].copy(), int(df.bowner[b]), (sl.stop - sl.start,))
        vecs.append((f, yE))
        my_ring = f[df.kfull:]
    # backward, root first
    for j in range(len(states) - 1, -1, -1):
        st = states[j]
        df, me, g, n = st.df, st.me, st.df.gsize, st.df.n
        u, yE = vecs[j]
        if j + 1 < len(states):
            par = states[j + 1]
            cmap = np.zeros(par.df.mc[par.cidx], dtype=int)
            inv = par.df.inv[par.cidx]
            cmap[inv[inv >= 0]] = np.flatnonzero(inv >= 0)
            u[df.kfull:] = vecs[j + 1][0][cmap]
        for sidx in range(df.nsteps - 1, -1, -1):
            col0, b1 = int(df.bstart[sidx]), int(df.bstart[sidx + 1])
            part = st.G[sidx].T @ u[st.mine[st.below[sidx]]] if len(st.below[sidx]) else np.zeros(b1 - col0, dtype=np.complex128)
            parts = [None] * g
            parts[me] = part
            for p in range(g):                          # all-gather, summed in rank order on every rank
                if p == me:
                    for q in range(g):
                        if q != me:
                            comm.send(part, df.gbase + q)
                else:
                    parts[p] = comm.recv((b1 - col0,), df.gbase + p)
            u[col0:b1] = yE[col0:b1] - sum(parts[1:], parts[0])
    st = states[0]
    inv = st.df.inv[st.cidx]
    cmap = np.zeros(st.df.mc[st.cidx], dtype=int)
    cmap[inv[inv >= 0]] = np.flatnonzero(inv >= 0)
    return vecs[0][0][cmap]
