"""Host-side plan for the structured direct solver (nested dissection on the periodic grid).

The reference hands ``A x = b`` to a general sparse LU (linalg.py:123-149, MKL Pardiso or
SuperLU), which re-derives an elimination order from the sparsity pattern on every call.
Here the pattern is known a priori -- a 5-point stencil on an Nx x Ny torus (the derivative
operators wrap, derivatives.py:21-33) -- so the elimination tree is built geometrically:

* the torus is cut by 2^ax vertical and 2^ay horizontal grid lines into leaf boxes that SHARE
  their boundary lines;
* every stencil entry is owned by exactly one leaf (half-open ownership, see ``leaf tables``),
  so a box's local matrix is "fully summed" on its interior and partial on its ring;
* a leaf eliminates its interior; each higher level merges boxes pairwise along x or y and
  eliminates the shared line(s); the last merge along an axis also closes the periodic wrap.

All boxes of a level have one of <= 4 shapes (interval lengths differ by at most one), so a
level is ONE batch of equally padded dense fronts  [E | R]  (E = eliminated, R = ring kept).
The plan only holds small per-shape index tables plus per-box (shape id, children, origin);
it contains no matrix values and is reusable for every eps_r / omega on the same grid shape.

This module is pure index bookkeeping (numpy, no arithmetic on field values).
"""
import numpy as np

WMIN = 3          # smallest leaf interval (cells); leaves are WMIN..2*WMIN-1 cells wide
SPLIT_MIN = 1000  # separators with at least this many nodes are eliminated in several steps
SPLIT_PARTS = -512   # > 0: that many steps; < 0: pieces of about -SPLIT_PARTS nodes (at most SPLIT_MAX_STEPS steps)
SPLIT_MAX_STEPS = 8       # single GPU; sharded trees use 16 (core.get_plan): on a distributed front the pivot-block inversions
                          # are the critical path, and sixteen 512-blocks invert faster than eight 1024-blocks (116 -> 105 ms on 8 GPUs)


def _parts(k, p):
    """Cut points of k items into p nearly equal consecutive pieces: [s_0=0, s_1, ..., s_p=k]."""
    base, extra = divmod(k, p)
    cuts = [0]
    for i in range(p):
        cuts.append(cuts[-1] + base + (1 if i < extra else 0))
    return cuts


def _halve(n, depth):
    """Interval lengths after ``depth`` recursive floor/ceil halvings of n (all within 1)."""
    sizes = [n]
    for _ in range(depth):
        nxt = []
        for s in sizes:
            nxt += [s // 2, s - s // 2]
        sizes = nxt
    return np.array(sizes, dtype=np.int64)


def _depth_for(n, wmin):
    if n < 4:
        raise ValueError("direct solver needs at least 4 cells per axis, got {}".format(n))
    d = 1
    while (n >> (d + 1)) >= wmin:
        d += 1
    return d


def _ring(w, h, xclosed, yclosed, nx, ny):
    """Ordered ring node list (lx, ly) of a box w x h cells (closed axes span the whole circle)."""
    pts = []
    xs = range(nx) if xclosed else range(w + 1)
    ys = range(ny) if yclosed else range(h + 1)
    if not yclosed:
        pts += [(x, 0) for x in xs]
        pts += [(x, h) for x in xs]
    if not xclosed:
        yy = ys if yclosed else range(1, h)
        pts += [(0, y) for y in yy]
        pts += [(w, y) for y in yy]
    return pts


def _class_table(sid, xs, ys):
    """Shape class of every leaf box (i, j) -> sid[(xs[i], ys[j])], without a Python loop over the boxes."""
    ux, uy = sorted(set(xs.tolist())), sorted(set(ys.tolist()))
    table = np.array([[sid[(w, h)] for h in uy] for w in ux], dtype=np.int32)
    xi = np.array([ux.index(w) for w in xs.tolist()])
    yi = np.array([uy.index(h) for h in ys.tolist()])
    return table[xi[:, None], yi[None, :]].ravel()


class Level:
    """One batch of fronts.  Attribute names are the ones the C ABI takes (include/fdfd_b200.h)."""
    pass


def build_plan(nx, ny, wmin=WMIN, split_min=None, split_parts=None, split_max_steps=None):
    """Return the list of levels, leaves first, root last.

    A merge whose separator has >= split_min nodes is emitted as a CHAIN of levels that eliminate the
    separator piece by piece (first level: the merge proper, eliminating piece 0 with the rest of the
    separator kept in the ring; then ``chain`` levels with a single child each).  Eliminating a block of k
    nodes in p pieces is blocked LDL^T in disguise: the explicit inverse of the whole k x k block and the
    full m x k coupling product are never formed (k^3/2 -> ~k^3/6 (1 + 1/p ...), m k^2 -> m k^2 (p+1)/2p),
    while the solve phase still streams plain matrix-vector products."""
    split_min = SPLIT_MIN if split_min is None else split_min
    split_parts = SPLIT_PARTS if split_parts is None else split_parts
    split_max_steps = SPLIT_MAX_STEPS if split_max_steps is None else split_max_steps
    ax, ay = _depth_for(nx, wmin), _depth_for(ny, wmin)
    levels = []

    # ---------------- leaves ----------------
    xs, ys = _halve(nx, ax), _halve(ny, ay)
    x0 = np.concatenate([[0], np.cumsum(xs)[:-1]])
    y0 = np.concatenate([[0], np.cumsum(ys)[:-1]])
    px, py = len(xs), len(ys)
    shapes = sorted({(int(w), int(h)) for w in set(xs) for h in set(ys)})
    sid = {s: i for i, s in enumerate(shapes)}
    lv = Level()
    lv.kind = "leaf"
    lv.px, lv.py, lv.nb = px, py, px * py
    W, H = np.meshgrid(xs, ys, indexing="ij")
    X0, Y0 = np.meshgrid(x0, y0, indexing="ij")
    lv.cls = _class_table(sid, xs, ys)
    lv.x0 = X0.ravel().astype(np.int32)
    lv.y0 = Y0.ravel().astype(np.int32)
    lv.k_cls = np.array([(w - 1) * (h - 1) for w, h in shapes], dtype=np.int32)
    lv.m_cls = np.array([2 * w + 2 * h for w, h in shapes], dtype=np.int32)
    lv.kmax, lv.mmax = int(lv.k_cls.max()), int(lv.m_cls.max())
    lv.nmax = lv.kmax + lv.mmax
    ncls = len(shapes)
    lv.ncls = ncls
    # leaf tables: per slot local coords, ownership and the slots of the +x / +y neighbours
    lv.slot_lx = np.full((ncls, lv.nmax), -1, dtype=np.int32)
    lv.slot_ly = np.full((ncls, lv.nmax), -1, dtype=np.int32)
    lv.slot_right = np.full((ncls, lv.nmax), -1, dtype=np.int32)   # -1: slot does not own entries
    lv.slot_up = np.full((ncls, lv.nmax), -1, dtype=np.int32)
    rings = {}
    for (w, h), c in sid.items():
        interior = [(x, y) for x in range(1, w) for y in range(1, h)]
        ring = _ring(w, h, False, False, nx, ny)
        rings[(w, h)] = ring
        pos = {p: i for i, p in enumerate(interior)}
        pos.update({p: lv.kmax + i for i, p in enumerate(ring)})
        assert len(pos) == (w + 1) * (h + 1)
        for (x, y), s in pos.items():
            lv.slot_lx[c, s], lv.slot_ly[c, s] = x, y
            if x < w and y < h:                      # half-open ownership
                lv.slot_right[c, s] = pos[(x + 1, y)]
                lv.slot_up[c, s] = pos[(x, y + 1)]
    levels.append(lv)

    # ---------------- merges ----------------
    dx, dy = ax, ay
    cur_shapes, cur_sid, cur_rings = shapes, sid, rings      # children description
    cur_xs, cur_ys = xs, ys
    xclosed = yclosed = False
    while dx > 0 or dy > 0:
        wx = nx if xclosed else int(cur_xs.max())
        wy = ny if yclosed else int(cur_ys.max())
        axis = 0 if (dx > 0 and (dy == 0 or wx <= wy)) else 1
        child_px, child_py = len(cur_xs), len(cur_ys)
        if axis == 0:
            dx -= 1
            new_xs, new_ys = cur_xs[0::2] + cur_xs[1::2], cur_ys
            closing = dx == 0
        else:
            dy -= 1
            new_xs, new_ys = cur_xs, cur_ys[0::2] + cur_ys[1::2]
            closing = dy == 0
        npx, npy = len(new_xs), len(new_ys)
        new_xclosed = xclosed or (axis == 0 and closing)
        new_yclosed = yclosed or (axis == 1 and closing)

        lv = Level()
        lv.kind = "merge"
        lv.axis = axis
        lv.px, lv.py, lv.nb = npx, npy, npx * npy
        I, J = np.meshgrid(np.arange(npx), np.arange(npy), indexing="ij")
        # shape key of front (i, j): (first child's extent, second child's extent, extent along the other axis)
        if axis == 0:
            c1 = (2 * I) * child_py + J
            c2 = (2 * I + 1) * child_py + J
            pairs, other = list(zip(cur_xs[0::2].tolist(), cur_xs[1::2].tolist())), cur_ys.tolist()
        else:
            c1 = I * child_py + 2 * J
            c2 = I * child_py + 2 * J + 1
            pairs, other = list(zip(cur_ys[0::2].tolist(), cur_ys[1::2].tolist())), cur_xs.tolist()
        lv.ch1 = c1.ravel().astype(np.int32)
        lv.ch2 = c2.ravel().astype(np.int32)
        pshapes = sorted({(a, b, o) for a, b in set(pairs) for o in set(other)})
        psid = {s: i for i, s in enumerate(pshapes)}
        upairs, uother = sorted(set(pairs)), sorted(set(other))
        table = np.array([[psid[(a, b, o)] for o in uother] for a, b in upairs], dtype=np.int32)
        pcode = np.array([upairs.index(pr) for pr in pairs])
        ocode = np.array([uother.index(o) for o in other])
        lv.cls = (table[pcode[:, None], ocode[None, :]] if axis == 0 else table[pcode[None, :], ocode[:, None]]).ravel()
        lv.ncls = len(pshapes)

        fronts = []
        new_rings, new_sid_shapes = {}, []
        for (a1, a2, other) in pshapes:
            if axis == 0:
                s1, s2 = (a1, other), (a2, other)
                pw, ph = a1 + a2, other
            else:
                s1, s2 = (other, a1), (other, a2)
                pw, ph = other, a1 + a2
            r1, r2 = cur_rings[s1], cur_rings[s2]
            if axis == 0:
                mod = nx if new_xclosed else None
                p1 = [(x, y) for x, y in r1]
                p2 = [((x + a1) % mod if mod else x + a1, y) for x, y in r2]
            else:
                mod = ny if new_yclosed else None
                p1 = [(x, y) for x, y in r1]
                p2 = [(x, (y + a1) % mod if mod else y + a1) for x, y in r2]
            pring = _ring(pw, ph, new_xclosed, new_yclosed, nx, ny)
            pset = set(pring)
            union = set(p1) | set(p2)
            assert pset <= union
            elim = sorted(union - pset)
            fronts.append((elim, pring, p1, p2))
            new_rings[(pw, ph)] = pring
        kfull = max(len(f[0]) for f in fronts)
        want = split_parts if split_parts > 0 else max(1, min(split_max_steps, int(round(kfull / float(-split_parts)))))
        nparts = want if (want > 1 and kfull >= split_min and min(len(f[0]) for f in fronts) >= want) else 1
        # pieces are identical for every shape class except the last one, which absorbs the (<= 2 node) size
        # differences: the intermediate chain levels then have no padded pivots and are factorised in place
        base = _parts(min(len(f[0]) for f in fronts), nparts)
        cuts = [base[:-1] + [len(f[0])] for f in fronts]                  # per class
        child_mmax = levels[-1].mmax
        for part in range(nparts):
            if part > 0:
                prev = lv
                lv = Level()
                lv.kind = "merge"
                lv.chain = True
                lv.axis = axis
                lv.px, lv.py, lv.nb = prev.px, prev.py, prev.nb
                lv.cls, lv.ncls = prev.cls, prev.ncls
                lv.ch1 = np.arange(lv.nb, dtype=np.int32)
                lv.ch2 = lv.ch1
                child_mmax = prev.mmax
            else:
                lv.chain = False
            els = [f[0][cuts[c][part]:cuts[c][part + 1]] for c, f in enumerate(fronts)]       # eliminated now
            rings = [f[0][cuts[c][part + 1]:] + f[1] for c, f in enumerate(fronts)]            # kept for later
            lv.k_cls = np.array([len(e) for e in els], dtype=np.int32)
            lv.m_cls = np.array([len(r) for r in rings], dtype=np.int32)
            lv.kmax, lv.mmax = int(lv.k_cls.max()), int(lv.m_cls.max())
            lv.nmax = lv.kmax + lv.mmax
            lv.child_mmax = child_mmax
            lv.c1map = np.full((lv.ncls, max(child_mmax, 1)), -1, dtype=np.int32)
            lv.c2map = np.full((lv.ncls, max(child_mmax, 1)), -1, dtype=np.int32)
            for c, (elim, pring, p1, p2) in enumerate(fronts):
                pos = {p: i for i, p in enumerate(els[c])}
                pos.update({p: lv.kmax + i for i, p in enumerate(rings[c])})
                if part == 0:
                    lv.c1map[c, :len(p1)] = [pos[p] for p in p1]
                    lv.c2map[c, :len(p2)] = [pos[p] for p in p2]
                else:                                        # single child: the previous step's ring
                    cring = elim[cuts[c][part]:] + pring
                    lv.c1map[c, :len(cring)] = [pos[p] for p in cring]
            levels.append(lv)
        cur_rings = new_rings
        cur_xs, cur_ys = new_xs, new_ys
        xclosed, yclosed = new_xclosed, new_yclosed
    assert levels[-1].nb == 1 and levels[-1].mmax == 0
    return levels


def plan_stats(levels):
    """(stored factor entries, peak transient front entries, complex MACs of the blocked sweep)."""
    stored = sum(lv.nb * (lv.kmax * lv.nmax + lv.mmax * lv.kmax) for lv in levels)
    transient = max(lv.nb * lv.nmax * lv.nmax for lv in levels)
    macs = sum(lv.nb * lv.kmax * lv.nmax * lv.nmax for lv in levels)
    return stored, transient, macs


DIST_RB = 512        # ring rows of a distributed front are dealt to the ranks in blocks of about this many


class DistFront:
    """One front of a level shared by ``gsize`` ranks (ranks gbase .. gbase + gsize - 1), stored and factorised
    DISTRIBUTED: its rows are dealt to the ranks of the group in blocks (the pieces of its separator, then pieces
    of its ring), block j to group rank ``bowner[j]`` in serpentine order 0 1 .. g-1 g-1 .. 1 0 (the Schur update of
    a block row grows linearly with its index, so pairs j, 2g-1-j carry equal work).

    Slots are COMPACT (no padded pivots: a distributed front is a single front, there is no batch to keep
    uniform): [piece 0 | piece 1 | ... | piece nsteps-1 | ring].  ``inv[c][p]`` is the position of parent slot p in
    child c's ring, -1 if that child does not reach it.  Levels level0 .. level0 + nsteps - 1 of the plan are this
    front's elimination steps."""
    pass


def _dist_front(levels, l0, nchain, fidx, gbase, gsize, rb):
    head = levels[l0]
    c = int(head.cls[fidx])
    steps = [levels[l0 + s] for s in range(nchain + 1)]
    ks = [int(lv.k_cls[c]) for lv in steps]
    m = int(steps[-1].m_cls[c])
    n = sum(ks) + m
    assert n == int(head.k_cls[c]) + int(head.m_cls[c])
    df = DistFront()
    df.level0, df.nsteps, df.gbase, df.gsize, df.n, df.m = l0, nchain + 1, gbase, gsize, n, m
    df.kfull = sum(ks)
    cuts = [0]
    for k in ks:
        cuts.append(cuts[-1] + k)
    if m > 0:
        pieces = max(1, int(round(m / float(rb))))
        cuts += [df.kfull + v for v in _parts(m, pieces)[1:]]
    df.bstart = np.array(cuts, dtype=np.int32)
    nblk = len(cuts) - 1
    pos = np.arange(nblk) % (2 * gsize)
    df.bowner = np.where(pos < gsize, pos, 2 * gsize - 1 - pos).astype(np.int32)
    # child ring position -> compact parent slot: padded slot s of the head is s (s < k_0) or s - (kmax - k_0)
    child = levels[l0 - 1]
    df.mc, df.inv = [], []
    for ch, cmap in ((head.ch1, head.c1map), (head.ch2, head.c2map)):
        cc = int(child.cls[ch[fidx]])
        mc = int(child.m_cls[cc])
        slots = cmap[c, :mc].astype(np.int64)
        assert slots.min() >= 0 and not ((slots >= ks[0]) & (slots < head.kmax)).any()
        slots = np.where(slots >= head.kmax, slots - (head.kmax - ks[0]), slots)
        inv = np.full(n, -1, dtype=np.int32)
        inv[slots] = np.arange(mc, dtype=np.int32)
        df.mc.append(mc)
        df.inv.append(inv)
    return df


def shard_plan(levels, world, rank, distribute=True, rb=None):
    """This rank's part of the elimination tree when one grid is split over ``world`` GPUs.

    ``distribute`` (default): the fronts of the log2(world) shared levels are DistFronts -- every rank of the group
    below a front owns block rows of it and takes part in its factorisation and substitution; the returned list
    carries them as ``levels.dist`` (attribute of the returned list object) and the shared levels themselves are
    empty (nb = 0) on every rank.  ``distribute=False`` keeps the older scheme described next, in which a shared
    front lives on the lowest rank of its group.

    The top log2(world) merges of the tree cut the torus into ``world`` slabs/blocks; each rank
    owns the whole subtree below one of them (no communication there).  Above, a front belongs to
    the lowest rank of the group of ranks below it; its second child lives on the rank half a
    group further on, which hands its Schur block (factorisation), its ring right-hand side
    (forward solve) and receives its ring solution (backward solve) -- one point-to-point
    exchange per shared level, a binary reduction tree over the ranks.

    Returns levels aligned one to one with ``levels``.  Local fronts are renumbered 0..nb-1;
    ``nb`` may be 0 on shared levels.  ``recv_from`` >= 0: before this level, slot ``nb_child``
    of the child batch is filled by that rank (``ch2`` already points at it).  ``send_to`` >= 0:
    before this level, this rank's single child front goes to that rank.
    """
    if world < 1 or world & (world - 1):
        raise ValueError("world size must be a power of two, got {}".format(world))
    nlev = len(levels)
    if world > 1 and (nlev < 2 or levels[0].nb < world):
        raise ValueError("grid too small to split over {} ranks".format(world))
    owner = [None] * nlev
    group = [1] * nlev
    owner[-1] = np.zeros(1, dtype=np.int64)
    group[-1] = world
    for l in range(nlev - 1, 0, -1):
        lv, g = levels[l], group[l]
        child = np.empty(levels[l - 1].nb, dtype=np.int64)
        if getattr(lv, "chain", False):              # same front, next piece of its separator: same owner
            child[lv.ch1] = owner[l]
            owner[l - 1], group[l - 1] = child, g
            continue
        child[lv.ch1] = owner[l]
        child[lv.ch2] = owner[l] + (g // 2 if g > 1 else 0)
        owner[l - 1], group[l - 1] = child, max(g // 2, 1)
    out = []
    loc_prev = None
    for l, lv in enumerate(levels):
        mine = np.flatnonzero(owner[l] == rank)
        loc = np.full(lv.nb, -1, dtype=np.int64)
        loc[mine] = np.arange(len(mine))
        nl = Level()
        nl.__dict__.update(lv.__dict__)
        nl.nb = int(len(mine))
        nl.gids = mine
        nl.cls = lv.cls[mine]
        nl.send_to = nl.recv_from = -1
        g = group[l]
        if lv.kind == "leaf":
            nl.x0, nl.y0 = lv.x0[mine], lv.y0[mine]
        else:
            nl.ch1 = loc_prev[lv.ch1[mine]].astype(np.int32)
            if getattr(lv, "chain", False):
                nl.ch2 = nl.ch1
            elif g > 1:
                assert nl.nb <= 1
                nl.ch2 = np.ones(nl.nb, dtype=np.int32)        # the slot after the single local child
                if nl.nb == 1:
                    nl.recv_from = rank + g // 2
                elif rank % g == g // 2:
                    nl.send_to = rank - g // 2
            else:
                nl.ch2 = loc_prev[lv.ch2[mine]].astype(np.int32)
            assert nl.nb == 0 or (nl.ch1.min() >= 0 and nl.ch2.min() >= 0)
        out.append(nl)
        loc_prev = loc
    out = PlanList(out)
    out.dist = []
    if distribute and world > 1:
        rb = DIST_RB if rb is None else rb
        l = 0
        while l < nlev:
            g = group[l]
            if g == 1 or getattr(levels[l], "chain", False):
                l += 1
                continue
            nchain = 0
            while l + 1 + nchain < nlev and getattr(levels[l + 1 + nchain], "chain", False):
                nchain += 1
            gbase = (rank // g) * g
            fidx = int(np.flatnonzero(owner[l] == gbase)[0])
            out.dist.append(_dist_front(levels, l, nchain, fidx, gbase, g, rb))
            for s in range(nchain + 1):                     # the shared levels hold no batched fronts any more
                nl = out[l + s]
                nl.nb, nl.send_to, nl.recv_from = 0, -1, -1
                nl.cls = nl.cls[:0]
                nl.ch1 = nl.ch1[:0]
                nl.ch2 = nl.ch2[:0]
                nl.gids = nl.gids[:0]
            l += nchain + 1
    return out


class PlanList(list):
    """A list of levels that can carry the distributed fronts of a sharded plan (``.dist``)."""
    dist = ()
