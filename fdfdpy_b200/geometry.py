"""Index helpers for source / probe planes (lines in 2-D)."""


def plane_slices(direction_normal, center, width, what="direction_normal"):
    """(x-slice, y-slice) of the one-cell-thick line through ``center`` that is ``width`` cells long
    and normal to ``direction_normal`` -- the indexing shared by mode.py:69-75 and
    simulation.py:270-277 of the reference (int() truncation of the half width included)."""
    half_lo = lambda c: int(c - width / 2)   # noqa: E731
    half_hi = lambda c: int(c + width / 2)   # noqa: E731
    if direction_normal == "x":
        return slice(center[0], center[0] + 1), slice(half_lo(center[1]), half_hi(center[1]))
    if direction_normal == "y":
        return slice(half_lo(center[0]), half_hi(center[0])), slice(center[1], center[1] + 1)
    raise ValueError("The value of {} is neither x nor y!".format(what))


def grow(sl, by=1):
    """The same slice extended by ``by`` cells at its upper end."""
    return slice(sl.start, sl.stop + by)
