"""Field plots (reference: fdfdpy/plot.py, simulation.py:329-436).  Visualisation only: matplotlib is an
optional dependency and nothing here is on the solve path."""
import numpy as np


def _pyplot():
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:      # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not installed") from e
    return plt


def _show(values, outline_val, cmap, vmin, vmax, label, cbar, outline, ax):
    plt = _pyplot()
    if ax is None:
        _, ax = plt.subplots(1, constrained_layout=True)
    h = ax.imshow(values.T, cmap=cmap, vmin=vmin, vmax=vmax, origin='lower')
    if cbar:
        plt.colorbar(h, label=label, ax=ax)
    if outline:
        for lw, col in ((1.0, 'w'), (0.5, 'k')):
            ax.contour(outline_val.T, levels=2, linewidths=lw, colors=col)
    ax.set_xticks([])
    ax.set_yticks([])
    return ax


def _primary(sim, nl, tiled_y):
    fld = (sim.fields_nl if nl else sim.fields)[sim.pol]
    if sim.fields[sim.pol] is None:
        raise ValueError("need to solve the simulation first")
    return np.hstack(tiled_y * [fld]), np.abs(np.hstack(tiled_y * [sim.eps_r]))


def plt_abs(sim, nl=False, cbar=True, outline=True, ax=None, vmax=None, tiled_y=1):
    fld, eps = _primary(sim, nl, tiled_y)
    val = np.abs(fld)
    return _show(val, eps, "magma", 0.0, val.max() if vmax is None else vmax, sim.pol, cbar, outline, ax)


def plt_re(sim, nl=False, cbar=True, outline=True, ax=None, tiled_y=1):
    fld, eps = _primary(sim, nl, tiled_y)
    val = np.real(fld)
    m = np.abs(fld).max()
    return _show(val, eps, "RdBu", -m, m, sim.pol, cbar, outline, ax)


def plt_diff(sim, cbar=True, outline=True, ax=None, vmax=None, tiled_y=1, normalize=True):
    lin = np.abs(np.hstack(tiled_y * [sim.fields[sim.pol]]))
    nl = np.abs(np.hstack(tiled_y * [sim.fields_nl[sim.pol]]))
    diff = lin - nl
    if normalize:
        diff = diff / lin.max()
    vmax = np.abs(diff).max() if vmax is None else vmax
    return _show(diff, np.abs(np.hstack(tiled_y * [sim.eps_r])), 'RdYlBu', -vmax, vmax, sim.pol, cbar, outline, ax)


def plt_eps(sim, cbar=True, outline=True, ax=None, tiled_y=1):
    eps = np.abs(np.hstack(tiled_y * [sim.eps_r]))
    return _show(eps, eps, "Greys", eps.min(), eps.max(), 'relative permittivity', cbar, outline, ax)
