"""Import the UNMODIFIED upstream fdfdpy from /root/reference for golden-vector generation.

Only usable in the build container (the GPU box has no /root/reference).  The reference
hard-imports two packages that are absent from this image, and has two bare-name bugs on
the paths we need; this module works around them WITHOUT touching the reference sources:

* ``pyMKL.pardisoSolver`` (linalg.py:4,11) -> stub backed by scipy SuperLU (``splu``), i.e.
  the same factorisation the reference's own ``solver='scipy'`` branch uses (linalg.py:139).
* ``matplotlib`` (plot.py:2-4)             -> inert stub, plotting is not on the path.
* ``eye`` (linalg.py:100) is used un-imported in the Hz branch (NameError upstream) ->
  ``scipy.sparse.eye`` is injected into the module namespace, the evident intent.
* ``spsolve`` / ``zeros`` (linalg.py:158,177) likewise injected.
"""
import sys
import types
import warnings

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

REFERENCE_ROOT = "/root/reference"


class _PardisoStub:
    def __init__(self, A, mtype=13):
        self._A = sp.csc_matrix(A)
        self._lu = None

    def factor(self):
        self._lu = spl.splu(self._A)

    def solve(self, b):
        return self._lu.solve(np.asarray(b))

    def clear(self):
        self._lu = None


def import_reference():
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if "pyMKL" not in sys.modules:
        m = types.ModuleType("pyMKL")
        m.pardisoSolver = _PardisoStub
        sys.modules["pyMKL"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        pylab = types.ModuleType("matplotlib.pylab")
        mpl.pyplot = plt
        mpl.pylab = pylab
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.pylab"] = pylab
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fdfdpy  # noqa: F401
    import fdfdpy.linalg as L
    L.eye = sp.eye
    L.spsolve = spl.spsolve
    L.zeros = np.zeros
    return fdfdpy
