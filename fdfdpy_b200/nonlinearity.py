"""Kerr nonlinearity container; interface of the reference's fdfdpy/nonlinearity.py:6-37."""
import numpy as np


class Nonlinearity:

    def __init__(self, chi, nl_region, nl_type='kerr', eps_scale=False, eps_max=None):
        self.chi = chi
        self.nl_region = nl_region
        self.nl_type = nl_type
        self.eps_scale = eps_scale
        self.eps_max = eps_max
        if nl_type != 'kerr':
            raise AssertionError("Only 'kerr' type nonlinearity is supported")
        if eps_scale and eps_max is None:
            raise AssertionError("Must provide eps_max when eps_scale is True")

    def _weight(self, eps_r):
        if self.eps_scale:
            return (eps_r - 1) / (self.eps_max - 1)
        return 1.0

    # eps_nl(e), d eps_nl / d e and d eps_nl / d eps  (nonlinearity.py:24-31)
    def eps_nl(self, e, eps_r):
        return 3 * self.chi * self.nl_region * np.square(np.abs(e)) * self._weight(eps_r)

    def dnl_de(self, e, eps_r):
        return 3 * self.chi * self.nl_region * np.conj(e) * self._weight(eps_r)

    def dnl_deps(self, e, eps_r):
        if self.eps_scale:
            return 3 * self.chi * self.nl_region * np.square(np.abs(e)) * (1 / (self.eps_max - 1))
        return 0
