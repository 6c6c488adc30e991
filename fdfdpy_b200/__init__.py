"""fdfdpy_b200: the fdfdpy 2-D FDFD hot path on B200 (sm_100a) behind the reference's Python API.

    from fdfdpy_b200 import Simulation        # drop-in for ``from fdfdpy import Simulation``
"""
from .simulation import Simulation  # noqa: F401

name = "fdfdpy_b200"
