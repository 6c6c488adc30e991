"""CPU tests: input validation of the public API, the elimination plan (executed by a numpy model
of the GPU algorithm against the oracle), the C-ABI symbol table, and the loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import fdfd_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constructor_rejects_bad_inputs():
    """tests/test_simulation.py of the reference, verbatim in intent: every bad argument -> ValueError."""
    from fdfdpy_b200 import Simulation
    good = dict(omega=100, eps_r=np.ones((100, 50)), dl=0.001, NPML=[10, 10], pol='Hz')

    def build(**kw):
        a = dict(good, **kw)
        return Simulation(a['omega'], a['eps_r'], a['dl'], a['NPML'], a['pol'])

    for bad in (dict(omega=-100), dict(omega=[100, 200, 300]), dict(eps_r=-np.ones((100, 50))),
                dict(eps_r=list(np.ones((100, 50)))), dict(dl=-0.001), dict(dl=[1e-4, 1e-5]),
                dict(NPML=10), dict(NPML=[10, 10, 10]), dict(NPML=[200, 200]), dict(pol=5),
                dict(pol='WrongPolarization')):
        with pytest.raises(ValueError):
            build(**bad)
    # the reference raises AssertionError for some of these (simulation.py:258-265): also accepted
    with pytest.raises(AssertionError):
        build(pol='TE')


def test_no_gpu_fails_loudly():
    from fdfdpy_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_int(0)
    if lib.fdfd_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    from fdfdpy_b200 import Simulation
    with pytest.raises(_lib.FdfdError):
        Simulation(100, np.ones((20, 20)), 0.001, [3, 3], 'Ez')


def test_cabi_exports_every_declared_symbol():
    from fdfdpy_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "fdfd_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fdfd_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "binding missing for " + name
    assert set(_lib.SIGNATURES) == declared
    assert lib.fdfd_version() >= 100


@pytest.mark.parametrize("shape,npml", [((16, 16), [3, 3]), ((23, 17), [4, 3]), ((37, 52), [0, 6]), ((8, 64), [2, 9])])
@pytest.mark.parametrize("pol", ["Ez", "Hz"])
def test_elimination_plan_numpy_model(shape, npml, pol):
    """The plan executed the way the kernels do (padded symmetric fronts, blocked inversion, level-wise solve)."""
    from fdfdpy_b200.ndplan import build_plan
    from tests.nd_model import factor, solve, row_scale
    nx, ny = shape
    rng = np.random.default_rng(0)
    eps = 1 + 5 * rng.random((nx, ny))
    omega = 2 * np.pi * 200e12
    planes = orc.stencil_planes(omega, eps, 0.04, npml, pol, 1e-6)
    levels = build_plan(nx, ny)
    isxf, _, isyf, _ = orc.pml_inverse_factors(omega, 1e-6, (nx, ny), npml, 0.04)
    d = row_scale(isxf, isyf)
    # the row-scaled operator is complex symmetric: that is what lets the fronts keep one triangle
    As = (orc.planes_to_csr(planes).multiply(d[:, None])).tocsr()
    assert abs(As - As.T).max() <= 1e-12 * abs(As).max()
    store = factor(levels, planes, nx, ny, d, tile=8)
    b = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    u = solve(levels, store, b, nx, ny, d)
    ref = orc.sparse_solve(orc.planes_to_csr(planes), b).reshape(nx, ny)
    assert np.linalg.norm(u - ref) / np.linalg.norm(ref) < 1e-11


@pytest.mark.parametrize("split_min,split_parts", [(6, 2), (6, 3), (5, -4)])
def test_separator_splitting_chain_levels(split_min, split_parts):
    """Separators eliminated piece by piece (chain levels, the production setting for >= 1000-node separators)
    give the same solution as the one-step elimination; exercised here with tiny thresholds."""
    from fdfdpy_b200.ndplan import build_plan
    from tests.nd_model import factor, solve, row_scale
    nx, ny, npml = 37, 29, [4, 3]
    rng = np.random.default_rng(4)
    eps = 1 + 5 * rng.random((nx, ny))
    omega = 2 * np.pi * 200e12
    planes = orc.stencil_planes(omega, eps, 0.04, npml, "Hz", 1e-6)
    isxf, _, isyf, _ = orc.pml_inverse_factors(omega, 1e-6, (nx, ny), npml, 0.04)
    d = row_scale(isxf, isyf)
    plain = build_plan(nx, ny, split_min=10 ** 9)
    levels = build_plan(nx, ny, split_min=split_min, split_parts=split_parts)
    assert len(levels) > len(plain) and any(getattr(lv, "chain", False) for lv in levels)
    assert sum(int(lv.k_cls[c]) for lv in levels for c in lv.cls) == nx * ny      # every node eliminated once
    b = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    u = solve(levels, factor(levels, planes, nx, ny, d, tile=8), b, nx, ny, d)
    ref = orc.sparse_solve(orc.planes_to_csr(planes), b).reshape(nx, ny)
    assert np.linalg.norm(u - ref) / np.linalg.norm(ref) < 1e-11


def test_plan_structure_invariants():
    from fdfdpy_b200.ndplan import build_plan, plan_stats
    for nx, ny in [(200, 200), (300, 100), (125, 87), (64, 1000)]:
        levels = build_plan(nx, ny)
        assert levels[0].kind == "leaf" and levels[-1].nb == 1 and levels[-1].mmax == 0
        # every grid node is eliminated exactly once
        total = sum(int(lv.k_cls[c]) for lv in levels for c in lv.cls)
        assert total == nx * ny
        for lv in levels[1:]:
            assert lv.c1map.max() < lv.nmax and lv.c2map.max() < lv.nmax
        stored, transient, macs = plan_stats(levels)
        assert stored > 0 and transient > 0 and macs > 0
    with pytest.raises(ValueError):
        build_plan(3, 50)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port timed on the host cores) prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "96"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "Mcell/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("Ez 4096x4096")


@pytest.mark.parametrize("n,tile", [(64, 64), (65, 64), (200, 64), (37, 8), (128, 32)])
def test_block_gauss_jordan_model(n, tile):
    """The out-of-place block Gauss-Jordan sweep of csrc/direct.cu (gj_update_kernel's four tile cases, ping-pong
    buffers, ragged last tile) inverts complex symmetric, diagonally strong matrices like the pivot blocks it is
    used on; the in-place formulation of the older model agrees with it."""
    from tests.nd_model import gj_inverse, gj_inverse_pingpong
    rng = np.random.default_rng(n)
    A = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    A = A + np.transpose(A, (0, 2, 1)) + 2 * n * np.eye(n)
    inv = np.linalg.inv(A)
    got = gj_inverse_pingpong(A, tile=tile)
    assert np.abs(got - inv).max() <= 1e-12 * np.abs(inv).max()
    assert np.abs(gj_inverse(A, tile=tile) - got).max() <= 1e-12 * np.abs(inv).max()
