"""Multi-GPU parity of the sharded elimination tree (needs >= 2 GPUs on the box; skipped otherwise):
one process per GPU under torch.distributed.run, NCCL point-to-point between the ranks."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import ctypes as C
    from fdfdpy_b200 import _lib
    n = C.c_int(0)
    return n.value if _lib.load().fdfd_device_count(C.byref(n)) == 0 else 0


@pytest.mark.parametrize("world,pol", [(2, "Ez"), (2, "Hz"), (4, "Ez")])
def test_sharded_direct_solver_vs_oracle(world, pol):
    if _gpu_count() < world:
        pytest.skip("needs {} GPUs".format(world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), os.path.join(ROOT, "tools", "dist_check.py"),
           "--parity", "200x160", "--pol", pol]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["rank0"]["parity_rel_l2_vs_oracle"] < 1e-8
    assert out["rank0"]["parity_relres"] < 1e-10
    for o in [out["rank0"]] + out["others"]:
        assert o["ranks_agree"]


@pytest.mark.parametrize("world,pol", [(2, "Ez"), (2, "Hz")])
def test_slab_stencil_and_krylov_vs_oracle(world, pol):
    """Slab operator over NCCL: halo exchange (overlapped with the interior rows), distributed BiCGSTAB / COCG
    with all-reduced inner products, against the oracle's matrix and sparse solve."""
    if _gpu_count() < world:
        pytest.skip("needs {} GPUs".format(world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "dist_check.py"),
           "--parity", "", "--slab-parity", "40x36", "--pol", pol]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    for o in [out["rank0"]] + out["others"]:
        assert o["slab_apply_rel_err"] < 1e-13
        assert o["slab_apply_fused_rel_err"] < 1e-13
        for method in ("slab_bicgstab", "slab_cocg"):
            assert o[method]["relres"] < 1e-10, (method, o[method])
            assert o[method]["rel_l2_vs_oracle"] < 1e-8, (method, o[method])
