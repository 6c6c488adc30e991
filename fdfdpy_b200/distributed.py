"""One process per GPU: the communicator the sharded paths run on (NCCL over NVLink / NVSwitch,
created inside libfdfd_b200.so; see include/fdfd_b200.h "one grid split over several GPUs").

The reference is single-process; this is the multi-GPU extension of its hot path:

* ``DirectSolver(op, comm=...)``   one grid's elimination tree split over the ranks (subtree per
  GPU, binary reduction tree above; ndplan.shard_plan);
* ``SlabOperator`` / ``slab_krylov``  the matrix-free stencil on slabs with halo exchange.

The 128-byte NCCL id has to reach every rank once; ``Communicator.from_torch()`` uses an
initialised ``torch.distributed`` process group for that (torch is plumbing here, nothing else),
``Communicator(rank, world, uid)`` takes an id delivered any other way.
"""
import ctypes as C
import glob
import os
import site

import numpy as np

from . import _lib
from ._lib import check


def _find_nccl():
    """The NCCL build torch ships (so both users share one copy when torch is loaded), else the system one."""
    cands = []
    for sp in site.getsitepackages() + [site.getusersitepackages()]:
        cands += glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"))
    return cands[0] if cands else None


class Communicator:
    def __init__(self, rank, world, uid):
        self.lib = _lib.load()
        _lib.require_gpu()
        path = os.environ.get("FDFD_NCCL_LIB") or _find_nccl()
        check(self.lib.fdfd_comm_load(path.encode() if path else None))
        self.rank, self.world = int(rank), int(world)
        uid = np.ascontiguousarray(np.frombuffer(bytes(uid), dtype=np.uint8))
        if uid.size != 128:
            raise ValueError("the communicator id is 128 bytes")
        self.h = C.c_void_p()
        check(self.lib.fdfd_comm_create(C.byref(self.h), _lib.ptr(uid), self.rank, self.world))

    @staticmethod
    def unique_id():
        lib = _lib.load()
        path = os.environ.get("FDFD_NCCL_LIB") or _find_nccl()
        check(lib.fdfd_comm_load(path.encode() if path else None))
        uid = np.zeros(128, dtype=np.uint8)
        check(lib.fdfd_comm_unique_id(_lib.ptr(uid)))
        return uid.tobytes()

    @classmethod
    def from_torch(cls):
        """Rank / world / id exchange through the default torch.distributed process group."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0])

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.fdfd_comm_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass
