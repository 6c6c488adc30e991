"""ctypes binding of libfdfd_b200.so (the C ABI in include/fdfd_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device answers, every
compute entry point raises.  ``load()`` only dlopens the library (works without a GPU, so the
symbol table can be checked on a CPU box).
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfdfd_b200.so")

c128 = np.complex128
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


class LevelDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("nb", C.c_int), ("kmax", C.c_int), ("mmax", C.c_int), ("ncls", C.c_int),
                ("child_mmax", C.c_int),
                ("cls", _ip), ("k_cls", _ip), ("ch1", _ip), ("ch2", _ip), ("c1map", _ip), ("c2map", _ip),
                ("x0", _ip), ("y0", _ip), ("slot_lx", _ip), ("slot_ly", _ip), ("slot_right", _ip),
                ("slot_up", _ip), ("send_to", C.c_int), ("recv_from", C.c_int)]


class DistFrontDesc(C.Structure):
    _fields_ = [("level0", C.c_int), ("nsteps", C.c_int), ("gbase", C.c_int), ("gsize", C.c_int), ("n", C.c_int),
                ("nblk", C.c_int), ("bstart", _ip), ("bowner", _ip), ("mc1", C.c_int), ("mc2", C.c_int),
                ("inv1", _ip), ("inv2", _ip)]


# name -> (restype, argtypes); mirrors include/fdfd_b200.h one to one
SIGNATURES = {
    "fdfd_version": (C.c_int, []),
    "fdfd_last_error": (C.c_char_p, []),
    "fdfd_device_count": (C.c_int, [_ip]),
    "fdfd_set_device": (C.c_int, [C.c_int]),
    "fdfd_mem_info": (C.c_int, [_dp, _dp]),
    "fdfd_malloc": (C.c_int, [C.POINTER(_vp), C.c_double]),
    "fdfd_free": (C.c_int, [_vp]),
    "fdfd_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_double]),
    "fdfd_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_double]),
    "fdfd_op_sync": (C.c_int, [_vp]),
    "fdfd_launch_count": (C.c_double, [C.c_int]),
    "fdfd_timer_start": (C.c_int, [_vp]),
    "fdfd_timer_stop": (C.c_int, [_vp, _dp]),
    "fdfd_gemm_timing": (C.c_int, [C.c_int]),
    "fdfd_gemm_timing_read": (C.c_int, [_vp]),
    "fdfd_gemm_timing_exec_flops": (C.c_int, [_dp]),
    "fdfd_dmma_peak": (C.c_int, [_dp]),
    "fdfd_phase_timing": (C.c_int, [C.c_int]),
    "fdfd_phase_timing_read": (C.c_int, [_vp]),
    "fdfd_phase_timing_read_levels": (C.c_int, [_vp, C.c_int]),
    "fdfd_dmma_probe": (C.c_int, [C.c_int, C.c_int, _dp]),
    "fdfd_dmma_probe_clocked": (C.c_int, [C.c_int, C.c_int, _vp]),
    "fdfd_dmma_pattern_probe": (C.c_int, [C.c_int, C.c_int, _vp]),
    "fdfd_dmma_smem_probe": (C.c_int, [C.c_int, _vp]),
    "fdfd_host_register": (C.c_int, [_vp, C.c_double]),
    "fdfd_host_unregister": (C.c_int, [_vp]),
    "fdfd_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_double]),
    "fdfd_host_free": (C.c_int, [_vp]),
    "fdfd_op_assemble_host_f64": (C.c_int, [_vp, _vp, C.c_int]),
    "fdfd_factor_solve_fields_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp, C.c_int,
                                                C.c_int, C.c_double, _dp, _ip, _dp]),
    "fdfd_op_eps_flags": (C.c_int, [_vp, _ip]),
    "fdfd_solve_fields_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp, C.c_int,
                                         C.c_int, C.c_double, _dp, _ip]),
    "fdfd_op_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_int, C.c_double]),
    "fdfd_op_destroy": (None, [_vp]),
    "fdfd_op_assemble_host": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "fdfd_op_assemble_dev": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "fdfd_op_get_sfactors_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "fdfd_op_get_planes_host": (C.c_int, [_vp, _vp]),
    "fdfd_op_apply_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "fdfd_op_apply_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "fdfd_op_derive_fields_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "fdfd_op_derive_fields_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "fdfd_direct_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int]),
    "fdfd_direct_add_level": (C.c_int, [_vp, C.POINTER(LevelDesc)]),
    "fdfd_direct_destroy": (None, [_vp]),
    "fdfd_direct_factor": (C.c_int, [_vp, _vp]),
    "fdfd_direct_stats": (C.c_int, [_vp, _dp, _dp]),
    "fdfd_direct_solve_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _dp, _ip]),
    "fdfd_direct_solve_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _dp, _ip]),
    "fdfd_comm_load": (C.c_int, [C.c_char_p]),
    "fdfd_comm_unique_id": (C.c_int, [_vp]),
    "fdfd_comm_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int]),
    "fdfd_comm_destroy": (None, [_vp]),
    "fdfd_comm_allreduce_sum_dev": (C.c_int, [_vp, _vp, _vp, C.c_double]),
    "fdfd_direct_set_comm": (C.c_int, [_vp, _vp]),
    "fdfd_direct_add_dist_front": (C.c_int, [_vp, C.POINTER(DistFrontDesc)]),
    "fdfd_comm_create_local": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "fdfd_comm_abort": (None, [_vp]),
    "fdfd_schwarz_sub_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int]),
    "fdfd_slab_set_schwarz": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "fdfd_slab_op_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_int, C.c_int, C.c_int, C.c_double]),
    "fdfd_krylov_solve_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                         _vp, C.c_int, _ip, _dp, _ip]),
    "fdfd_krylov_solve_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                        _vp, C.c_int, _ip, _dp, _ip]),
    "fdfd_nl_solve_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_int, _vp, _ip, _ip]),
    "fdfd_op_apply_host_c64": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "fdfd_op_apply_dev_c64": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "fdfd_krylov_solve_host_c64": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _ip, _dp,
                                             _ip]),
    "fdfd_krylov_solve_dev_c64": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _ip, _dp,
                                            _ip]),
    "fdfd_zgemm_batched_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int]),
    "fdfd_stencil_set_variant": (C.c_int, [C.c_int, C.c_int]),
    "fdfd_stencil_set_hz_variant": (C.c_int, [C.c_int, C.c_int]),
    "fdfd_zgemm_set_variant": (C.c_int, [C.c_int]),
    "fdfd_direct_set_small_fronts": (C.c_int, [C.c_int]),
    "fdfd_zgemm_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "fdfd_mode_solve_host": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                       C.c_int, C.c_int, _vp, _vp]),
}

_lib = None


class FdfdError(RuntimeError):
    pass


def load():
    """dlopen the library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdfdError("libfdfd_b200.so is not built (run `python -m fdfdpy_b200.build` or "
                        "__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise FdfdError(load().fdfd_last_error().decode("utf-8", "replace"))


def require_gpu():
    lib = load()
    n = C.c_int(0)
    if lib.fdfd_device_count(C.byref(n)) != 0 or n.value < 1:
        raise FdfdError("no CUDA device available: fdfdpy_b200 has no CPU path ({})".format(
            lib.fdfd_last_error().decode("utf-8", "replace")))
    return n.value


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=c128)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- page-locked result arrays -------------------------------------------------------------------
# Fields handed back to the caller are numpy arrays over cudaHostAlloc'ed buffers: the device->host
# copy runs at full PCIe rate and skips the first-touch page faults of a fresh np.empty.  A buffer
# returns to the free list when the last array viewing it is garbage collected, so an optimisation
# loop that overwrites its fields every iteration keeps re-using the same few buffers.
PINNED_RESULTS = True               # False: plain numpy result arrays (a program that KEEPS every result pays
#                                     ~0.13 s per 268 MB to page-lock a fresh buffer; recycled buffers are free)
PINNED_MIN_BYTES = 1 << 20          # smaller arrays are ordinary numpy memory
PINNED_KEEP_BYTES = 8 << 30         # free-list cap; beyond it released buffers go back to the driver
_pool_lock = threading.Lock()
_pool_free = {}                     # nbytes -> [address, ...]
_pool_free_bytes = 0


def _pinned_release(addr, nbytes):
    global _pool_free_bytes
    with _pool_lock:
        if _pool_free_bytes + nbytes <= PINNED_KEEP_BYTES:
            _pool_free.setdefault(nbytes, []).append(addr)
            _pool_free_bytes += nbytes
            return
    try:
        _lib.fdfd_host_free(_vp(addr))
    except Exception:
        pass


def pinned_empty(shape, dtype=c128):
    """np.empty over page-locked memory from the pool (plain np.empty for small arrays or when the
    driver refuses to pin more memory)."""
    global _pool_free_bytes
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < PINNED_MIN_BYTES or not PINNED_RESULTS:
        return np.empty(shape, dtype=dtype)
    lib = load()
    addr = None
    with _pool_lock:
        lst = _pool_free.get(nbytes)
        if lst:
            addr = lst.pop()
            _pool_free_bytes -= nbytes
    if addr is None:
        p = _vp()
        if lib.fdfd_host_alloc(C.byref(p), float(nbytes)) != 0 or not p.value:
            return np.empty(shape, dtype=dtype)
        addr = p.value
    buf = (C.c_char * nbytes).from_address(addr)
    weakref.finalize(buf, _pinned_release, addr, nbytes)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
