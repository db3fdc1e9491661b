"""ctypes binding of libarnoldi_b200.so (the C ABI in include/arnoldi_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible,
every n-length operation raises.  The library is built in-tree by
``__graft_entry__.build()`` / ``make -C arnoldi-py_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AB200_LIB_PATH") or os.path.join(_HERE, "libarnoldi_b200.so")

ABI_VERSION = 2
OK, EINVAL, ECUDA, ENOMEM, ESTATE, ECOMM = 0, -1, -2, -3, -4, -5
F64, C128 = 0, 1
ORTHO_CGS2, ORTHO_MGS = 0, 1
SPMV_AUTO, SPMV_VECTOR, SPMV_STREAM, SPMV_MERGE = 0, 1, 2, 3
SPMV_ALGOS = {"auto": SPMV_AUTO, "vector": SPMV_VECTOR, "stream": SPMV_STREAM, "merge": SPMV_MERGE}
# C callback type of ab200_set_operator: int fn(void *user, const void *x, void *y, int64 n,
#                                              int is_real, void *stream)
APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p)


class Stats(C.Structure):
    """Mirror of ``ab200_stats``."""

    _fields_ = (
        [(n, C.c_double) for n in ("spmv_ms", "ortho_pass1_ms", "ortho_pass2_ms",
                                   "ortho_fused_ms", "mgs_ms", "restart_ms", "spmv_bytes",
                                   "ortho_pass1_bytes", "ortho_pass2_bytes", "ortho_fused_bytes",
                                   "mgs_bytes", "restart_bytes")]
        + [(n, C.c_int64) for n in ("spmv_launches", "ortho_pass1_launches",
                                    "ortho_pass2_launches", "ortho_fused_launches",
                                    "mgs_launches", "restart_launches",
                                    "arnoldi_steps", "ortho_rounds", "second_rounds",
                                    "kernel_launches", "real_storage")]
    )

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DeviceError(RuntimeError):
    """CUDA / allocation / communicator failure reported by the library."""


# every exported symbol of include/arnoldi_b200.h: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
SIGNATURES = {
    "ab200_abi_version": (C.c_int, []),
    "ab200_last_error": (C.c_char_p, []),
    "ab200_device_count": (C.c_int, []),
    "ab200_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "ab200_destroy": (C.c_int, [_P]),
    "ab200_set_csr": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int64, C.c_int]),
    "ab200_set_operator": (C.c_int, [_P, _P, _P, C.c_int]),
    "ab200_set_columns": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64]),
    "ab200_get_columns": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64]),
    "ab200_expand": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _P, _I, _I]),
    "ab200_restart": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int]),
    "ab200_combine": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "ab200_orthonormalize_column": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double,
                                              C.c_int, _D, _I]),
    "ab200_project": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ab200_spmv": (C.c_int, [_P, _P, _P]),
    "ab200_ortho": (C.c_int, [_P, C.c_int, _P, _P, C.c_double, C.c_double, C.c_int, _D, _I]),
    "ab200_comm_export": (C.c_int, [_P, _P]),
    "ab200_comm_connect": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, _P]),
    "ab200_set_halo": (C.c_int, [_P, _P, C.c_int64]),
    "ab200_comm_disconnect": (C.c_int, [_P]),
    "ab200_comm_bench": (C.c_int, [_P, C.c_int, _D]),
    "ab200_halo_export": (C.c_int, [_P, _P]),
    "ab200_halo_connect": (C.c_int, [_P, C.c_char_p, _P, _P, _P]),
    "ab200_set_timing": (C.c_int, [_P, C.c_int]),
    "ab200_reset_stats": (C.c_int, [_P]),
    "ab200_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "ab200_synchronize": (C.c_int, [_P]),
    "ab200_timer_start": (C.c_int, [_P]),
    "ab200_timer_stop": (C.c_int, [_P, _D]),
    "ab200_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "ab200_host_schur": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "ab200_host_schur_real": (C.c_int, [_P, C.c_int, _P, _P]),
    "ab200_host_eigh_real": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "ab200_host_reorder": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "ab200_host_alloc": (C.c_int, [C.POINTER(_P), C.c_int64]),
    "ab200_host_free": (C.c_int, [_P]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C arnoldi-py_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.ab200_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI version mismatch: library {lib.ab200_abi_version()}, "
                          f"binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    """Map a status code to the exception the reference would raise at that point."""
    if rc == OK:
        return
    msg = load().ab200_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise AssertionError(msg)  # the reference validates arguments with `assert`
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise DeviceError(f"[{rc}] {msg}")
