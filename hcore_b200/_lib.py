"""ctypes loader for ``hcore_b200/lib/libhcore_b200.so`` (the C ABI in ``include/hcore_b200.h``).

The library is the product; this module only declares argument types.  There is no
Python or CPU fallback: if the shared library is missing, importing any compute entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HCORE_B200_LIB selects another build of the same library (the instrumented `make prof` one)
LIB_PATH = os.environ.get("HCORE_B200_LIB") or os.path.join(_HERE, "lib", "libhcore_b200.so")

#: every symbol ``include/hcore_b200.h`` / ``include/hcore_d.h`` declare
EXPORTED_SYMBOLS = [
    "hcore_b200_create", "hcore_b200_destroy", "hcore_b200_set_chunk", "hcore_b200_last_error",
    "hcore_b200_version", "hcore_b200_launch_count", "hcore_b200_shard_range",
    "HCORE_dgemm_batched", "HCORE_dgemm_batched_strided", "HCORE_dgemm_batched_host",
    "HCORE_sgemm_batched_host", "HCORE_zgemm_batched_host", "hcore_b200_gather_status",
    "HCORE_zgemm_batched", "HCORE_dgemmcd_batched", "HCORE_dsyrk_batched", "HCORE_dcompress_batched_strided",
    "HCORE_sgemm_batched", "HCORE_sgemm_batched_strided", "HCORE_sgemm",
    "HCORE_zgemm_batched_strided", "HCORE_zgemm",
    "sgemm_use_trmm", "sgemm_print_index", "sgemm_print_mat",
    "HCORE_dgemmcd_batched_strided", "HCORE_dsyrk_batched_strided", "hcore_b200_gemm_flops",
    "hcore_b200_debug_fetch", "hcore_b200_debug_set_stages", "hcore_b200_set_stage_timing",
    "hcore_b200_stage_ms",
    "HCORE_dgemm", "HCORE_dgemmcd", "HCORE_dsyrk",
    "HCORE_dgemm_qr_svd", "HCORE_dgemm_ormqr", "HCORE_dgemm_qr_svd_b_dense",
    "HCORE_dgemm_batched_b_dense", "HCORE_dgemm_batched_qr_svd_b_dense", "HCORE_dgemm_batched_b_dense_strided",
    "HCORE_dgemm_batched_qr_svd", "HCORE_dgemm_batched_ormqr",
    "HCORE_dgemm_batched_qr_svd_strided", "HCORE_dgemm_batched_ormqr_strided",
    "use_trmm", "gemm_print_index", "gemm_print_mat", "hc_nelm_limit", "syrk_print_index",
    "gemmcd_print_index", "gemmcd_print_mat", "gemmcd_use_scratch",
    # the remaining tile kernels of one TLR-Cholesky step + HCORE_dgemm_fast
    "HCORE_dgemm_fast", "HCORE_dtrsm", "HCORE_dpotrf", "HCORE_dgemmbdcd", "HCORE_duncompress", "HCORE_zuncompress",
    "HCORE_dtrsm_batched_strided", "HCORE_dpotrf_batched_strided", "HCORE_dgemmbdcd_batched_strided",
    "HCORE_duncompress_batched_strided", "HCORE_zuncompress_batched_strided",
    "gemmbdcd_print_index", "gemmbdcd_print_mat", "gemmfrk_print_index", "gemmfrk_print_mat",
]

_lib = None


class DComplex(C.Structure):
    """``hcore_b200_dcomplex``: complex double scalar passed by value (layout of C99 double _Complex)."""
    _fields_ = [("re", C.c_double), ("im", C.c_double)]

    @classmethod
    def of(cls, z):
        z = complex(z)
        return cls(z.real, z.imag)


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hcore_b200/csrc`. hcore_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, i64, d = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.hcore_b200_create.argtypes = [C.POINTER(vp), i]
    L.hcore_b200_destroy.argtypes = [vp]
    L.hcore_b200_set_chunk.argtypes = [vp, i]
    L.hcore_b200_last_error.argtypes = [vp]
    L.hcore_b200_last_error.restype = C.c_char_p
    L.hcore_b200_version.restype = C.c_char_p
    L.hcore_b200_launch_count.argtypes = [vp]
    L.hcore_b200_launch_count.restype = i64
    L.hcore_b200_shard_range.argtypes = [i64, i, i, C.POINTER(i64), C.POINTER(i64)]
    L.hcore_b200_shard_range.restype = None
    L.hcore_b200_gemm_flops.argtypes = [i, i, i, i, i]
    L.hcore_b200_gemm_flops.restype = C.c_uint64
    L.HCORE_dgemm_batched.argtypes = [vp, i, i, d, vp, vp, vp, i, vp, vp, vp, i, d, vp, vp, vp, i, i, i, d, vp, vp]
    L.HCORE_dgemm_batched_strided.argtypes = [vp, i, i, d, vp, vp, vp, i, i64, vp, vp, vp, i, i64, d,
                                              vp, vp, vp, i, i64, i, i, d, vp, vp]
    L.HCORE_dgemm_batched_qr_svd.argtypes = [vp, i, i, d, vp, vp, vp, i, vp, vp, vp, i, d, vp, vp, vp, i, i, i, d, vp, vp, vp]
    L.HCORE_dgemm_batched_ormqr.argtypes = [vp, i, i, vp, vp, vp, i, vp]
    L.HCORE_dgemm_batched_qr_svd_strided.argtypes = [vp, i, i, d, vp, vp, vp, i, i64, vp, vp, vp, i, i64, d,
                                                     vp, vp, vp, i, i64, i, i, d, vp, vp, vp]
    L.HCORE_dgemm_batched_ormqr_strided.argtypes = [vp, i, i, vp, vp, vp, i, i64, vp]
    L.HCORE_dgemm_batched_b_dense.argtypes = [vp, i, i, d, vp, vp, vp, i, vp, i, d, vp, vp, vp, i, i, i, d, vp, vp]
    L.HCORE_dgemm_batched_qr_svd_b_dense.argtypes = [vp, i, i, d, vp, vp, vp, i, vp, i, d, vp, vp, vp, i, i, i, d, vp, vp, vp]
    L.HCORE_dgemm_batched_b_dense_strided.argtypes = [vp, i, i, d, vp, vp, vp, i, i64, vp, i, i64, d,
                                                      vp, vp, vp, i, i64, i, i, d, vp, vp]
    f = C.c_float
    L.HCORE_sgemm_batched.argtypes = [vp, i, i, f, vp, vp, vp, i, vp, vp, vp, i, f, vp, vp, vp, i, i, i, f, vp, vp]
    L.HCORE_sgemm_batched_strided.argtypes = [vp, i, i, f, vp, vp, vp, i, i64, vp, vp, vp, i, i64, f,
                                              vp, vp, vp, i, i64, i, i, f, vp, vp]
    L.HCORE_zgemm_batched_strided.argtypes = [vp, i, i, DComplex, vp, vp, vp, i, i64, vp, vp, vp, i, i64, DComplex,
                                              vp, vp, vp, i, i64, i, i, d, vp, vp]
    L.HCORE_dgemm_batched_host.argtypes = [vp, i, i, d, vp, vp, vp, i, i64, vp, vp, vp, i, i64, d,
                                           vp, vp, vp, i, i64, i, i, d, vp]
    L.HCORE_sgemm_batched_host.argtypes = [vp, i, i, f, vp, vp, vp, i, i64, vp, vp, vp, i, i64, f,
                                           vp, vp, vp, i, i64, i, i, f, vp]
    L.HCORE_zgemm_batched_host.argtypes = [vp, i, i, DComplex, vp, vp, vp, i, i64, vp, vp, vp, i, i64, DComplex,
                                           vp, vp, vp, i, i64, i, i, d, vp]
    L.HCORE_zgemm_batched.argtypes = [vp, i, i, DComplex, vp, vp, vp, i, vp, vp, vp, i, DComplex, vp, vp, vp, i, i, i, d, vp, vp]
    L.HCORE_dgemmcd_batched.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.HCORE_dsyrk_batched.argtypes = [vp, i, i, i, d, vp, vp, vp, vp, vp]
    L.HCORE_dcompress_batched_strided.argtypes = [vp, i, i, vp, i, i64, d, i, i, vp, vp, i, i64, vp, vp, vp]
    L.hcore_b200_gather_status.argtypes = [vp, vp, i, i, i64, vp, vp, vp, vp, vp]
    L.HCORE_dgemmcd_batched_strided.argtypes = [vp, i, i, vp, vp, vp, i64, vp, vp, vp, i64, vp, i64, vp]
    L.HCORE_dsyrk_batched_strided.argtypes = [vp, i, i, d, vp, vp, vp, i64, vp, i64, vp]
    L.hcore_b200_debug_fetch.argtypes = [vp, i, i, vp, C.c_size_t]
    L.hcore_b200_debug_set_stages.argtypes = [vp, i]
    L.hcore_b200_set_stage_timing.argtypes = [vp, i]
    L.hcore_b200_stage_ms.argtypes = [vp, i]
    L.hcore_b200_stage_ms.restype = d
    # legacy drop-in symbols (reference src/include/hcore_d.h:25-46, 68-86, 107-116)
    dp = C.POINTER(C.c_double)
    L.HCORE_dgemm.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, dp, dp, i, d, dp, dp, dp, i, i, i, d, dp, vp]
    L.HCORE_dgemm.restype = None
    # split-phase legacy symbols (reference src/include/hcore_d.h:190-268): ranks are int*
    ip, dpp = C.POINTER(C.c_int), C.POINTER(dp)
    L.HCORE_dgemm_qr_svd.argtypes = [i, i, i, i, d, dp, dp, ip, i, dp, dp, ip, i, d, dp, dp, ip, i, i, i, i, d, dp, vp,
                                     dpp, dpp, dpp, ip, ip, dpp, ip, dpp, ip, dpp, ip, dpp, ip, dpp, ip, dpp, dpp, ip, dpp]
    L.HCORE_dgemm_qr_svd.restype = None
    L.HCORE_dgemm_qr_svd_b_dense.argtypes = [i, i, i, i, d, dp, dp, ip, i, dp, i, d, dp, dp, ip, i, i, i, i, d, dp, vp,
                                             dpp, dpp, dpp, ip, ip, dpp, ip, dpp, ip, dpp, ip, dpp, ip, dpp, ip, dpp, dpp, ip, dpp]
    L.HCORE_dgemm_qr_svd_b_dense.restype = None
    L.HCORE_dgemm_ormqr.argtypes = [i, i, i, i, d, dp, dp, ip, i, d, dp, dp, ip, i, i, i, i, d, dp, vp,
                                    dp, dp, i, i, dp, i, dp, i, dp, i, dp, i, dp, i, dp, dp, i, dp]
    L.HCORE_dgemm_ormqr.restype = None
    fp = C.POINTER(C.c_float)
    # reference src/include/hcore_s.h:25-46
    L.HCORE_sgemm.argtypes = [i, i, i, i, f, fp, fp, fp, i, fp, fp, fp, i, f, fp, fp, fp, i, i, i, f, fp, vp]
    L.HCORE_sgemm.restype = None
    # reference src/include/hcore_z.h:27-48
    L.HCORE_zgemm.argtypes = [i, i, i, i, DComplex, vp, vp, dp, i, vp, vp, dp, i, DComplex, vp, vp, dp, i, i, i, d, vp, vp]
    L.HCORE_zgemm.restype = None
    L.HCORE_dgemmcd.argtypes = [i, i, i, d, dp, dp, i, i, dp, dp, i, i, d, dp, i, i, dp, vp]
    L.HCORE_dgemmcd.restype = None
    L.HCORE_dsyrk.argtypes = [i, i, i, i, d, dp, i, dp, i, d, dp, i, dp, vp]
    L.HCORE_dsyrk.restype = None
    # reference src/include/hcore_d.h:47-67, 87-106, 117-138; hcore_z.h:76-88
    L.HCORE_dgemm_fast.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, dp, dp, i, d, dp, dp, dp, i, i, i, d, dp]
    L.HCORE_dgemm_fast.restype = None
    L.HCORE_dtrsm.argtypes = [i, i, i, i, i, i, d, dp, i, dp, i, vp]
    L.HCORE_dtrsm.restype = None
    L.HCORE_dpotrf.argtypes = [i, i, dp, i, ip, vp]
    L.HCORE_dpotrf.restype = None
    L.HCORE_dgemmbdcd.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, i, d, dp, i, dp, vp]
    L.HCORE_dgemmbdcd.restype = None
    L.HCORE_duncompress.argtypes = [i, i, i, i, d, dp, dp, i, dp, dp, i, d, dp, i]
    L.HCORE_duncompress.restype = None
    L.HCORE_zuncompress.argtypes = [i, i, i, i, DComplex, vp, dp, i, vp, dp, i, DComplex, vp, i]
    L.HCORE_zuncompress.restype = None
    L.HCORE_dgemmbdcd_batched_strided.argtypes = [vp, i, i, i, d, vp, vp, vp, i, i64, vp, i, i64, d, vp, i, i64, vp]
    L.HCORE_duncompress_batched_strided.argtypes = [vp, i, i, i, vp, vp, i, i64, vp, vp, i, i64, vp, i, i64, vp, vp]
    L.HCORE_zuncompress_batched_strided.argtypes = [vp, i, i, i, vp, vp, i, i64, vp, vp, i, i64, vp, i, i64, vp, vp]
    L.HCORE_dtrsm_batched_strided.argtypes = [vp, i, i, i, i, i, i, i, d, vp, i, i64, vp, vp, i, i64, vp]
    L.HCORE_dpotrf_batched_strided.argtypes = [vp, i, i, i, vp, i, i64, vp, vp]
    _lib = L
    return L
