"""TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_ref/libhcore_ref.so`` -- the unmodified reference
(ecrc/hcore) compiled from ``/root/reference/src`` by ``oracle/Makefile`` against
SciPy's OpenBLAS.  Used to pin ``oracle/tlr_oracle.py``, to generate
``tests/golden`` fixtures and as the CPU baseline (``cpu_baseline.kind="reference"``).

Prototypes bound (reference file:line):
  HCORE_dgemm   src/include/hcore_d.h:25-46     HCORE_sgemm  src/include/hcore_s.h:25-46
  HCORE_zgemm   src/include/hcore_z.h:27-48     HCORE_dgemmcd src/include/hcore_d.h:68-86
  HCORE_dsyrk   src/include/hcore_d.h:107-116   flop_counter src/misc/include/flop_util_structs.h:5-36
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhcore_ref.so")

CblasNoTrans, CblasTrans = 111, 112
HCORE_Lower = 122


class FlopCounter(C.Structure):
    """Mirror of ``flop_counter`` (reference src/misc/include/flop_util_structs.h:5-36)."""
    _fields_ = [("pad1", C.c_ulong), ("pad2", C.c_ulong), ("copy", C.c_ulong), ("flop", C.c_ulong),
                ("tid", C.c_int), ("potrf", C.c_ulong), ("trsm", C.c_ulong), ("syrk", C.c_ulong),
                ("update", C.c_ulong), ("geqrfu", C.c_ulong), ("geqrfv", C.c_ulong),
                ("geqrfv_gemm", C.c_ulong), ("gesvd", C.c_ulong), ("gesvd_gemm", C.c_ulong),
                ("gesvd_lasetu", C.c_ulong), ("gesvd_lasetv", C.c_ulong), ("gesvd_lacpyu", C.c_ulong),
                ("gesvd_lacpyv", C.c_ulong), ("gesvd_dscal", C.c_ulong), ("ormqru", C.c_ulong),
                ("ormqru_laset", C.c_ulong), ("ormqru_lacpy", C.c_ulong), ("ormqrv", C.c_ulong),
                ("ormqrv_laset", C.c_ulong), ("ormqrv_lacpy", C.c_ulong), ("omatcopy", C.c_ulong),
                ("numthreads", C.c_int), ("pad3", C.c_ulong), ("pad4", C.c_ulong)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        dp, fp, vp, i, d, f = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p, C.c_int,
                               C.c_double, C.c_float)
        fl = C.POINTER(FlopCounter)
        L.HCORE_dgemm.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, dp, dp, i, d, dp, dp, dp, i, i, i, d, dp, fl]
        L.HCORE_dgemm.restype = None
        L.HCORE_sgemm.argtypes = [i, i, i, i, f, fp, fp, fp, i, fp, fp, fp, i, f, fp, fp, fp, i, i, i, f, fp, fl]
        L.HCORE_sgemm.restype = None
        # double _Complex scalars are passed by value: two doubles in xmm0/xmm1 on SysV x86-64,
        # which is what (c_double, c_double) produces for the leading floating-point arguments.
        L.HCORE_zgemm.argtypes = [i, i, i, i, d, d, vp, vp, dp, i, vp, vp, dp, i, d, d, vp, vp, dp, i, i, i, d, vp, fl]
        L.HCORE_zgemm.restype = None
        L.HCORE_dgemmcd.argtypes = [i, i, i, d, dp, dp, i, i, dp, dp, i, i, d, dp, i, i, dp, fl]
        L.HCORE_dgemmcd.restype = None
        L.HCORE_dsyrk.argtypes = [i, i, i, i, d, dp, i, dp, i, d, dp, i, dp, fl]
        L.HCORE_dsyrk.restype = None
        # reference src/include/hcore_d.h:47-67, 87-106, 117-138; src/include/hcore_z.h:76-88
        L.HCORE_dgemm_fast.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, dp, dp, i, d, dp, dp, dp, i, i, i, d, dp]
        L.HCORE_dgemm_fast.restype = None
        L.HCORE_dtrsm.argtypes = [i, i, i, i, i, i, d, dp, i, dp, i, fl]
        L.HCORE_dtrsm.restype = None
        L.HCORE_dpotrf.argtypes = [i, i, dp, i, C.POINTER(C.c_int), fl]
        L.HCORE_dpotrf.restype = None
        L.HCORE_dgemmbdcd.argtypes = [i, i, i, i, d, dp, dp, dp, i, dp, i, d, dp, i, dp, fl]
        L.HCORE_dgemmbdcd.restype = None
        L.HCORE_duncompress.argtypes = [i, i, i, i, d, dp, dp, i, dp, dp, i, d, dp, i]
        L.HCORE_duncompress.restype = None
        L.HCORE_zuncompress.argtypes = [i, i, i, i, d, d, vp, dp, i, vp, dp, i, d, d, vp, i]
        L.HCORE_zuncompress.restype = None
        L.oracle_ref_set_blas_threads.argtypes = [i]
        L.oracle_ref_blas_config.restype = C.c_char_p
        _lib = L
    return _lib


def set_blas_threads(n: int):
    lib().oracle_ref_set_blas_threads(int(n))


def blas_config() -> str:
    return lib().oracle_ref_blas_config().decode()


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def hcore_gemm(au, av, bu, bv, cu, cv, crk, *, alpha, beta, rk, maxrk, acc):
    """Run the reference ``HCORE_{d,s}gemm`` on ONE tile.

    ``au, av`` are (a, nb) arrays (column-major nb x a), likewise ``bu, bv``;
    ``cu, cv`` are (>=maxrk, nb) buffers whose first ``crk`` columns hold C.
    Returns ``(cu_out, cv_out, new_rank, flops_update)`` (copies; inputs untouched).
    """
    L = lib()
    dt = au.dtype
    nb = au.shape[1]
    a, b = au.shape[0], bu.shape[0]
    cu2 = np.zeros((max(maxrk, cu.shape[0]), nb), dt); cu2[:crk] = cu[:crk]
    cv2 = np.zeros_like(cu2); cv2[:crk] = cv[:crk]
    au, av, bu, bv = (np.ascontiguousarray(x) for x in (au, av, bu, bv))
    fl = FlopCounter()
    if dt == np.float64:
        rks = [np.array([float(x)]) for x in (a, b, crk)]
        L.HCORE_dgemm(CblasNoTrans, CblasTrans, nb, nb, float(alpha), _p(au, C.c_double), _p(av, C.c_double),
                      _p(rks[0], C.c_double), nb, _p(bu, C.c_double), _p(bv, C.c_double), _p(rks[1], C.c_double), nb,
                      float(beta), _p(cu2, C.c_double), _p(cv2, C.c_double), _p(rks[2], C.c_double), nb,
                      int(rk), int(maxrk), float(acc), None, C.byref(fl))
    elif dt == np.float32:
        rks = [np.array([float(x)], np.float32) for x in (a, b, crk)]
        L.HCORE_sgemm(CblasNoTrans, CblasTrans, nb, nb, float(alpha), _p(au, C.c_float), _p(av, C.c_float),
                      _p(rks[0], C.c_float), nb, _p(bu, C.c_float), _p(bv, C.c_float), _p(rks[1], C.c_float), nb,
                      float(beta), _p(cu2, C.c_float), _p(cv2, C.c_float), _p(rks[2], C.c_float), nb,
                      int(rk), int(maxrk), float(acc), None, C.byref(fl))
    elif dt == np.complex128:
        rks = [np.array([float(x)]) for x in (a, b, crk)]
        al, be = complex(alpha), complex(beta)
        L.HCORE_zgemm(CblasNoTrans, CblasTrans, nb, nb, al.real, al.imag, au.ctypes.data, av.ctypes.data,
                      _p(rks[0], C.c_double), nb, bu.ctypes.data, bv.ctypes.data, _p(rks[1], C.c_double), nb,
                      be.real, be.imag, cu2.ctypes.data, cv2.ctypes.data, _p(rks[2], C.c_double), nb,
                      int(rk), int(maxrk), float(acc), None, C.byref(fl))
    else:
        raise TypeError(dt)
    r = int(rks[2][0])
    return cu2, cv2, r, int(fl.update)


def hcore_dgemmcd(au, av, bu, bv, cd):
    """Reference ``HCORE_dgemmcd`` on one tile; ``cd`` is (nb, nb) column-major. Returns a copy."""
    L = lib()
    nb = cd.shape[0]
    a, b = au.shape[0], bu.shape[0]
    out = np.array(cd, dtype=np.float64, order="C", copy=True)
    au, av, bu, bv = (np.ascontiguousarray(x, np.float64) for x in (au, av, bu, bv))
    fl = FlopCounter()
    L.HCORE_dgemmcd(CblasNoTrans, CblasTrans, nb, -1.0, _p(au, C.c_double), _p(av, C.c_double), a, nb,
                    _p(bu, C.c_double), _p(bv, C.c_double), b, nb, 1.0, _p(out, C.c_double), nb,
                    max(a, b), None, C.byref(fl))
    return out


def hcore_dsyrk(au, av, cd, alpha=-1.0):
    """Reference ``HCORE_dsyrk`` on one tile. Returns ``(cd_out, flops_syrk)``."""
    L = lib()
    nb = cd.shape[0]
    k = au.shape[0]
    out = np.array(cd, dtype=np.float64, order="C", copy=True)
    au, av = (np.ascontiguousarray(x, np.float64) for x in (au, av))
    fl = FlopCounter()
    L.HCORE_dsyrk(HCORE_Lower, CblasNoTrans, nb, k, float(alpha), _p(au, C.c_double), nb,
                  _p(av, C.c_double), nb, 1.0, _p(out, C.c_double), nb, None, C.byref(fl))
    return out, int(fl.syrk)


# ---- the remaining tile kernels of one TLR-Cholesky step + HCORE_dgemm_fast ----------------

def hcore_dgemm_fast(au, av, bu, bv, cu, cv, crk, *, alpha, beta, rk, maxrk, acc):
    """Reference ``HCORE_dgemm_fast`` on ONE tile (needs the globals use_trmm = use_scratch = 1, which this sets and
    restores, and a scratch array: reference src/compute/hcore_dgemm_fast.c:82-83, :117-300)."""
    L = lib()
    nb, a, b = au.shape[1], au.shape[0], bu.shape[0]
    n = a + crk
    cu2 = np.zeros((max(2 * maxrk, cu.shape[0], n), nb)); cu2[:crk] = cu[:crk]
    cv2 = np.zeros_like(cu2); cv2[:crk] = cv[:crk]
    au, av, bu, bv = (np.ascontiguousarray(x, np.float64) for x in (au, av, bu, bv))
    work = np.zeros(4 * nb * 2 * maxrk + 2 * nb + maxrk * maxrk + 3 * n * n + 2 * nb * n + n + 64 * n + 4096)
    rks = [np.array([float(x)]) for x in (a, b, crk)]
    trmm, scr = C.c_int.in_dll(L, "use_trmm"), C.c_int.in_dll(L, "use_scratch")
    old = trmm.value, scr.value
    trmm.value, scr.value = 1, 1
    try:
        L.HCORE_dgemm_fast(CblasNoTrans, CblasTrans, nb, nb, float(alpha), _p(au, C.c_double), _p(av, C.c_double),
                           _p(rks[0], C.c_double), nb, _p(bu, C.c_double), _p(bv, C.c_double), _p(rks[1], C.c_double), nb,
                           float(beta), _p(cu2, C.c_double), _p(cv2, C.c_double), _p(rks[2], C.c_double), nb,
                           int(rk), int(maxrk), float(acc), _p(work, C.c_double))
    finally:
        trmm.value, scr.value = old
    return cu2, cv2, int(rks[2][0])


def hcore_dtrsm(side, uplo, trans, diag, alpha, a, bv):
    """Reference ``HCORE_dtrsm``; ``a`` (n, n) and ``bv`` (Brank, M) column-major arrays. Returns ``(bv_out, flops_trsm)``."""
    L = lib()
    brank, m = bv.shape
    a = np.ascontiguousarray(a, np.float64)
    out = np.array(bv, dtype=np.float64, order="C", copy=True)
    fl = FlopCounter()
    L.HCORE_dtrsm(int(side), int(uplo), int(trans), int(diag), m, brank, float(alpha), _p(a, C.c_double), a.shape[1],
                  _p(out, C.c_double), m, C.byref(fl))
    return out, int(fl.trsm)


def hcore_dpotrf(uplo, a):
    """Reference ``HCORE_dpotrf``; ``a`` (N, N) column-major. Returns ``(a_out, info, flops_potrf)``."""
    L = lib()
    n = a.shape[0]
    out = np.array(a, dtype=np.float64, order="C", copy=True)
    info = C.c_int(-1)
    fl = FlopCounter()
    L.HCORE_dpotrf(int(uplo), n, _p(out, C.c_double), n, C.byref(info), C.byref(fl))
    return out, info.value, int(fl.potrf)


def hcore_dgemmbdcd(trans_a, alpha, au, av, bd, beta, cd):
    """Reference ``HCORE_dgemmbdcd``; ``au, av`` (a, M), ``bd``, ``cd`` (N, M) column-major. Returns ``(cd_out, flops_update)``."""
    L = lib()
    a, m = au.shape
    n = bd.shape[0]
    au, av, bd = (np.ascontiguousarray(x, np.float64) for x in (au, av, bd))
    out = np.array(cd, dtype=np.float64, order="C", copy=True)
    ark = np.array([float(a)])
    fl = FlopCounter()
    L.HCORE_dgemmbdcd(int(trans_a), CblasNoTrans, m, n, float(alpha), _p(au, C.c_double), _p(av, C.c_double),
                      _p(ark, C.c_double), m, _p(bd, C.c_double), m, float(beta), _p(out, C.c_double), m, None, C.byref(fl))
    return out, int(fl.update)


def hcore_uncompress(au, bv):
    """Reference ``HCORE_duncompress`` / ``HCORE_zuncompress`` (by dtype); ``au`` (k, M), ``bv`` (k, N). Returns CD (N, M)."""
    L = lib()
    k, m = au.shape
    n = bv.shape[1]
    dt = au.dtype
    au, bv = np.ascontiguousarray(au), np.ascontiguousarray(bv)
    out = np.full((n, m), np.nan, dt)
    rk = np.array([float(k)]), np.array([float(k)])
    if dt == np.float64:
        L.HCORE_duncompress(CblasNoTrans, CblasTrans, m, n, 7.0, _p(au, C.c_double), _p(rk[0], C.c_double), m,
                            _p(bv, C.c_double), _p(rk[1], C.c_double), n, 3.0, _p(out, C.c_double), m)
    else:
        L.HCORE_zuncompress(CblasNoTrans, CblasTrans, m, n, 7.0, 1.0, au.ctypes.data, _p(rk[0], C.c_double), m,
                            bv.ctypes.data, _p(rk[1], C.c_double), n, 3.0, -1.0, out.ctypes.data, m)
    return out
