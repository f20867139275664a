"""Host-side mirror of the reference interface for the TLR update path.

Names and argument meaning follow the reference C API (``src/include/hcore_d.h``):
``hcore_dgemm`` <-> ``HCORE_dgemm`` (:25-46), ``hcore_dgemmcd`` <-> ``HCORE_dgemmcd`` (:68-86),
``hcore_dsyrk`` <-> ``HCORE_dsyrk`` (:107-116).  The legacy functions take HOST numpy
arrays and call the drop-in C symbols of ``libhcore_b200.so`` directly; ``Context`` wraps
the batched device-pointer entry points for torch CUDA tensors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

HCORE_NoTrans, HCORE_Trans = 111, 112
HCORE_Upper, HCORE_Lower = 121, 122
HCORE_NonUnit, HCORE_Unit = 131, 132
HCORE_Left, HCORE_Right = 141, 142

STATUS_NAMES = {0: "OK", 1: "ZERO_RANK", 2: "MAXRK", 3: "UNSUPPORTED", 4: "NEWRANK", 5: "NO_CONVERGE", 6: "RANK_MISMATCH"}


class FlopCounter(C.Structure):
    """``flop_counter`` (reference src/misc/include/flop_util_structs.h:5-36; include/hcore_b200_flops.h)."""
    _fields_ = [("pad1", C.c_ulong), ("pad2", C.c_ulong), ("copy", C.c_ulong), ("flop", C.c_ulong),
                ("tid", C.c_int), ("potrf", C.c_ulong), ("trsm", C.c_ulong), ("syrk", C.c_ulong),
                ("update", C.c_ulong), ("geqrfu", C.c_ulong), ("geqrfv", C.c_ulong),
                ("geqrfv_gemm", C.c_ulong), ("gesvd", C.c_ulong), ("gesvd_gemm", C.c_ulong),
                ("gesvd_lasetu", C.c_ulong), ("gesvd_lasetv", C.c_ulong), ("gesvd_lacpyu", C.c_ulong),
                ("gesvd_lacpyv", C.c_ulong), ("gesvd_dscal", C.c_ulong), ("ormqru", C.c_ulong),
                ("ormqru_laset", C.c_ulong), ("ormqru_lacpy", C.c_ulong), ("ormqrv", C.c_ulong),
                ("ormqrv_laset", C.c_ulong), ("ormqrv_lacpy", C.c_ulong), ("omatcopy", C.c_ulong),
                ("numthreads", C.c_int), ("pad3", C.c_ulong), ("pad4", C.c_ulong)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check_colmajor(x, name):
    if x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name}: need a C-contiguous float64 array holding the column-major matrix as (cols, ld)")


# ---- legacy single-tile calls (host arrays) -------------------------------------------------

def hcore_dgemm(transA, transB, M, N, alpha, AU, AV, Ark, LDA, BU, BV, Brk, LDB, beta, CU, CV, Crk, LDC,
                rk, maxrk, acc, work=None, flops=None):
    """``HCORE_dgemm`` (reference src/compute/hcore_dgemm.c:703-1012), same arguments.

    Arrays are numpy ``(cols, ld)`` float64 (i.e. the column-major matrices the C code sees);
    ``Ark, Brk, Crk`` are 1-element float64 arrays (the reference passes ranks as ``double*``);
    ``CU, CV`` are overwritten with the new factors and ``Crk[0]`` with the new rank.
    Errors terminate the process like the reference (``exit(-1)``)."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
        _check_colmajor(x, nm)
    fl = C.byref(flops) if flops is not None else None
    L.HCORE_dgemm(int(transA), int(transB), int(M), int(N), float(alpha), _dp(AU), _dp(AV), _dp(Ark), int(LDA),
                  _dp(BU), _dp(BV), _dp(Brk), int(LDB), float(beta), _dp(CU), _dp(CV), _dp(Crk), int(LDC),
                  int(rk), int(maxrk), float(acc), None, fl)


def hcore_sgemm(transA, transB, M, N, alpha, AU, AV, Ark, LDA, BU, BV, Brk, LDB, beta, CU, CV, Crk, LDC,
                rk, maxrk, acc, work=None, flops=None):
    """``HCORE_sgemm`` (reference src/compute/hcore_sgemm.c:703-1012): float32 arrays, ranks as
    1-element float32 arrays, otherwise as :func:`hcore_dgemm`."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV), ("Ark", Ark), ("Brk", Brk),
                  ("Crk", Crk)):
        if x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{nm}: need a C-contiguous float32 array")
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    L.HCORE_sgemm(int(transA), int(transB), int(M), int(N), float(alpha), fp(AU), fp(AV), fp(Ark), int(LDA),
                  fp(BU), fp(BV), fp(Brk), int(LDB), float(beta), fp(CU), fp(CV), fp(Crk), int(LDC),
                  int(rk), int(maxrk), float(acc), None, C.byref(flops) if flops is not None else None)


def hcore_zgemm(transA, transB, M, N, alpha, AU, AV, Ark, LDA, BU, BV, Brk, LDB, beta, CU, CV, Crk, LDC,
                rk, maxrk, acc, work=None, flops=None):
    """``HCORE_zgemm`` (reference src/compute/hcore_zgemm.c:663-975): complex128 arrays, complex
    ``alpha``/``beta``, ranks as 1-element float64 arrays; tiles are ``U @ V.T`` (plain transpose)."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
        if x.dtype != np.complex128 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{nm}: need a C-contiguous complex128 array")
    vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    L.HCORE_zgemm(int(transA), int(transB), int(M), int(N), _lib.DComplex.of(alpha), vp(AU), vp(AV), _dp(Ark), int(LDA),
                  vp(BU), vp(BV), _dp(Brk), int(LDB), _lib.DComplex.of(beta), vp(CU), vp(CV), _dp(Crk), int(LDC),
                  int(rk), int(maxrk), float(acc), None, C.byref(flops) if flops is not None else None)


def hcore_dgemmcd(transA, transB, M, alpha, Au, Av, Arank, LDA, Bu, Bv, Brank, LDB, beta, CD, LDC, maxrank,
                  work=None, flops=None):
    """``HCORE_dgemmcd`` (reference src/compute/hcore_dgemmcd.c:36-107): ``CD -= (Au Av^T)(Bu Bv^T)^T``."""
    L = _lib.load()
    for nm, x in (("Au", Au), ("Av", Av), ("Bu", Bu), ("Bv", Bv), ("CD", CD)):
        _check_colmajor(x, nm)
    L.HCORE_dgemmcd(int(transA), int(transB), int(M), float(alpha), _dp(Au), _dp(Av), int(Arank), int(LDA),
                    _dp(Bu), _dp(Bv), int(Brank), int(LDB), float(beta), _dp(CD), int(LDC), int(maxrank), None,
                    C.byref(flops) if flops is not None else None)


def hcore_dsyrk(uplo, trans, N, K, alpha, AU, LDAU, AV, LDAV, beta, CD, LDCD, work=None, flops=None):
    """``HCORE_dsyrk`` (reference src/compute/hcore_dsyrk.c:47-189)."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("CD", CD)):
        _check_colmajor(x, nm)
    L.HCORE_dsyrk(int(uplo), int(trans), int(N), int(K), float(alpha), _dp(AU), int(LDAU), _dp(AV), int(LDAV),
                  float(beta), _dp(CD), int(LDCD), None, C.byref(flops) if flops is not None else None)


def hcore_dgemm_fast(transA, transB, M, N, alpha, AU, AV, Ark, LDA, BU, BV, Brk, LDB, beta, CU, CV, Crk, LDC,
                     rk, maxrk, acc, work=None):
    """``HCORE_dgemm_fast`` (reference src/compute/hcore_dgemm_fast.c:59-426): arguments of :func:`hcore_dgemm`
    without the flop counter; needs the library global ``use_trmm`` set to 1 (reference :82)."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
        _check_colmajor(x, nm)
    L.HCORE_dgemm_fast(int(transA), int(transB), int(M), int(N), float(alpha), _dp(AU), _dp(AV), _dp(Ark), int(LDA),
                       _dp(BU), _dp(BV), _dp(Brk), int(LDB), float(beta), _dp(CU), _dp(CV), _dp(Crk), int(LDC),
                       int(rk), int(maxrk), float(acc), None)


def set_global(name, value):
    """Sets one of the reference's link-level tunables (``use_trmm`` ...) in the loaded library."""
    C.c_int.in_dll(_lib.load(), name).value = int(value)


def hcore_dtrsm(side, uplo, transA, diag, M, Brank, alpha, A, LDA, BV, LDBV, flops=None):
    """``HCORE_dtrsm`` (reference src/compute/hcore_dtrsm.c:24-48): ``BV`` (Brank, LDBV) column-major M x Brank, in place."""
    L = _lib.load()
    _check_colmajor(A, "A"); _check_colmajor(BV, "BV")
    L.HCORE_dtrsm(int(side), int(uplo), int(transA), int(diag), int(M), int(Brank), float(alpha), _dp(A), int(LDA),
                  _dp(BV), int(LDBV), C.byref(flops) if flops is not None else None)


def hcore_dpotrf(uplo, N, A, LDA, flops=None):
    """``HCORE_dpotrf`` (reference src/compute/hcore_dpotrf.c:23-35): factors ``A`` in place, returns ``info``."""
    L = _lib.load()
    _check_colmajor(A, "A")
    info = C.c_int(-1)
    L.HCORE_dpotrf(int(uplo), int(N), _dp(A), int(LDA), C.byref(info), C.byref(flops) if flops is not None else None)
    return info.value


def hcore_dgemmbdcd(transA, transB, M, N, alpha, AU, AV, Ark, LDA, BD, LDB, beta, CD, LDC, work=None, flops=None):
    """``HCORE_dgemmbdcd`` (reference src/compute/hcore_dgemmbdcd.c:35-122): ``CD <- alpha AU (AV^T BD) + beta CD``."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("AV", AV), ("BD", BD), ("CD", CD)):
        _check_colmajor(x, nm)
    L.HCORE_dgemmbdcd(int(transA), int(transB), int(M), int(N), float(alpha), _dp(AU), _dp(AV), _dp(Ark), int(LDA),
                      _dp(BD), int(LDB), float(beta), _dp(CD), int(LDC), None, C.byref(flops) if flops is not None else None)


def hcore_duncompress(transA, transB, M, N, alpha, AU, Ark, LDA, BV, Brk, LDB, beta, CD, LDC):
    """``HCORE_duncompress`` (reference src/compute/hcore_duncompress.c:33-85): ``CD <- AU BV^T``."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("BV", BV), ("CD", CD)):
        _check_colmajor(x, nm)
    L.HCORE_duncompress(int(transA), int(transB), int(M), int(N), float(alpha), _dp(AU), _dp(Ark), int(LDA), _dp(BV), _dp(Brk),
                        int(LDB), float(beta), _dp(CD), int(LDC))


def hcore_zuncompress(transA, transB, M, N, alpha, AU, Ark, LDA, BV, Brk, LDB, beta, CD, LDC):
    """``HCORE_zuncompress`` (reference src/compute/hcore_zuncompress.c:35-95): complex128 arrays, plain transpose."""
    L = _lib.load()
    for nm, x in (("AU", AU), ("BV", BV), ("CD", CD)):
        if x.dtype != np.complex128 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{nm}: need a C-contiguous complex128 array")
    vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    L.HCORE_zuncompress(int(transA), int(transB), int(M), int(N), _lib.DComplex.of(alpha), vp(AU), _dp(Ark), int(LDA), vp(BV),
                        _dp(Brk), int(LDB), _lib.DComplex.of(beta), vp(CD), int(LDC))


def gemm_flops(m, a, b, c, r):
    """Reference flop model of one ``HCORE_dgemm`` call (``flops->update``)."""
    return int(_lib.load().hcore_b200_gemm_flops(int(m), int(a), int(b), int(c), int(r)))


def shard_range(ntiles, world, rank):
    first, count = C.c_int64(), C.c_int64()
    _lib.load().hcore_b200_shard_range(int(ntiles), int(world), int(rank), C.byref(first), C.byref(count))
    return first.value, count.value


# ---- batched context -----------------------------------------------------------------------------

class HcoreError(RuntimeError):
    pass


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Context:
    """Owns a ``hcore_b200_handle_t`` (device workspace) on one CUDA device."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise HcoreError("hcore_b200 needs a CUDA device (no CPU fallback)")
        self._L = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        rc = self._L.hcore_b200_create(C.byref(h), self.device)
        if rc != 0:
            raise HcoreError(f"hcore_b200_create failed ({rc})")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.hcore_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers --
    def _check(self, rc):
        if rc != 0:
            raise HcoreError(f"hcore_b200 error {rc}: {self._L.hcore_b200_last_error(self._h).decode()}")

    def _dev_f64(self, t, name):
        import torch
        if t.dtype != torch.float64 or not t.is_cuda or t.device.index != self.device or not t.is_contiguous():
            raise TypeError(f"{name}: need a contiguous float64 tensor on cuda:{self.device}")

    def _dev_i32(self, t, name):
        import torch
        if t.dtype != torch.int32 or not t.is_cuda or t.device.index != self.device or not t.is_contiguous():
            raise TypeError(f"{name}: need a contiguous int32 tensor on cuda:{self.device}")

    @staticmethod
    def _stream_ptr(stream):
        import torch
        s = torch.cuda.current_stream() if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    @property
    def launches(self):
        return int(self._L.hcore_b200_launch_count(self._h))

    def set_chunk(self, n):
        self._check(self._L.hcore_b200_set_chunk(self._h, int(n)))

    def set_stage_timing(self, on):
        self._check(self._L.hcore_b200_set_stage_timing(self._h, 1 if on else 0))

    def stage_ms(self):
        return [float(self._L.hcore_b200_stage_ms(self._h, i)) for i in range(4)]

    def debug_set_stages(self, n):
        self._check(self._L.hcore_b200_debug_set_stages(self._h, int(n)))

    def debug_fetch(self, what, tile, nelem):
        out = np.zeros(nelem, np.float64)
        self._check(self._L.hcore_b200_debug_fetch(self._h, int(what), int(tile), out.ctypes.data, nelem))
        return out

    # -- batched TLR update --
    def dgemm_batched(self, alpha, AU, AV, Ark, BU, BV, Brk, beta, CU, CV, Crk, *, rk, maxrk, acc,
                      status=None, stream=None):
        """Batched ``HCORE_dgemm`` on strided device tensors.

        ``AU, AV``: (T, ka, nb) float64 CUDA tensors (tile t is the column-major nb x ka matrix),
        ``BU, BV``: (T, kb, nb); ``CU, CV``: (T, >=maxrk, nb) in/out; ``Ark, Brk, Crk``: (T,) int32
        (``Crk`` in/out); ``status``: optional (T,) int32 out."""
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
            self._dev_f64(x, nm)
        for nm, x in (("Ark", Ark), ("Brk", Brk), ("Crk", Crk)):
            self._dev_i32(x, nm)
        if status is not None:
            self._dev_i32(status, "status")
        kb, kc = BU.shape[1], CU.shape[1]
        if kc < maxrk:
            raise ValueError("CU/CV need room for maxrk columns")
        rc = self._L.HCORE_dgemm_batched_strided(
            self._h, T, nb, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), nb, ka * nb,
            _ptr(BU), _ptr(BV), _ptr(Brk), nb, kb * nb, float(beta), _ptr(CU), _ptr(CV), _ptr(Crk), nb, kc * nb,
            int(rk), int(maxrk), float(acc), _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    # -- split-phase batched update (reference HCORE_dgemm_qr_svd / HCORE_dgemm_ormqr) --
    def dgemm_batched_qr_svd(self, alpha, AU, AV, Ark, BU, BV, Brk, beta, CU, CV, Crk, newrank, *, rk, maxrk, acc,
                             status=None, stream=None):
        """First half: stack, QR, SVD, truncation.  Writes the revealed ranks to ``newrank`` (T,) int32 and
        leaves ``CU, CV, Crk`` untouched; the reflectors stay in the context until ``dgemm_batched_ormqr``."""
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
            self._dev_f64(x, nm)
        for nm, x in (("Ark", Ark), ("Brk", Brk), ("Crk", Crk), ("newrank", newrank)):
            self._dev_i32(x, nm)
        if status is not None:
            self._dev_i32(status, "status")
        kb, kc = BU.shape[1], CU.shape[1]
        rc = self._L.HCORE_dgemm_batched_qr_svd_strided(
            self._h, T, nb, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), nb, ka * nb,
            _ptr(BU), _ptr(BV), _ptr(Brk), nb, kb * nb, float(beta), _ptr(CU), _ptr(CV), _ptr(Crk), nb, kc * nb,
            int(rk), int(maxrk), float(acc), _ptr(newrank), _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    def dgemm_batched_ormqr(self, CU, CV, Crk, stream=None):
        """Second half: forms the new factors in ``CU, CV`` ((T, >= new rank, nb), may be freshly allocated at
        the exact size) and sets ``Crk``."""
        T, kc, nb = CU.shape
        self._dev_f64(CU, "CU"); self._dev_f64(CV, "CV"); self._dev_i32(Crk, "Crk")
        self._check(self._L.HCORE_dgemm_batched_ormqr_strided(self._h, T, nb, _ptr(CU), _ptr(CV), _ptr(Crk), nb, kc * nb,
                                                              self._stream_ptr(stream)))

    def dgemm_batched_b_dense(self, alpha, AU, AV, Ark, B, beta, CU, CV, Crk, *, rk, maxrk, acc, status=None, stream=None):
        """B-dense variant (reference ``HCORE_dgemm_qr_svd_b_dense`` + ``HCORE_dgemm_ormqr``): ``B`` is (T, nb, nb),
        tile t the column-major dense nb x nb matrix; ``C <- alpha (AU AV^T) B^T + beta C``."""
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("B", B), ("CU", CU), ("CV", CV)):
            self._dev_f64(x, nm)
        for nm, x in (("Ark", Ark), ("Crk", Crk)):
            self._dev_i32(x, nm)
        if status is not None:
            self._dev_i32(status, "status")
        kc = CU.shape[1]
        rc = self._L.HCORE_dgemm_batched_b_dense_strided(
            self._h, T, nb, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), nb, ka * nb, _ptr(B), nb, nb * nb, float(beta),
            _ptr(CU), _ptr(CV), _ptr(Crk), nb, kc * nb, int(rk), int(maxrk), float(acc), _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    def sgemm_batched(self, alpha, AU, AV, Ark, BU, BV, Brk, beta, CU, CV, Crk, *, rk, maxrk, acc,
                      status=None, stream=None):
        """Batched ``HCORE_sgemm``: as :meth:`dgemm_batched` with float32 factor tensors."""
        import torch
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
            if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous():
                raise TypeError(f"{nm}: need a contiguous float32 CUDA tensor")
        for nm, x in (("Ark", Ark), ("Brk", Brk), ("Crk", Crk)):
            self._dev_i32(x, nm)
        kb, kc = BU.shape[1], CU.shape[1]
        if kc < maxrk:
            raise ValueError("CU/CV need room for maxrk columns")
        rc = self._L.HCORE_sgemm_batched_strided(
            self._h, T, nb, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), nb, ka * nb,
            _ptr(BU), _ptr(BV), _ptr(Brk), nb, kb * nb, float(beta), _ptr(CU), _ptr(CV), _ptr(Crk), nb, kc * nb,
            int(rk), int(maxrk), float(acc), _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    def zgemm_batched(self, alpha, AU, AV, Ark, BU, BV, Brk, beta, CU, CV, Crk, *, rk, maxrk, acc,
                      status=None, stream=None):
        """Batched ``HCORE_zgemm``: complex128 factor tensors (T, k, nb), complex scalars."""
        import torch
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CU", CU), ("CV", CV)):
            if x.dtype != torch.complex128 or not x.is_cuda or not x.is_contiguous():
                raise TypeError(f"{nm}: need a contiguous complex128 CUDA tensor")
        for nm, x in (("Ark", Ark), ("Brk", Brk), ("Crk", Crk)):
            self._dev_i32(x, nm)
        kb, kc = BU.shape[1], CU.shape[1]
        if kc < maxrk:
            raise ValueError("CU/CV need room for maxrk columns")
        rc = self._L.HCORE_zgemm_batched_strided(
            self._h, T, nb, _lib.DComplex.of(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), nb, ka * nb,
            _ptr(BU), _ptr(BV), _ptr(Brk), nb, kb * nb, _lib.DComplex.of(beta), _ptr(CU), _ptr(CV), _ptr(Crk), nb,
            kc * nb, int(rk), int(maxrk), float(acc), _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    def dgemm_batched_host(self, alpha, AU, AV, Ark, BU, BV, Brk, beta, CU, CV, Crk, *, rk, maxrk, acc, status=None):
        """Same on HOST tensors/arrays (pinned torch tensors recommended): the end-to-end path.  float32 arrays go to
        ``HCORE_sgemm_batched_host``, complex128 ones to ``HCORE_zgemm_batched_host``."""
        T, ka, nb = AU.shape
        kb, kc = BU.shape[1], CU.shape[1]

        def hp(x):
            return None if x is None else C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data)

        dt = str(AU.dtype).replace("torch.", "")
        if dt == "complex128":
            fn, al, be = self._L.HCORE_zgemm_batched_host, _lib.DComplex.of(alpha), _lib.DComplex.of(beta)
        elif dt == "float32":
            fn, al, be = self._L.HCORE_sgemm_batched_host, float(alpha), float(beta)
        else:
            fn, al, be = self._L.HCORE_dgemm_batched_host, float(alpha), float(beta)
        rc = fn(self._h, T, nb, al, hp(AU), hp(AV), hp(Ark), nb, ka * nb, hp(BU), hp(BV), hp(Brk), nb, kb * nb,
                be, hp(CU), hp(CV), hp(Crk), nb, kc * nb, int(rk), int(maxrk), float(acc), hp(status))
        self._check(rc)

    def gather_status(self, nccl_comm, world, rank, ntiles, rank_local, status_local, rank_all, status_all, stream=None):
        """``hcore_b200_gather_status``: the NCCL all-gather of {new rank, status} through the C ABI (``nccl_comm`` is
        a raw ``ncclComm_t`` as an integer / void pointer, or None with world == 1)."""
        self._check(self._L.hcore_b200_gather_status(self._h, C.c_void_p(nccl_comm) if nccl_comm else None, int(world), int(rank),
                                                     int(ntiles), _ptr(rank_local), _ptr(status_local), _ptr(rank_all),
                                                     _ptr(status_all), self._stream_ptr(stream)))

    # -- batched dense-C updates --
    def dgemmcd_batched(self, AU, AV, Ark, BU, BV, Brk, CD, stream=None):
        """Batched ``HCORE_dgemmcd``: ``CD[t] -= (AU AV^T)(BU BV^T)^T``; ``CD``: (T, nb, nb) column-major."""
        T, ka, nb = AU.shape
        kb = BU.shape[1]
        for nm, x in (("AU", AU), ("AV", AV), ("BU", BU), ("BV", BV), ("CD", CD)):
            self._dev_f64(x, nm)
        rc = self._L.HCORE_dgemmcd_batched_strided(self._h, T, nb, _ptr(AU), _ptr(AV), _ptr(Ark), ka * nb,
                                                   _ptr(BU), _ptr(BV), _ptr(Brk), kb * nb, _ptr(CD), nb * nb,
                                                   self._stream_ptr(stream))
        self._check(rc)

    def dsyrk_batched(self, alpha, AU, AV, Ark, CD, stream=None):
        """Batched ``HCORE_dsyrk``: ``CD[t] += alpha (AU AV^T)(AU AV^T)^T``, strict upper triangle zeroed."""
        T, ka, nb = AU.shape
        for nm, x in (("AU", AU), ("AV", AV), ("CD", CD)):
            self._dev_f64(x, nm)
        rc = self._L.HCORE_dsyrk_batched_strided(self._h, T, nb, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark),
                                                 ka * nb, _ptr(CD), nb * nb, self._stream_ptr(stream))
        self._check(rc)

    # -- the remaining tile kernels of one TLR-Cholesky step --
    def dtrsm_batched(self, side, uplo, transA, diag, alpha, A, BV, Brk=None, stream=None):
        """Batched ``HCORE_dtrsm``: ``A`` (T or 1, n, n) triangular tiles (one shared tile when its first dimension
        is 1), ``BV`` (T, kmax, M) in place; ``Brk`` optional (T,) int32 ranks."""
        T, kb, M = BV.shape
        self._dev_f64(A, "A"); self._dev_f64(BV, "BV")
        if Brk is not None:
            self._dev_i32(Brk, "Brk")
        na = A.shape[-1]
        strideA = 0 if A.shape[0] == 1 else na * na
        rc = self._L.HCORE_dtrsm_batched_strided(self._h, T, int(side), int(uplo), int(transA), int(diag), M, kb, float(alpha),
                                                 _ptr(A), na, strideA, _ptr(BV), _ptr(Brk), M, kb * M, self._stream_ptr(stream))
        self._check(rc)

    def dpotrf_batched(self, uplo, A, info, stream=None):
        """Batched ``HCORE_dpotrf``: ``A`` (T, N, N) column-major tiles in place, ``info`` (T,) int32 out."""
        T, N, _ = A.shape
        self._dev_f64(A, "A"); self._dev_i32(info, "info")
        self._check(self._L.HCORE_dpotrf_batched_strided(self._h, T, int(uplo), N, _ptr(A), N, N * N, _ptr(info),
                                                         self._stream_ptr(stream)))

    def dgemmbdcd_batched(self, alpha, AU, AV, Ark, BD, beta, CD, stream=None):
        """Batched ``HCORE_dgemmbdcd``: ``CD[t] <- alpha AU (AV^T BD) + beta CD``; ``BD``, ``CD``: (T, N, M) column-major M x N."""
        T, ka, M = AU.shape
        N = BD.shape[1]
        for nm, x in (("AU", AU), ("AV", AV), ("BD", BD), ("CD", CD)):
            self._dev_f64(x, nm)
        self._dev_i32(Ark, "Ark")
        rc = self._L.HCORE_dgemmbdcd_batched_strided(self._h, T, M, N, float(alpha), _ptr(AU), _ptr(AV), _ptr(Ark), M, ka * M,
                                                     _ptr(BD), M, N * M, float(beta), _ptr(CD), M, N * M, self._stream_ptr(stream))
        self._check(rc)

    def uncompress_batched(self, AU, Ark, BV, Brk, CD, status=None, stream=None):
        """Batched ``HCORE_duncompress`` / ``HCORE_zuncompress`` (by dtype): ``CD[t] <- AU BV^T``;
        ``AU`` (T, k, M), ``BV`` (T, k, N), ``CD`` (T, N, M)."""
        import torch
        T, ka, M = AU.shape
        kb, N = BV.shape[1], BV.shape[2]
        fn = self._L.HCORE_zuncompress_batched_strided if AU.dtype == torch.complex128 else self._L.HCORE_duncompress_batched_strided
        for nm, x in (("AU", AU), ("BV", BV), ("CD", CD)):
            if x.dtype != AU.dtype or not x.is_cuda or not x.is_contiguous():
                raise TypeError(f"{nm}: need a contiguous CUDA tensor of the same dtype")
        self._dev_i32(Ark, "Ark"); self._dev_i32(Brk, "Brk")
        rc = fn(self._h, T, M, N, _ptr(AU), _ptr(Ark), M, ka * M, _ptr(BV), _ptr(Brk), N, kb * N, _ptr(CD), M, N * M,
                _ptr(status), self._stream_ptr(stream))
        self._check(rc)

    def dcompress_batched(self, A, acc, maxrk, U, V, rank, status=None, power_iters=1, stream=None):
        """``HCORE_dcompress_batched_strided``: dense tiles ``A`` (T, nb, nb) column-major -> ``U, V`` (T, maxrk, nb) with
        ``A[t] ~ U[t].T @ V[t]`` in numpy orientation (U orthonormal), ``rank`` (T,) int32 out."""
        T, nb, _ = A.shape
        for nm, x in (("A", A), ("U", U), ("V", V)):
            self._dev_f64(x, nm)
        self._dev_i32(rank, "rank")
        if U.shape[1] < maxrk:
            raise ValueError("U/V need room for maxrk columns")
        rc = self._L.HCORE_dcompress_batched_strided(self._h, T, nb, _ptr(A), nb, nb * nb, float(acc), int(maxrk), int(power_iters),
                                                     _ptr(U), _ptr(V), nb, U.shape[1] * nb, _ptr(rank), _ptr(status),
                                                     self._stream_ptr(stream))
        self._check(rc)
