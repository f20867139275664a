"""GPU parity of the remaining TLR-Cholesky tile kernels and HCORE_dgemm_fast (row a10),
called through the C ABI: legacy symbols with the reference's prototypes against committed outputs of the compiled
reference (tests/golden/ref_cholstep.npz), batched entry points against the oracle at the benchmark sizes, and
size-independent properties (residuals, L L^T = A).  Floating point: tolerances are written at each assert.
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import cholstep_cases as cc
from hcore_b200 import api
from oracle import tlr_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available()
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_cholstep.npz"))


def cm(x):
    """math orientation -> the column-major array the C code sees"""
    return np.ascontiguousarray(x.T)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda:0")


# ---- HCORE_dtrsm ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("side,uplo,tr,dg", cc.TRSM_VARIANTS)
def test_dtrsm_legacy_all_variants_vs_reference(gold, side, uplo, tr, dg):
    a, b = cc.trsm_case(70, 9, side, uplo)
    A, B = cm(a), cm(b)
    fl = api.FlopCounter()
    api.hcore_dtrsm(side, uplo, tr, dg, 70, 9, 0.7, A, A.shape[1], B, 70, fl)
    ref = gold[f"trsm_{side}_{uplo}_{tr}_{dg}"]
    # the diagonal blocks are applied through their explicit inverses: a few ulps of cond(A) ~ 10 away from cblas_dtrsm
    assert np.abs(B.T - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    assert fl.trsm == int(gold[f"trsm_{side}_{uplo}_{tr}_{dg}_flops"])
    assert np.array_equal(A, cm(a))                                   # A (and its unused triangle) is read only


@pytest.mark.parametrize("m,k,side,uplo,tr", [(200, 20, 141, 122, 111), (200, 20, 142, 121, 112), (130, 70, 142, 122, 111),
                                              (257, 33, 141, 121, 111)])
def test_dtrsm_legacy_multiblock_vs_reference(gold, m, k, side, uplo, tr):
    a, b = cc.trsm_case(m, k, side, uplo)
    A, B = cm(a), cm(b)
    api.hcore_dtrsm(side, uplo, tr, 131, m, k, -1.25, A, A.shape[1], B, m)
    ref = gold[f"trsmbig_{m}_{k}_{side}_{uplo}_{tr}"]
    assert np.abs(B.T - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_dtrsm_legacy_padded_leading_dimensions():
    m, k, lda, ldb = 90, 11, 97, 101
    a, b = cc.trsm_case(m, k, 141, 122)
    A = np.full((m, lda), 7.0); A[:, :m] = a.T
    B = np.full((k, ldb), 9.0); B[:, :m] = b.T
    api.hcore_dtrsm(141, 122, 111, 131, m, k, 1.0, A, lda, B, ldb)
    x, _ = orc.trsm(141, 122, 111, 131, 1.0, a, b)
    assert np.abs(B[:, :m].T - x).max() < 1e-13
    assert (B[:, m:] == 9.0).all()                                     # padding rows untouched


@pytest.mark.parametrize("M,k,T", [(1024, 64, 6), (1000, 40, 3), (2048, 32, 2), (512, 128, 2)])
def test_dtrsm_batched_shared_diagonal_tile(ctx, M, k, T):
    """The call of a TLR-Cholesky step: ONE lower triangular diagonal tile, T low-rank tiles' V factors with ragged ranks."""
    rng = np.random.default_rng(M + k)
    L = cc.tri_matrix(rng, M, 122)
    ranks = np.array([k - (3 * t) % max(1, k // 2) for t in range(T)], np.int32)
    BV = rng.random((T, k, M)) - 0.5
    bt = dev(BV)
    before = BV.copy()
    ctx.dtrsm_batched(141, 122, 111, 131, 1.0, dev(cm(L)[None]), bt, dev(ranks))
    torch.cuda.synchronize()
    out = bt.cpu().numpy()
    Lt = np.tril(L)
    for t in range(T):
        r = int(ranks[t])
        x, _ = orc.trsm(141, 122, 111, 131, 1.0, L, BV[t, :r].T)
        assert np.abs(out[t, :r].T - x).max() <= 1e-12 * max(1.0, np.abs(x).max())
        # residual: || L X - B || small relative to ||L|| ||X||
        assert np.abs(Lt @ out[t, :r].T - BV[t, :r].T).max() <= 64 * EPS * M * np.abs(out[t, :r]).max()
        assert np.array_equal(out[t, r:], before[t, r:])               # columns beyond the tile's rank untouched


def test_dtrsm_batched_one_triangle_per_tile_right_side(ctx):
    T, M, k = 3, 300, 24
    rng = np.random.default_rng(9)
    A = np.stack([cm(cc.tri_matrix(rng, k, 121)) for _ in range(T)])
    BV = rng.random((T, k, M)) - 0.5
    bt = dev(BV)
    ctx.dtrsm_batched(142, 121, 112, 132, 0.5, dev(A), bt)
    torch.cuda.synchronize()
    out = bt.cpu().numpy()
    for t in range(T):
        x, _ = orc.trsm(142, 121, 112, 132, 0.5, A[t].T, BV[t].T)
        assert np.abs(out[t].T - x).max() < 1e-12


# ---- HCORE_dpotrf ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [100, 64, 200])
@pytest.mark.parametrize("ch", ["L", "U"])
def test_dpotrf_legacy_vs_reference(gold, n, ch):
    s = cc.spd_matrix(n)
    A = cm(s)
    fl = api.FlopCounter()
    info = api.hcore_dpotrf(ord(ch), n, A, n, fl)
    ref = gold[f"potrf_{n}_{ch}"]
    assert info == 0 and fl.potrf == int(gold[f"potrf_{n}_{ch}_flops"])
    assert np.abs(A.T - ref).max() <= 1e-13 * n
    other = np.triu(A.T, 1) if ch == "L" else np.tril(A.T, -1)
    assert np.array_equal(other, np.triu(s, 1) if ch == "L" else np.tril(s, -1))    # the other triangle is left untouched
    # the enum values are accepted too (the reference itself only works with the character codes, hcore_dpotrf.c:29-32)
    A2 = cm(s)
    assert api.hcore_dpotrf(api.HCORE_Lower if ch == "L" else api.HCORE_Upper, n, A2, n) == 0
    assert np.array_equal(A2, A)


def test_dpotrf_legacy_not_positive_definite(gold):
    s = cc.spd_matrix(100); s[50, 50] = -1.0
    A = cm(s)
    info = api.hcore_dpotrf(ord("L"), 100, A, 100)
    assert info == int(gold["potrf_notpd_info"]) == 51
    ref = gold["potrf_notpd"]
    # what LAPACK defines on failure: the Cholesky factor of the positive definite leading minor of order info - 1
    assert np.abs(np.tril(A.T)[:50, :50] - np.tril(ref)[:50, :50]).max() < 1e-12


def test_dpotrf_legacy_invalid_uplo_like_lapack():
    s = cc.spd_matrix(40)
    A = cm(s)
    assert api.hcore_dpotrf(ord("x"), 40, A, 40) == -2 and np.array_equal(A, cm(s))


@pytest.mark.parametrize("N,T,uplo", [(1024, 3, 122), (1000, 2, 122), (500, 4, 121), (2048, 1, 122)])
def test_dpotrf_batched_vs_oracle(ctx, N, T, uplo):
    A = np.stack([cm(cc.spd_matrix(N, seed=3 + t)) for t in range(T)])
    if T > 1:
        A[1, N // 2, N // 2] = -5.0                                    # tile 1 is not positive definite
    at, info = dev(A), torch.full((T,), -9, dtype=torch.int32, device="cuda:0")
    ctx.dpotrf_batched(uplo, at, info)
    torch.cuda.synchronize()
    out, inf = at.cpu().numpy(), info.cpu().numpy()
    for t in range(T):
        c, info_o, _ = orc.potrf(ord("L") if uplo == 122 else ord("U"), A[t].T)
        assert int(inf[t]) == info_o
        if info_o == 0:
            F = np.tril(out[t].T) if uplo == 122 else np.triu(out[t].T)
            Fo = np.tril(c) if uplo == 122 else np.triu(c)
            assert np.abs(F - Fo).max() <= 1e-12 * N
            rec = F @ F.T if uplo == 122 else F.T @ F
            sym = np.tril(A[t].T) + np.tril(A[t].T, -1).T if uplo == 122 else np.triu(A[t].T) + np.triu(A[t].T, 1).T
            assert np.abs(rec - sym).max() <= 64 * EPS * N              # L L^T = A


# ---- HCORE_dgemmbdcd / uncompress ------------------------------------------------------------------------------

@pytest.mark.parametrize("m,n,k", [(96, 96, 7), (130, 80, 12)])
def test_dgemmbdcd_legacy_vs_reference(gold, m, n, k):
    au, av, bd, cd = cc.bdcd_case(m, n, k)
    for ta in (111, 112):
        for beta in (0.0, 1.0, -0.5):
            CD = cm(cd)
            if beta == 0.0:
                CD[:] = np.nan                                          # beta = 0: CD must not be read
            fl = api.FlopCounter()
            api.hcore_dgemmbdcd(ta, 111, m, n, 0.3, cm(au), cm(av), np.array([float(k)]), m, cm(bd), m, beta, CD, m, None, fl)
            ref = gold[f"bdcd_{m}_{n}_{k}_{ta}_{beta}"]
            assert np.abs(CD.T - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
            assert fl.update == int(gold[f"bdcd_{m}_{n}_{k}_{ta}_{beta}_flops"])


@pytest.mark.parametrize("M,N,k,T", [(1024, 1024, 64, 3), (1000, 900, 40, 2), (257, 129, 9, 2)])
def test_dgemmbdcd_batched_vs_oracle(ctx, M, N, k, T):
    rng = np.random.default_rng(M + N)
    AU, AV = rng.random((T, k, M)) - 0.5, rng.random((T, k, M)) - 0.5
    BD, CD = rng.random((T, N, M)) - 0.5, rng.random((T, N, M)) - 0.5
    ranks = np.array([k - t for t in range(T)], np.int32)
    ct = dev(CD)
    ctx.dgemmbdcd_batched(-1.0, dev(AU), dev(AV), dev(ranks), dev(BD), 1.0, ct)
    torch.cuda.synchronize()
    out = ct.cpu().numpy()
    for t in range(T):
        r = int(ranks[t])
        ref, _ = orc.gemmbdcd(111, -1.0, AU[t, :r].T, AV[t, :r].T, BD[t].T, 1.0, CD[t].T)
        assert np.abs(out[t].T - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("nm,dt", [("d", np.float64), ("z", np.complex128)])
@pytest.mark.parametrize("m,n,k", [(96, 80, 7), (130, 200, 33)])
def test_uncompress_legacy_vs_reference(gold, nm, dt, m, n, k):
    au, bv = cc.uncompress_case(m, n, k, dt)
    CD = np.full((n, m), np.nan, dt)
    rk = np.array([float(k)])
    if nm == "d":
        api.hcore_duncompress(111, 112, m, n, 7.0, cm(au), rk, m, cm(bv), rk.copy(), n, 3.0, CD, m)
    else:
        api.hcore_zuncompress(111, 112, m, n, 7.0 + 1j, cm(au), rk, m, cm(bv), rk.copy(), n, 3.0 - 1j, CD, m)
    assert np.abs(CD.T - gold[f"unc_{nm}_{m}_{n}_{k}"]).max() <= 1e-13     # alpha / beta are ignored, like the reference


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_uncompress_batched_vs_oracle_and_rank_mismatch(ctx, dt):
    T, M, N, k = 4, 1024, 1000, 64
    rng = np.random.default_rng(4)
    f = (lambda s: rng.random(s) - 0.5 + 1j * (rng.random(s) - 0.5)) if dt == np.complex128 else (lambda s: rng.random(s) - 0.5)
    AU, BV = f((T, k, M)).astype(dt), f((T, k, N)).astype(dt)
    ark = np.array([64, 33, 1, 20], np.int32)
    brk = np.array([64, 33, 1, 21], np.int32)                           # tile 3: ranks differ (the reference asserts)
    CD = np.full((T, N, M), 5.0, dt)
    ct, st = dev(CD), torch.full((T,), -9, dtype=torch.int32, device="cuda:0")
    ctx.uncompress_batched(dev(AU), dev(ark), dev(BV), dev(brk), ct, status=st)
    torch.cuda.synchronize()
    out = ct.cpu().numpy()
    assert st.cpu().numpy().tolist() == [0, 0, 0, 6]
    for t in range(3):
        r = int(ark[t])
        ref = orc.uncompress(AU[t, :r].T, BV[t, :r].T)
        assert np.abs(out[t].T - ref).max() <= 256 * EPS * np.abs(ref).max()
    assert (out[3] == 5.0).all()                                        # mismatched tile left untouched


# ---- HCORE_dgemm_fast -----------------------------------------------------------------------------------------------

def test_dgemm_fast_legacy_vs_reference(gold):
    api.set_global("use_trmm", 1)                                       # the reference asserts use_trmm == 1 (hcore_dgemm_fast.c:82)
    try:
        for (name, au, av, bu, bv, cu, cv, acc) in cc.fast_cases():
            m, a, b, c = au.shape[0], au.shape[1], bu.shape[1], cu.shape[1]
            mk = 2 * max(a, b, c)
            CU = np.zeros((mk, m)); CV = np.zeros((mk, m)); CU[:c] = cu.T; CV[:c] = cv.T
            Crk = np.array([float(c)])
            api.hcore_dgemm_fast(111, 112, m, m, -1.0, cm(au), cm(av), np.array([float(a)]), m, cm(bu), cm(bv), np.array([float(b)]), m,
                                 1.0, CU, CV, Crk, m, 0, mk, acc)
            r = int(Crk[0])
            assert r == int(gold[f"fast_{name}_rank"])                  # scan from index 2: rank 2 on the crafted tiles
            ref = gold[f"fast_{name}_cu"] @ gold[f"fast_{name}_cv"].T
            assert np.abs(CU[:r].T @ CV[:r] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
            # the same tile through HCORE_dgemm: scan from index 1
            CU2 = np.zeros((mk, m)); CV2 = np.zeros((mk, m)); CU2[:c] = cu.T; CV2[:c] = cv.T
            Crk2 = np.array([float(c)])
            api.hcore_dgemm(111, 112, m, m, -1.0, cm(au), cm(av), np.array([float(a)]), m, cm(bu), cm(bv), np.array([float(b)]), m,
                            1.0, CU2, CV2, Crk2, m, 0, mk, acc, None, api.FlopCounter())
            assert int(Crk2[0]) == int(gold[f"fast_{name}_rank_dgemm"])
    finally:
        api.set_global("use_trmm", 0)


# ---- padded tiles through the legacy dense symbols (ADVICE: LDAU / LDCD honoured) -----------------------------------

def test_dsyrk_legacy_padded_leading_dimensions_vs_live_rule():
    """hcore_dsyrk.c:57-58: LDAU is the leading dimension of AU and AV AND the contraction length of AV^T AV; LDCD the
    leading dimension of CD (:165-181)."""
    n, k, ld, ldc = 50, 6, 56, 60
    rng = np.random.default_rng(8)
    AU = rng.random((k, ld)) - 0.5; AV = rng.random((k, ld)) - 0.5
    CD = rng.random((n, ldc)) - 0.5
    out = CD.copy()
    fl = api.FlopCounter()
    api.hcore_dsyrk(api.HCORE_Lower, api.HCORE_NoTrans, n, k, -1.0, AU, ld, AV, ld, 1.0, out, ldc, None, fl)
    w = AV @ AV.T                                                       # contraction over all ld rows
    ref = CD[:, :n].T - (AU[:, :n].T @ w) @ AU[:, :n]
    got = out[:, :n].T
    assert np.abs(np.tril(got) - np.tril(ref)).max() < 1e-13
    assert np.abs(np.triu(got, 1)).max() == 0.0
    assert np.array_equal(out[:, n:], CD[:, n:])                        # padding rows of CD untouched
    assert fl.syrk == 2 * k * k * ld + 2 * n * k * k + 2 * n * n * k


def test_dense_batched_ranks_above_128(ctx):
    """ADVICE: the X workspace is sized from the factor stride, so ranks above 128 are fine."""
    T, nb, ka, kb = 2, 384, 150, 200
    rng = np.random.default_rng(12)
    AU, AV = rng.random((T, ka, nb)) - 0.5, rng.random((T, ka, nb)) - 0.5
    BU, BV = rng.random((T, kb, nb)) - 0.5, rng.random((T, kb, nb)) - 0.5
    CD = rng.random((T, nb, nb)) - 0.5
    ark = np.array([ka, 0], np.int32); brk = np.array([kb, kb - 3], np.int32)       # tile 1: empty A -> CD unchanged
    c1 = dev(CD)
    ctx.dgemmcd_batched(dev(AU), dev(AV), dev(ark), dev(BU), dev(BV), dev(brk), c1)
    torch.cuda.synchronize()
    ref = orc.gemmcd(AU[0].T, AV[0].T, BU[0].T, BV[0].T, CD[0].T)
    assert np.abs(c1[0].cpu().numpy().T - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.array_equal(c1[1].cpu().numpy(), CD[1])


@pytest.mark.parametrize("nb,T", [(128, 1), (384, 5), (1024, 4)])
def test_dense_streaming_kernel_cases(ctx, nb, T):
    """Whole 128 x 32 blocks, contraction lengths <= 32, aligned strided batches go through the persistent TMA
    streaming kernel (dense_stream_kernel): varying ranks per tile, empty factors, the triangle rule of HCORE_dsyrk
    (reference hcore_dsyrk.c:174-181) on every strip, a single tile, write-only uncompress with a flagged tile."""
    k = 24
    rng = np.random.default_rng(100 + nb)
    f = lambda *s: rng.random(s) - 0.5
    AU, AV, BU, BV, CD = f(T, k, nb), f(T, k, nb), f(T, k, nb), f(T, k, nb), f(T, nb, nb)
    ark = np.array(([k, 7, 0, 24, 1] * 2)[:T], np.int32)
    brk = np.array(([k, 0, 5, 13, 24] * 2)[:T], np.int32)
    c1, c2 = dev(CD), dev(CD)
    ctx.dgemmcd_batched(dev(AU), dev(AV), dev(ark), dev(BU), dev(BV), dev(brk), c1)
    ctx.dsyrk_batched(0.75, dev(AU), dev(AV), dev(ark), c2)
    torch.cuda.synchronize()
    for i in range(T):
        a, b = int(ark[i]), int(brk[i])
        out = c1[i].cpu().numpy().T
        if a == 0 or b == 0:
            assert np.array_equal(out, CD[i].T)                         # nothing to add: the tile is not touched
        else:
            ref = orc.gemmcd(AU[i, :a].T, AV[i, :a].T, BU[i, :b].T, BV[i, :b].T, CD[i].T)
            assert np.abs(out - ref).max() <= 1000 * EPS * max(1.0, np.abs(ref).max())
        out = c2[i].cpu().numpy().T
        if a > 0:
            ref = orc.syrk(AU[i, :a].T, AV[i, :a].T, CD[i].T, 0.75)
            assert np.abs(out - ref).max() <= 1000 * EPS * max(1.0, np.abs(ref).max())
            assert np.abs(np.triu(out, 1)).max() == 0.0
    # write-only form (beta = 0: no tensor loads): CD = AU BV^T, one tile with mismatched ranks, one with rank 0
    M, N = nb, max(128, nb // 2)
    AUm, BVm = f(T, k, M), f(T, k, N)
    ra = np.array(([k, 3, 0, 9, 24] * 2)[:T], np.int32)
    rb = ra.copy()
    if T > 3:
        rb[3] = 10
    ct, st = dev(np.full((T, N, M), 5.0)), torch.full((T,), -9, dtype=torch.int32, device="cuda:0")
    ctx.uncompress_batched(dev(AUm), dev(ra), dev(BVm), dev(rb), ct, status=st)
    torch.cuda.synchronize()
    out = ct.cpu().numpy()
    for i in range(T):
        if ra[i] != rb[i]:
            assert int(st[i].item()) == 6 and (out[i] == 5.0).all()
        else:
            r = int(ra[i])
            ref = orc.uncompress(AUm[i, :r].T, BVm[i, :r].T) if r > 0 else np.zeros((M, N))
            assert int(st[i].item()) == 0 and np.abs(out[i].T - ref).max() <= 256 * EPS * max(1.0, np.abs(ref).max())


# ---- pointer-array forms (HCORE_{d,s,z}gemm_batched, HCORE_dgemmcd_batched, HCORE_dsyrk_batched) -----------

def _ptr_table(tensors):
    return torch.tensor([x.data_ptr() for x in tensors], dtype=torch.int64, device="cuda:0")


def test_pointer_array_forms_equal_strided_forms(ctx):
    import ctypes as C
    from hcore_b200 import _lib
    from hcore_b200.tilegen import make_lr_batch, make_dense_batch
    L = _lib.load()
    vp = lambda t: C.c_void_p(t.data_ptr())
    T, nb, k = 5, 160, 10
    # complex update: tiles scattered in separately allocated arrays, in reverse order
    b = make_lr_batch(T, nb, k, base_seed=91, dtype=np.complex128)
    t = {n: dev(getattr(b, n)) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    ctx.zgemm_batched(-1.0 + 0.5j, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, t["cu"], t["cv"], t["crk"], rk=0,
                      maxrk=2 * k, acc=1e-8)
    torch.cuda.synchronize()
    sep = {n: [dev(getattr(b, n)[i]) for i in range(T)] for n in ("au", "av", "bu", "bv", "cu", "cv")}
    tabs = {n: _ptr_table(sep[n]) for n in sep}
    crk2, st = dev(b.crk), torch.full((T,), -3, dtype=torch.int32, device="cuda:0")
    rc = L.HCORE_zgemm_batched(ctx._h, T, nb, _lib.DComplex.of(-1.0 + 0.5j), vp(tabs["au"]), vp(tabs["av"]), vp(t["ark"]), nb,
                               vp(tabs["bu"]), vp(tabs["bv"]), vp(t["brk"]), nb, _lib.DComplex.of(1.0), vp(tabs["cu"]), vp(tabs["cv"]),
                               vp(crk2), nb, 0, 2 * k, 1e-8, vp(st), None)
    torch.cuda.synchronize()
    assert rc == 0 and (st.cpu().numpy() == 0).all() and torch.equal(crk2, t["crk"])
    for i in range(T):
        r = int(crk2[i].item())
        assert torch.equal(sep["cu"][i][:r], t["cu"][i, :r]) and torch.equal(sep["cv"][i][:r], t["cv"][i, :r])
    # dense updates
    bd = make_lr_batch(T, nb, k, base_seed=92)
    td = {n: dev(getattr(bd, n)) for n in ("au", "av", "bu", "bv", "ark", "brk")}
    cd = make_dense_batch(T, nb)
    c1, c2 = dev(cd), dev(cd)
    ctx.dgemmcd_batched(td["au"], td["av"], td["ark"], td["bu"], td["bv"], td["brk"], c1)
    ctx.dsyrk_batched(-1.0, td["au"], td["av"], td["ark"], c2)
    sepd = {n: [dev(getattr(bd, n)[i]) for i in range(T)] for n in ("au", "av", "bu", "bv")}
    cds1, cds2 = [dev(cd[i]) for i in range(T)], [dev(cd[i]) for i in range(T)]
    tb = {n: _ptr_table(sepd[n]) for n in sepd}
    t1, t2 = _ptr_table(cds1), _ptr_table(cds2)
    assert L.HCORE_dgemmcd_batched(ctx._h, T, nb, k, vp(tb["au"]), vp(tb["av"]), vp(td["ark"]), vp(tb["bu"]), vp(tb["bv"]), vp(td["brk"]),
                                   vp(t1), None) == 0
    assert L.HCORE_dsyrk_batched(ctx._h, T, nb, k, -1.0, vp(tb["au"]), vp(tb["av"]), vp(td["ark"]), vp(t2), None) == 0
    torch.cuda.synchronize()
    for i in range(T):
        assert torch.equal(cds1[i], c1[i]) and torch.equal(cds2[i], c2[i])
