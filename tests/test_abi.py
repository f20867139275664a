"""CPU tests: the C-ABI library loads without a GPU and exports every symbol the headers
declare; ABI structs match the reference; host-side helpers agree with the oracle."""
import ctypes as C
import os
import re

import pytest

from hcore_b200 import _lib, api
from oracle import ref_binding as rb
from oracle import tlr_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b((?:HCORE|hcore_b200)_[A-Za-z0-9_]+)\s*\(", txt))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared_functions("hcore_b200.h") | _declared_functions("hcore_d.h") | _declared_functions("hcore_s.h") | _declared_functions("hcore_z.h")
    assert {"HCORE_dgemm", "HCORE_dgemmcd", "HCORE_dsyrk", "HCORE_dgemm_batched"} <= declared
    for sym in declared | set(_lib.EXPORTED_SYMBOLS):
        assert hasattr(L, sym), sym
    assert declared <= set(_lib.EXPORTED_SYMBOLS), declared - set(_lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in L.hcore_b200_version()


def test_flop_counter_layout_matches_reference():
    """flop_counter must be layout-identical to reference src/misc/include/flop_util_structs.h:5-36."""
    assert C.sizeof(api.FlopCounter) == C.sizeof(rb.FlopCounter) == 232
    for f in ("copy", "potrf", "syrk", "update", "numthreads"):
        assert getattr(api.FlopCounter, f).offset == getattr(rb.FlopCounter, f).offset
    hdr = open(os.path.join(ROOT, "include", "hcore_b200_flops.h")).read()
    names = re.findall(r"X\((?:unsigned long int|int), (\w+)\)", hdr)
    assert names == ["__pad1", "__pad2", "copy", "flop", "tid", "potrf", "trsm", "syrk", "update", "geqrfu",
                     "geqrfv", "geqrfv_gemm", "gesvd", "gesvd_gemm", "gesvd_lasetu", "gesvd_lasetv",
                     "gesvd_lacpyu", "gesvd_lacpyv", "gesvd_dscal", "ormqru", "ormqru_laset", "ormqru_lacpy",
                     "ormqrv", "ormqrv_laset", "ormqrv_lacpy", "omatcopy", "numthreads", "__pad3", "__pad4"]


@pytest.mark.parametrize("m,a,b,c,r", [(1024, 64, 64, 64, 93), (1024, 64, 64, 64, 64), (500, 25, 25, 25, 37),
                                       (10, 4, 4, 4, 6), (96, 5, 7, 6, 9)])
def test_flop_model_matches_reference_counter(m, a, b, c, r):
    assert api.gemm_flops(m, a, b, c, r) == orc.flop_model(m, a, b, c, r)
    if (m, r) == (1024, 93):
        assert api.gemm_flops(m, a, b, c, r) == 219370048       # probe-exact value, SURVEY 8(d)


def test_shard_range_partitions_by_tile_index():
    for ntiles in (0, 1, 7, 16, 131072, 100003):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, count = api.shard_range(ntiles, world, r)
                assert first == nxt and count >= 0
                nxt += count
            assert nxt == ntiles
            counts = [api.shard_range(ntiles, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.HcoreError):
        api.Context(0)
    h = C.c_void_p()
    assert _lib.load().hcore_b200_create(C.byref(h), 0) != 0


def test_product_does_not_import_oracle():
    """Nothing under hcore_b200/ may reference the oracle (it is test infrastructure only)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hcore_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
