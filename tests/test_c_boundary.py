"""The boundary proved in C: the reference's own test programs (testing/testing_{dgemm,sgemm,dgemmcd,dsyrk}.c),
UNMODIFIED, compiled against this repo's drop-in headers (include/) and linked with libhcore_b200.so instead of the
reference library by `make -C oracle b200-ctests` (run by __graft_entry__.build() where /root/reference exists; the
binaries travel to the GPU box in oracle/_ref/).  Their own asserts are the contract (e.g. testing_dgemm.c:377-378)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")
PROGS = ["testing_dgemm", "testing_sgemm", "testing_dgemmcd", "testing_dsyrk"]


def _have():
    return all(os.path.exists(os.path.join(BIN, "b200_" + p)) for p in PROGS)


@pytest.mark.gpu
@pytest.mark.parametrize("prog", PROGS)
def test_reference_test_program_links_and_passes_on_the_gpu(prog):
    exe = os.path.join(BIN, "b200_" + prog)
    assert os.path.exists(exe), "build with `make -C oracle b200-ctests` where /root/reference exists"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.skipif(not _have(), reason="oracle/_ref/b200_testing_* not built (needs /root/reference)")
def test_programs_resolve_the_hcore_symbols_from_libhcore_b200():
    """CPU: the programs are linked against libhcore_b200.so (not the reference library) and every HCORE_* symbol they
    import is exported by it."""
    from hcore_b200 import _lib
    lib = _lib.load()
    for prog in PROGS:
        exe = os.path.join(BIN, "b200_" + prog)
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libhcore_b200.so" in out and "libhcore_ref" not in out
        und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
        syms = [ln.split()[-1] for ln in und.splitlines() if "HCORE_" in ln]
        assert syms, prog
        for s_ in syms:
            getattr(lib, s_)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/b200_testing_* not built (needs /root/reference)")
def test_no_cpu_fallback_without_a_gpu():
    """CPU: without a CUDA device the drop-in symbol fails loudly (message + non-zero exit), it never computes on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([os.path.join(BIN, "b200_testing_dgemmcd")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout)
