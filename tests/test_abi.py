"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports exactly what
include/ptoy_b200.h declares; without a GPU it refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import ptoy_b200
from ptoy_b200 import particles as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ptoy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptoy_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(ptoy_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ptoy_b200.load_library()


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ptoy_b200.h but not exported"


def test_python_binding_covers_the_header(lib):
    bound = set(P.SIGNATURES) | set(P.FREE_FUNCTIONS)
    assert bound == set(header_symbols())


def test_abi_version(lib):
    assert lib.ptoy_abi_version() == 2


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.ptoy_create(C.byref(h), 0)
    assert rc == -1 and not h.value          # PTOY_ERR_CUDA
    assert b"no CPU fallback" in lib.ptoy_last_error()
    with pytest.raises(ptoy_b200.PtoyError):
        ptoy_b200.Particles()


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ptoy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|libptoy_oracle|libptoy_ref", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_null_handle_is_an_error(lib):
    assert lib.ptoy_step_n(None, 1) == -2     # PTOY_ERR_ARG


def test_stage_kernels_sass_keeps_the_exact_arithmetic_contract(lib):
    """The library's SASS: the stage kernels are TMA-fed (UBLKCP), use the packed fp32 forms for
    differences / products (FADD2, FMUL2) and contain NO packed fused multiply-add - ptxas contracts
    mul.rn.f32x2 + add.rn.f32x2, which would round once where the reference rounds twice."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-sass", ptoy_b200.LIB_PATH], capture_output=True, text=True, check=True).stdout
    stage = "".join(part for part in out.split("Function : ") if part.startswith("_ZN4ptoy17stage_tile_kernel"))
    assert stage, "no stage_tile_kernel in the library"
    assert "UBLKCP" in stage and "SYNCS" in stage
    assert "FADD2" in stage and "FMUL2" in stage
    assert "FFMA2" not in stage
