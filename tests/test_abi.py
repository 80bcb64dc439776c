"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/cosmolike.h declares,
the ctypes struct mirrors the C struct, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cosmology_model_fit_b200 as pkg
from cosmology_model_fit_b200 import build as cbuild
from cosmology_model_fit_b200 import engine
from cosmology_model_fit_b200.spec import ClSpec
from cases import spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    cbuild.build()
    return engine.load_library()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "cosmolike.h")).read()
    declared = set(re.findall(r"\b(cl_[a-z_0-9]+)\s*\(", header))
    assert declared == set(engine.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_ctypes_struct_matches_c_layout(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cosmolike.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(cl_spec), offsetof(cl_spec, z_grid), offsetof(cl_spec, bao_qty), offsetof(cl_spec, gl_x),'
                   'offsetof(cl_spec, guard_value));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(ClSpec), ClSpec.z_grid.offset, ClSpec.bao_qty.offset, ClSpec.gl_x.offset, ClSpec.guard_value.offset]
    assert got == want


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.EngineError, match="no CUDA device|CPU fallback"):
        pkg.Engine(spec("sn_union3_1"))


def test_spec_validation_is_reported(lib):
    s = spec("sn_union3_1").c_spec()
    s.abi_version = 999
    ctx = C.c_void_p()
    rc = lib.cl_create(C.byref(s), 0, C.byref(ctx))
    assert rc == -1 and b"abi_version" in lib.cl_last_error(None)


def test_package_does_not_import_oracle():
    import sys
    code = "import sys, cosmology_model_fit_b200; sys.exit(any(m.split('.')[0] == 'oracle' for m in sys.modules))"
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0
    for fn in os.listdir(os.path.join(ROOT, "cosmology_model_fit_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "cosmology_model_fit_b200", fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|libcosmo_oracle|oracle/_", src, re.M), fn
