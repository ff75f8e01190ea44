"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/m3b200.h declares, and fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import pytest

from mach3_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "m3b200.h")).read()
    return sorted(set(re.findall(r"M3B_API\s+[\w\s\*]+?\b(m3b_\w+)\s*\(", hdr)))


def test_header_declares_what_python_binds():
    assert _declared() == sorted(lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = lib.load()
    for name in _declared():
        assert hasattr(L, name), name
    assert L.m3b_abi_version() == 1


def test_no_torch_types_in_signatures():
    hdr = open(os.path.join(ROOT, "include", "m3b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # strip comments
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert "#include <stdint.h>" in code


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_create_fails_loudly_without_a_device():
    with pytest.raises(lib.M3BError) as ei:
        lib.Handle()
    assert ei.value.code == 6            # M3B_ERR_NODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_null_arguments_are_rejected_not_crashed():
    L = lib.load()
    assert L.m3b_create(None, None) == 1
    assert L.m3b_step(None, None, None, None) == 1
    assert L.m3b_llh(None, None, None) == 1
    assert L.m3b_get_info(None, None) == 1
    L.m3b_destroy(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mach3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f == "build.py":      # compiles the checker (allowed); never loads or calls it
                continue
            if f.endswith((".py", ".cu", ".h", ".c", ".cuh", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "m3_oracle" not in src, f
                assert "libm3oracle" not in src, f
