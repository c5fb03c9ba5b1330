"""The C-ABI boundary without a GPU: libspades_b200.so loads, exports every function include/sb200.h declares, and fails
loudly (no CPU fallback) when no sm_100 device is present."""
import ctypes as C
import os
import re

import pytest

from spades_for_blackbird_b200.host import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(B.LIB_PATH)
    names = declared_functions()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    names = set(declared_functions())
    assert names == set(B.EXPORTS), (names ^ set(B.EXPORTS))
    B.load_library()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B.Sb200Error, match="no CPU fallback|sm_100a only"):
        B.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spades_for_blackbird_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "sb200_oracle" not in src and "libsb200_oracle" not in src, f
