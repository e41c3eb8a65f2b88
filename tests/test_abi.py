"""The C-ABI library loads and exports every symbol include/dkg_b200.h declares (no compute)."""
import ctypes
import os
import re

from decoupledbo_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dkg_b200.h")).read()
    return re.findall(r"DKG_API\s+[\w\s\*]+?\b(dkg_\w+)\s*\(", text)


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    assert set(names) == set(_native.EXPORTED_SYMBOLS), (names, _native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_native.LIB_PATH), "build with decoupled-kg_b200/build.sh (or __graft_entry__.build())"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.dkg_abi_version.restype = ctypes.c_int
    assert lib.dkg_abi_version() == _native.ABI_VERSION
    assert _native.load_library().dkg_abi_version() == _native.ABI_VERSION


def test_no_torch_types_in_the_header():
    text = open(os.path.join(ROOT, "include", "dkg_b200.h")).read()
    assert "at::" not in text and "#include <torch" not in text and "Tensor" not in text


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "decoupled-kg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
