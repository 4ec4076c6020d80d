"""The C-ABI library loads and exports every symbol include/mpopis_b200.h declares (no compute)."""
import ctypes
import re
from pathlib import Path

from mpopis_b200 import _abi, _lib

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    txt = (ROOT / "include" / "mpopis_b200.h").read_text()
    return sorted(set(re.findall(r"\b(mpopis_b200_\w+)\s*\(", txt)))


def test_header_and_bindings_agree():
    assert header_symbols() == sorted(_abi.exported_symbols())


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} not exported"
    lib.mpopis_b200_abi_version.restype = ctypes.c_int
    assert lib.mpopis_b200_abi_version() == _abi.ABI_VERSION


def test_oracle_mirrors_the_abi(orc):
    b = orc.bound()
    for name in _abi.SIGNATURES:
        assert hasattr(b.lib, "orc_" + name)


def test_cfg_struct_layout_matches_header():
    # mpopis_cfg_t: 4 x i32, 3 x i64, 4 x f64, 6 x i32, 4 x i32 reserved
    assert ctypes.sizeof(_abi.Cfg) == 16 + 24 + 32 + 24 + 16
    assert ctypes.sizeof(_abi.Cma) == 9 * 8


def test_product_path_never_imports_the_oracle():
    for py in (ROOT / "mpopis_b200").glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
    for cu in (ROOT / "mpopis_b200" / "csrc").glob("*.cu*"):
        assert not re.search(r"#include[^\n]*oracle", cu.read_text()), cu
    build_py = (ROOT / "mpopis_b200" / "build.py").read_text()
    assert "oracle" not in build_py  # the product .so never links the checker
