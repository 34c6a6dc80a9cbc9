"""The C-ABI library loads and exports every symbol include/sininn.h declares (no GPU needed)."""
import ctypes
import os
import re

from sin_inn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sininn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sininn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"libsininn.so does not export {s}"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in sin_inn_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)


def test_load_and_version():
    lib = _lib.load()
    assert lib.sininn_version() >= 100
    assert isinstance(lib.sininn_last_error(), bytes)


def test_struct_layout_matches_header():
    # field order/type drift between the header structs and the ctypes mirrors would corrupt calls silently
    assert ctypes.sizeof(_lib.ConvDesc) == 200 and ctypes.sizeof(_lib.WgradDesc) == 544      # (gcc/nvcc sizeof of the header structs)
    assert ctypes.sizeof(_lib.Subnet1x1Desc) == 216 and ctypes.sizeof(_lib.Subnet1x1BwdDesc) == 208 and ctypes.sizeof(_lib.WgradSegment) == 48
