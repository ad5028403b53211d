"""The C-ABI library builds, loads and exports every symbol include/gk_typing.h declares."""
import os
import re

from kir_graph_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols() -> set[str]:
    text = open(os.path.join(ROOT, "include", "gk_typing.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(gk_[a-z0-9_]+)\s*\(", text))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = _cabi.load()
    names = declared_symbols()
    assert names == set(_cabi.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert lib.gk_abi_version() == 2
    assert lib.gk_sizeof(b"nonsense") == -1


def test_struct_layouts_match():
    lib = _cabi.load()
    for name, dtype in _cabi._STRUCTS.items():
        assert lib.gk_sizeof(name.encode()) == dtype.itemsize
