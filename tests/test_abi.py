"""The C-ABI shared library loads and exports every symbol ``include/modaltune_b200.h`` declares (no compute calls)."""
import ctypes
import os
import re

from modaltune_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    text = open(os.path.join(ROOT, "include", "modaltune_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    assert set(_declared()) == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m modaltune_b200.build` (or __graft_entry__.build()) first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_binding_loads_and_reports_version():
    lib = _lib.load()
    assert lib.mt_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_argument_counts_match_header():
    text = open(os.path.join(ROOT, "include", "modaltune_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_geometry_struct_layout():
    g = _lib.DilatedGeometry()
    assert ctypes.sizeof(g) == 4 * 4 + 2 * 4 * _lib.MT_MAX_BRANCHES


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "modaltune_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), f
