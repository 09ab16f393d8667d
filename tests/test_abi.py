"""The C-ABI library builds, loads and exports every symbol include/fac_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

from fac_via_ppg_b200 import _ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(fac_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = _ext.load()
    names = declared_symbols()
    assert names, "no declarations parsed"
    for name in names:
        assert hasattr(lib, name), name
    assert names == set(_ext.SIGNATURES), names ^ set(_ext.SIGNATURES)


def test_version_and_error_string():
    lib = _ext.load()
    assert lib.fac_version() >= 200
    assert isinstance(lib.fac_last_error(), bytes)


def test_struct_sizes_match_header():
    import ctypes as C
    # pointer tables: the C structs are plain pointers/ints, so sizes are predictable
    assert C.sizeof(_ext.ConvSrc) == 8 + 3 * 8 + 6 * 4
    assert C.sizeof(_ext.WgWorkspace) == 4 * 8
    assert C.sizeof(_ext.WgFlow) == 8 + 5 * 8 + 4 * 16 * 8
    assert C.sizeof(_ext.WgModel) == 10 * 4 + 2 * 8 + 16 * C.sizeof(_ext.WgFlow)
    assert C.sizeof(_ext.TcConv) == 10 * 8 + 3 * 8 + 12 * 4 + 8 + 8
    assert C.sizeof(_ext.ConvEpilogue) == 2 * 4 + 8 + 2 * 8 + 3 * 8 + 2 * 4 + 8
    assert C.sizeof(_ext.TacoDecoderState) == 9 * 8


def test_build_digest_does_not_depend_on_the_checkout_path(monkeypatch):
    """The library built here must be accepted as up to date in a copy of the repository elsewhere (the GPU box):
    otherwise every rank of a torchrun launch rebuilds it at import time, concurrently."""
    from fac_via_ppg_b200 import build
    here = build._digest()
    link = os.path.join(ROOT, "fac-via-ppg_b200")              # the same package through its symlinked name
    assert os.path.isdir(link)
    monkeypatch.setattr(build, "CSRC", os.path.join(link, "csrc"))
    assert build._digest() == here
    assert build.build() == build.LIB_PATH                      # up to date: no compiler involved
