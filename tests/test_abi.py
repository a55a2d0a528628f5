"""The C-ABI library loads and exports every symbol include/insr_b200.h declares; host-only
entry points (sizes, validation) behave.  No compute call is made here (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from insr_pde_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    path = build.build_library()
    assert os.path.exists(path)
    return _lib.Library(path)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "insr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(insr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert getattr(lib.cdll, n) is not None


def test_version_and_sizes(lib):
    assert lib.version() == 1
    d = _lib.make_desc(2, 1, 32, 3)
    assert lib.theta_size(d) == 3297          # fluid pressure net
    assert lib.theta_size(_lib.make_desc(2, 2, 32, 3)) == 3330
    assert lib.theta_size(_lib.make_desc(1, 1, 20, 2)) == 901
    assert lib.theta_size(_lib.make_desc(2, 2, 68, 3)) == 14418
    assert lib.theta_size(_lib.make_desc(3, 3, 66, 3)) == 13731
    assert lib.theta_size(_lib.make_desc(3, 3, 128, 3)) == 50435
    assert lib.workspace_bytes(_lib.make_desc(2, 1, 256, 3, flags=_lib.FLAG_FORCE_GENERIC), 1000, 2, True) > 0


@pytest.mark.parametrize("bad", [dict(D=0), dict(D=4), dict(O=0), dict(O=4), dict(H=0), dict(H=513), dict(L=-1), dict(L=17)])
def test_bad_shapes_are_errors_not_fallbacks(lib, bad):
    kw = dict(D=2, O=1, H=32, L=3)
    kw.update(bad)
    d = _lib.make_desc(kw["D"], kw["O"], kw["H"], kw["L"])
    assert lib.theta_size(d) == 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    p += (-p) % 16
    rc = lib.cdll.insr_siren_forward(ctypes.byref(d), p, p, 1, 0, p, None, None, None, 0, None)
    assert rc == -2
    assert b"outside" in lib.cdll.insr_last_error()


def test_argument_validation_order(lib):
    d = _lib.make_desc(2, 1, 32, 3)
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    p += (-p) % 16
    assert lib.cdll.insr_siren_forward(ctypes.byref(d), p, p, 1, 7, p, None, None, None, 0, None) == -3
    assert lib.cdll.insr_siren_forward(ctypes.byref(d), None, p, 1, 0, p, None, None, None, 0, None) == -1
    assert lib.cdll.insr_siren_forward(ctypes.byref(d), p, p, 1, 1, p, None, None, None, 0, None) == -1   # jac missing
    assert lib.cdll.insr_siren_forward(ctypes.byref(d), p, p + 4, 1, 0, p, None, None, None, 0, None) == -4
    with pytest.raises(_lib.InsrError):
        lib.check(-4)
