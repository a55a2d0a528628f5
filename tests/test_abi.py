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


def test_auxiliary_entry_points_validate_before_touching_the_device(lib):
    """the closure / sampler / optimiser entry points reject bad arguments with their insr_status before any CUDA call
    (this box has no GPU: reaching the device check would return a different code)"""
    c = lib.cdll
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    p += (-p) % 16
    t = _lib.ElasticTermsDesc()
    t.n, t.n_left, t.n_right, t.dt = 4, 0, 0, 0.05
    assert c.insr_elastic_terms(None, 2, p, None, p, p, p, p, p, None, None) == -1
    assert c.insr_elastic_terms(ctypes.byref(t), 4, p, None, p, p, p, p, p, None, None) == -2          # d must be 2 or 3
    t.r_arap = 1.0
    assert c.insr_elastic_terms(ctypes.byref(t), 2, p, None, p, p, p, p, p, None, None) == -1          # arap needs J and gJ
    assert b"arap / volume need J" in c.insr_last_error()
    t.r_arap, t.dt = 0.0, 0.0
    assert c.insr_elastic_terms(ctypes.byref(t), 2, p, None, p, p, p, p, p, None, None) == -2          # dt must be positive
    assert c.insr_sample_mesh(p, p, p, 0, 4, 10, 3, 0, None, None, 0, p, None) == -2                   # no elements
    assert c.insr_sample_mesh(p, p, p, 5, 5, 10, 3, 0, None, None, 0, p, None) == -2                   # neither triangles nor tets
    assert c.insr_sample_mesh(p, p, p, 5, 4, 10, 3, 0, p, None, 0, p, None) == -1                      # counter without ticket
    assert c.insr_sample_boxes(0, 2, p, p, p, 0, None, None, 0, p, None) == -2
    assert c.insr_svd_small(p, 3, 4, None, p, None, None) == -2
    arr = (ctypes.c_void_p * 1)(p)
    sizes = (ctypes.c_int64 * 1)(8)
    args = [ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(0.1), 500,
            ctypes.c_float(1e-4), ctypes.c_float(1e-8), ctypes.c_float(1e-8), 1, 0, None]
    assert c.insr_iteration_update(0, arr, arr, arr, arr, sizes, p, p, 1, 0, None, 0, None, p, *args) == -2    # no slots
    assert c.insr_iteration_update(1, arr, arr, arr, arr, sizes, p, p, 2, 2, None, 0, None, p, *args) == -2    # main index out of range
    assert c.insr_iteration_update(1, arr, arr, arr, arr, sizes, p, p, 1, 0, p, 0, None, p, *args) == -1       # log without index word
    # peer memory: argument validation happens before any CUDA call
    bases = (ctypes.c_void_p * 2)(p, p)
    assert c.insr_peer_alloc(-1, ctypes.byref(ctypes.c_void_p()), ctypes.create_string_buffer(64)) == -2
    assert c.insr_peer_alloc(16, None, None) == -1
    assert c.insr_peer_open(None, None) == -1
    assert c.insr_peer_allreduce(0, 0, bases, 1024, 8, 1.0, p, None) == -2          # world must be 1..16
    assert c.insr_peer_allreduce(2, 2, bases, 1024, 8, 1.0, p, None) == -2          # rank outside the world
    assert c.insr_peer_allreduce(2, 0, bases, 512, 8, 1.0, p, None) == -2           # data offset inside the header
    assert c.insr_peer_allreduce(2, 0, bases, 1026, 8, 1.0, p, None) == -2          # data offset not 16-byte aligned
    assert c.insr_peer_allreduce(2, 0, bases, 1024, 8, 1.0, None, None) == -1
    assert c.insr_peer_allreduce(2, 0, (ctypes.c_void_p * 2)(p, None), 1024, 8, 1.0, p, None) == -1   # a peer mapping is missing
    pargs = [ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(0.1), 500,
             ctypes.c_float(1e-4), ctypes.c_float(1e-8), ctypes.c_float(1e-8), 1, 1, None]
    # gradient / loss slots must lie inside the own peer allocation (behind its header)
    assert c.insr_iteration_update_peer(2, 0, bases, 8192, 0.5, 1, arr, arr, arr, arr, sizes, p, p, 1, 0, p, None, 0, None, *pargs) == -2
    assert b"not inside this rank's peer allocation" in c.insr_last_error()
    # tape: offered for the tiled family on one workspace chunk only, never for the H <= 32 family or forced-generic
    assert lib.tape_supported(_lib.make_desc(2, 2, 68, 3), 20000, 1)
    assert not lib.tape_supported(_lib.make_desc(2, 1, 32, 3), 20000, 2)
    assert not lib.tape_supported(_lib.make_desc(2, 2, 68, 3, flags=_lib.FLAG_FORCE_GENERIC), 20000, 1)
    assert not lib.tape_supported(_lib.make_desc(2, 1, 512, 5), 1 << 22, 2)
