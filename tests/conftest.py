import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def emu_library():
    """tests/emu: the kernel sources compiled for the host SIMT emulator (debug harness)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from insr_pde_b200 import _lib
    return _lib.Library(build_emu.build_emu())


@pytest.fixture()
def emu_backend(emu_library, monkeypatch):
    """Route the package's C-ABI calls to the emulation build so that the *host-side* logic
    (autograd boundary, drop-in modules, diff_ops) can be tested on CPU tensors.  Test-only
    monkeypatching: the product has no such switch."""
    from insr_pde_b200 import _lib, _ops
    monkeypatch.setattr(_lib, "_LIB", emu_library)
    monkeypatch.setattr(_ops, "_require_cuda", lambda t: None)
    monkeypatch.setattr(_ops, "_stream", lambda device: None)
    return emu_library
