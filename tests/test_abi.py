"""The C-ABI library loads and exports every symbol include/gudni_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from gudni_b200 import _build
from gudni_b200.raster import ABI_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gudni_b200.h")).read()
    return sorted(set(re.findall(r"\b(gudni_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_binding_agree():
    assert header_symbols() == sorted(ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_build.LIB_CUDA):
        _build.build_cuda()
    lib = ctypes.CDLL(_build.LIB_CUDA)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_init_without_device_fails_loudly():
    """No GPU in the CPU container: init must report GUDNI_ERR_NO_DEVICE, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gudni_b200.raster import GudniError, setup_rasterizer
    with pytest.raises(GudniError):
        setup_rasterizer()
