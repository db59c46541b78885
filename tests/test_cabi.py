"""The C-ABI boundary without a GPU: both shared libraries load, export every function `include/*.h` declares, and the product
fails loudly (no CPU fallback) when no device is usable. No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ifx_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("header,lib", [("infidex_gpu.h", "gpu"), ("infidex_host.h", "host")])
def test_library_exports_every_declared_symbol(header, lib):
    path = _build.build_gpu() if lib == "gpu" else _build.build_host()
    handle = C.CDLL(path)
    names = _declared(header)
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, "declared in include/%s but not exported by %s: %s" % (header, os.path.basename(path), missing)


def test_no_torch_types_in_the_boundary():
    for header in ("infidex_gpu.h", "infidex_host.h"):
        src = open(os.path.join(ROOT, "include", header), encoding="utf-8").read()
        assert 'extern "C"' in src
        assert "torch" not in src.lower() and "at::" not in src and "std::" not in src


def test_fails_loudly_without_a_device():
    lib = C.CDLL(_build.build_gpu()); lib.ifx_device_count.restype = C.c_int
    if lib.ifx_device_count() > 0:
        pytest.skip("a GPU is present")
    eng = ib.SearchEngine.CreateDefault()
    with pytest.raises(ib.engine.NativeError):        # index creation is the first call that needs the device: IFX_ERR_NO_DEVICE, never a CPU path
        eng.IndexColumns(np.arange(3), [ib.Field("content")], [["alpha beta", "beta gamma", "gamma delta"]])
