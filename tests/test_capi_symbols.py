"""The C-ABI library loads on a box without a GPU and exports every symbol include/vp8gpu.h
declares; record layouts match the header.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from alfalfa_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vp8gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vp8gpu_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    assert _declared() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in _declared():
        assert hasattr(L, name), name


def test_record_layouts():
    assert C.sizeof(capi.FrameDesc) == 80
    assert capi.MB_DTYPE.itemsize == 32
    assert capi.MB_DTYPE.fields["b_modes"][1] == 24
    assert capi.FrameDesc.quant.offset == 16 and capi.FrameDesc.n_tokens.offset == 64


def test_no_device_fails_loudly_not_silently():
    """on the CPU-only build box creating a context must raise, never fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from alfalfa_b200 import Context, CudaError
    with pytest.raises(CudaError):
        Context(320, 240)


def test_cxx_host_mirror_compiles_and_links():
    """alfalfa_b200/host/alfalfa_gpu.hh (Decoder / DecoderState / References / RasterHandle mirror)"""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "tests", "host_mirror_compile.cc")
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-shared", "-fPIC", src, "-o",
                               os.path.join(d, "m.so"), "-L" + os.path.join(ROOT, "alfalfa_b200"), "-l:libvp8gpu.so"])
