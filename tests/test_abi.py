"""CPU suite: the C-ABI library loads and exports every symbol include/tfr.h declares; argument
validation and the no-CPU-fallback rule work without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import tfrec_b200 as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tfrec_b200 import build
    build.build()
    return tb.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "tfr.h")).read()
    declared = set(re.findall(r"\b(tfr_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(tb.ABI_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    assert lib.tfr_abi_version() == 2


def test_struct_sizes_match_header():
    # sizes the C side static_asserts / memcpy's against
    assert C.sizeof(tb.Config) == 40
    assert C.sizeof(tb.Frame) == 112
    assert C.sizeof(tb.Record) == 72
    assert C.sizeof(tb.BlockTrace) == 12
    assert C.sizeof(tb.Stats) == 144


def test_create_validates_arguments(lib):
    h = C.c_void_p()
    cfg = tb.Config(0, 0, 7, 0, 0, 1, 0, 0, 0)       # wrong struct_size
    assert lib.tfr_create(C.byref(cfg), C.byref(h)) == -1
    assert b"struct_size" in lib.tfr_last_error()
    cfg = tb.Config(C.sizeof(tb.Config), 0, 7, 0, 0, 0, 0, 0, 0)   # n_streams = 0
    assert lib.tfr_create(C.byref(cfg), C.byref(h)) == -1


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = tb.Config(C.sizeof(tb.Config), 0, 7, 0, 0, 1, 0, 0, 0)
    rc = lib.tfr_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value     # TFR_E_NODEVICE
    with pytest.raises(tb.TfrError):
        tb.Receiver()


def test_downconvert_validates_arguments_without_a_device(lib):
    import numpy as np
    iq = np.full(64, 128, np.uint8)
    out = np.zeros(64, np.int16)
    call = lambda passes, mem=tb.MEM_HOST, n=iq.size: lib.tfr_downconvert(0, iq.ctypes.data, n, passes, 0, out.ctypes.data, mem, 1, None)
    assert call(0) == -1 and b"passes" in lib.tfr_last_error()        # TFR_E_INVAL
    assert call(9) == -1
    assert call(2, mem=7) == -1
    assert call(2, n=2) == -1                                           # fewer than two IQ pairs
    assert lib.tfr_downconvert(0, None, 64, 2, 0, out.ctypes.data, tb.MEM_HOST, 1, None) == -1
    import torch
    if not torch.cuda.is_available():
        assert call(2) == -2 and b"no CPU fallback" in lib.tfr_last_error()   # TFR_E_NODEVICE


def test_streaming_downconvert_object_without_a_device(lib, tmp_path):
    """tfr_dc_create validates passes and fails loudly without a device; the host mirror host/dsp_stuff.h (the
    reference's `downconvert` class, dsp_stuff.h:46-56) compiles against the ABI and throws instead of falling back"""
    import subprocess
    import torch
    h = C.c_void_p()
    assert lib.tfr_dc_create(0, 0, C.byref(h)) == -1 and lib.tfr_dc_create(0, 6, C.byref(h)) == -1      # TFR_E_INVAL
    assert lib.tfr_dc_create(0, 2, None) == -1
    src = tmp_path / "dc.cpp"
    src.write_text('#include "tfrec_b200/host/dsp_stuff.h"\n'
                   'int main() { try { downconvert dc(2); int16_t b[64] = {0}; return dc.process_iq(b, 64, 0) == 16 ? 0 : 1; }\n'
                   '             catch (const std::runtime_error &) { return 2; } }\n')
    exe = tmp_path / "dc"
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", ROOT, "-o", str(exe), str(src), "-L" + os.path.join(ROOT, "tfrec_b200"),
                    "-ltfrb200", "-Wl,-rpath," + os.path.join(ROOT, "tfrec_b200")], check=True)
    rc = subprocess.run([str(exe)]).returncode
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc == 2 and lib.tfr_dc_create(0, 2, C.byref(h)) == -2                                     # TFR_E_NODEVICE


def test_product_does_not_touch_the_oracle():
    # the product path must never import, link or execute anything under oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tfrec_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "tfrec_oracle" not in txt, f
