"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU it refuses to
compute instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from brutefir_b200 import _abi, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b((?:bfcuda|convolver)_[a-z0-9_]+)\s*\(", text):
        names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    declared = declared_functions("bfcuda.h") | declared_functions("bfcuda_convolver.h")
    assert len(declared) >= 50
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert missing == []
    assert set(_abi.ENGINE_SYMBOLS) | set(_abi.CONVOLVER_SYMBOLS) == declared


def test_ctypes_structs_match_the_header_layout():
    # struct sample_format: 4 ints, double, int (dai.h:21-28) -> 32 bytes with padding; buffer_format adds 2 ints
    assert C.sizeof(_abi.SampleFormatC) == 32 and C.sizeof(_abi.BufferFormatC) == 40
    assert C.sizeof(_abi.OverflowC) == 24           # struct bfoverflow, bfmod.h:99-104
    assert _abi.OverflowC.largest.offset == 8 and _abi.OverflowC.max.offset == 16
    # struct dither_state, dither.h:17-22: int, pointer, float[2], double[2]
    assert C.sizeof(_abi.DitherStateC) == 40 and _abi.DitherStateC.randtab.offset == 8
    assert _abi.DitherStateC.sf.offset == 16 and _abi.DitherStateC.sd.offset == 24


def test_no_cpu_fallback_without_a_device():
    lib = _abi.load_library()
    if lib.bfcuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg, keep = configs.config_c5().to_config()
    h = C.c_void_p()
    rc = lib.bfcuda_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value                 # BFCUDA_ENODEV
    assert b"no CPU fallback" in lib.bfcuda_strerror()
    from brutefir_b200 import convolver
    assert convolver.convolver_init("", 64, 4) is False


def test_invalid_configurations_are_rejected_like_convolver_init():
    lib = _abi.load_library()
    h = C.c_void_p()
    g = configs.config_c5()
    for mutate, needle in ((lambda c: setattr(c, "realsize", 6), b"Invalid real size"),
                           (lambda c: setattr(c, "filter_length", 48), b"Invalid length"),
                           (lambda c: setattr(c, "filter_length", 1 << 23), b"beyond the four-step transform")):
        cfg, keep = g.to_config()
        mutate(cfg)
        rc = lib.bfcuda_create(C.byref(cfg), C.byref(h))
        assert rc < 0 and needle in lib.bfcuda_strerror()
