"""The reference's per-call convolver interface (convolver.h), executed by the CUDA library.

Function for function the surface of /root/reference/convolver.h:16-152 as exported by libbfcuda.so
(include/bfcuda_convolver.h), on numpy buffers: same names, argument order and buffer layouts, so the
parity tests read like calls into the reference.  Every call runs on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .formats import BufferFormat

MIXMODE_INPUT, MIXMODE_INPUT_ADD, MIXMODE_OUTPUT = 1, 2, 3

_EXIT_CB = C.CFUNCTYPE(None, C.c_int)
_state = {"exit_status": None, "cb": None, "realsize": 0, "L": 0}


def _lib():
    lib = _abi.load_library()
    if _state["cb"] is None:
        def on_exit(status):
            _state["exit_status"] = status
        _state["cb"] = _EXIT_CB(on_exit)
        lib.bfcuda_convolver_set_host.argtypes = [_EXIT_CB, C.c_int, C.c_double]
        lib.bfcuda_convolver_set_host(_state["cb"], 1, 0.0)
        lib.bfcuda_convolver_last_error.restype = C.c_char_p
        lib.convolver_coeffs2cbuf.restype = C.c_void_p
        lib.convolver_coeffs2cbuf.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        lib.convolver_fftplan.restype = C.c_void_p
        lib.convolver_td_new.restype = C.c_void_p
        lib.convolver_init.argtypes = [C.c_char_p, C.c_int, C.c_int]
    return lib


def set_safety_limit(limit: float):
    lib = _lib()
    lib.bfcuda_convolver_set_host(_state["cb"], 1, limit)


def exit_status(reset: bool = True):
    """Status the convolver passed to the host's bf_exit() hook since the last call (None = none)."""
    s = _state["exit_status"]
    if reset:
        _state["exit_status"] = None
    return s


def _dtype():
    return np.float32 if _state["realsize"] == 4 else np.float64


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _bf(bf: BufferFormat) -> _abi.BufferFormatC:
    c = _abi.BufferFormatC()
    c.sf.isfloat, c.sf.swap, c.sf.bytes, c.sf.sbytes = int(bf.sf.isfloat), int(bf.sf.swap), bf.sf.bytes, bf.sf.sbytes
    c.sf.scale, c.sf.format = bf.sf.scale, bf.sf.format
    c.sample_spacing, c.byte_offset = bf.sample_spacing, bf.byte_offset
    return c


def convolver_init(config_filename: str, length: int, realsize: int) -> bool:
    ok = bool(_lib().convolver_init(config_filename.encode(), length, realsize))
    if ok:
        _state["realsize"], _state["L"] = realsize, length
    return ok


def convolver_cbufsize() -> int:
    return _lib().convolver_cbufsize()


def new_cbuf(n=None) -> np.ndarray:
    return np.zeros(2 * _state["L"] if n is None else n, _dtype())


def convolver_raw2cbuf(rawbuf: np.ndarray, cbuf: np.ndarray, next_cbuf: np.ndarray, bf: BufferFormat):
    c = _bf(bf)
    _lib().convolver_raw2cbuf(_p(rawbuf), _p(cbuf), _p(next_cbuf), C.byref(c), None, None)


def convolver_time2freq(input_cbuf: np.ndarray, output_cbuf: np.ndarray):
    _lib().convolver_time2freq(_p(input_cbuf), _p(output_cbuf))


def convolver_freq2time(input_cbuf: np.ndarray, output_cbuf: np.ndarray):
    _lib().convolver_freq2time(_p(input_cbuf), _p(output_cbuf))


def convolver_mixnscale(input_cbufs, output_cbuf: np.ndarray, scales, mixmode: int):
    arr = (C.c_void_p * len(input_cbufs))(*[b.ctypes.data for b in input_cbufs])
    sc = (C.c_double * len(input_cbufs))(*scales)
    _lib().convolver_mixnscale(arr, _p(output_cbuf), sc, len(input_cbufs), mixmode)


def convolver_convolve_inplace(cbuf, coeffs):
    _lib().convolver_convolve_inplace(_p(cbuf), _p(coeffs))


def convolver_convolve(input_cbuf, coeffs, output_cbuf):
    _lib().convolver_convolve(_p(input_cbuf), _p(coeffs), _p(output_cbuf))


def convolver_convolve_add(input_cbuf, coeffs, output_cbuf):
    _lib().convolver_convolve_add(_p(input_cbuf), _p(coeffs), _p(output_cbuf))


def convolver_crossfade_inplace(input_cbuf, crossfade_cbuf, buffer_cbuf):
    _lib().convolver_crossfade_inplace(_p(input_cbuf), _p(crossfade_cbuf), _p(buffer_cbuf))


def convolver_dirac_convolve(input_cbuf, output_cbuf):
    _lib().convolver_dirac_convolve(_p(input_cbuf), _p(output_cbuf))


def convolver_dirac_convolve_inplace(cbuf):
    _lib().convolver_dirac_convolve_inplace(_p(cbuf))


def convolver_convolve_eval(input_cbuf, buffer_cbuf, output_cbuf):
    _lib().convolver_convolve_eval(_p(input_cbuf), _p(buffer_cbuf), _p(output_cbuf))


def convolver_cbuf2raw(cbuf, outbuf, bf: BufferFormat, apply_dither: bool, overflow: _abi.OverflowC):
    c = _bf(bf)
    _lib().convolver_cbuf2raw(_p(cbuf), _p(outbuf), C.byref(c), int(apply_dither), None, C.byref(overflow))


def set_dither_table(table: np.ndarray):
    """bfcuda_convolver_set_dither_table: the host's dither_randtab (int8, kept alive and written by the preloop)."""
    assert table.dtype == np.int8 and table.flags.c_contiguous
    lib = _lib()
    lib.bfcuda_convolver_set_dither_table.restype = None
    lib.bfcuda_convolver_set_dither_table.argtypes = [C.c_void_p, C.c_int]
    _state["dither_table"] = table
    lib.bfcuda_convolver_set_dither_table(C.c_void_p(table.ctypes.data), len(table))


def convolver_cbuf2raw_dither(cbuf, outbuf, bf: BufferFormat, state: _abi.DitherStateC, overflow: _abi.OverflowC):
    """convolver_cbuf2raw with apply_dither = true and the channel's struct dither_state."""
    b = _bf(bf)
    _lib().convolver_cbuf2raw(_p(cbuf), _p(outbuf), C.byref(b), 1, C.byref(state), C.byref(overflow))


def convolver_coeffs2cbuf(coeffs, scale: float, optional_dest: np.ndarray):
    """Returns optional_dest, or None where the reference returns NULL (NaN/Inf among the taps)."""
    coeffs = np.ascontiguousarray(coeffs, _dtype())
    r = _lib().convolver_coeffs2cbuf(_p(coeffs), len(coeffs), scale, _p(optional_dest))
    return optional_dest if r else None


def convolver_runtime_coeffs2cbuf(src, dest):
    src = np.ascontiguousarray(src, _dtype())   # keep alive across the call
    _lib().convolver_runtime_coeffs2cbuf(_p(src), _p(dest))


def convolver_verify_cbuf(cbufs) -> bool:
    arr = (C.c_void_p * len(cbufs))(*[b.ctypes.data for b in cbufs])
    return bool(_lib().convolver_verify_cbuf(arr, len(cbufs)))


def convolver_td_block_length(n_coeffs: int) -> int:
    return _lib().convolver_td_block_length(n_coeffs)


def convolver_td_new(coeffs: np.ndarray):
    """convolver_td_new (convolver.h:140-142): opaque handle or None; coefficients in the convolver's real type."""
    lib = _lib()
    lib.convolver_td_new.restype = C.c_void_p
    lib.convolver_td_new.argtypes = [C.c_void_p, C.c_int]
    taps = np.ascontiguousarray(coeffs, _dtype())
    return lib.convolver_td_new(_p(taps), len(taps))


def convolver_td_convolve(tdc, overlap_block: np.ndarray):
    """convolver_td_convolve (convolver.h:144-146): in place on the caller's 2 * block_length reals."""
    lib = _lib()
    lib.convolver_td_convolve.restype = None
    lib.convolver_td_convolve.argtypes = [C.c_void_p, C.c_void_p]
    assert overlap_block.dtype == _dtype() and overlap_block.flags.c_contiguous
    lib.convolver_td_convolve(C.c_void_p(tdc), _p(overlap_block))


def convolver_td_delete(tdc):
    lib = _lib()
    lib.bfcuda_convolver_td_delete.restype = None
    lib.bfcuda_convolver_td_delete.argtypes = [C.c_void_p]
    lib.bfcuda_convolver_td_delete(C.c_void_p(tdc))
