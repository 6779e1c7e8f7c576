"""ctypes front end of the two CPU checkers -- TEST INFRASTRUCTURE, not product code.

  kind="oracle" : oracle/libbforacle.so, the plain-C restatement (bf_oracle.c) driven by
                  bf_blockdriver.c.
  kind="ref"    : oracle/_ref/libbfref.so, the reference's own fftw_convolver.c / convolver_xmm.c /
                  raw2real.h / real2raw.h compiled from /root/reference, driven by the same driver.

Both export the same entry points with prefix bfo_ / bfref_.  The convolver keeps its sizes in file
scope statics (fftw_convolver.c:36-49), so each library serves ONE (filter_length, realsize) at a
time: creating a new driver or calling cv_init() re-initialises it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from brutefir_b200 import _abi
from brutefir_b200.graph import FilterGraph

HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {"oracle": os.path.join(HERE, "libbforacle.so"), "ref": os.path.join(HERE, "_ref", "libbfref.so")}
_PREFIX = {"oracle": "bfo_", "ref": "bfref_"}
_libs = {}


def build(force: bool = False) -> None:
    """Compile the checkers (oracle/Makefile).  _ref is only (re)built where /root/reference exists."""
    args = ["make", "-C", HERE] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


class _Lib:
    def __init__(self, kind: str):
        if not os.path.exists(_PATHS[kind]):
            raise FileNotFoundError(f"{_PATHS[kind]} not built (make -C oracle)")
        self.kind = kind
        self.dll = C.CDLL(_PATHS[kind], mode=os.RTLD_LOCAL | os.RTLD_NOW)
        p = _PREFIX[kind]
        V, I, D = C.c_void_p, C.c_int, C.c_double

        def fn(name, restype, *argtypes):
            f = getattr(self.dll, p + name)
            f.restype = restype
            f.argtypes = list(argtypes)
            setattr(self, name, f)

        fn("create", I, C.POINTER(_abi.ConfigC), I, C.POINTER(V))
        fn("destroy", None, V)
        fn("coeff_from_taps", I, V, I, V, I, D)
        fn("coeff_set_block", I, V, I, I, V)
        fn("coeff_get_block", I, V, I, I, V)
        fn("coeff_runtime_block", I, V, I, I, V)
        fn("set_control", I, V, I, C.POINTER(_abi.FilterControlC))
        fn("set_mute", I, V, I, I, I)
        fn("set_subdelay", I, V, I, I, V, I)
        fn("get_overflow", I, V, I, C.POINTER(_abi.OverflowC))
        fn("debug_read", I, V, I, I, I, V)
        fn("run", D, V, I, V, C.c_size_t, V, C.c_size_t)
        fn("process_block", I, V, V, V)
        fn("cv_init", I, I, I)
        fn("cv_set_safety_limit", None, D)
        fn("cv_failed", I, I)
        fn("cv_cbufsize", I)
        fn("cv_raw2cbuf", None, V, V, V, C.POINTER(_abi.BufferFormatC))
        fn("cv_cbuf2raw", None, V, V, C.POINTER(_abi.BufferFormatC), C.POINTER(_abi.OverflowC))
        fn("cv_dither_init", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
        fn("cv_cbuf2raw_dither", None, V, V, C.POINTER(_abi.BufferFormatC), C.POINTER(_abi.OverflowC), C.c_int)
        fn("cv_time2freq", None, V, V)
        fn("cv_freq2time", None, V, V)
        fn("cv_mixnscale", None, C.POINTER(V), V, C.POINTER(D), I, I)
        fn("cv_convolve", None, V, V, V)
        fn("cv_convolve_add", None, V, V, V)
        fn("cv_convolve_inplace", None, V, V)
        fn("cv_dirac_convolve", None, V, V)
        fn("cv_dirac_convolve_inplace", None, V)
        fn("cv_crossfade_inplace", None, V, V, V)
        fn("cv_convolve_eval", None, V, V, V)
        fn("cv_coeffs2cbuf", I, V, I, D, V)
        fn("cv_runtime_coeffs2cbuf", None, V, V)
        fn("cv_dither_table", V, C.POINTER(I))
        fn("cv_dither_state", V, I)
        fn("cv_td_block_length", I, I)
        fn("cv_td_new", V, V, I)
        fn("cv_td_convolve", None, V, V)


MAX_LENGTH = 16384


def lib(kind: str = "oracle") -> _Lib:
    if kind not in _libs:
        l = _Lib(kind)
        if kind == "ref":
            # The reference's convolver_runtime_coeffs2cbuf keeps a `static` scratch buffer sized by the
            # FIRST configuration it runs under (fftw_convolver.c:579-584).  This process re-initialises
            # the convolver with many sizes, so size that scratch for the largest one up front.
            l.cv_init(MAX_LENGTH, 8)
            src = np.zeros(MAX_LENGTH, np.float64)
            dst = np.zeros(2 * MAX_LENGTH, np.float64)
            l.cv_runtime_coeffs2cbuf(C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data))
        _libs[kind] = l
    return _libs[kind]


def _ptr(a: np.ndarray) -> C.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def buffer_format_c(bf) -> _abi.BufferFormatC:
    c = _abi.BufferFormatC()
    c.sf.isfloat, c.sf.swap, c.sf.bytes, c.sf.sbytes = int(bf.sf.isfloat), int(bf.sf.swap), bf.sf.bytes, bf.sf.sbytes
    c.sf.scale, c.sf.format = bf.sf.scale, bf.sf.format
    c.sample_spacing, c.byte_offset = bf.sample_spacing, bf.byte_offset
    return c


class Convolver:
    """convolver.h, one call at a time, on numpy buffers (real dtype = float32 | float64)."""

    MIXMODE_INPUT, MIXMODE_OUTPUT = 1, 3

    def __init__(self, kind: str, length: int, realsize: int):
        self.l = lib(kind)
        if not self.l.cv_init(length, realsize):
            raise ValueError("convolver_init failed")
        self.L, self.N, self.realsize = length, 2 * length, realsize
        self.dtype = np.float32 if realsize == 4 else np.float64
        self.l.cv_failed(1)

    def new(self, n=None):
        return np.zeros(self.N if n is None else n, self.dtype)

    def failed(self) -> bool:
        return bool(self.l.cv_failed(1))

    def raw2cbuf(self, raw, cbuf, next_cbuf, bf):
        self.l.cv_raw2cbuf(_ptr(raw), _ptr(cbuf), _ptr(next_cbuf), C.byref(buffer_format_c(bf)))

    def cbuf2raw(self, cbuf, out, bf, overflow: _abi.OverflowC):
        self.l.cv_cbuf2raw(_ptr(cbuf), _ptr(out), C.byref(buffer_format_c(bf)), C.byref(overflow))

    def dither_init(self, n_channels, sample_rate, max_size=0):
        """dither_init (dither.c:75-139) for the per-call cbuf2raw_dither; one table per library."""
        assert self.l.cv_dither_init(n_channels, sample_rate, self.realsize, max_size, self.L) == 1

    def cbuf2raw_dither(self, cbuf, out, bf, overflow: _abi.OverflowC, index: int):
        self.l.cv_cbuf2raw_dither(_ptr(cbuf), _ptr(out), C.byref(buffer_format_c(bf)), C.byref(overflow), index)

    def time2freq(self, x):
        out = self.new()
        src = np.ascontiguousarray(x, self.dtype).copy()    # keep alive across the call
        self.l.cv_time2freq(_ptr(src), _ptr(out))
        return out

    def freq2time(self, x):
        out = self.new()
        src = np.ascontiguousarray(x, self.dtype).copy()
        self.l.cv_freq2time(_ptr(src), _ptr(out))
        return out

    def mixnscale(self, bufs, scales, mode):
        bufs = [np.ascontiguousarray(b, self.dtype) for b in bufs]
        arr = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        sc = (C.c_double * len(bufs))(*scales)
        out = self.new()
        self.l.cv_mixnscale(arr, _ptr(out), sc, len(bufs), mode)
        return out

    def convolve(self, b, c):
        out = self.new()
        self.l.cv_convolve(_ptr(b), _ptr(c), _ptr(out))
        return out

    def convolve_add(self, b, c, d):
        self.l.cv_convolve_add(_ptr(b), _ptr(c), _ptr(d))
        return d

    def convolve_inplace(self, b, c):
        self.l.cv_convolve_inplace(_ptr(b), _ptr(c))
        return b

    def dirac_convolve(self, b):
        out = self.new()
        self.l.cv_dirac_convolve(_ptr(b), _ptr(out))
        return out

    def crossfade_inplace(self, new_cbuf, old_cbuf):
        scratch = self.new(2 * self.N)
        self.l.cv_crossfade_inplace(_ptr(new_cbuf), _ptr(old_cbuf), _ptr(scratch))
        return new_cbuf

    def convolve_eval(self, inp, state):
        out = self.new()
        self.l.cv_convolve_eval(_ptr(inp), _ptr(state), _ptr(out))
        return out

    def coeffs2cbuf(self, taps, scale=1.0):
        taps = np.ascontiguousarray(taps, self.dtype)
        out = self.new()
        ok = self.l.cv_coeffs2cbuf(_ptr(taps), len(taps), scale, _ptr(out))
        return out if ok else None

    def runtime_coeffs2cbuf(self, taps_L):
        out = self.new()
        src = np.ascontiguousarray(taps_L, self.dtype)
        self.l.cv_runtime_coeffs2cbuf(_ptr(src), _ptr(out))
        return out

    def dither_table(self) -> np.ndarray:
        """A copy of the dither table (dither_randtab, dither.c:22-24) after dither_init()."""
        size = C.c_int(0)
        ptr = self.l.cv_dither_table(C.byref(size))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int8)), shape=(size.value,)).copy()

    def dither_state(self, index: int):
        """(randtab_ptr, sf[2], sd[2]) of channel `index` (struct dither_state, dither.h:17-22)."""
        st = C.cast(self.l.cv_dither_state(index), C.POINTER(_abi.DitherStateC)).contents
        return st.randtab_ptr, (st.sf[0], st.sf[1]), (st.sd[0], st.sd[1])

    def td_block_length(self, n_coeffs: int) -> int:
        return self.l.cv_td_block_length(n_coeffs)

    def td_new(self, taps):
        """convolver_td_new: an opaque handle (never freed, as in the reference), or None."""
        taps = np.ascontiguousarray(taps, self.dtype)
        return self.l.cv_td_new(_ptr(taps), len(taps))

    def td_convolve(self, tdc, overlap_block):
        """In place on a copy of the 2 * blocklen reals; returns the copy."""
        blk = np.array(overlap_block, self.dtype)
        self.l.cv_td_convolve(C.c_void_p(tdc), _ptr(blk))
        return blk


class BlockDriver:
    """filter_process() replay (bf_blockdriver.c) for a FilterGraph; same surface as brutefir_b200.Engine."""

    def __init__(self, kind: str, graph: FilterGraph, n_threads: int = 1):
        self.l = lib(kind)
        self.graph = graph
        self.dtype = np.float32 if graph.realsize == 4 else np.float64
        cfg, keep = graph.to_config()
        h = C.c_void_p()
        rc = self.l.create(C.byref(cfg), n_threads, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"driver create failed: {rc}")
        self.h = h

    def close(self):
        if self.h:
            self.l.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def coeff_from_taps(self, coeff, taps, scale=1.0):
        taps = np.ascontiguousarray(taps, self.dtype)
        rc = self.l.coeff_from_taps(self.h, coeff, _ptr(taps), len(taps), scale)
        if rc != 0:
            raise RuntimeError(f"coeff_from_taps: {rc}")

    def coeff_set_block(self, coeff, block, cbuf):
        src = np.ascontiguousarray(cbuf, self.dtype)
        assert self.l.coeff_set_block(self.h, coeff, block, _ptr(src)) == 0

    def coeff_get_block(self, coeff, block):
        out = np.zeros(self.graph.n_fft, self.dtype)
        assert self.l.coeff_get_block(self.h, coeff, block, _ptr(out)) == 0
        return out

    def set_control(self, filt, coeff, delayblocks=0, in_scales=None, out_scales=None, fscales=None):
        c = _abi.FilterControlC()
        c.coeff, c.delayblocks = coeff, delayblocks
        keep = []
        for io, s in ((0, in_scales), (1, out_scales)):
            if s is not None:
                arr = (C.c_double * len(s))(*s)
                keep.append(arr)
                c.scale[io] = C.cast(arr, C.POINTER(C.c_double))
        if fscales is not None:
            farr = (C.c_double * len(fscales))(*fscales)
            keep.append(farr)
            c.fscale = C.cast(farr, C.POINTER(C.c_double))
        assert self.l.set_control(self.h, filt, C.byref(c)) == 0

    def set_mute(self, io, channel, muted):
        assert self.l.set_mute(self.h, io, channel, 1 if muted else 0) == 0

    def set_subdelay(self, io, channel, taps=None):
        if taps is None:
            assert self.l.set_subdelay(self.h, io, channel, None, 0) == 0
        else:
            t = np.ascontiguousarray(taps, self.dtype)
            assert self.l.set_subdelay(self.h, io, channel, _ptr(t), len(t)) == 0

    def process_block(self, raw_in: np.ndarray) -> np.ndarray:
        out = np.zeros(self.graph.out_bytes, np.uint8)
        rc = self.l.process_block(self.h, _ptr(raw_in), _ptr(out))
        if rc != 0:
            raise RuntimeError("reference aborted (NaN/Inf or safety limit)")
        return out

    def run(self, raw_in_blocks: np.ndarray) -> np.ndarray:
        """raw_in_blocks: uint8[n_blocks, in_bytes] -> uint8[n_blocks, out_bytes]."""
        n = raw_in_blocks.shape[0]
        out = np.zeros((n, self.graph.out_bytes), np.uint8)
        t = self.l.run(self.h, n, _ptr(raw_in_blocks), raw_in_blocks.shape[1], _ptr(out), out.shape[1])
        if t < 0:
            raise RuntimeError("reference aborted (NaN/Inf or safety limit)")
        return out

    def run_timed(self, raw_in: np.ndarray, n_blocks: int) -> float:
        """Process the same input block n_blocks times; returns wall seconds."""
        out = np.zeros(self.graph.out_bytes, np.uint8)
        return self.l.run(self.h, n_blocks, _ptr(raw_in), 0, _ptr(out), 0)

    def overflow(self, out_channel) -> _abi.OverflowC:
        o = _abi.OverflowC()
        assert self.l.get_overflow(self.h, out_channel, C.byref(o)) == 0
        return o

    def debug_read(self, what, index, slot=0):
        out = np.zeros(self.graph.n_fft, self.dtype)
        assert self.l.debug_read(self.h, what, index, slot, _ptr(out)) == 0
        return out
