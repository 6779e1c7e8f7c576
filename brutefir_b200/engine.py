"""Block-level engine: Python face of include/bfcuda.h.

One :class:`Engine` replaces the per-block body of the reference's ``filter_process``
(/root/reference/bfrun.c:1420-2083) for a :class:`~brutefir_b200.graph.FilterGraph`: raw input block in,
raw output block out, with the delay lines, coefficient spectra and overflow counters resident in HBM.
All arithmetic happens in the CUDA library; this module only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import check
from .graph import FilterGraph


class PinnedBuffer:
    """Page-locked host memory (bfcuda_host_alloc) viewed as a numpy uint8 array."""

    def __init__(self, n_bytes: int, device: Optional[int] = None):
        lib = _abi.load_library()
        self._lib = lib
        # device given: on the NUMA node that GPU hangs off (one process per GPU on a multi-socket host)
        self.ptr = lib.bfcuda_host_alloc(n_bytes) if device is None else lib.bfcuda_host_alloc_near(device, n_bytes)
        if not self.ptr:
            raise MemoryError("bfcuda_host_alloc failed")
        self.n_bytes = n_bytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * n_bytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.bfcuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, graph: FilterGraph, device: int = 0, flags: int = 0, mac_split: int = 0, max_batch: int = 1):
        self.lib = _abi.load_library()
        self.graph = graph
        self.max_batch = max_batch
        self.dtype = np.float32 if graph.realsize == 4 else np.float64
        cfg, keep = graph.to_config(device=device, flags=flags, mac_split=mac_split, max_batch=max_batch)
        h = C.c_void_p()
        check(self.lib.bfcuda_create(C.byref(cfg), C.byref(h)))
        self.h = h

    # ---- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.bfcuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- coefficients (bfconf.c:1992-2019, fftw_convolver.c:526-596) ----------------------------
    def coeff_from_taps(self, coeff: int, taps, scale: float = 1.0):
        taps = np.ascontiguousarray(taps, self.dtype)
        check(self.lib.bfcuda_coeff_from_taps(self.h, coeff, taps.ctypes.data, len(taps), scale))

    def coeff_set_block(self, coeff: int, block: int, cbuf):
        cbuf = np.ascontiguousarray(cbuf, self.dtype)
        assert cbuf.size == self.graph.n_fft
        check(self.lib.bfcuda_coeff_set_block(self.h, coeff, block, cbuf.ctypes.data))

    def coeff_get_block(self, coeff: int, block: int) -> np.ndarray:
        out = np.zeros(self.graph.n_fft, self.dtype)
        check(self.lib.bfcuda_coeff_get_block(self.h, coeff, block, out.ctypes.data))
        return out

    def coeff_runtime_block(self, coeff: int, block: int, taps_L):
        taps = np.ascontiguousarray(taps_L, self.dtype)
        assert taps.size == self.graph.filter_length
        check(self.lib.bfcuda_coeff_runtime_block(self.h, coeff, block, taps.ctypes.data))

    # ---- control (bfmod.h:128-133, bfrun.c:1462-1478) ---------------------------------------------
    def set_control(self, filt: int, coeff: int, delayblocks: int = 0,
                    in_scales: Optional[Sequence[float]] = None, out_scales: Optional[Sequence[float]] = None,
                    fscales: Optional[Sequence[float]] = None):
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
        check(self.lib.bfcuda_set_control(self.h, filt, C.byref(c)))

    def set_subdelay(self, io: int, channel: int, taps=None):
        """Sub-sample delay FIR of one channel (taps = the windowed sinc of its delay step), None = off."""
        if taps is None:
            check(self.lib.bfcuda_set_subdelay(self.h, io, channel, None, 0))
        else:
            t = np.ascontiguousarray(taps, self.dtype)
            check(self.lib.bfcuda_set_subdelay(self.h, io, channel, t.ctypes.data, len(t)))

    def set_mute(self, io: int, channel: int, muted: bool):
        check(self.lib.bfcuda_set_mute(self.h, io, channel, 1 if muted else 0))

    def overflow(self, out_channel: int) -> _abi.OverflowC:
        o = _abi.OverflowC()
        check(self.lib.bfcuda_get_overflow(self.h, out_channel, C.byref(o)))
        return o

    def reset_overflow(self):
        check(self.lib.bfcuda_reset_overflow(self.h))

    # ---- block step ---------------------------------------------------------------------------
    def process_block(self, raw_in: np.ndarray, raw_out: Optional[np.ndarray] = None) -> np.ndarray:
        assert raw_in.dtype == np.uint8 and raw_in.size == self.graph.in_bytes and raw_in.flags["C_CONTIGUOUS"]
        if raw_out is None:
            raw_out = np.zeros(self.graph.out_bytes, np.uint8)
        check(self.lib.bfcuda_process_block(self.h, raw_in.ctypes.data, raw_out.ctypes.data))
        return raw_out

    def process_block_async(self, raw_in: np.ndarray, raw_out: np.ndarray):
        check(self.lib.bfcuda_process_block_async(self.h, raw_in.ctypes.data, raw_out.ctypes.data))

    def synchronize(self):
        check(self.lib.bfcuda_synchronize(self.h))

    def process_blocks_async(self, raw_in: np.ndarray, raw_out: np.ndarray, n_blocks: int):
        check(self.lib.bfcuda_process_blocks_async(self.h, n_blocks, raw_in.ctypes.data, raw_out.ctypes.data))

    def run(self, raw_in_blocks: np.ndarray) -> np.ndarray:
        """uint8[n_blocks, in_bytes] -> uint8[n_blocks, out_bytes], pipelined, max_batch blocks per call."""
        n = raw_in_blocks.shape[0]
        out = np.zeros((n, self.graph.out_bytes), np.uint8)
        raw_in_blocks = np.ascontiguousarray(raw_in_blocks)
        b = 0
        while b < n:
            nb = min(self.max_batch, n - b)
            self.process_blocks_async(raw_in_blocks[b:b + nb], out[b:b + nb], nb)
            b += nb
        self.synchronize()
        return out

    def upload_input(self, raw_in: np.ndarray):
        check(self.lib.bfcuda_upload_input(self.h, raw_in.ctypes.data))

    def process_block_device(self):
        check(self.lib.bfcuda_process_block_device(self.h))

    def upload_inputs(self, raw_in_blocks: np.ndarray):
        check(self.lib.bfcuda_upload_inputs(self.h, raw_in_blocks.shape[0], raw_in_blocks.ctypes.data))

    def process_blocks_device(self, n_blocks: int):
        check(self.lib.bfcuda_process_blocks_device(self.h, n_blocks))

    def download_output(self) -> np.ndarray:
        out = np.zeros(self.graph.out_bytes, np.uint8)
        check(self.lib.bfcuda_download_output(self.h, out.ctypes.data))
        return out

    # ---- measurement ----------------------------------------------------------------------------
    def timer_start(self):
        check(self.lib.bfcuda_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(self.lib.bfcuda_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def copy_baseline(self, raw_in: np.ndarray, raw_out: np.ndarray, n_blocks: int, reps: int):
        """(ms per round, H2D GB/s, D2H GB/s) of the host-buffer path's copies alone (no kernels)."""
        ms, a, b = C.c_double(), C.c_double(), C.c_double()
        check(self.lib.bfcuda_copy_baseline(self.h, n_blocks, raw_in.ctypes.data, raw_out.ctypes.data, reps,
                                            C.byref(ms), C.byref(a), C.byref(b)))
        return ms.value, a.value, b.value

    def stage_times(self):
        ms = (C.c_double * 3)()
        nb, nl = C.c_long(), C.c_long()
        check(self.lib.bfcuda_stage_times(self.h, ms, C.byref(nb), C.byref(nl)))
        return list(ms), nb.value, nl.value

    def set_stage_timing(self, on: bool):
        """BFCUDA_FLAG_STAGE_TIMING at run time (per-stage CUDA events; costs a few percent of throughput)."""
        check(self.lib.bfcuda_set_stage_timing(self.h, 1 if on else 0))

    def set_serial_stages(self, on: bool):
        """BFCUDA_FLAG_SERIAL_STAGES at run time: no overlap between the stages of consecutive launches, so that
        stage_times() reports each stage running alone (roofline measurements)."""
        check(self.lib.bfcuda_set_serial_stages(self.h, 1 if on else 0))

    def info(self) -> _abi.InfoC:
        i = _abi.InfoC()
        check(self.lib.bfcuda_get_info(self.h, C.byref(i)))
        return i

    def debug_read(self, what: int, index: int, slot: int = 0) -> np.ndarray:
        out = np.zeros(self.graph.n_fft, self.dtype)
        check(self.lib.bfcuda_debug_read(self.h, what, index, slot, out.ctypes.data))
        return out

    # ---- multi-GPU ------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(_abi.load_library().bfcuda_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.bfcuda_comm_init(self.h, rank, n_ranks, buf))

    def comm_shared_outputs(self, out_channels: Sequence[int]):
        arr = (C.c_int * max(1, len(out_channels)))(*out_channels)
        check(self.lib.bfcuda_comm_shared_outputs(self.h, len(out_channels), arr))
