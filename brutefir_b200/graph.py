"""The filter graph the convolution engine runs: inputs, outputs, filters, coefficient sets.

This is the host-side mirror of what the reference keeps in ``struct bfconf`` for the hot path
(/root/reference/bfconf.h:22-80): ``filter_length``/``n_blocks``/``realsize``, the per-channel
``struct buffer_format`` (dai.h:30-34), ``struct bffilter`` + its initial ``struct bffilter_control``
(bfmod.h:118-133) and ``struct bfcoeff.n_blocks`` (bfmod.h:106-111).  ``to_config()`` lowers it to the
``struct bfcuda_config`` of include/bfcuda.h.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Sequence

from . import _abi
from .formats import BufferFormat


@dataclasses.dataclass
class Filter:
    """One ``filter { ... }`` block: ``from_inputs``, ``to_outputs`` with linear multipliers,
    ``coeff`` (-1 = none/dirac), ``delay`` in blocks, ``crossfade`` (brutefir.html filter section)."""
    inputs: Sequence[int]
    outputs: Sequence[int]
    in_scales: Sequence[float] | None = None
    out_scales: Sequence[float] | None = None
    coeff: int = -1
    delayblocks: int = 0
    crossfade: bool = False
    from_filters: Sequence[int] = ()
    fscales: Sequence[float] | None = None

    def __post_init__(self):
        self.inputs = list(self.inputs)
        self.outputs = list(self.outputs)
        self.in_scales = [1.0] * len(self.inputs) if self.in_scales is None else list(self.in_scales)
        self.out_scales = [1.0] * len(self.outputs) if self.out_scales is None else list(self.out_scales)
        self.from_filters = list(self.from_filters)
        self.fscales = [1.0] * len(self.from_filters) if self.fscales is None else list(self.fscales)


@dataclasses.dataclass
class FilterGraph:
    filter_length: int
    n_blocks: int
    realsize: int
    in_formats: List[BufferFormat]
    out_formats: List[BufferFormat]
    in_bytes: int
    out_bytes: int
    filters: List[Filter]
    coeff_n_blocks: List[int]
    safety_limit: float = 0.0
    sampling_rate: int = 48000
    apply_dither: Sequence[bool] | None = None      # per output channel (`dither: true` of its device)
    max_dither_table_size: int = 0
    powersave: bool = False                 # bfconf->powersave
    analog_powersave: float = 1.0           # linear level; >= 1.0: only exact zeros are silent (bfrun.c:722-772)
    out_physical: Sequence[int] | None = None   # virt2phys[OUT]: outputs with equal ids share a physical channel

    @property
    def n_fft(self) -> int:
        return 2 * self.filter_length

    def validate(self) -> None:
        L = self.filter_length
        if L < 4 or L & (L - 1):
            raise ValueError("filter_length must be a power of two >= 4")      # bfconf.c:1495-1520
        if self.realsize not in (4, 8):
            raise ValueError("realsize must be 4 or 8")                         # fftw_convolver.c:796-799
        for f in self.filters:
            if any(c < 0 or c >= len(self.in_formats) for c in f.inputs):
                raise ValueError("filter input channel out of range")
            if any(c < 0 or c >= len(self.out_formats) for c in f.outputs):
                raise ValueError("filter output channel out of range")
            if f.coeff >= len(self.coeff_n_blocks):
                raise ValueError("coefficient index out of range")

    def to_config(self, device: int = 0, flags: int = 0, mac_split: int = 0, max_batch: int = 1):
        """Build ``struct bfcuda_config``.  Returns (config, keepalive): ``keepalive`` owns the arrays
        the config points into and must outlive the bfcuda_create call."""
        self.validate()
        keep = []

        def fmt_array(bfs):
            arr = (_abi.BufferFormatC * max(1, len(bfs)))()
            for i, bf in enumerate(bfs):
                arr[i].sf.isfloat = int(bf.sf.isfloat)
                arr[i].sf.swap = int(bf.sf.swap)
                arr[i].sf.bytes = bf.sf.bytes
                arr[i].sf.sbytes = bf.sf.sbytes
                arr[i].sf.scale = bf.sf.scale
                arr[i].sf.format = bf.sf.format
                arr[i].sample_spacing = bf.sample_spacing
                arr[i].byte_offset = bf.byte_offset
            keep.append(arr)
            return arr

        def int_array(vals):
            arr = (C.c_int * max(1, len(vals)))(*vals)
            keep.append(arr)
            return arr

        def dbl_array(vals):
            arr = (C.c_double * max(1, len(vals)))(*vals)
            keep.append(arr)
            return arr

        cfg = _abi.ConfigC()
        cfg.filter_length = self.filter_length
        cfg.n_blocks = self.n_blocks
        cfg.realsize = self.realsize
        cfg.n_channels[0] = len(self.in_formats)
        cfg.n_channels[1] = len(self.out_formats)
        cfg.formats[0] = C.cast(fmt_array(self.in_formats), C.POINTER(_abi.BufferFormatC))
        cfg.formats[1] = C.cast(fmt_array(self.out_formats), C.POINTER(_abi.BufferFormatC))
        cfg.n_bytes[0] = self.in_bytes
        cfg.n_bytes[1] = self.out_bytes
        farr = (_abi.FilterC * max(1, len(self.filters)))()
        keep.append(farr)
        for i, f in enumerate(self.filters):
            farr[i].crossfade = int(f.crossfade)
            farr[i].n_channels[0] = len(f.inputs)
            farr[i].n_channels[1] = len(f.outputs)
            farr[i].channels[0] = C.cast(int_array(f.inputs), C.POINTER(C.c_int))
            farr[i].channels[1] = C.cast(int_array(f.outputs), C.POINTER(C.c_int))
            farr[i].scale[0] = C.cast(dbl_array(f.in_scales), C.POINTER(C.c_double))
            farr[i].scale[1] = C.cast(dbl_array(f.out_scales), C.POINTER(C.c_double))
            farr[i].n_filters_in = len(f.from_filters)
            farr[i].filters_in = C.cast(int_array(f.from_filters), C.POINTER(C.c_int))
            farr[i].fscale = C.cast(dbl_array(f.fscales), C.POINTER(C.c_double))
            farr[i].coeff = f.coeff
            farr[i].delayblocks = f.delayblocks
        cfg.n_filters = len(self.filters)
        cfg.filters = C.cast(farr, C.POINTER(_abi.FilterC))
        cfg.n_coeffs = len(self.coeff_n_blocks)
        cfg.coeff_n_blocks = C.cast(int_array(self.coeff_n_blocks), C.POINTER(C.c_int))
        cfg.safety_limit = self.safety_limit
        cfg.device = device
        cfg.flags = flags
        cfg.mac_split = mac_split
        if self.apply_dither is not None:
            assert len(self.apply_dither) == len(self.out_formats)
            cfg.apply_dither = C.cast(int_array([int(bool(v)) for v in self.apply_dither]), C.POINTER(C.c_int))
        cfg.sampling_rate = self.sampling_rate
        cfg.max_dither_table_size = self.max_dither_table_size
        cfg.max_batch = max_batch
        cfg.powersave = int(bool(self.powersave))
        cfg.analog_powersave = float(self.analog_powersave)
        if self.out_physical is not None:
            assert len(self.out_physical) == len(self.out_formats)
            cfg.out_physical = C.cast(int_array(list(self.out_physical)), C.POINTER(C.c_int))
        return cfg, keep

    # ---- derived figures used by bench.py -------------------------------------
    def block_seconds(self) -> float:
        return self.filter_length / float(self.sampling_rate)

    def taps_per_filter(self) -> int:
        return self.filter_length * self.n_blocks

    def gtap_mac_per_realtime(self) -> float:
        """Gtap-MAC/s delivered at realtime multiple 1."""
        return len(self.filters) * self.taps_per_filter() * self.sampling_rate / 1e9
