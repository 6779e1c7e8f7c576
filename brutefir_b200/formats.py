"""Sample formats of the raw PCM blocks entering and leaving the convolver.

Mirrors the table in the reference's ``parse_sample_format`` (/root/reference/bfconf.c:358-533) and the
``struct sample_format`` / ``struct buffer_format`` descriptors (/root/reference/dai.h:21-34) that
``convolver_raw2cbuf`` / ``convolver_cbuf2raw`` receive.  The host is little endian (x86-64), so
``swap`` is set for the ``_BE`` formats.
"""
from __future__ import annotations

import dataclasses
import sys

import numpy as np

assert sys.byteorder == "little", "the reference's swap flags below assume a little-endian host"

# BF_SAMPLE_FORMAT_* ids, /root/reference/bfmod.h:33-48
FORMAT_IDS = {
    "S8": 1, "S16_LE": 2, "S16_BE": 3, "S24_LE": 6, "S24_BE": 7, "S24_4LE": 8, "S24_4BE": 9,
    "S32_LE": 10, "S32_BE": 11, "FLOAT_LE": 12, "FLOAT_BE": 13, "FLOAT64_LE": 14, "FLOAT64_BE": 15,
}

_ALIASES = {"S24_3LE": "S24_LE", "S24_3BE": "S24_BE", "S16_NE": "S16_LE", "S24_NE": "S24_LE",
            "S24_3NE": "S24_LE", "S24_4NE": "S24_4LE", "S32_NE": "S32_LE"}


@dataclasses.dataclass(frozen=True)
class SampleFormat:
    name: str
    isfloat: bool
    swap: bool
    bytes: int      # storage bytes per sample
    sbytes: int     # significant bytes
    format: int

    @property
    def scale(self) -> float:
        """bfconf.c:473-477: 1.0 for floats, 2^-(8*sbytes-1) for integers."""
        return 1.0 if self.isfloat else 1.0 / float(1 << (8 * self.sbytes - 1))

    @property
    def bits(self) -> int:
        return 8 * self.sbytes

    @property
    def overflow_max(self) -> float:
        """bfrun.c:2264-2279."""
        return 1.0 if self.isfloat else float((1 << (self.bits - 1)) - 1)


def parse_sample_format(name: str) -> SampleFormat:
    key = name.upper()
    key = _ALIASES.get(key, key)
    if key in ("FLOAT_NE", "FLOAT64_NE"):
        # bfconf.c:449-453, 465-469: the reference never sets isfloat for these (and gives FLOAT64_NE
        # 4 bytes), so they silently run as S32.  Refuse instead of reproducing the accident.
        raise ValueError(f"{name}: ambiguous in the reference (parsed as S32 there); use an explicit _LE/_BE format")
    table = {
        "S8": (False, 1, 1, False), "S16_LE": (False, 2, 2, True), "S16_BE": (False, 2, 2, False),
        "S24_LE": (False, 3, 3, True), "S24_BE": (False, 3, 3, False),
        "S24_4LE": (False, 4, 3, True), "S24_4BE": (False, 4, 3, False),
        "S32_LE": (False, 4, 4, True), "S32_BE": (False, 4, 4, False),
        "FLOAT_LE": (True, 4, 4, True), "FLOAT_BE": (True, 4, 4, False),
        "FLOAT64_LE": (True, 8, 8, True), "FLOAT64_BE": (True, 8, 8, False),
    }
    if key not in table:
        raise ValueError(f"Unknown sample format {name!r}")
    isfloat, nbytes, sbytes, little = table[key]
    return SampleFormat(key, isfloat, not little, nbytes, sbytes, FORMAT_IDS[key])


@dataclasses.dataclass(frozen=True)
class BufferFormat:
    """One channel inside a raw block: dai.h:30-34."""
    sf: SampleFormat
    sample_spacing: int     # in samples
    byte_offset: int        # in bytes


def interleaved_layout(n_channels: int, fmt: str | SampleFormat, fragsize: int, alignment: int = 32):
    """Buffer formats and block size of one interleaved device, as ``calc_buffer_format`` lays it out
    (/root/reference/dai.c:537-576): channel c at byte offset c*bytes, spacing n_channels, block
    padded to ALIGNMENT (sysarch.h:10)."""
    sf = parse_sample_format(fmt) if isinstance(fmt, str) else fmt
    bfs = [BufferFormat(sf, n_channels, c * sf.bytes) for c in range(n_channels)]
    n_bytes = n_channels * sf.bytes * fragsize
    if n_bytes % alignment:
        n_bytes += alignment - n_bytes % alignment
    return bfs, n_bytes


def planar_layout(n_channels: int, fmt: str | SampleFormat, fragsize: int, alignment: int = 32):
    """Non-interleaved device (dai.c:559-563): channel c occupies its own contiguous run."""
    sf = parse_sample_format(fmt) if isinstance(fmt, str) else fmt
    bfs = [BufferFormat(sf, 1, c * sf.bytes * fragsize) for c in range(n_channels)]
    n_bytes = n_channels * sf.bytes * fragsize
    if n_bytes % alignment:
        n_bytes += alignment - n_bytes % alignment
    return bfs, n_bytes


def encode_samples(values: np.ndarray, sf: SampleFormat) -> np.ndarray:
    """Pack sample values (integers in LSB units, or floats) into raw bytes of format ``sf``.
    Returns uint8 of shape values.shape + (sf.bytes,).  Test/bench helper: the inverse of raw2real."""
    v = np.asarray(values)
    if sf.isfloat:
        raw = v.astype("<f4" if sf.bytes == 4 else "<f8").view(np.uint8).reshape(v.shape + (sf.bytes,))
    else:
        i = v.astype(np.int64)
        if sf.bytes == 3:
            raw = (i.astype("<i4").view(np.uint8).reshape(v.shape + (4,)))[..., :3]
        else:
            raw = i.astype({1: "i1", 2: "<i2", 4: "<i4"}[sf.bytes]).view(np.uint8).reshape(v.shape + (sf.bytes,))
    raw = np.ascontiguousarray(raw)
    return raw[..., ::-1].copy() if sf.swap else raw


def decode_samples(raw: np.ndarray, sf: SampleFormat) -> np.ndarray:
    """Inverse of :func:`encode_samples`: uint8[..., bytes] -> int64 or float64 values."""
    r = np.asarray(raw, dtype=np.uint8)
    if sf.swap:
        r = r[..., ::-1]
    r = np.ascontiguousarray(r)
    if sf.isfloat:
        return r.view("<f4" if sf.bytes == 4 else "<f8")[..., 0].astype(np.float64)
    if sf.bytes == 3:
        pad = np.zeros(r.shape[:-1] + (4,), np.uint8)
        pad[..., 1:] = r
        return (pad.view("<i4")[..., 0] >> 8).astype(np.int64)
    return r.view({1: "i1", 2: "<i2", 4: "<i4"}[sf.bytes])[..., 0].astype(np.int64)


def pack_block(channels: np.ndarray, bfs, n_bytes: int) -> np.ndarray:
    """channels[c, n] sample values -> one raw block (uint8[n_bytes]) laid out per ``bfs``."""
    out = np.zeros(n_bytes, np.uint8)
    for c, bf in enumerate(bfs):
        raw = encode_samples(channels[c], bf.sf)
        b = bf.sf.bytes
        idx = bf.byte_offset + np.arange(raw.shape[0])[:, None] * (bf.sample_spacing * b) + np.arange(b)[None, :]
        out[idx] = raw
    return out


def unpack_block(block: np.ndarray, bfs, fragsize: int) -> np.ndarray:
    """Inverse of :func:`pack_block`; returns float64[c, n] (integers exactly representable)."""
    res = np.zeros((len(bfs), fragsize), np.float64)
    for c, bf in enumerate(bfs):
        b = bf.sf.bytes
        idx = bf.byte_offset + np.arange(fragsize)[:, None] * (bf.sample_spacing * b) + np.arange(b)[None, :]
        res[c] = decode_samples(block[idx], bf.sf)
    return res
