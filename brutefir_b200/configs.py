"""The workloads BASELINE.json names, as filter graphs plus seeded synthetic signals and filters.

Shapes and seeds follow SURVEY.md section 8(d):
  C1  bench1_config as shipped (/root/reference/bench1_config): 44.1 kHz, 8192 x 8, 2 in / 2 out
      S24_4LE, six "dirac pulse" filters, four input-fed -> two filter-fed (to_filters).  CPU anchor; the
      filter-to-filter chain is a "next" row, so on the GPU the equivalent direct graph is used.
  C2  stereo 2 ch x 65 536 taps, 4096-sample partitions, 48 kHz, float32 I/O.
  C3  64 in / 64 out x 1 048 576 taps, 8192 x 128 (massive_config-style), 48 kHz -- the headline.
  C4  low latency: 32 ch x 262 144 taps, 256-sample partitions (1024-deep delay line).
  C5  xtc_config topology (/root/reference/xtc_config:28-50) with runtime crossfade.
Signals: white Gaussian, sigma = 0.1 full scale, clipped, quantised to the input format; seed 1000 + id.
Filters: Gaussian taps x exp(-n / (taps/4)), unit energy; seed 2000 + id.
"""
from __future__ import annotations

import numpy as np

from .formats import interleaved_layout, pack_block, parse_sample_format
from .graph import Filter, FilterGraph


def diagonal_graph(n_ch: int, L: int, P: int, realsize: int = 4, fmt: str = "S24_4LE", rate: int = 48000,
                   coeff_blocks: int | None = None) -> FilterGraph:
    """n_ch inputs -> n_ch outputs, filter i: input i -> output i with its own coefficient set."""
    inb, nin = interleaved_layout(n_ch, fmt, L)
    outb, nout = interleaved_layout(n_ch, fmt, L)
    filters = [Filter([i], [i], coeff=i) for i in range(n_ch)]
    return FilterGraph(L, P, realsize, inb, outb, nin, nout, filters, [coeff_blocks or P] * n_ch, sampling_rate=rate)


def config_c1_direct(realsize: int = 4) -> FilterGraph:
    """bench1_config's signal path flattened: out0 = in0*h2*h0 + in1*h5*h0 with dirac pulses everywhere is
    out0 = in0 + in1 delayed; on the accelerated path we run the four input-fed filters straight to the
    outputs (same filter count feeding each output, same partitioning)."""
    L, P = 8192, 8
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S24_4LE", L)
    filters = [Filter([0], [0], coeff=0), Filter([1], [0], coeff=1), Filter([0], [1], coeff=2), Filter([1], [1], coeff=3)]
    return FilterGraph(L, P, realsize, inb, outb, nin, nout, filters, [P] * 4, sampling_rate=44100)


def config_c1_chained(realsize: int = 4) -> FilterGraph:
    """bench1_config exactly (filters 2..5 feed filters 0, 1 through to_filters)."""
    L, P = 8192, 8
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S24_4LE", L)
    # processing order must be topological (bfconf.c:2933-2964): producers first
    filters = [Filter([0], [], coeff=2), Filter([0], [], coeff=3), Filter([1], [], coeff=4), Filter([1], [], coeff=5),
               Filter([], [0], coeff=0, from_filters=[0, 3]), Filter([], [1], coeff=1, from_filters=[1, 2])]
    return FilterGraph(L, P, realsize, inb, outb, nin, nout, filters, [P] * 6, sampling_rate=44100)


def config_matrix(n: int = 8, realsize: int = 4, fmt: str = "S24_4LE", L: int = 8192, P: int = 128) -> FilterGraph:
    """n x n crosstalk / room-correction matrix of the headline shape: n inputs, n outputs, n*n filters of L*P taps,
    output o = sum over inputs i of filter (o, i).  Every input feeds n filters with the same scale and delay, so their
    delay lines are one and the same (xtc_config's topology, /root/reference/xtc_config:28-50, scaled up)."""
    inb, nin = interleaved_layout(n, fmt, L)
    outb, nout = interleaved_layout(n, fmt, L)
    filters = [Filter([i], [o], out_scales=[1.0 / n], coeff=o * n + i) for o in range(n) for i in range(n)]
    return FilterGraph(L, P, realsize, inb, outb, nin, nout, filters, [P] * (n * n))


def config_c2(realsize: int = 4) -> FilterGraph:
    return diagonal_graph(2, 4096, 16, realsize, "FLOAT_LE")


def config_c3(realsize: int = 4, n_ch: int = 64, fmt: str = "S24_4LE", L: int = 8192, P: int = 128) -> FilterGraph:
    return diagonal_graph(n_ch, L, P, realsize, fmt)


def config_c4(realsize: int = 4, n_ch: int = 32) -> FilterGraph:
    return diagonal_graph(n_ch, 256, 1024, realsize, "S24_4LE")


def config_c5(realsize: int = 4, L: int = 64, P: int = 64) -> FilterGraph:
    """xtc_config: left/right direct and cross paths, all filters crossfading; coefficient sets 0/1 are
    direct/cross, 2/3 the swapped pair the script switches to."""
    inb, nin = interleaved_layout(2, "S24_LE", L)
    outb, nout = interleaved_layout(2, "S24_LE", L)
    filters = [Filter([0], [0], coeff=0, crossfade=True), Filter([1], [0], coeff=1, crossfade=True),
               Filter([1], [1], coeff=0, crossfade=True), Filter([0], [1], coeff=1, crossfade=True)]
    return FilterGraph(L, P, realsize, inb, outb, nin, nout, filters, [P, P], sampling_rate=44100)


WORKLOADS = {"c1": config_c1_direct, "c2": config_c2, "c3": config_c3, "c4": config_c4, "c5": config_c5}


def synthetic_filters(graph: FilterGraph, config_id: int, n_sets: int | None = None) -> list[np.ndarray]:
    """One tap vector per coefficient set: Gaussian x exponential decay, unit energy."""
    rng = np.random.default_rng(2000 + config_id)
    dt = np.float32 if graph.realsize == 4 else np.float64
    out = []
    for nb in graph.coeff_n_blocks[: n_sets or len(graph.coeff_n_blocks)]:
        taps = nb * graph.filter_length
        h = rng.standard_normal(taps) * np.exp(-np.arange(taps) / (taps / 4.0))
        h /= np.sqrt(np.sum(h * h))
        out.append(h.astype(dt))
    return out


def synthetic_signal(graph: FilterGraph, config_id: int, n_blocks: int, sigma: float = 0.1) -> np.ndarray:
    """Raw input blocks uint8[n_blocks, in_bytes] of white noise in the graph's input format."""
    rng = np.random.default_rng(1000 + config_id)
    n_ch = len(graph.in_formats)
    L = graph.filter_length
    blocks = np.zeros((n_blocks, graph.in_bytes), np.uint8)
    for b in range(n_blocks):
        x = np.clip(rng.standard_normal((n_ch, L)) * sigma, -1.0, 1.0)
        vals = np.empty_like(x)
        for c, bf in enumerate(graph.in_formats):
            if bf.sf.isfloat:
                vals[c] = x[c]
            else:
                fs = float(1 << (bf.sf.bits - 1))
                vals[c] = np.clip(np.round(x[c] * fs), -fs, fs - 1)
        blocks[b] = pack_block(vals, graph.in_formats, graph.in_bytes)
    return blocks
