"""Partitions longer than one thread block's transform: the four-step FFT path (brutefir_b200/csrc/bf_fft4.cu).

The reference takes any power-of-two filter_length (fftw_convolver.c:784-808) and ships bench3_config with
`filter_length: 65536` (/root/reference/bench3_config:2).  Compared with the oracle's replay of filter_process() on the
same seeded inputs, north_star tolerances (1 LSB at 24 bit for float_bits 32, identical samples at float_bits 64)."""
import numpy as np
import pytest

from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu


def run_pair(g, taps, sig, script=None, B=1):
    with Engine(g, max_batch=B) as e:
        d = po.BlockDriver("oracle", g, n_threads=4)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        if script is None:
            got, ref = e.run(sig), d.run(sig)
        else:
            got, ref = [], []
            for b in range(sig.shape[0]):
                for filt, coeff in script.get(b, ()):
                    e.set_control(filt, coeff)
                    d.set_control(filt, coeff)
                got.append(e.process_block(sig[b]))
                ref.append(d.process_block(sig[b]))
            got, ref = np.stack(got), np.stack(ref)
        d.close()
    L = g.filter_length
    return unpack_run(got, g.out_formats, L), unpack_run(ref, g.out_formats, L)


@pytest.mark.parametrize("rs,L,P,B", [(4, 32768, 3, 1), (4, 65536, 2, 1), (4, 131072, 1, 2), (8, 16384, 3, 1), (8, 32768, 2, 2)])
def test_long_partitions_against_the_oracle(gpu_lib, oracle_libs, rs, L, P, B):
    g = configs.diagonal_graph(3, L, P, rs, "S24_4LE")
    taps = configs.synthetic_filters(g, 17)
    sig = configs.synthetic_signal(g, 17, P + 3, sigma=0.01)
    y, r = run_pair(g, taps, sig, B=B)
    assert np.abs(r).max() > 1e4
    assert np.abs(y - r).max() <= (1 if rs == 4 else 0)


def test_bench3_config_shape_unit_pulses(gpu_lib):
    """bench3_config: 26 channels, filter_length 65536 unpartitioned, every filter the "dirac pulse" coefficient:
    the output is the input, exactly (integer samples)."""
    g = configs.diagonal_graph(26, 65536, 1, 4, "S24_4LE", rate=44100)
    g.filters = [Filter([i], [i], coeff=0) for i in range(26)]
    g.coeff_n_blocks = [1]
    pulse = np.zeros(65536, np.float32)
    pulse[0] = 1.0
    sig = configs.synthetic_signal(g, 3, 3, sigma=0.004)     # -48 dBFS: the float32 round trip stays below 1/2 LSB
    with Engine(g) as e:
        e.coeff_from_taps(0, pulse)
        out = e.run(sig)
    x, y = unpack_run(sig, g.in_formats, 65536), unpack_run(out, g.out_formats, 65536)
    assert np.abs(x).max() > 5e4 and np.array_equal(x, y)


def test_long_partitions_mix_and_crossfade(gpu_lib, oracle_libs):
    """two inputs mixed into one filter, two filters into one output, a crossfaded coefficient swap: the generic
    destination / output-mix / two-pass crossfade branches of the four-step path."""
    L, P = 32768, 2
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S24_4LE", L)
    filters = [Filter([0, 1], [0], in_scales=[0.6, -0.3], coeff=0, crossfade=True),
               Filter([1], [0, 1], out_scales=[0.5, 1.0], coeff=1, crossfade=True)]
    g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, P])
    taps = configs.synthetic_filters(g, 19)
    sig = configs.synthetic_signal(g, 19, 7, sigma=0.01)
    y, r = run_pair(g, taps, sig, script={3: [(0, 2)], 5: [(1, 0), (0, 1)]})
    assert np.abs(r).max() > 1e4 and np.abs(y - r).max() <= 1
