"""Virtual -> physical output mixing, mute and sub-sample delay on the engine path
(/root/reference/bfrun.c:1503-1526, 1918-2002; delay.c:415-442), against the oracle's replay of the same lines."""
import numpy as np
import pytest

from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu
IN, OUT = 0, 1


def sinc_taps(half, frac, beta=9.0, dtype=np.float32):
    """a Kaiser-windowed sinc delayed by half + frac samples (what firwindow.c's sample_sinc produces for delay.c)"""
    n = np.arange(2 * half + 1)
    x = n - half - frac
    return (np.sinc(x) * np.kaiser(2 * half + 1, beta)).astype(dtype)


def run_scripted(g, taps, sig, script, B=1):
    """script: block -> list of (method name, args) applied to both the engine and the oracle before that block"""
    with Engine(g, max_batch=B) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref = [], []
        b = 0
        n = sig.shape[0]
        while b < n:
            for name, args in script.get(b, ()):
                getattr(e, name)(*args)
                getattr(d, name)(*args)
            nxt = min([k for k in script if k > b] + [n])
            nb = min(B, nxt - b)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            e.process_blocks_async(np.ascontiguousarray(sig[b:b + nb]), out, nb)
            e.synchronize()
            got.extend(out)
            for k in range(nb):
                ref.append(d.process_block(sig[b + k]))
            b += nb
        of = [(e.overflow(o), d.overflow(o)) for o in range(len(g.out_formats))]
        d.close()
    L = g.filter_length
    return unpack_run(np.stack(got), g.out_formats, L), unpack_run(np.stack(ref), g.out_formats, L), of


@pytest.mark.parametrize("rs,B", [(4, 1), (4, 4), (8, 1)])
def test_several_virtual_outputs_on_one_physical_channel(gpu_lib, oracle_libs, rs, B):
    """outputs 0, 1, 2 share physical channel 0, output 3 has its own; a mute comes and goes"""
    L, P = 256, 4
    inb, nin = interleaved_layout(3, "S24_4LE", L)
    phys, nout = interleaved_layout(2, "S24_4LE", L)
    outb = [phys[0], phys[0], phys[0], phys[1]]
    filters = [Filter([0], [0], coeff=0), Filter([1], [1], coeff=1), Filter([2], [2], out_scales=[0.5], coeff=2),
               Filter([0, 1], [3], in_scales=[0.4, 0.4], coeff=1)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, P, P], out_physical=[0, 0, 0, 1])
    taps = configs.synthetic_filters(g, 37)
    sig = configs.synthetic_signal(g, 37, 16, sigma=0.02)
    script = {5: [("set_mute", (OUT, 1, True))], 9: [("set_mute", (OUT, 1, False)), ("set_mute", (IN, 2, True))],
              13: [("set_mute", (IN, 2, False))]}
    y, r, of = run_scripted(g, taps, sig, script, B)
    assert np.abs(r).max() > 1e4
    assert np.abs(y - r).max() <= (1 if rs == 4 else 0)
    for a, b in of:
        assert a.n_overflows == b.n_overflows
    assert (of[0][0].intlargest, of[1][0].intlargest) == (of[2][0].intlargest, of[2][0].intlargest)   # one shared record


@pytest.mark.parametrize("rs", [4, 8])
def test_subsample_delay_on_inputs_and_outputs(gpu_lib, oracle_libs, rs):
    """channel 0: input side, channel 1: output side, channel 2: none; the delay step changes at run time (cfid / cfod)"""
    L, P = 1024, 3
    g = configs.diagonal_graph(3, L, P, rs, "S24_4LE")
    taps = configs.synthetic_filters(g, 41)
    sig = configs.synthetic_signal(g, 41, 14, sigma=0.02)
    dt = np.float32 if rs == 4 else np.float64
    script = {0: [("set_subdelay", (IN, 0, sinc_taps(15, 0.25, dtype=dt))), ("set_subdelay", (OUT, 1, sinc_taps(15, -0.4, dtype=dt)))],
              6: [("set_subdelay", (IN, 0, sinc_taps(15, 0.7, dtype=dt)))],
              10: [("set_subdelay", (OUT, 1, None))]}
    y, r, _ = run_scripted(g, taps, sig, script)
    assert np.abs(r).max() > 1e4
    # the reference runs the filter through its FFT-based td convolver, the engine as a direct FIR: rounding-level
    # differences (31 taps), then one quantiser -- 1 LSB at float_bits 32, and at float_bits 64 a tie may flip
    assert np.abs(y - r).max() <= 1
    assert np.mean(y != r) < (0.05 if rs == 4 else 1e-4)
    # the delayed channels really are delayed: against the undelayed engine output they differ
    with Engine(g) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        plain = unpack_run(e.run(sig), g.out_formats, L)
    assert np.abs(plain[0] - y[0]).max() > 100 and np.abs(plain[1] - y[1]).max() > 100
    assert np.array_equal(plain[2], y[2])
