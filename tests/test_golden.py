"""The oracle restatement against vectors the reference itself produced (tests/golden/make_golden.py).
Bit-exact everywhere: the restatement and the reference share the FFT shim."""
import os

import numpy as np
import pytest

from brutefir_b200 import _abi, configs
from brutefir_b200.formats import BufferFormat, interleaved_layout, parse_sample_format
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fn():
    return np.load(os.path.join(HERE, "golden", "functions.npz"))


@pytest.fixture(scope="module")
def blk():
    return np.load(os.path.join(HERE, "golden", "blocks.npz"))


@pytest.mark.parametrize("rs", [4, 8])
def test_functions_against_reference_vectors(oracle_libs, fn, rs):
    cv = po.Convolver("oracle", 32, rs)
    t = f"f{rs * 8}"
    assert np.array_equal(cv.time2freq(fn[f"{t}_time"]), fn[f"{t}_time2freq"])
    assert np.array_equal(cv.freq2time(fn[f"{t}_time2freq"]), fn[f"{t}_freq2time"])
    a, b, c = fn[f"{t}_mix_in"]
    sc = list(fn[f"{t}_mix_scales"])
    assert np.array_equal(cv.mixnscale([a, b, c], sc, cv.MIXMODE_INPUT), fn[f"{t}_mix_INPUT"])
    assert np.array_equal(cv.mixnscale([a, b, c], sc, cv.MIXMODE_OUTPUT), fn[f"{t}_mix_OUTPUT"])
    assert np.array_equal(cv.convolve(a, b), fn[f"{t}_convolve"])
    assert np.array_equal(cv.convolve_add(a, b, c.copy()), fn[f"{t}_convolve_add"])
    assert np.array_equal(cv.dirac_convolve(a), fn[f"{t}_dirac"])
    assert np.array_equal(cv.coeffs2cbuf(fn[f"{t}_taps"], 0.75), fn[f"{t}_coeffs2cbuf"])
    if rs == 4:
        assert np.array_equal(cv.crossfade_inplace(a.copy(), b.copy()), fn[f"{t}_crossfade"])
    bf = BufferFormat(parse_sample_format("S24_4LE"), 1, 0)
    raw = np.zeros(32 * 4, np.uint8)
    of = _abi.OverflowC(0, 0, 0.0, float((1 << 23) - 1))
    cv.cbuf2raw(np.concatenate([fn[f"{t}_quant_in"], np.zeros(32, cv.dtype)]), raw, bf, of)
    assert np.array_equal(raw, fn[f"{t}_quant_raw"])
    assert [of.n_overflows, of.intlargest, of.largest, of.max] == list(fn[f"{t}_quant_overflow"])
    # the corner cases spelled out in SURVEY.md: q(-0.5)=0, q(-1.0)=-1, q(-1.5)=-2, q(-2.5)=-3, q(0.5)=1,
    # q(1.5)=2, q(3.8)=4, q(-3.2)=-3; clip to [-2^23, 2^23-1]
    q = raw.view("<i4")[:14]
    assert list(q[:8]) == [0, -1, -2, -3, 1, 2, 4, -3]
    assert q[12] == (1 << 23) - 1 and q[13] == -(1 << 23)
    # the sub-sample delay's small convolver
    td = cv.td_new(fn[f"{t}_td_taps"])
    x = fn[f"{t}_td_in"]
    for k in range(3):
        assert np.array_equal(cv.td_convolve(td, x[k * 32:(k + 2) * 32]), fn[f"{t}_td_out"][k])
    # HP-TPDF dither with error feedback through a wrap of the table pointer and a clipping block
    cv.dither_init(2, 100)
    bf16 = BufferFormat(parse_sample_format("S16_LE"), 1, 0)
    of = _abi.OverflowC(0, 0, 0.0, 32767.0)
    for k in range(40):
        raw = np.zeros(32 * 2, np.uint8)
        cv.cbuf2raw_dither(np.concatenate([fn[f"{t}_dither_in"][k], np.zeros(32, cv.dtype)]), raw, bf16, of, 1)
        assert np.array_equal(raw, fn[f"{t}_dither_raw"][k]), k
    assert [of.n_overflows, of.intlargest, of.largest, of.max] == list(fn[f"{t}_dither_overflow"])
    assert of.n_overflows > 0


def golden_graph_a():
    L, P = 16, 4
    inb, nin = interleaved_layout(2, "S24_LE", L)
    outb, nout = interleaved_layout(2, "S16_BE", L)
    return FilterGraph(L, P, 4, inb, outb, nin, nout,
                       [Filter([0], [0], coeff=0, delayblocks=1),
                        Filter([1, 0], [1, 0], in_scales=[0.5, 0.25], out_scales=[1.0, 0.125], coeff=1)], [P, 2])


def golden_graph_b():
    g = configs.config_c5(L=16, P=8)
    g.coeff_n_blocks = [8, 8]
    return g


def golden_graph_c():
    """to_filters chaining: four input-fed filters, two fed by them, one fed by those (bfrun.c:1603-1660)."""
    L, P = 32, 3
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(3, "S24_4LE", L)
    filters = [Filter([0], [], coeff=2), Filter([0], [], coeff=3), Filter([1], [2], coeff=4), Filter([1], [], coeff=5),
               Filter([], [0], coeff=0, from_filters=[0, 3], fscales=[1.0, -0.5]),
               Filter([1], [1], in_scales=[0.25], coeff=1, from_filters=[1, 2], delayblocks=1),
               Filter([], [2], out_scales=[0.5], coeff=-1, from_filters=[4, 5], fscales=[0.5, 0.25])]
    return FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, 2, P, 1, P])


def golden_graph_d():
    """dither: true on two S16_LE outputs (dither.c, dither_funs.h:7-68); small table so that it wraps."""
    L, P = 64, 2
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S16_LE", L)
    return FilterGraph(L, P, 4, inb, outb, nin, nout, [Filter([0], [0], coeff=0), Filter([1], [1], coeff=1)], [P, P],
                       sampling_rate=100, apply_dither=[True, True])


def run_golden_b(engine_like, sig):
    outs = []
    for b in range(16):
        if b in (5, 10):
            swap = (b // 5) % 2
            for f in range(4):
                engine_like.set_control(f, (f % 2 + swap) % 2)
        if b == 13:
            engine_like.set_control(0, -1)
        outs.append(engine_like.process_block(sig[b]))
    return np.stack(outs)


def test_block_sequences_against_reference_vectors(oracle_libs, blk):
    d = po.BlockDriver("oracle", golden_graph_a())
    d.coeff_from_taps(0, blk["a_taps0"])
    d.coeff_from_taps(1, blk["a_taps1"])
    assert np.array_equal(d.run(blk["a_sig"]), blk["a_out"])
    d.close()
    d = po.BlockDriver("oracle", golden_graph_b())
    d.coeff_from_taps(0, blk["b_taps0"])
    d.coeff_from_taps(1, blk["b_taps1"])
    assert np.array_equal(run_golden_b(d, blk["b_sig"]), blk["b_out"])
    d.close()
    gc = golden_graph_c()
    d = po.BlockDriver("oracle", gc)
    for c in range(len(gc.coeff_n_blocks)):
        d.coeff_from_taps(c, blk[f"c_taps{c}"])
    assert np.array_equal(d.run(blk["c_sig"]), blk["c_out"])
    d.close()
    d = po.BlockDriver("oracle", golden_graph_d())
    d.coeff_from_taps(0, blk["d_taps0"])
    d.coeff_from_taps(1, blk["d_taps1"])
    assert np.array_equal(d.run(blk["d_sig"]), blk["d_out"])
    assert [[d.overflow(o).n_overflows, d.overflow(o).intlargest, d.overflow(o).largest] for o in range(2)] == \
        blk["d_overflow"].tolist()
    d.close()
