"""The oracle restatement (oracle/bf_oracle.c) against the reference's own sources compiled here
(oracle/_ref/libbfref.so): every restated function must agree BIT FOR BIT on random inputs.
This is what pins the oracle; tests/test_golden.py pins it again against committed vectors so the check
survives where /root/reference (and hence a rebuild of _ref) is absent."""
import numpy as np
import pytest

from brutefir_b200 import _abi
from brutefir_b200.formats import BufferFormat, interleaved_layout, pack_block, parse_sample_format
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po

ALL_FORMATS = ["S8", "S16_LE", "S16_BE", "S24_LE", "S24_BE", "S24_4LE", "S24_4BE", "S32_LE", "S32_BE",
               "FLOAT_LE", "FLOAT_BE", "FLOAT64_LE", "FLOAT64_BE"]


@pytest.fixture(scope="module")
def both(oracle_libs):
    if oracle_libs["ref"] is None:
        pytest.skip("oracle/_ref/libbfref.so not built (needs /root/reference)")
    return oracle_libs


def pair(L, rs):
    return po.Convolver("oracle", L, rs), po.Convolver("ref", L, rs)


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("L", [4, 8, 64, 1024])
def test_frequency_domain_functions_bit_exact(both, L, rs):
    o, r = pair(L, rs)
    rng = np.random.default_rng(L * rs)
    bufs = [rng.standard_normal(o.N).astype(o.dtype) for _ in range(6)]
    for n in (1, 2, 3, 4, 6):
        scales = list(rng.standard_normal(n))
        for mode in (o.MIXMODE_INPUT, o.MIXMODE_OUTPUT):
            assert np.array_equal(o.mixnscale(bufs[:n], scales, mode), r.mixnscale(bufs[:n], scales, mode))
    a, b, c = bufs[:3]
    assert np.array_equal(o.convolve(a, b), r.convolve(a, b))
    assert np.array_equal(o.convolve_add(a, b, c.copy()), r.convolve_add(a, b, c.copy()))
    assert np.array_equal(o.convolve_inplace(a.copy(), b), r.convolve_inplace(a.copy(), b))
    assert np.array_equal(o.dirac_convolve(a), r.dirac_convolve(a))
    assert np.array_equal(o.time2freq(a), r.time2freq(a))          # same FFT shim behind both
    assert np.array_equal(o.freq2time(a), r.freq2time(a))
    taps = rng.standard_normal(L - 1).astype(o.dtype)
    assert np.array_equal(o.coeffs2cbuf(taps, 0.3), r.coeffs2cbuf(taps, 0.3))
    assert np.array_equal(o.runtime_coeffs2cbuf(bufs[3][:L]), r.runtime_coeffs2cbuf(bufs[3][:L]))
    so, sr = np.zeros(3 * L, o.dtype), np.zeros(3 * L, o.dtype)
    for k in range(3):
        assert np.array_equal(o.convolve_eval(bufs[k], so), r.convolve_eval(bufs[k], sr))
    bad = taps.copy()
    bad[1] = np.nan
    assert o.coeffs2cbuf(bad) is None and r.coeffs2cbuf(bad) is None


@pytest.mark.parametrize("L", [8, 256])
def test_crossfade_float_bit_exact(both, L):
    # float_bits 32 only: the reference's double branch reads past its buffer
    o, r = pair(L, 4)
    rng = np.random.default_rng(L)
    new, old = rng.standard_normal(o.N).astype(np.float32), rng.standard_normal(o.N).astype(np.float32)
    assert np.array_equal(o.crossfade_inplace(new.copy(), old.copy()), r.crossfade_inplace(new.copy(), old.copy()))


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_sample_conversion_bit_exact(both, fmt, rs):
    L = 32
    o, r = pair(L, rs)
    sf = parse_sample_format(fmt)
    bfs, n_bytes = interleaved_layout(3, sf, L)
    rng = np.random.default_rng(len(fmt) + rs)
    if sf.isfloat:
        vals = rng.standard_normal((3, L)) * 2.0
    else:
        vals = rng.integers(-(1 << (sf.bits - 1)), 1 << (sf.bits - 1), (3, L)).astype(np.float64)
    raw = pack_block(vals, bfs, n_bytes)
    for ch in range(3):
        co, no, cr, nr = o.new(), o.new(), r.new(), r.new()
        o.raw2cbuf(raw, co, no, bfs[ch])
        r.raw2cbuf(raw, cr, nr, bfs[ch])
        assert np.array_equal(co, cr) and np.array_equal(no, nr)
        if not sf.isfloat:
            assert np.array_equal(no[:L], vals[ch].astype(o.dtype))     # integers are NOT scaled
    # real2raw: values around the clip points, halves, negatives
    fs = sf.overflow_max
    x = np.concatenate([rng.standard_normal(L - 12) * fs * 0.7,
                        [-0.5, -1.0, -1.5, -2.5, 0.5, 1.5, 3.8, -3.2, fs + 0.4, fs + 0.6, -fs - 1.4, -fs - 1.6]]).astype(o.dtype)
    cb = np.concatenate([x, np.zeros(L, o.dtype)])
    for ch in (0, 2):
        ro, rr = np.full(n_bytes, 0xAA, np.uint8), np.full(n_bytes, 0xAA, np.uint8)
        ofo, ofr = _abi.OverflowC(3, 7, 1.5, fs), _abi.OverflowC(3, 7, 1.5, fs)
        o.cbuf2raw(cb, ro, bfs[ch], ofo)
        r.cbuf2raw(cb, rr, bfs[ch], ofr)
        assert np.array_equal(ro, rr)
        assert (ofo.n_overflows, ofo.intlargest, ofo.largest) == (ofr.n_overflows, ofr.intlargest, ofr.largest)


def _graph(rs, fmt_in="S24_4LE", fmt_out="S24_LE", L=32, P=6):
    inb, nin = interleaved_layout(3, fmt_in, L)
    outb, nout = interleaved_layout(3, fmt_out, L)
    filters = [Filter([0], [0], coeff=0), Filter([1, 2], [1], in_scales=[0.7, -0.2], coeff=1, delayblocks=2),
               Filter([2], [1, 2], out_scales=[0.5, 2.0], coeff=-1), Filter([0], [2], coeff=2, crossfade=True)]
    return FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, 3, 1])


@pytest.mark.parametrize("rs", [4, 8])
def test_block_driver_bit_exact_and_thread_invariant(both, rs):
    g = _graph(rs)
    rng = np.random.default_rng(99)
    dt = np.float32 if rs == 4 else np.float64
    taps = [rng.standard_normal(g.filter_length * nb).astype(dt) / 6 for nb in g.coeff_n_blocks]
    x = np.round(rng.standard_normal((14, 3, g.filter_length)) * 0.1 * (1 << 23))
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(14)])
    outs = {}
    for kind, nt in (("oracle", 1), ("ref", 1), ("ref", 3), ("oracle", 2)):
        d = po.BlockDriver(kind, g, n_threads=nt)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        res = []
        for b in range(14):
            if b == 6 and rs == 4:      # crossfading coefficient change on filter 3 (float only)
                d.set_control(3, 0)
            if b == 9:
                d.set_control(1, 1, delayblocks=0, in_scales=[0.1, 0.3])
            res.append(d.process_block(sig[b]))
        outs[(kind, nt)] = (np.stack(res), [tuple([d.overflow(o).n_overflows, d.overflow(o).intlargest]) for o in range(3)])
        d.close()
    base = outs[("ref", 1)]
    for k, v in outs.items():
        assert np.array_equal(v[0], base[0]), k
        assert v[1] == base[1], k


def test_filter_chain_eval_bit_exact(both):
    """to_filters / convolver_convolve_eval (bench1_config topology) -- oracle only, a "next" row on the GPU."""
    from brutefir_b200 import configs
    g = configs.config_c1_chained()
    g.filter_length, g.n_blocks, g.coeff_n_blocks = 64, 3, [3] * 6
    g.in_formats, g.in_bytes = interleaved_layout(2, "S24_4LE", 64)
    g.out_formats, g.out_bytes = interleaved_layout(2, "S24_4LE", 64)
    rng = np.random.default_rng(5)
    taps = [rng.standard_normal(64 * 3).astype(np.float32) / 8 for _ in range(6)]
    x = np.round(rng.standard_normal((8, 2, 64)) * 0.05 * (1 << 23))
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(8)])
    res = []
    for kind in ("oracle", "ref"):
        d = po.BlockDriver(kind, g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        res.append(d.run(sig))
        d.close()
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("rs,fmt", [(4, "S16_LE"), (8, "S16_LE"), (8, "S24_LE"), (4, "S8")])
def test_dither_hp_tpdf_bit_exact(both, rs, fmt):
    """HP-TPDF dither with error feedback (dither.c:37-139, dither.h:28-38, dither_funs.h:7-68) through whole block
    sequences, long enough for the random-table pointer to wrap; eligibility rules of bfconf.c:3173-3217 (output 2 is
    not dithered although requested when it is wider than 16 bit at float_bits 32 -- here it is a float channel)."""
    L, P, nb = 64, 2, 40
    sf_i, sf_f = parse_sample_format(fmt), parse_sample_format("FLOAT_LE")
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb = [BufferFormat(sf_i, 1, 0), BufferFormat(sf_i, 1, L * sf_i.bytes), BufferFormat(sf_f, 1, 2 * L * sf_i.bytes)]
    nout = 2 * L * sf_i.bytes + 4 * L + 32
    filters = [Filter([0], [0], coeff=0), Filter([1], [1], coeff=1), Filter([0, 1], [2], in_scales=[0.5, 0.5], coeff=0)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, P], sampling_rate=100,
                    apply_dither=[True, True, True])
    rng = np.random.default_rng(77)
    taps = [rng.standard_normal(L * P).astype(np.float32) / 6 for _ in range(2)]
    x = np.round(rng.standard_normal((nb, 2, L)) * 0.1 * (1 << 23))
    x[5:8] *= 40                                            # a few blocks that clip: overflow bookkeeping with dither
    x = np.clip(x, -(1 << 23), (1 << 23) - 1)
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(nb)])
    res = []
    for kind in ("oracle", "ref"):
        d = po.BlockDriver(kind, g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        out = d.run(sig)
        res.append((out, [(d.overflow(o).n_overflows, d.overflow(o).intlargest, d.overflow(o).largest) for o in range(3)]))
        d.close()
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1] and res[0][1][0][0] > 0
    # and dither really is on: the same run without it gives different samples on the integer outputs
    g.apply_dither = None
    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    plain = d.run(sig)
    d.close()
    assert not np.array_equal(plain[:, :2 * L * sf_i.bytes], res[0][0][:, :2 * L * sf_i.bytes])
    assert np.array_equal(plain[:, 2 * L * sf_i.bytes:], res[0][0][:, 2 * L * sf_i.bytes:])


@pytest.mark.parametrize("rs", [4, 8])
def test_td_convolver_bit_exact_and_is_a_linear_convolution(both, rs):
    """convolver_td_* (fftw_convolver.c:682-782, the sub-sample delay's convolver): the restatement against the
    reference build bit for bit, and both against the definition -- with the coefficients parked in the second
    half of the frame, the first blocklen outputs of a [previous | current] block are the linear convolution."""
    o, r = pair(64, rs)
    rng = np.random.default_rng(77 + rs)
    assert [o.td_block_length(n) for n in (0, 1, 2, 3, 199, 256, 257)] == [-1, 1, 2, 4, 256, 256, 512]
    # one coefficient is undefined in the reference (log2_roof(1) = -1, then 1 << -1; log2.h:28-43): not compared
    assert [r.td_block_length(n) for n in (0, 2, 3, 199, 256, 257)] == [-1, 2, 4, 256, 256, 512]
    for n_coeffs in (2, 3, 5, 31, 199, 256):
        B = o.td_block_length(n_coeffs)
        taps = rng.standard_normal(n_coeffs).astype(o.dtype)
        to, tr = o.td_new(taps), r.td_new(taps)
        assert to and tr
        x = rng.standard_normal(5 * B).astype(o.dtype)
        want = np.convolve(x.astype(np.float64), taps.astype(np.float64))
        for k in range(1, 5):
            blk = x[(k - 1) * B:(k + 1) * B]
            yo, yr = o.td_convolve(to, blk), r.td_convolve(tr, blk)
            assert np.array_equal(yo, yr), (n_coeffs, k)
            tol = (2e-5 if rs == 4 else 1e-12) * max(1.0, np.abs(want).max())
            assert np.abs(yo[:B] - want[k * B:(k + 1) * B]).max() <= tol, (n_coeffs, k)
