"""GPU parity, one convolver.h call at a time (include/bfcuda_convolver.h == /root/reference/convolver.h):
the CUDA kernels against the CPU oracle on the same buffers, in the reference's own layouts.

Bit-exact: everything without an FFT in it (raw2cbuf, mixnscale, convolve, convolve_add, convolve_inplace,
dirac, cbuf2raw incl. overflow accounting).  FFT-bearing calls (time2freq, freq2time, crossfade, coeffs2cbuf,
convolve_eval) are compared within a few ulp * sqrt(log2 N) of the data's magnitude: FFTW's own rounding is
not reproducible either, only the transform's definition is."""
import numpy as np
import pytest

from brutefir_b200 import _abi
from brutefir_b200 import convolver as cv
from brutefir_b200.formats import BufferFormat, interleaved_layout, pack_block, parse_sample_format
from oracle import pyoracle as po
from helpers import ulp_tol

pytestmark = pytest.mark.gpu

ALL_FORMATS = ["S8", "S16_LE", "S16_BE", "S24_LE", "S24_BE", "S24_4LE", "S24_4BE", "S32_LE", "S32_BE",
               "FLOAT_LE", "FLOAT_BE", "FLOAT64_LE", "FLOAT64_BE"]


def init(L, rs):
    assert cv.convolver_init(".fftw3wisdom", L, rs)
    assert cv.convolver_cbufsize() == 2 * L * rs
    return po.Convolver("oracle", L, rs)


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("L", [4, 64, 4096])
def test_elementwise_calls_bit_exact(gpu_lib, oracle_libs, L, rs):
    o = init(L, rs)
    rng = np.random.default_rng(L + rs)
    bufs = [rng.standard_normal(o.N).astype(o.dtype) for _ in range(6)]
    for n in (1, 2, 3, 5):
        scales = list(rng.standard_normal(n) * 3)
        for mode in (cv.MIXMODE_INPUT, cv.MIXMODE_OUTPUT):
            out = cv.new_cbuf()
            cv.convolver_mixnscale(bufs[:n], out, scales, mode)
            assert np.array_equal(out, o.mixnscale(bufs[:n], scales, mode)), (n, mode)
    a, b, c = bufs[:3]
    out = cv.new_cbuf()
    cv.convolver_convolve(a, b, out)
    assert np.array_equal(out, o.convolve(a, b))
    acc_g, acc_o = c.copy(), c.copy()
    for k in range(4):      # a short partition sum, like bfrun.c:1745-1754
        cv.convolver_convolve_add(bufs[k], bufs[k + 1], acc_g)
        o.convolve_add(bufs[k], bufs[k + 1], acc_o)
    assert np.array_equal(acc_g, acc_o)
    ip = a.copy()
    cv.convolver_convolve_inplace(ip, b)
    assert np.array_equal(ip, o.convolve_inplace(a.copy(), b))
    cv.convolver_dirac_convolve(a, out)
    assert np.array_equal(out, o.dirac_convolve(a))
    ip = a.copy()
    cv.convolver_dirac_convolve_inplace(ip)
    assert np.array_equal(ip, o.dirac_convolve(a))
    assert cv.convolver_verify_cbuf([a, b]) and not cv.convolver_verify_cbuf([a, np.full(o.N, np.inf, o.dtype)])


def test_invalid_mixmode_calls_bf_exit(gpu_lib, oracle_libs):
    init(16, 4)
    out = cv.new_cbuf()
    cv.exit_status()
    cv.convolver_mixnscale([cv.new_cbuf()], out, [1.0], cv.MIXMODE_INPUT_ADD)   # fftw_convfuns.h:496-499
    assert cv.exit_status() == 1                                                 # BF_EXIT_OTHER


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_sample_conversion_bit_exact(gpu_lib, oracle_libs, fmt, rs):
    L = 64
    o = init(L, rs)
    sf = parse_sample_format(fmt)
    bfs, n_bytes = interleaved_layout(3, sf, L)
    rng = np.random.default_rng(len(fmt) * rs)
    vals = rng.standard_normal((3, L)) * 2 if sf.isfloat else \
        rng.integers(-(1 << (sf.bits - 1)), 1 << (sf.bits - 1), (3, L)).astype(np.float64)
    raw = pack_block(vals, bfs, n_bytes)
    for ch in range(3):
        cg, ng, co, no = cv.new_cbuf(), cv.new_cbuf(), o.new(), o.new()
        cv.convolver_raw2cbuf(raw, cg, ng, bfs[ch])
        o.raw2cbuf(raw, co, no, bfs[ch])
        assert np.array_equal(cg, co) and np.array_equal(ng, no)
    fs = sf.overflow_max
    edge = [-0.5, -1.0, -1.5, -2.5, 0.5, 1.5, 3.8, -3.2, fs + 0.4, fs + 0.6, -fs - 1.4, -fs - 1.6, fs * 3, -fs * 3]
    if sf.sbytes == 4 and not sf.isfloat:
        # a sample quantising to exactly INT32_MIN makes the reference negate INT32_MIN (dither_funs.h:93-95,
        # undefined behaviour); keep the 32-bit negative clip-edge cases clear of it
        edge[10] = -fs * 0.5
        edge[11] = -fs * 0.25
    x = np.concatenate([rng.standard_normal(L - len(edge)) * fs * 0.7, edge]).astype(o.dtype)
    cb = np.concatenate([x, np.zeros(L, o.dtype)])
    for ch in (0, 2):
        rg, ro = np.full(n_bytes, 0xAA, np.uint8), np.full(n_bytes, 0xAA, np.uint8)
        og, oo = _abi.OverflowC(3, 7, 1.5, fs), _abi.OverflowC(3, 7, 1.5, fs)
        cv.convolver_cbuf2raw(cb, rg, bfs[ch], False, og)
        o.cbuf2raw(cb, ro, bfs[ch], oo)
        assert np.array_equal(rg, ro)
        assert (og.n_overflows, og.intlargest, og.largest) == (oo.n_overflows, oo.intlargest, oo.largest)
        assert og.n_overflows > 3


def test_nonfinite_output_and_safety_limit(gpu_lib, oracle_libs):
    L = 32
    init(L, 4)
    bfs, n_bytes = interleaved_layout(1, "S16_LE", L)
    cb = np.zeros(2 * L, np.float32)
    cb[5] = np.nan
    cv.exit_status()
    cv.convolver_cbuf2raw(cb, np.zeros(n_bytes, np.uint8), bfs[0], False, _abi.OverflowC(0, 0, 0.0, 32767.0))
    assert cv.exit_status() == -5          # the reference abort()s here (real2raw.h:27-31)
    cb[5] = 30000.0
    cv.set_safety_limit(0.5)               # bfconf->safety_limit, real2raw.h:32-41
    cv.convolver_cbuf2raw(cb, np.zeros(n_bytes, np.uint8), bfs[0], False, _abi.OverflowC(0, 0, 0.0, 32767.0))
    assert cv.exit_status() == 1
    cv.set_safety_limit(0.0)
    cv.convolver_cbuf2raw(cb, np.zeros(n_bytes, np.uint8), bfs[0], False, _abi.OverflowC(0, 0, 0.0, 32767.0))
    assert cv.exit_status() is None


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("L", [4, 8, 16, 32, 256, 1024, 4096, 8192])
def test_transforms_within_fft_tolerance(gpu_lib, oracle_libs, L, rs):
    o = init(L, rs)
    rng = np.random.default_rng(L * 3 + rs)
    x = rng.standard_normal(o.N).astype(o.dtype)
    X = cv.new_cbuf()
    cv.convolver_time2freq(x, X)
    Xo = o.time2freq(x)
    assert np.abs(X - Xo).max() <= ulp_tol(o.dtype, o.N, np.sqrt(o.N) * 4)
    truth = np.fft.rfft(x.astype(np.float64))
    assert np.abs(X[: L + 1] - truth.real).max() <= ulp_tol(o.dtype, o.N, np.sqrt(o.N) * 4)
    y = cv.new_cbuf()
    cv.convolver_freq2time(Xo, y)
    assert np.abs(y - o.freq2time(Xo)).max() <= ulp_tol(o.dtype, o.N, o.N * 4.0)
    inplace = x.copy()
    cv.convolver_time2freq(inplace, inplace)     # fftw plans "inplace" (fftw_convolver.c:202-206)
    assert np.array_equal(inplace, X)
    taps = rng.standard_normal(L - 1).astype(o.dtype)
    H = cv.new_cbuf()
    assert cv.convolver_coeffs2cbuf(taps, 0.5, H) is H
    assert np.abs(H - o.coeffs2cbuf(taps, 0.5)).max() <= ulp_tol(o.dtype, o.N, np.sqrt(L) * 2 / o.N) * 4
    assert cv.convolver_coeffs2cbuf(np.array([1.0, np.nan]), 1.0, H) is None     # fftw_convolver.c:543-546
    Hr = cv.new_cbuf()
    cv.convolver_runtime_coeffs2cbuf(taps.tolist() + [0.25], Hr)
    assert np.abs(Hr - o.runtime_coeffs2cbuf(np.array(taps.tolist() + [0.25]))).max() <= \
        ulp_tol(o.dtype, o.N, np.sqrt(L) * 4 / o.N) * 4
    sg, so = np.zeros(3 * L, o.dtype), np.zeros(3 * L, o.dtype)
    for k in range(3):
        spec = o.time2freq(rng.standard_normal(o.N).astype(o.dtype)) / o.N
        og = cv.new_cbuf()
        cv.convolver_convolve_eval(spec.astype(o.dtype), sg, og)
        oo = o.convolve_eval(spec.astype(o.dtype), so)
        assert np.abs(og - oo).max() <= ulp_tol(o.dtype, o.N, np.sqrt(o.N) * 8)


@pytest.mark.parametrize("L", [8, 64, 2048])
def test_crossfade_within_fft_tolerance(gpu_lib, oracle_libs, L):
    o = init(L, 4)
    rng = np.random.default_rng(L)
    # realistic magnitudes: spectra of +-1 signals scaled by 1/N, blocked layout
    def spec():
        x = rng.standard_normal(o.N).astype(np.float32)
        return o.mixnscale([o.time2freq(x)], [1.0 / o.N], o.MIXMODE_INPUT)
    new, old = spec(), spec()
    g_new, g_old, scratch = new.copy(), old.copy(), cv.new_cbuf(2 * o.N)
    cv.convolver_crossfade_inplace(g_new, g_old, scratch)
    ref = o.crossfade_inplace(new.copy(), old.copy())
    assert np.abs(g_new - ref).max() <= ulp_tol(np.float32, o.N, 4.0 / np.sqrt(o.N)) * 4


def test_fftplan_reports_not_supported(gpu_lib):
    lib = gpu_lib
    assert not lib.convolver_fftplan(10, 0, 0)
    assert b"FFTW" in lib.bfcuda_convolver_last_error()


@pytest.mark.parametrize("rs", [4, 8])
def test_td_convolver_matches_oracle(gpu_lib, oracle_libs, rs):
    """convolver_td_* (convolver.h:134-146; fftw_convolver.c:682-782): the sub-sample delay's small convolver on the
    device against the oracle, block by block the way delay.c:415-442 slides its [previous | current] frames.
    Tolerance (north_star): 1e-6 of full scale in float32, 1e-12 in float64 -- full scale = the largest output."""
    o = init(64, rs)
    rng = np.random.default_rng(500 + rs)
    assert cv.convolver_td_block_length(5) == 8 and cv.convolver_td_block_length(0) == -1   # log2.h:28-43
    assert cv.convolver_td_block_length(199) == 256
    assert not cv.convolver_td_new(np.zeros(0, o.dtype))
    for n_coeffs in (1, 2, 3, 5, 31, 199, 256, 1000):
        B = cv.convolver_td_block_length(n_coeffs)
        # sinc-like taps of a fractional delay: unit DC gain, decaying tails (delay.c:444-506)
        k = np.arange(n_coeffs) - (n_coeffs - 1) / 2.0 - 0.37
        taps = (np.sinc(k) * np.hanning(n_coeffs + 2)[1:-1] if n_coeffs > 2 else rng.standard_normal(n_coeffs)).astype(o.dtype)
        tg, to = cv.convolver_td_new(taps), o.td_new(taps)
        assert tg and to
        x = rng.uniform(-1, 1, 6 * B).astype(o.dtype)
        for j in range(1, 6):
            blk = x[(j - 1) * B:(j + 1) * B].copy()      # convolved in place: never the stream itself
            want = o.td_convolve(to, blk)
            cv.convolver_td_convolve(tg, blk)
            fs = max(1.0, np.abs(want).max())
            assert np.abs(blk - want).max() <= (1e-6 if rs == 4 else 1e-12) * fs, (n_coeffs, j)
        cv.convolver_td_delete(tg)


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("fmt", ["S16_LE", "S24_4BE", "S8"])
def test_cbuf2raw_with_dither_bit_exact(gpu_lib, oracle_libs, rs, fmt):
    """convolver_cbuf2raw(apply_dither = true) (fftw_convolver.c:489-499; HP-TPDF dither with error feedback,
    dither_funs.h:7-68, dither.h:28-38): the device quantiser against the oracle on the same time-domain blocks, with
    the oracle's own dither table and starting state handed to the library the way a host hands over dither_randtab and
    its struct dither_state -- bytes, overflow record and the state after every block are identical, through a wrap of
    the table pointer and through clipping blocks."""
    L, n_ch, rate = 64, 2, 100
    o = init(L, rs)
    o.dither_init(n_ch, rate)
    table = o.dither_table()
    cv.set_dither_table(table)
    sf = parse_sample_format(fmt)
    bf = BufferFormat(sf, 2, sf.bytes)          # second channel of an interleaved pair
    rng = np.random.default_rng(900 + rs)
    index = 1
    ptr, sf0, sd0 = o.dither_state(index)
    st = _abi.DitherStateC()
    st.randtab_ptr = ptr
    st.sf[0], st.sf[1], st.sd[0], st.sd[1] = sf0[0], sf0[1], sd0[0], sd0[1]
    of_g, of_o = _abi.OverflowC(), _abi.OverflowC()
    of_g.max = of_o.max = float((1 << (8 * sf.sbytes - 1)) - 1)
    wraps = 0
    for blk in range(60):
        amp = (1 << (8 * sf.sbytes - 1)) * (3.0 if blk in (7, 8, 30) else 0.2)
        x = np.zeros(2 * L, o.dtype)
        x[:L] = rng.standard_normal(L) * amp
        raw_g = np.full(2 * L * sf.bytes, 0x5A, np.uint8)       # the other channel's bytes must survive
        raw_o = raw_g.copy()
        before = st.randtab_ptr
        cv.convolver_cbuf2raw_dither(x, raw_g, bf, st, of_g)
        o.cbuf2raw_dither(x, raw_o, bf, of_o, index)
        wraps += st.randtab_ptr < before
        assert np.array_equal(raw_g, raw_o), blk
        assert (of_g.n_overflows, of_g.intlargest, of_g.largest) == (of_o.n_overflows, of_o.intlargest, of_o.largest)
        ptr, sfo, sdo = o.dither_state(index)
        assert st.randtab_ptr == ptr
        assert (st.sf[0], st.sf[1]) == sfo if rs == 4 else (st.sd[0], st.sd[1]) == sdo
    assert wraps >= 1 and of_g.n_overflows > 0


@pytest.mark.parametrize("rs", [4, 8])
def test_td_and_dither_against_reference_vectors(gpu_lib, oracle_libs, rs):
    """The committed vectors the reference build itself produced (tests/golden/make_golden.py), no oracle in the
    comparison: convolver_td_* within the float tolerance, the dithered quantiser byte for byte (only the dither TABLE
    is taken from the oracle library, whose generator the same vectors pin in tests/test_golden.py)."""
    import os
    fn = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "functions.npz"))
    t = f"f{rs * 8}"
    o = init(32, rs)
    td = cv.convolver_td_new(fn[f"{t}_td_taps"])
    assert td
    x = fn[f"{t}_td_in"]
    for k in range(3):
        blk = x[k * 32:(k + 2) * 32].copy()
        cv.convolver_td_convolve(td, blk)
        want = fn[f"{t}_td_out"][k]
        assert np.abs(blk - want).max() <= (1e-6 if rs == 4 else 1e-12) * max(1.0, np.abs(want).max())
    cv.convolver_td_delete(td)
    o.dither_init(2, 100)
    table = o.dither_table()
    cv.set_dither_table(table)
    ptr, _, _ = o.dither_state(1)
    st = _abi.DitherStateC()
    st.randtab_ptr = ptr
    bf16 = BufferFormat(parse_sample_format("S16_LE"), 1, 0)
    of = _abi.OverflowC(0, 0, 0.0, 32767.0)
    for k in range(40):
        raw = np.zeros(32 * 2, np.uint8)
        cv.convolver_cbuf2raw_dither(np.concatenate([fn[f"{t}_dither_in"][k], np.zeros(32, o.dtype)]), raw, bf16, st, of)
        assert np.array_equal(raw, fn[f"{t}_dither_raw"][k]), k
    assert [of.n_overflows, of.intlargest, of.largest, of.max] == list(fn[f"{t}_dither_overflow"])
