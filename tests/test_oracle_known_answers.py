"""Analytic pins of the oracle: FFTW's r2r definition, the layout box, dirac identity,
`coeff: -1` == "dirac pulse", float64 direct convolution, the shipped xtc coefficients."""
import os

import numpy as np
import pytest

from brutefir_b200 import configs
from brutefir_b200.formats import interleaved_layout, pack_block
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import blocked_to_complex, hc_to_complex, unpack_run

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("L", [4, 32, 2048])
def test_fft_matches_fftw_definition(oracle_libs, L, rs):
    cv = po.Convolver("oracle", L, rs)
    rng = np.random.default_rng(L)
    x = rng.standard_normal(cv.N).astype(cv.dtype)
    X = np.fft.rfft(x.astype(np.float64))
    tol = (1e-5 if rs == 4 else 1e-13) * np.sqrt(cv.N)
    assert np.abs(hc_to_complex(cv.time2freq(x).astype(np.float64)) - X).max() < tol
    y = cv.freq2time(cv.time2freq(x))       # unnormalised both ways: N * x
    assert np.abs(y / cv.N - x).max() < tol


def test_layout_box_unit_tap(oracle_libs):
    """SURVEY.md preamble: a unit tap at index 1 gives R[0..3] = Re X_0..3, R[4] = Re X_{N/2} = -1, R[5..7] = Im X_1..3."""
    L = 8
    cv = po.Convolver("oracle", L, 8)
    taps = np.zeros(L)
    taps[1] = 1.0
    R = cv.coeffs2cbuf(taps)
    N = 2 * L
    frame = np.zeros(N)
    frame[L + 1] = 1.0
    X = np.fft.rfft(frame) / N
    assert np.allclose(R[0:4], X[0:4].real, atol=1e-15)
    assert np.isclose(R[4], X[L].real) and np.isclose(R[4] * N, -1.0)
    assert np.allclose(R[5:8], X[1:4].imag, atol=1e-15)
    assert np.allclose(blocked_to_complex(R), X, atol=1e-15)


@pytest.mark.parametrize("rs", [4, 8])
def test_input_then_output_mixmode_is_identity(oracle_libs, rs):
    cv = po.Convolver("oracle", 64, rs)
    x = np.random.default_rng(1).standard_normal(cv.N).astype(cv.dtype)
    assert np.array_equal(cv.mixnscale([cv.mixnscale([x], [1.0], cv.MIXMODE_INPUT)], [1.0], cv.MIXMODE_OUTPUT), x)


@pytest.mark.parametrize("kind", ["oracle", "ref"])
@pytest.mark.parametrize("fmt", ["S24_4LE", "S16_LE", "S32_LE"])
def test_dirac_pulse_identity_and_coeff_minus_one(oracle_libs, kind, fmt):
    """"dirac pulse" coefficients (bfconf.c:1905-1913) reproduce the integer input exactly, and the
    coeff:-1 short cut (bfrun.c:1779-1837) equals them."""
    if oracle_libs[kind] is None:
        pytest.skip("reference build unavailable")
    L, P = 64, 4
    inb, nin = interleaved_layout(2, fmt, L)
    outb, nout = interleaved_layout(2, fmt, L)
    g = FilterGraph(L, P, 8, inb, outb, nin, nout, [Filter([0], [0], coeff=0), Filter([1], [1], coeff=-1)], [P])
    d = po.BlockDriver(kind, g)
    pulse = np.zeros(L * P)
    pulse[0] = 1.0
    d.coeff_from_taps(0, pulse)
    rng = np.random.default_rng(3)
    bits = inb[0].sf.bits
    x = rng.integers(-(1 << (bits - 2)), 1 << (bits - 2), (6, 2, L)).astype(np.float64)
    sig = np.stack([pack_block(x[b], inb, nin) for b in range(6)])
    out = d.run(sig)
    d.close()
    assert np.array_equal(out[:, :nin], sig[:, :nin])
    y = unpack_run(out, outb, L)
    assert np.array_equal(y[0], x[:, 0].reshape(-1)) and np.array_equal(y[1], x[:, 1].reshape(-1))


@pytest.mark.parametrize("rs,tol_lsb", [(4, 1.5), (8, 0.5001)])
def test_direct_convolution_float64(oracle_libs, rs, tol_lsb):
    # distance to the UNROUNDED float64 truth: 0.5 LSB of quantisation plus, at float_bits 32, the float32
    # arithmetic error of the reference algorithm itself (peaks ~3.5e6 LSB carry an ulp of 0.25)
    g = configs.diagonal_graph(2, 128, 8, rs, "S24_4LE")
    taps = configs.synthetic_filters(g, 7)
    sig = configs.synthetic_signal(g, 7, 20)
    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    y = unpack_run(d.run(sig), g.out_formats, 128)
    d.close()
    x = unpack_run(sig, g.in_formats, 128)
    for c in range(2):
        truth = np.convolve(x[c], taps[c].astype(np.float64))[: x.shape[1]]
        assert np.abs(y[c] - truth).max() <= tol_lsb


def test_xtc_topology_with_shipped_coefficients(oracle_libs):
    """xtc_config (xtc_config:28-50) with directpath.txt / crosspath.txt against float64 direct convolution."""
    t = np.load(os.path.join(HERE, "golden", "xtc_taps.npz"))
    L, P = 64, 64
    g = configs.config_c5(realsize=8, L=L, P=P)
    for f in g.filters:
        f.crossfade = False
    d = po.BlockDriver("oracle", g)
    d.coeff_from_taps(0, t["directpath"])
    d.coeff_from_taps(1, t["crosspath"])
    sig = configs.synthetic_signal(g, 5, 80)
    y = unpack_run(d.run(sig), g.out_formats, L)
    d.close()
    x = unpack_run(sig, g.in_formats, L)
    n = x.shape[1]
    left = np.convolve(x[0], t["directpath"])[:n] + np.convolve(x[1], t["crosspath"])[:n]
    right = np.convolve(x[1], t["directpath"])[:n] + np.convolve(x[0], t["crosspath"])[:n]
    assert np.abs(y[0] - left).max() <= 0.5001 and np.abs(y[1] - right).max() <= 0.5001


def test_crossfade_ramps_old_to_new(oracle_libs):
    """SURVEY.md preamble: 1000 -> 250 crossfade is y[0] = 1000, y[L-1] = 250, linear in between."""
    L = 64
    cv = po.Convolver("oracle", L, 4)

    def const_spectrum(v):
        x = np.full(cv.N, v, np.float32)
        return cv.mixnscale([cv.time2freq(x)], [1.0 / cv.N], cv.MIXMODE_INPUT)

    res = cv.crossfade_inplace(const_spectrum(250.0), const_spectrum(1000.0))
    y = cv.freq2time(cv.mixnscale([res], [1.0], cv.MIXMODE_OUTPUT))
    n = np.arange(L)
    assert np.allclose(y[:L], 1000.0 + (250.0 - 1000.0) * n / (L - 1), atol=2e-3)
    assert np.allclose(y[L:], 250.0, atol=2e-3)
