"""Parity AT the BASELINE.json shapes, every partition live, against the reference's own convolver.

The checker is oracle/_ref (the reference's fftw_convolver.c / convolver_xmm.c compiled unmodified, driven by the
filter_process() replay of oracle/bf_blockdriver.c on all host threads); where that build is absent (it travels to the
GPU box prebuilt) the plain-C restatement, which tests/test_oracle_vs_ref.py pins to it bit for bit, stands in.

  c3  64 ch x 1 048 576 taps, 8192 x 128, S24_4LE: >= 136 blocks (the whole delay line holds data for the last 8),
      random unit-energy filters on ALL 64 channels; engine block by block (B = 1, the reference's schedule) AND 8 blocks
      per call (B = 8, bench.py's headline mode) against the reference.
  c4  32 ch x 262 144 taps, 256 x 1024 with the engine's AUTOMATIC partition split (a different summation tree than the
      reference's left-to-right sum): >= 1032 blocks.
  c5  xtc_config topology (/root/reference/xtc_config:28-50) at L 64 x P 64 with the shipped directpath / crosspath taps
      (tests/golden/xtc_taps.npz) and crossfaded coefficient swaps (bench5_config's cfc mechanism), one GPU.

Criteria (north_star; the same rules as tests/test_gpu_engine.py::assert_parity):
  sigma = 0.01 (-40 dBFS): |gpu - reference| <= 1 LSB at 24 bit on every sample, strictly.
  sigma = 0.1  (-20 dBFS): the output peaks near 2^22 LSB where a float32 ulp is 0.25-0.5 LSB and the reference itself is
      > 1 LSB from the exact result: |diff| <= 2 LSB, > 1 LSB on < 0.01 % of the samples, and the
      GPU no further from the float64 truth than the reference is (rms and max) -- a stated deviation from "1 LSB".
"""
import os

import numpy as np
import pytest

from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def checker_kind():
    return "ref" if po.available("ref") else "oracle"


def host_threads():
    return max(1, len(os.sched_getaffinity(0)))


def fast_unit_energy_filters(graph, seed, gain=1.0):
    """Gaussian taps x exp(-n / (taps/4)), unit energy; float32 generation keeps 64 x 1 Mi taps short."""
    rng = np.random.default_rng(seed)
    out = []
    for nb in graph.coeff_n_blocks:
        taps = nb * graph.filter_length
        env = np.exp(-np.arange(taps, dtype=np.float32) / (taps / 4.0))
        h = rng.standard_normal(taps, dtype=np.float32) * env
        h *= gain / np.sqrt(np.sum(h.astype(np.float64) ** 2))
        out.append(h)
    return out


def reference_run(graph, taps, sig, script=None):
    d = po.BlockDriver(checker_kind(), graph, n_threads=host_threads())
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    if script is None:
        out = d.run(sig)
    else:
        out = np.stack([_scripted(d, script, b, sig[b]) for b in range(sig.shape[0])])
    d.close()
    return out


def _scripted(obj, script, b, raw):
    for filt, coeff in script.get(b, ()):
        obj.set_control(filt, coeff)
    return obj.process_block(raw)


def engine_run(graph, taps, sig, B, mac_split=0, script=None):
    with Engine(graph, max_batch=B, mac_split=mac_split) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        if script is None:
            out = e.run(sig)
        else:
            out = np.stack([_scripted(e, script, b, sig[b]) for b in range(sig.shape[0])])
        info = e.info()
    return out, info


def truth_tail(graph, taps, sig, channels, n_tail):
    """float64 linear convolution of the chosen diagonal channels; the last n_tail samples."""
    L = graph.filter_length
    x = unpack_run(sig, graph.in_formats, L)
    out = {}
    for c in channels:
        h = taps[graph.filters[c].coeff].astype(np.float64)
        n = x.shape[1] + len(h) - 1
        nfft = 1 << int(np.ceil(np.log2(n)))
        y = np.fft.irfft(np.fft.rfft(x[graph.filters[c].inputs[0]], nfft) * np.fft.rfft(h, nfft), nfft)[: x.shape[1]]
        out[c] = y[-n_tail:]                       # integer input samples are LSB units already; unit-energy filters
    return out


def check(graph, got, ref, strict, truth=None, n_tail=None):
    L = graph.filter_length
    g, r = unpack_run(got, graph.out_formats, L), unpack_run(ref, graph.out_formats, L)
    diff = np.abs(g - r)
    assert np.abs(r).max() > 1e4                               # the comparison is not vacuous
    if strict:
        assert diff.max() <= 1, diff.max()
    else:
        assert diff.max() <= 2 and np.mean(diff > 1) < 1e-4, (diff.max(), np.mean(diff > 1))
    if truth is not None:
        for c, t in truth.items():
            eg, er = g[c, -n_tail:] - t, r[c, -n_tail:] - t
            assert np.sqrt(np.mean(eg ** 2)) <= 1.05 * np.sqrt(np.mean(er ** 2)), c
            assert np.abs(eg).max() <= np.abs(er).max() + 0.25, c
    return float(diff.max()), float(np.mean(diff > 0))


@pytest.mark.parametrize("sigma", [0.01, 0.1])
def test_c3_all_partitions_live_against_the_reference(gpu_lib, oracle_libs, sigma):
    g = configs.config_c3()
    n_blocks = 136                                              # P + 8: the last 8 blocks read 128 live partitions
    taps = fast_unit_energy_filters(g, 2003)
    sig = configs.synthetic_signal(g, 3, n_blocks, sigma=sigma)
    ref = reference_run(g, taps, sig)
    truth = None
    n_tail = 4 * g.filter_length
    if sigma > 0.05:
        truth = truth_tail(g, taps, sig, [0, 21, 42, 63], n_tail)
    for B in (1, 8):
        got, info = engine_run(g, taps, sig, B)
        assert info.mac_split == 1                              # the reference's summation order
        check(g, got, ref, strict=sigma < 0.05, truth=truth, n_tail=n_tail)
        if B == 1:
            first = got
        else:
            assert np.array_equal(first, got)                   # batched == block by block, byte for byte


@pytest.mark.parametrize("sigma", [0.01, 0.03, 0.1])
def test_c4_automatic_partition_split_against_the_reference(gpu_lib, oracle_libs, sigma):
    """32 filters x 256 bins cannot fill 148 SMs, so the engine splits the 1024-deep partition sum (37 ways block by
    block, 5 ways at 8 blocks per call): a different float32 summation tree than the reference's left-to-right one.
    Up to -30 dBFS that is invisible (<= 1 LSB, strictly).  At -20 dBFS -- outputs near 2^22 LSB, 1024 float32 terms per
    bin -- the reference's own rounding noise is several LSB and ANY other order lands elsewhere: measured
    (profiles/r2_diag_c4.txt) max 4 LSB, 1.2 % of the samples beyond 1 LSB, < 1e-4 beyond 2 LSB.  What is required
    there: <= 4 LSB, < 2 % beyond 1, < 0.05 % beyond 2, and the GPU no further from the float64 truth than the reference
    (rms and max).  mac_split = 1 keeps the reference's order (bit-exact MAC stage) and meets the same rule as the
    headline shape (<= 2 LSB, < 0.05 % beyond 1 at this depth); it is the documented way to ask for it."""
    g = configs.config_c4()
    n_blocks = 1032                                             # P + 8
    taps = fast_unit_energy_filters(g, 2004)
    sig = configs.synthetic_signal(g, 4, n_blocks, sigma=sigma)
    ref = reference_run(g, taps, sig)
    L = g.filter_length
    r = unpack_run(ref, g.out_formats, L)
    assert np.abs(r).max() > 1e4
    n_tail = 64 * L
    truth = truth_tail(g, taps, sig, [0, 13, 31], n_tail) if sigma > 0.05 else None

    def near_truth(y):
        for c, t in truth.items():
            eg, er = y[c, -n_tail:] - t, r[c, -n_tail:] - t
            assert np.sqrt(np.mean(eg ** 2)) <= 1.05 * np.sqrt(np.mean(er ** 2)), c
            assert np.abs(eg).max() <= np.abs(er).max() + 0.25, c

    for B in (1, 8):
        got, info = engine_run(g, taps, sig, B)
        assert info.mac_split > 1                               # the sum is split
        d = np.abs(unpack_run(got, g.out_formats, L) - r)
        if sigma < 0.05:
            assert d.max() <= 1, d.max()
        else:
            assert d.max() <= 4 and np.mean(d > 1) < 0.02 and np.mean(d > 2) < 5e-4, (d.max(), np.mean(d > 1), np.mean(d > 2))
            near_truth(unpack_run(got, g.out_formats, L))
    # the split forced off: the reference's summation order
    got, info = engine_run(g, taps, sig, 1, mac_split=1)
    assert info.mac_split == 1
    d = np.abs(unpack_run(got, g.out_formats, L) - r)
    if sigma < 0.05:
        assert d.max() <= 1, d.max()
    else:
        assert d.max() <= 2 and np.mean(d > 1) < 5e-4, (d.max(), np.mean(d > 1))
        near_truth(unpack_run(got, g.out_formats, L))


def xtc_setup(L=64, P=64):
    """config 5: coefficient sets 0 / 1 = the shipped direct / cross path (first L * P taps of the 4096)."""
    g = configs.config_c5(L=L, P=P)
    t = np.load(os.path.join(HERE, "golden", "xtc_taps.npz"))
    n = L * P
    taps = [t["directpath"][:n].astype(np.float32), t["crosspath"][:n].astype(np.float32)]
    # bench5_config:5-9 style script: every 16 blocks swap which set the direct / cross filters use (all four filters
    # crossfade), the first swap while the delay line is still filling, later ones with all 64 partitions live
    script = {}
    state = 0
    for b in range(10, 200, 16):
        state ^= 1
        script[b] = [(f, (f % 2) ^ state) for f in range(4)]
    return g, taps, script


def test_c5_xtc_shipped_taps_with_crossfaded_swaps(gpu_lib, oracle_libs):
    g, taps, script = xtc_setup()
    sig = configs.synthetic_signal(g, 5, 200, sigma=0.05)
    ref = reference_run(g, taps, sig, script)
    got, info = engine_run(g, taps, sig, 1, script=script)
    L = g.filter_length
    y, r = unpack_run(got, g.out_formats, L), unpack_run(ref, g.out_formats, L)
    assert np.abs(r).max() > 1e4
    assert np.abs(y - r).max() <= 1
    # batched calls split the control changes and crossfade blocks off (bfcuda_process_blocks): same bytes
    with Engine(g, max_batch=8) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        out = np.zeros((200, g.out_bytes), np.uint8)
        b = 0
        while b < 200:
            nxt = min([k for k in script if k > b] + [200])
            for filt, coeff in script.get(b, ()):
                e.set_control(filt, coeff)
            nb = min(8, nxt - b)
            e.process_blocks_async(sig[b:b + nb], out[b:b + nb], nb)
            b += nb
        e.synchronize()
    assert np.array_equal(out, got)
