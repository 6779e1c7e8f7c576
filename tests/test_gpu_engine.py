"""GPU parity of the block-level engine (include/bfcuda.h) against the CPU oracle's replay of
filter_process() on the same seeded inputs.

Criteria (BASELINE.json north_star, dither off):
  * integer output: within 1 LSB of the reference at float_bits 32 (|diff| <= 1), identical at float_bits 64;
  * float32 output: max abs error <= 1e-6 of full scale; float64: <= 1e-12;
  * the multiply-accumulate stage alone: BIT-EXACT against the reference's convolve / convolve_add applied to
    the very same delay-line and coefficient buffers (split 1 keeps the reference's summation order)."""
import os

import numpy as np
import pytest

from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import BufferFormat, interleaved_layout, pack_block, parse_sample_format, planar_layout, unpack_block
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import unpack_run
from test_golden import golden_graph_a, golden_graph_b, golden_graph_c, golden_graph_d, run_golden_b

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_both(graph, taps, sig, mac_split=0, scale=1.0, flags=0):
    with Engine(graph, mac_split=mac_split, flags=flags) as e:
        d = po.BlockDriver("oracle", graph)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h, scale)
            d.coeff_from_taps(c, h, scale)
        got, ref = e.run(sig), d.run(sig)
        of = [(e.overflow(o), d.overflow(o)) for o in range(len(graph.out_formats))]
        d.close()
    return got, ref, of


def float64_truth(graph, taps, sig):
    """Exact (float64, unrounded) output of a diagonal graph, for judging float32 error levels."""
    L = graph.filter_length
    x = unpack_run(sig, graph.in_formats, L)
    out = []
    for f in graph.filters:
        h = taps[f.coeff].astype(np.float64)
        n = x.shape[1] + len(h) - 1
        nfft = 1 << int(np.ceil(np.log2(n)))
        out.append(np.fft.irfft(np.fft.rfft(x[f.inputs[0]], nfft) * np.fft.rfft(h, nfft), nfft)[: x.shape[1]])
    return np.stack(out)


def assert_parity(graph, got, ref, truth=None):
    """north_star tolerances.  `truth` switches the float_bits-32 integer criterion to its
    float32-physics form (see below)."""
    L = graph.filter_length
    g, r = unpack_run(got, graph.out_formats, L), unpack_run(ref, graph.out_formats, L)
    for c, bf in enumerate(graph.out_formats):
        diff = np.abs(g[c] - r[c])
        if bf.sf.isfloat:
            tol = 1e-6 if graph.realsize == 4 or bf.sf.bytes == 4 else 1e-12
            assert diff.max() <= tol, (c, diff.max())
        elif graph.realsize == 8:
            assert diff.max() == 0, (c, diff.max())
        elif truth is None:
            assert diff.max() <= 1, (c, diff.max())        # 1 LSB
        else:
            # At -20 dBFS with N = 16384 the output peaks near 2^22 LSB where one float32 ulp is 0.25-0.5 LSB:
            # the reference ITSELF is up to ~1.4 LSB from the exact result there, so two
            # float32 implementations can land 2 LSB apart on a handful of samples.  Require: never more than
            # 2 LSB, more than 1 LSB on < 0.01 % of the samples, and the GPU no further from the truth than
            # the reference is.
            assert diff.max() <= 2 and np.mean(diff > 1) < 1e-4, (c, diff.max(), np.mean(diff > 1))
            eg, er = g[c] - truth[c], r[c] - truth[c]
            assert np.sqrt(np.mean(eg ** 2)) <= 1.05 * np.sqrt(np.mean(er ** 2))
            assert np.abs(eg).max() <= np.abs(er).max() + 0.25


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("L,P", [(4, 3), (64, 16), (1024, 8), (8192, 4)])
@pytest.mark.parametrize("fmt", ["S24_4LE", "FLOAT_LE"])
def test_diagonal_graph(gpu_lib, oracle_libs, L, P, fmt, rs):
    g = configs.diagonal_graph(3, L, P, rs, fmt)
    taps = configs.synthetic_filters(g, 11)
    sig = configs.synthetic_signal(g, 11, P + 8)
    got, ref, of = run_both(g, taps, sig, mac_split=1)
    assert_parity(g, got, ref, truth=float64_truth(g, taps, sig) if (L >= 4096 and rs == 4 and fmt == "S24_4LE") else None)
    if L >= 4096 and rs == 4 and fmt == "S24_4LE":
        # the same shape at -40 dBFS: strictly within 1 LSB
        sig = configs.synthetic_signal(g, 11, P + 8, sigma=0.01)
        got, ref, _ = run_both(g, taps, sig, mac_split=1)
        assert_parity(g, got, ref)
    for a, b in of:
        assert a.n_overflows == b.n_overflows and a.max == b.max


@pytest.mark.parametrize("fmt_in,fmt_out", [("S16_LE", "S32_LE"), ("S24_LE", "S16_BE"), ("S32_BE", "S24_BE"),
                                            ("FLOAT64_LE", "FLOAT64_BE"), ("S8", "S8"), ("FLOAT_BE", "S24_4BE")])
@pytest.mark.parametrize("L,rs", [(128, 8), (1024, 4)])
def test_sample_formats_end_to_end(gpu_lib, oracle_libs, fmt_in, fmt_out, L, rs):
    """L = 128 / float_bits 64: the fused generic kernels; L = 1024 / float_bits 32: the transposing k_unpack / k_pack
    with their per-sample generic conversion (3 channels: a ragged 32-channel tile), planar output."""
    P = 4
    nch = 2 if L == 128 else 3
    inb, nin = interleaved_layout(nch, fmt_in, L)
    outb, nout = planar_layout(nch, fmt_out, L)
    filters = [Filter([c], [c], coeff=c % 2) for c in range(nch)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, 2])
    taps = configs.synthetic_filters(g, 12)
    sig = configs.synthetic_signal(g, 12, 10, sigma=0.1 if rs == 8 else 0.02)
    got, ref, of = run_both(g, taps, sig)
    if rs == 4 and not g.out_formats[0].sf.isfloat and g.out_formats[0].sf.sbytes == 4:
        # 32-bit samples out of a float32 engine: one float32 ulp is 16-32 LSB at -20 dBFS, the "1 LSB" of north_star
        # is a 24-bit statement; require a few ulp instead
        y, r = unpack_run(got, g.out_formats, L), unpack_run(ref, g.out_formats, L)
        assert np.abs(y - r).max() <= 4 * 2.0 ** -23 * np.abs(r).max()
    else:
        assert_parity(g, got, ref)
    if rs == 8 and not g.out_formats[0].sf.isfloat:     # float_bits 64, integer output: identical, counters included
        for a, b in of:
            assert (a.n_overflows, a.intlargest, a.largest) == (b.n_overflows, b.intlargest, b.largest)


@pytest.mark.parametrize("rs", [4, 8])
def test_mixing_delays_scales_and_dirac(gpu_lib, oracle_libs, rs):
    """Several inputs per filter (mixnscale INPUT, n_bufs > 1), several filters per output, one filter to two
    outputs, block delays, attenuations, a coeff:-1 filter, a short coefficient set, an unused output.
    (The dirac filter's attenuation is 1/3: integer input x 0.5 (or x 0.53 = 53/100) lands exactly on the
    quantiser's .5 ties, where 1e-16 of FFT noise decides the rounding and "identical at float_bits 64" cannot hold;
    x/3 never does.)"""
    L, P = 256, 6
    inb, nin = interleaved_layout(3, "S24_4LE", L)
    outb, nout = interleaved_layout(4, "S24_LE", L)
    filters = [Filter([0], [0], coeff=0), Filter([1, 2], [1], in_scales=[0.7, -0.2], coeff=1, delayblocks=2),
               Filter([2], [1, 2], out_scales=[1.0 / 3.0, 2.0], coeff=-1), Filter([0, 1, 2], [2], in_scales=[0.3, 0.3, 0.3], coeff=2),
               Filter([1], [0], out_scales=[-0.25], coeff=0, delayblocks=7)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, 3, 1])
    taps = configs.synthetic_filters(g, 13)
    # -34 dBFS: the x2.0 output attenuation and the three-filter sums stay below 2^21 LSB, where float32
    # resolves 1/8 LSB and the strict 1 LSB criterion is meaningful
    sig = configs.synthetic_signal(g, 13, 16, sigma=0.02)
    got, ref, _ = run_both(g, taps, sig, mac_split=1, scale=0.9)
    assert_parity(g, got, ref)
    # the unused output stays silent (bfrun.c never mixes into it)
    assert np.all(unpack_run(got, g.out_formats, L)[3] == 0)


@pytest.mark.parametrize("split", [2, 5])
def test_split_partition_sum_stays_within_tolerance(gpu_lib, oracle_libs, split):
    g = configs.diagonal_graph(2, 64, 40, 4, "S24_4LE")
    taps = configs.synthetic_filters(g, 14)
    sig = configs.synthetic_signal(g, 14, 50)
    got, ref, _ = run_both(g, taps, sig, mac_split=split)
    assert_parity(g, got, ref)


def test_runtime_control_and_crossfade(gpu_lib, oracle_libs):
    """cfc / cfia / cfoa / cfd at block boundaries (bfrun.c:1462-1478) including crossfaded coefficient
    switches to another set, to 'no coefficients' and back (bfrun.c:1726-1837), float_bits 32."""
    L, P = 128, 8
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S24_4LE", L)
    filters = [Filter([0], [0], coeff=0, crossfade=True), Filter([1], [0], coeff=1, crossfade=True),
               Filter([1], [1], coeff=2, crossfade=False), Filter([0], [1], coeff=-1, crossfade=True)]
    g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, 4])
    taps = configs.synthetic_filters(g, 15)
    sig = configs.synthetic_signal(g, 15, 30)
    script = {4: [(0, dict(coeff=1))], 9: [(0, dict(coeff=-1)), (2, dict(coeff=0))],
              12: [(3, dict(coeff=2)), (1, dict(coeff=1, in_scales=[0.5], out_scales=[1.5]))],
              13: [(3, dict(coeff=0, delayblocks=3))], 20: [(0, dict(coeff=0)), (1, dict(coeff=0, delayblocks=1))]}
    with Engine(g, mac_split=1) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref = [], []
        for b in range(30):
            for filt, kw in script.get(b, []):
                e.set_control(filt, **kw)
                d.set_control(filt, **kw)
            got.append(e.process_block(sig[b]))
            ref.append(d.process_block(sig[b]))
        d.close()
    assert_parity(g, np.stack(got), np.stack(ref))


@pytest.mark.parametrize("rs,B", [(8, 1), (4, 1), (4, 4)])
def test_stacked_delay_changes_follow_the_reference_ring(gpu_lib, oracle_libs, rs, B):
    """Delay changes (cfd) that follow each other within P blocks, up and down, mixed with 'no coefficients': the
    reference's P-slot ring (written at (t + delay) % P, read at (t - i) % P, bfrun.c:1579-1600, 1745-1754) then
    reads slots that alias blocks written under EARLIER delays, several changes deep.  The engine's longer ring is
    repaired from an exact P-slot mirror during such transitions (bf_engine.cu, begin_transitions); float_bits 64
    must match the oracle bit for bit, and the delay line slot for slot."""
    L, P = 256, 11
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "FLOAT64_LE" if rs == 8 else "S24_4LE", L)
    filters = [Filter([1, 0], [1, 0], coeff=-1, delayblocks=4), Filter([0], [0], coeff=0, delayblocks=2)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [10, 3])
    taps = configs.synthetic_filters(g, 31)
    nblk = 44
    sig = configs.synthetic_signal(g, 31, nblk, sigma=0.01)
    script = {2: [(0, dict(coeff=-1, delayblocks=0))], 5: [(0, dict(coeff=0, delayblocks=7))],
              8: [(0, dict(coeff=-1, delayblocks=10)), (1, dict(coeff=1, delayblocks=9))],
              11: [(0, dict(coeff=0, delayblocks=0))], 13: [(1, dict(coeff=0, delayblocks=1))],
              14: [(0, dict(coeff=-1, delayblocks=2))], 17: [(0, dict(coeff=0, delayblocks=3))],
              20: [(0, dict(coeff=1, delayblocks=6)), (1, dict(coeff=0, delayblocks=10))],
              23: [(0, dict(coeff=0, delayblocks=7))], 24: [(0, dict(coeff=0, delayblocks=1))],
              25: [(0, dict(coeff=0, delayblocks=8))], 26: [(1, dict(coeff=0, delayblocks=0))]}
    with Engine(g, mac_split=1, max_batch=B) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        ref, got = [], np.zeros((nblk, g.out_bytes), np.uint8)
        for b in range(nblk):
            for filt, kw in script.get(b, []):
                d.set_control(filt, **kw)
            ref.append(d.process_block(sig[b]))
        b = 0
        while b < nblk:
            for filt, kw in script.get(b, []):
                e.set_control(filt, **kw)
            k = 1
            while k < B and b + k < nblk and (b + k) not in script:
                k += 1
            e.process_blocks_async(sig[b:b + k], got[b:b + k], k)
            b += k
        e.synchronize()
        if rs == 8:
            # the delay lines in the reference's slot numbering, a few blocks after the last change
            # (which block sits in which slot; the spectra themselves differ by FFT rounding only)
            for filt in range(2):
                for slot in range(P):
                    a, r = e.debug_read(_abi.DBG_DELAYLINE, filt, slot), d.debug_read(_abi.DBG_DELAYLINE, filt, slot)
                    assert np.abs(a - r).max() <= 1e-12 * max(1.0, np.abs(r).max()), (filt, slot)
        d.close()
    assert_parity(g, got, np.stack(ref))


@pytest.mark.parametrize("rs", [4, 8])
def test_low_latency_schedule(gpu_lib, oracle_libs, rs):
    """BFCUDA_FLAG_LOW_LATENCY: partitions 1 .. P-1 of the next block are summed ahead of time, partition 0 when the
    input arrives (head + tail instead of the reference's left-to-right sum: the tolerance of a split partition sum).
    Control changes, a crossfade, a coefficient upload and a batched call in the middle must all discard or bypass the
    ahead-of-time sum."""
    L, P = 256, 8
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb, nout = interleaved_layout(2, "S24_4LE" if rs == 4 else "FLOAT64_LE", L)
    filters = [Filter([0], [0], coeff=0, crossfade=True), Filter([1], [0], coeff=1, delayblocks=2),
               Filter([1, 0], [1], in_scales=[0.5, -0.25], coeff=2), Filter([0], [1], coeff=-1)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, P, 3])
    taps = configs.synthetic_filters(g, 41)
    nblk = 40
    sig = configs.synthetic_signal(g, 41, nblk, sigma=0.01)
    script = {6: [(0, dict(coeff=1))], 11: [(1, dict(coeff=0, delayblocks=5))], 12: [(1, dict(coeff=0, delayblocks=1))],
              19: [(2, dict(coeff=2, in_scales=[0.25, 0.5]))], 25: [(0, dict(coeff=-1))], 31: [(3, dict(coeff=2))]}
    new_taps = configs.synthetic_filters(g, 42)[1]

    def run(engine):
        out = np.zeros((nblk, g.out_bytes), np.uint8)
        cur = {f: dict(coeff=flt.coeff, delayblocks=flt.delayblocks) for f, flt in enumerate(filters)}
        b = 0
        while b < nblk:
            for filt, kw in script.get(b, []):
                cur[filt] = dict(dict(delayblocks=0), **kw)
            for filt, kw in cur.items():            # the host's snapshot: every filter, every block (bfrun.c:1462-1478)
                engine.set_control(filt, **kw)
            if b == 15:
                engine.coeff_from_taps(1, new_taps)         # between two blocks
            if b == 21:                                     # a batched call in the middle
                engine.process_blocks_async(sig[b:b + 3], out[b:b + 3], 3)
                engine.synchronize()
                b += 3
                continue
            out[b] = engine.process_block(sig[b])
            b += 1
        return out

    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    ref = []
    for b in range(nblk):
        for filt, kw in script.get(b, []):
            d.set_control(filt, **kw)
        if b == 15:
            d.coeff_from_taps(1, new_taps)
        ref.append(d.process_block(sig[b]))
    d.close()
    outs = []
    for flags in (_abi.FLAG_LOW_LATENCY, 0):
        with Engine(g, flags=flags, mac_split=2, max_batch=4) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            outs.append(run(e))
    y, r = unpack_run(outs[0], g.out_formats, L), unpack_run(np.stack(ref), g.out_formats, L)
    if rs == 4:
        assert np.abs(y - r).max() <= 1
    else:
        assert np.abs(y - r).max() <= 1e-12
    # versus the plain schedule with an even two-way split: same head + tail structure only where P = 2; here the
    # split points differ (1 | P-1 against P/2 | P/2), so this is again a tolerance statement
    y2 = unpack_run(outs[1], g.out_formats, L)
    assert np.abs(y - y2).max() <= (1 if rs == 4 else 1e-12)


def test_golden_block_sequences(gpu_lib):
    """The committed vectors the reference itself produced (tests/golden/make_golden.py), no oracle involved."""
    blk = np.load(os.path.join(HERE, "golden", "blocks.npz"))
    ga = golden_graph_a()
    with Engine(ga) as e:
        e.coeff_from_taps(0, blk["a_taps0"])
        e.coeff_from_taps(1, blk["a_taps1"])
        assert_parity(ga, e.run(blk["a_sig"]), blk["a_out"])
    gb = golden_graph_b()
    with Engine(gb) as e:
        e.coeff_from_taps(0, blk["b_taps0"])
        e.coeff_from_taps(1, blk["b_taps1"])
        assert_parity(gb, run_golden_b(e, blk["b_sig"]), blk["b_out"])
    gc = golden_graph_c()
    with Engine(gc, mac_split=1) as e:
        for c in range(len(gc.coeff_n_blocks)):
            e.coeff_from_taps(c, blk[f"c_taps{c}"])
        assert_parity(gc, e.run(blk["c_sig"]), blk["c_out"])
    gd = golden_graph_d()
    with Engine(gd) as e:
        e.coeff_from_taps(0, blk["d_taps0"])
        e.coeff_from_taps(1, blk["d_taps1"])
        y, r = unpack_run(e.run(blk["d_sig"]), gd.out_formats, 64), unpack_run(blk["d_out"], gd.out_formats, 64)
        # Dither on: last-bit FFT differences flip a few decisions, each costing three +-1 LSB samples.  Blocks 5-7 clip
        # hard; there the error-feedback state reaches ~1e6 LSB where float32 resolves 0.1 LSB, the two implementations'
        # states part by that much, and the recurrence (poles on the unit circle) never forgets it: afterwards the
        # outputs agree within +-2 LSB but no longer sample for sample -- for ANY two float32 FFTs, the reference
        # against itself with another FFTW included.
        assert np.abs(y - r).max() <= 2
        assert np.mean(y[:, :5 * 64] != r[:, :5 * 64]) < 0.10


@pytest.mark.parametrize("variant", ["0", "1"])
@pytest.mark.parametrize("rs", [4, 8])
def test_mac_stage_bit_exact(gpu_lib, oracle_libs, rs, variant, monkeypatch):
    """Read the delay line and the coefficient blocks back from the device in the reference's blocked layout,
    run the reference's own convolve / convolve_add order on exactly those buffers (bfrun.c:1737-1754), and
    require the engine's filter outputs to match bit for bit -- for both data paths of the MAC kernel."""
    monkeypatch.setenv("BFCUDA_MAC_VARIANT", variant)
    L, P, nb = 1024, 12, 17
    g = configs.diagonal_graph(2, L, P, rs, "S24_4LE")
    g.filters[1].coeff = 0
    g.filters.append(Filter([0], [1], coeff=-1))
    g.coeff_n_blocks = [P, 5]
    taps = configs.synthetic_filters(g, 16)
    sig = configs.synthetic_signal(g, 16, nb)
    o = po.Convolver("oracle", L, rs)
    with Engine(g, mac_split=1) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        for b in range(nb):
            e.process_block(sig[b])
        t = nb - 1
        H = [[e.coeff_get_block(c, i) for i in range(g.coeff_n_blocks[c])] for c in range(2)]
        for f in range(3):
            coeff = g.filters[f].coeff
            got = e.debug_read(_abi.DBG_FILTER_OUTPUT, f)
            slot = lambda i: e.debug_read(_abi.DBG_DELAYLINE, f, (t - i) % P)
            if coeff < 0:
                want = o.dirac_convolve(slot(0))
            else:
                want = o.convolve(slot(0), H[coeff][0])
                for i in range(1, g.coeff_n_blocks[coeff]):
                    o.convolve_add(slot(i), H[coeff][i], want)
            assert np.array_equal(got, want), (f, np.abs(got - want).max())


def test_coefficient_layouts_round_trip(gpu_lib, oracle_libs):
    """bfaccess->coeffs_data / the "processed" coefficient format stay in the reference's blocked layout."""
    L, P = 512, 3
    g = configs.diagonal_graph(1, L, P, 8, "S24_4LE")
    o = po.Convolver("oracle", L, 8)
    taps = configs.synthetic_filters(g, 17)[0]
    with Engine(g) as e:
        e.coeff_from_taps(0, taps, 0.5)
        for i in range(P):
            want = o.coeffs2cbuf(taps[i * L:(i + 1) * L], 0.5)
            got = e.coeff_get_block(0, i)
            assert np.abs(got - want).max() <= 1e-15
            e.coeff_set_block(0, i, want)
            assert np.array_equal(e.coeff_get_block(0, i), want)         # processed blocks survive verbatim
        e.coeff_runtime_block(0, 1, taps[:L])
        assert np.abs(e.coeff_get_block(0, 1) - o.runtime_coeffs2cbuf(taps[:L])).max() <= 1e-15
        with pytest.raises(_abi.BfcudaError) as err:
            bad = taps.copy()
            bad[3] = np.inf
            e.coeff_from_taps(0, bad)
        assert err.value.code == -5


@pytest.mark.parametrize("L,rs,B", [(16384, 4, 1), (16384, 4, 2), (8192, 8, 1)])
def test_maximum_partition_sizes(gpu_lib, oracle_libs, L, rs, B):
    """The largest partitions the single-block transforms take: 16384 samples at float_bits 32 (N = 32768, the
    1024-thread size-specialised kernels with twiddles left in global memory), 8192 at float_bits 64."""
    g = configs.diagonal_graph(2, L, 2, rs, "S24_4LE")
    taps = configs.synthetic_filters(g, 51)
    sig = configs.synthetic_signal(g, 51, 6, sigma=0.01)
    with Engine(g, mac_split=1, max_batch=B) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got = np.zeros((6, g.out_bytes), np.uint8)
        for b in range(0, 6, B):
            e.process_blocks_async(sig[b:b + B], got[b:b + B], B)
        e.synchronize()
        ref = d.run(sig)
        d.close()
    assert_parity(g, got, ref)


def test_maximum_channel_and_filter_counts(gpu_lib, oracle_libs):
    """BF_MAXCHANNELS = BF_MAXFILTERS = 256 (bfmod.h:22-23): 256 inputs, 256 outputs, 256 filters; every fourth
    filter reads its neighbour's input too and every output but the last is also fed by the next filter."""
    n, L, P = 256, 64, 2
    inb, nin = interleaved_layout(n, "S16_LE", L)
    outb, nout = interleaved_layout(n, "S24_LE", L)
    filters = []
    for f in range(n):
        ins = [f, (f + 1) % n] if f % 4 == 0 else [f]
        outs = [f] if f == 0 else [f, f - 1]
        filters.append(Filter(ins, outs, in_scales=[0.5] * len(ins), out_scales=[0.3183098861837907] * len(outs),
                              coeff=f % 3, delayblocks=f % 2))
    g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, 1, P])
    taps = configs.synthetic_filters(g, 52)
    sig = configs.synthetic_signal(g, 52, 8, sigma=0.02)
    got, ref, _ = run_both(g, taps, sig, mac_split=1)
    assert_parity(g, got, ref)


def test_errors_and_limits(gpu_lib):
    g = configs.config_c1_chained()
    g.filters[0].from_filters, g.filters[0].fscales = [4], [1.0]      # a source with a higher index: not topological
    with pytest.raises(_abi.BfcudaError) as err:
        Engine(g)
    assert err.value.code == -1 and "processing order" in str(err.value)      # bfconf.c:2933-2964
    g = configs.diagonal_graph(1, 1 << 23, 1, 4, "S16_LE")
    with pytest.raises(_abi.BfcudaError) as err:
        Engine(g)
    assert err.value.code == -7                                         # partition beyond the four-step transform
    g = configs.config_c1_chained()
    g.filter_length = 32768                                             # chained filters: one-block transforms only
    g.in_formats, g.in_bytes = interleaved_layout(2, "S24_4LE", 32768)
    g.out_formats, g.out_bytes = interleaved_layout(2, "S24_4LE", 32768)
    with pytest.raises(_abi.BfcudaError) as err:
        Engine(g)
    assert err.value.code == -7
    g = configs.diagonal_graph(1, 64, 2, 4, "S16_LE")
    with Engine(g) as e:
        e.coeff_from_taps(0, np.full(128, 1e3, np.float32))
        x = np.full((1, 64), 30000.0)
        raw = pack_block(x, g.in_formats, g.in_bytes)
        out = e.process_block(raw)
        of = e.overflow(0)
        assert of.n_overflows == 64 and of.max == 32767.0 and of.largest > 32767.0
        assert np.all(unpack_run(out[None], g.out_formats, 64)[0] == 32767)
        e.reset_overflow()
        assert e.overflow(0).n_overflows == 0
        e.coeff_from_taps(0, np.full(128, 3e38, np.float32))
        with pytest.raises(_abi.BfcudaError) as err:
            e.process_block(raw)
        assert err.value.code == -5             # NaN/Inf in the output: the reference abort()s


@pytest.mark.parametrize("rs,B", [(4, 2), (4, 4), (4, 8), (4, 16), (8, 3), (8, 4)])
def test_batched_launches_are_bit_identical_to_block_by_block(gpu_lib, oracle_libs, rs, B):
    """max_batch > 1 (offline throughput mode): up to B blocks per launch, coefficient and delay-line spectra
    reused in registers across the batch.  Every output byte and every overflow counter must equal the
    block-by-block engine's, including control changes and crossfades falling between and inside batches, a
    ragged last batch, mixed inputs, delays and a split partition sum."""
    L, P = 256, 10
    inb, nin = interleaved_layout(3, "S24_4LE", L)
    outb, nout = interleaved_layout(3, "S16_LE", L)
    filters = [Filter([0], [0], coeff=0, crossfade=True), Filter([1, 2], [1], in_scales=[0.7, -0.2], coeff=1, delayblocks=2),
               Filter([2], [1, 2], out_scales=[1.0 / 3.0, 2.0], coeff=-1), Filter([0], [2], coeff=2, delayblocks=9)]
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, 3, 7])
    taps = configs.synthetic_filters(g, 18)
    nb = 37
    sig = configs.synthetic_signal(g, 18, nb, sigma=0.3)        # loud: exercises clipping and overflow counters
    script = {5: [(0, dict(coeff=2))], 6: [(1, dict(coeff=1, delayblocks=0, in_scales=[0.1, 0.3]))],
              16: [(0, dict(coeff=-1)), (3, dict(coeff=0))],
              22: [(1, dict(coeff=1, delayblocks=6, in_scales=[0.1, 0.3]))],      # delay increase: stale-slot reads
              29: [(0, dict(coeff=1))],
              33: [(1, dict(coeff=1, delayblocks=2, in_scales=[0.1, 0.3]))]}      # decrease: "ahead" blocks alias

    def run(max_batch, split):
        with Engine(g, mac_split=split, max_batch=max_batch) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h, 40.0)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            b = 0
            while b < nb:
                for filt, kw in script.get(b, []):
                    e.set_control(filt, **kw)
                # largest batch that does not run past the next scripted change
                n = 1
                while n < max_batch and b + n < nb and (b + n) not in script:
                    n += 1
                e.process_blocks_async(sig[b:b + n], out[b:b + n], n)
                b += n
            e.synchronize()
            stats = [(e.overflow(o).n_overflows, e.overflow(o).intlargest, e.overflow(o).largest) for o in range(3)]
        return out, stats

    for split in (1, 3):
        ref_out, ref_stats = run(1, split)
        got_out, got_stats = run(B, split)
        assert np.array_equal(got_out, ref_out), split
        assert got_stats == ref_stats and ref_stats[1][0] > 0
        if split == 1 and B == 2:
            # ... and the block-by-block engine itself follows the reference's P-slot ring through every delay
            # change (the engine's ring is longer; bf_engine.cu begin_transitions)
            d = po.BlockDriver("oracle", g)
            for c, h in enumerate(taps):
                d.coeff_from_taps(c, h, 40.0)
            want = []
            for b in range(nb):
                for filt, kw in script.get(b, []):
                    d.set_control(filt, **kw)
                want.append(d.process_block(sig[b]))
            d.close()
            assert_parity(g, ref_out, np.stack(want))


def chained_graph(L, P, rs, fmt="S24_4LE", crossfade=False):
    """bench1_config's topology (four input-fed filters feeding two filter-fed ones through to_filters,
    /root/reference/bench1_config:28-57) made harder: a third level, a consumer that also mixes an input channel,
    source multipliers, a delayed consumer and crossfading sources."""
    inb, nin = interleaved_layout(2, fmt, L)
    outb, nout = interleaved_layout(3, fmt, L)
    filters = [Filter([0], [], coeff=2, crossfade=crossfade), Filter([0], [], coeff=3), Filter([1], [2], coeff=4),
               Filter([1], [], coeff=5, crossfade=crossfade),
               Filter([], [0], coeff=0, from_filters=[0, 3], fscales=[1.0, -0.5]),
               Filter([1], [1], in_scales=[0.25], coeff=1, from_filters=[1, 2], delayblocks=1),
               Filter([], [2], out_scales=[0.5], coeff=-1, from_filters=[4, 5], fscales=[0.5, 0.25])]
    return FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, P, 2, P, 1, P])


@pytest.mark.parametrize("L,P,rs", [(64, 3, 4), (64, 3, 8), (1024, 4, 4), (2048, 3, 8)])
def test_filter_chaining_against_oracle(gpu_lib, oracle_libs, L, P, rs):
    """to_filters / from_filters on the device (bfrun.c:1603-1660, convolver_convolve_eval): the sources' outputs are
    evaluated in the time domain per block and become an input spectrum of the consumer, level by level."""
    g = chained_graph(L, P, rs)
    taps = [t * 0.5 for t in configs.synthetic_filters(g, 21)]
    sig = configs.synthetic_signal(g, 21, 3 * P + 4, sigma=0.02)
    got, ref, of = run_both(g, taps, sig, mac_split=1)
    assert_parity(g, got, ref)
    assert np.abs(unpack_run(ref, g.out_formats, L)).max() > 1e3


def test_filter_chaining_control_crossfade_and_batches(gpu_lib, oracle_libs):
    """Chained filters through coefficient switches with crossfade on the SOURCE filters (the consumer must see the
    blended block), source-multiplier changes, and the batched mode (bit-identical to block by block)."""
    L, P, nb = 256, 5, 26
    g = chained_graph(L, P, 4, crossfade=True)
    taps = [t * 0.5 for t in configs.synthetic_filters(g, 22)]
    sig = configs.synthetic_signal(g, 22, nb, sigma=0.02)
    script = {4: [(0, dict(coeff=3))], 9: [(3, dict(coeff=-1)), (4, dict(coeff=0, fscales=[0.7, 0.1]))],
              15: [(0, dict(coeff=2)), (5, dict(coeff=1, delayblocks=0, in_scales=[0.25]))], 21: [(3, dict(coeff=5))]}

    def run_engine(max_batch):
        with Engine(g, mac_split=1, max_batch=max_batch) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            b = 0
            while b < nb:
                for filt, kw in script.get(b, []):
                    e.set_control(filt, **kw)
                n = 1
                while n < max_batch and b + n < nb and (b + n) not in script:
                    n += 1
                e.process_blocks_async(sig[b:b + n], out[b:b + n], n)
                b += n
            e.synchronize()
        return out

    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    want = []
    for b in range(nb):
        for filt, kw in script.get(b, []):
            d.set_control(filt, **kw)
        want.append(d.process_block(sig[b]))
    d.close()
    one = run_engine(1)
    assert_parity(g, one, np.stack(want))
    assert np.array_equal(run_engine(4), one)


def dither_graph(L, P, rs, fmt, rate=100):
    sf_i, sf_f = parse_sample_format(fmt), parse_sample_format("FLOAT_LE")
    inb, nin = interleaved_layout(2, "S24_4LE", L)
    outb = [BufferFormat(sf_i, 1, 0), BufferFormat(sf_i, 1, L * sf_i.bytes), BufferFormat(sf_f, 1, 2 * L * sf_i.bytes)]
    nout = 2 * L * sf_i.bytes + 4 * L + 32
    filters = [Filter([0], [0], coeff=0), Filter([1], [1], coeff=1), Filter([0, 1], [2], in_scales=[0.5, 0.5], coeff=0)]
    return FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P, P], sampling_rate=rate,
                       apply_dither=[True, True, True])


@pytest.mark.parametrize("L,rs,fmt,B", [(64, 4, "S16_LE", 1), (64, 8, "S24_LE", 1), (1024, 4, "S16_LE", 1), (1024, 4, "S16_BE", 4)])
def test_dither_quantiser_bit_exact_on_the_devices_own_samples(gpu_lib, oracle_libs, L, rs, fmt, B):
    """HP-TPDF dither + error feedback (dither.c, dither_funs.h:7-68).  The recurrence amplifies
    nothing but it does turn last-bit differences of the inverse FFT into +-1 LSB decisions, so the quantiser is pinned
    on its own: the time-domain blocks the device produced (read back) go through the ORACLE's convolver_cbuf2raw
    with dither, and the device's raw bytes and overflow counters must equal that bit for bit -- through table
    wraps, clipping blocks, and batches."""
    P, nb = 2, 40
    g = dither_graph(L, P, rs, fmt, rate=100 if L == 64 else 1500)
    rng = np.random.default_rng(78)
    taps = [rng.standard_normal(L * P).astype(np.float32) / 6 for _ in range(2)]
    x = np.round(rng.standard_normal((nb, 2, L)) * 0.1 * (1 << 23))
    x[5:8] *= 40
    x = np.clip(x, -(1 << 23), (1 << 23) - 1)
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(nb)])
    o = po.Convolver("oracle", L, rs)
    o.dither_init(2, g.sampling_rate)
    ofs = [_abi.OverflowC(), _abi.OverflowC()]
    for k in range(2):
        ofs[k].max = float((1 << (8 * g.out_formats[k].sf.sbytes - 1)) - 1)
    with Engine(g, max_batch=B) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        for b0 in range(0, nb, B):
            got = np.zeros((B, g.out_bytes), np.uint8)
            if B == 1:
                got[0] = e.process_block(sig[b0])
                times = [[e.debug_read(_abi.DBG_OUTPUT_TIME, k)[:L] for k in range(2)]]
            else:
                # batches: the time-domain blocks of a batch come from a second engine run block by block
                e.process_blocks_async(sig[b0:b0 + B], got, B)
                e.synchronize()
                times = None
            if times is not None:
                want = np.zeros(g.out_bytes, np.uint8)
                for k in range(2):
                    cbuf = np.zeros(2 * L, o.dtype)
                    cbuf[:L] = times[0][k]
                    o.cbuf2raw_dither(cbuf, want, g.out_formats[k], ofs[k], k)
                nbytes = 2 * L * g.out_formats[0].sf.bytes
                assert np.array_equal(got[0][:nbytes], want[:nbytes]), b0
            else:
                batched = got.copy()
                if b0 == 0:
                    ref_eng = Engine(g)
                    for c, h in enumerate(taps):
                        ref_eng.coeff_from_taps(c, h)
                for i in range(B):
                    assert np.array_equal(ref_eng.process_block(sig[b0 + i]), batched[i]), (b0, i)
        if B == 1:
            for k in range(2):
                dev = e.overflow(k)
                assert (dev.n_overflows, dev.intlargest, dev.largest) == (ofs[k].n_overflows, ofs[k].intlargest, ofs[k].largest)
                assert dev.n_overflows > 0
        else:
            for k in range(2):
                a, b = e.overflow(k), ref_eng.overflow(k)
                assert (a.n_overflows, a.intlargest, a.largest) == (b.n_overflows, b.intlargest, b.largest)
            ref_eng.close()


def test_dither_end_to_end_against_oracle(gpu_lib, oracle_libs):
    """Whole path with dither on against the oracle's block sequence: identical except where a last-bit difference of
    the inverse FFT flips a quantiser decision, which the error feedback undoes within three samples."""
    L, P, nb = 1024, 3, 12
    g = dither_graph(L, P, 4, "S16_LE", rate=48000)
    taps = [t * 0.5 for t in configs.synthetic_filters(g, 31)]
    sig = configs.synthetic_signal(g, 31, nb, sigma=0.05)
    got, ref, of = run_both(g, taps, sig)
    y, r = unpack_run(got, g.out_formats, L), unpack_run(ref, g.out_formats, L)
    for k in range(2):
        diff = np.abs(y[k] - r[k])
        # float32 FFT noise at 16 bit is ~1e-2 LSB: ~1 % of the decisions flip, each costing three +-1 LSB samples
        assert diff.max() <= 2 and np.mean(diff > 0) < 0.10, (k, diff.max(), np.mean(diff > 0))
        assert np.abs(r[k]).max() > 3000
    assert np.abs(y[2] - r[2]).max() <= 1e-6
    # dither is on: the undithered run differs on most samples' last bit pattern
    g2 = dither_graph(L, P, 4, "S16_LE", rate=48000)
    g2.apply_dither = None
    with Engine(g2) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        plain = unpack_run(e.run(sig), g2.out_formats, L)
    assert np.mean(plain[0] != y[0]) > 0.2


def matrix_graph(n, L, P, rs, in_scale=1.0):
    """n inputs x n outputs, one filter per (input, output) pair: every input feeds n filters (crosstalk / room
    correction matrices, xtc_config's topology scaled up)."""
    inb, nin = interleaved_layout(n, "S24_4LE", L)
    outb, nout = interleaved_layout(n, "S24_4LE", L)
    filters = [Filter([i], [o], in_scales=[in_scale], out_scales=[1.0 / n], coeff=o * n + i, crossfade=(i == 0))
               for o in range(n) for i in range(n)]
    return FilterGraph(L, P, rs, inb, outb, nin, nout, filters, [P] * (n * n))


@pytest.mark.parametrize("L,P,rs,B", [(64, 6, 4, 1), (1024, 5, 4, 1), (1024, 5, 4, 4), (256, 4, 8, 2)])
def test_delay_lines_shared_across_filters(gpu_lib, oracle_libs, L, P, rs, B):
    """Filters fed by the same input with the same scale and delay share one delay line ("input spectra reused across
    filters"): n rings instead of n^2.  Results must not depend on it -- byte-identical to the engine with sharing off
    and within tolerance of the oracle -- through run-time changes that split a filter off its ring (input scale,
    delay), the delay transition on the split-off copy, crossfaded coefficient switches, and re-merging once the rings
    are identical again."""
    n, nb = 3, 8 * P + 20
    g = matrix_graph(n, L, P, rs)
    taps = [t * 0.8 for t in configs.synthetic_filters(g, 41)]
    sig = configs.synthetic_signal(g, 41, nb, sigma=0.03)
    script = {3: [(1, dict(coeff=1, in_scales=[0.5]))],                    # filter 1 leaves input 1's ring
              5: [(0, dict(coeff=4))],                                     # crossfade on a ring owner
              7: [(3, dict(coeff=3, delayblocks=2))],                      # owner of input 0's ring changes delay
              9: [(6, dict(coeff=6, delayblocks=2))],                      # ... a follower follows (own ring, transition)
              12: [(1, dict(coeff=1, in_scales=[1.0]))],                   # back to the common scale: re-merge later
              14: [(3, dict(coeff=3, delayblocks=0)), (6, dict(coeff=6, delayblocks=0))]}

    def run_engine(flags):
        rings = []
        with Engine(g, mac_split=1, max_batch=B, flags=flags) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            b = 0
            while b < nb:
                for filt, kw in script.get(b, []):
                    e.set_control(filt, **kw)
                k = 1
                while k < B and b + k < nb and (b + k) not in script:
                    k += 1
                e.process_blocks_async(sig[b:b + k], out[b:b + k], k)
                e.synchronize()
                rings.append((b, e.info().n_streams))
                b += k
        return out, dict(rings)

    shared, rings = run_engine(0)
    private, rings_off = run_engine(_abi.FLAG_NO_STREAM_SHARING)
    assert np.array_equal(shared, private)
    assert rings[0] == n and rings_off[0] == n * n                  # three rings for nine filters
    assert max(rings.values()) >= n + 3                             # the splits happened ...
    assert rings[max(rings)] == n                                   # ... and the rings merged again at the end
    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    want = []
    for b in range(nb):
        for filt, kw in script.get(b, []):
            d.set_control(filt, **kw)
        want.append(d.process_block(sig[b]))
    d.close()
    assert_parity(g, shared, np.stack(want))


@pytest.mark.parametrize("n_ch,B", [(64, 1), (64, 4), (8, 4), (40, 2), (4, 1)])
def test_packed_s24_tiles_equal_the_4_byte_layout(gpu_lib, n_ch, B):
    """massive_config's own sample format, packed "S24_LE" on interleaved channels (a multiple of four of them), takes a
    tile path of its own in k_unpack / k_pack (24 lanes move the 96 bytes of a 32-channel row, shuffles pick the three
    bytes; ragged last tiles and shards of 8 channels included): the same samples in the S24_4LE layout (pinned against
    the oracle elsewhere) must come out as the same integers, overflow counters included -- loud enough to clip."""
    L, P, nb = 256, 3, 9
    outs, stats = [], []
    for fmt in ("S24_4LE", "S24_LE"):
        g = configs.diagonal_graph(n_ch, L, P, 4, fmt)
        taps = configs.synthetic_filters(g, 31)
        sig = configs.synthetic_signal(g, 31, nb, sigma=0.4)
        with Engine(g, max_batch=B) as e:
            assert e.lib is not None
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h, 3.0)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            b = 0
            while b < nb:
                n = min(B, nb - b)
                e.process_blocks_async(sig[b:b + n], out[b:b + n], n)
                b += n
            e.synchronize()
            stats.append([(e.overflow(o).n_overflows, e.overflow(o).intlargest, e.overflow(o).largest) for o in range(n_ch)])
        outs.append(np.stack([unpack_block(blk, g.out_formats, L) for blk in out]))
    assert np.array_equal(outs[0], outs[1])
    assert stats[0] == stats[1] and sum(s[0] for s in stats[0]) > 0 and np.abs(outs[0]).max() == 2 ** 23


@pytest.mark.parametrize("which", ["1", "2"])
def test_shared_ring_mac_kernels_bit_identical(gpu_lib, oracle_libs, which):
    """bf_mac_tile.cu: the batched MAC with a block-shared operand ring -- bulk-copy staged (BFCUDA_MAC_TILE=1) and
    cooperative cp.async (=2; the default for 16-block launches of small shards).  The library reads the variable
    once, so the batched == block-by-block test (control changes, crossfades, delays, split sums, ragged batches) and
    the headline-size byte-identity test run in a child interpreter with the kernel forced wherever it applies."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(HERE, "test_gpu_engine.py"),
                        os.path.join(HERE, "test_gpu_fullsize.py"), "-k",
                        "batched_launches_are_bit_identical or c3_batched or c3_mac_stage"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, BFCUDA_MAC_TILE=which))
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("calls,B", [(600, 1), (300, 4)])
def test_pipelined_stages_equal_serialised_stages(gpu_lib, calls, B):
    """The engine overlaps the stages of consecutive launches on three streams (bf_engine.cu).  tests/checks/soak_pipeline.py
    drives hundreds of asynchronous calls with random coefficient / delay / scale changes, crossfades, chained filters
    and shared delay lines once with the pipeline and once with BFCUDA_FLAG_SERIAL_STAGES: byte-identical outputs at
    every checkpoint, i.e. no stage ever touches a buffer a neighbouring launch still uses."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(HERE, "checks", "soak_pipeline.py"), str(calls), str(B)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("env,cases,seed", [({}, 40, 7), ({"FUZZ_BIG": "1"}, 24, 9002), ({"FUZZ_PMAX": "24"}, 40, 9010),
                                            ({"FUZZ_WIDE": "1"}, 24, 9020), ({"FUZZ_LL": "1", "FUZZ_B": "1"}, 24, 9031)])
def test_randomised_graphs_against_oracle(gpu_lib, oracle_libs, env, cases, seed):
    """tests/checks/fuzz_parity.py: random graphs (mixes, chaining, delays, crossfade, formats, partition counts from
    1), random run-time control scripts, random batch sizes, engine against the oracle under the north_star
    tolerances.  It found the early-block case of delay changes (partitions older than the first block must stay
    unread) and, at 11-12 partitions (FUZZ_BIG, seed 9002), delay changes stacked within P blocks of each other."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(HERE, "checks", "fuzz_parity.py"), str(cases), str(seed)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0 and f"{cases}/{cases}" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
