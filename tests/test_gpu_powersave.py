"""powersave (bfconf `powersave`, /root/reference/bfrun.c:1541-1552, 1613-1700, 722-772) on the engine path.

With exact-zero detection the reference only SKIPS work on zero blocks, so results cannot change: the engine with
powersave on must produce the bytes of the engine with powersave off (it then does not read the flagged delay-line slots
nor the coefficient blocks they would meet).  With an analog level a frame whose peak is below it is made truly zero:
compared with the oracle's replay, which restates test_silent()."""
import numpy as np
import pytest

from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import pack_block
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu


def gated_signal(g, seed, n_blocks, quiet_level=0.0):
    """white noise with silent stretches: channel 0 always on, channel 1 silent from block 3 on, channel 2 bursts,
    channel 3 silent throughout (quiet_level > 0: low-level noise instead of digital silence)."""
    rng = np.random.default_rng(seed)
    L, n_ch = g.filter_length, len(g.in_formats)
    fs = float(1 << 23)
    blocks = np.zeros((n_blocks, g.in_bytes), np.uint8)
    for b in range(n_blocks):
        x = rng.standard_normal((n_ch, L)) * 0.02
        gate = np.ones(n_ch)
        if n_ch > 1 and b >= 3:
            gate[1] = 0
        if n_ch > 2 and (b // 3) % 2 == 1:
            gate[2] = 0
        if n_ch > 3:
            gate[3] = 0
        q = rng.standard_normal((n_ch, L)) * quiet_level
        v = np.where(gate[:, None] > 0, x, q)
        blocks[b] = pack_block(np.clip(np.round(v * fs), -fs, fs - 1), g.in_formats, g.in_bytes)
    return blocks


@pytest.mark.parametrize("L,P,B,rs", [(1024, 6, 1, 4), (1024, 6, 4, 4), (256, 12, 8, 4), (4096, 3, 2, 8), (32768, 2, 1, 4)])
def test_digital_powersave_changes_no_byte(gpu_lib, L, P, B, rs):
    g = configs.diagonal_graph(4, L, P, rs, "S24_4LE")
    taps = configs.synthetic_filters(g, 23)
    sig = gated_signal(g, 23, 3 * P + 5)
    outs = []
    for ps in (False, True):
        g.powersave = ps
        with Engine(g, max_batch=B) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            outs.append(e.run(sig))
    assert np.array_equal(outs[0], outs[1])
    y = unpack_run(outs[1], g.out_formats, L)
    assert np.abs(y[0]).max() > 1e4 and np.all(y[3] == 0)
    assert np.all(y[1, (3 + P) * L:] == 0)                # channel 1: silent once its delay line has drained


@pytest.mark.parametrize("B", [1, 4])
def test_analog_powersave_matches_the_reference_rule(gpu_lib, oracle_libs, B):
    """-60 dB level: the low-level noise on the gated channels (-80 dBFS) is below it, so their frames are zeroed --
    but only frames whose PREVIOUS block is quiet too (the test looks at the whole cbuf, bfrun.c:1541-1546)."""
    L, P = 1024, 5
    g = configs.diagonal_graph(4, L, P, 4, "S24_4LE")
    g.powersave, g.analog_powersave = True, 1e-3
    taps = configs.synthetic_filters(g, 29)
    sig = gated_signal(g, 29, 22, quiet_level=1e-4)
    with Engine(g, max_batch=B) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref = e.run(sig), d.run(sig)
        d.close()
    y, r = unpack_run(got, g.out_formats, L), unpack_run(ref, g.out_formats, L)
    assert np.abs(y - r).max() <= 1
    assert np.all(r[3] == 0) and np.abs(r[0]).max() > 1e4
    # and the level matters: without powersave the quiet channel is NOT silent
    g.powersave = False
    with Engine(g) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        plain = unpack_run(e.run(sig), g.out_formats, L)
    assert np.abs(plain[3]).max() > 100


def test_powersave_with_shared_rings_mixes_and_delay_changes(gpu_lib, oracle_libs):
    """flags must follow the delay-line bookkeeping: filters sharing a ring, a multi-input mix (never skipped), a
    run-time delay change (flags ignored until the repaired slots are out of reach)."""
    from brutefir_b200.formats import interleaved_layout
    from brutefir_b200.graph import Filter, FilterGraph
    L, P = 256, 6
    inb, nin = interleaved_layout(3, "S24_4LE", L)
    outb, nout = interleaved_layout(3, "S24_4LE", L)
    filters = [Filter([1], [0], coeff=0), Filter([1], [1], coeff=1), Filter([0, 2], [2], in_scales=[0.5, 0.5], coeff=2),
               Filter([2], [0], coeff=1, delayblocks=1)]
    g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, P])
    g.powersave = True
    taps = configs.synthetic_filters(g, 31)
    sig = gated_signal(g, 31, 40)
    script = {9: (3, 3), 20: (3, 0), 27: (0, 2)}        # block -> (filter, new delay in blocks)
    with Engine(g) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref = [], []
        for b in range(sig.shape[0]):
            if b in script:
                f, dl = script[b]
                e.set_control(f, g.filters[f].coeff, delayblocks=dl)
                d.set_control(f, g.filters[f].coeff, delayblocks=dl)
            got.append(e.process_block(sig[b]))
            ref.append(d.process_block(sig[b]))
        d.close()
    y, r = unpack_run(np.stack(got), g.out_formats, L), unpack_run(np.stack(ref), g.out_formats, L)
    assert np.abs(r).max() > 1e4 and np.abs(y - r).max() <= 1
