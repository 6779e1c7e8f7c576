"""Generate the golden fixtures of tests/golden/ from THE REFERENCE ITSELF.

Runs the reference's own convolver sources compiled from /root/reference (oracle/_ref/libbfref.so: its
fftw_convolver.c, fftw_convfuns.h, convolver_xmm.c, raw2real.h, real2raw.h, dither_funs.h, unmodified,
against oracle/shim's stand-in for the absent FFTW3) and stores seeded inputs and the outputs it produced.
The reference ships no tests or vectors of its own, so these are the pins.

Re-run from the repo root in the authoring container:   python tests/golden/make_golden.py
Also converts the reference's shipped xtc coefficient files (directpath.txt / crosspath.txt, plain text
taps) into xtc_taps.npz so the crosstalk-cancellation case runs where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from brutefir_b200 import _abi, configs  # noqa: E402
from brutefir_b200.formats import BufferFormat, interleaved_layout, pack_block, parse_sample_format  # noqa: E402
from brutefir_b200.graph import Filter, FilterGraph  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_golden import golden_graph_c, golden_graph_d  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def golden_functions():
    """Per-function vectors at L = 32 (N = 64), both precisions."""
    out = {}
    rng = np.random.default_rng(4242)
    for rs in (4, 8):
        cv = po.Convolver("ref", 32, rs)
        dt, N = cv.dtype, cv.N
        t = f"f{rs * 8}"
        x = rng.standard_normal(N).astype(dt)
        hc = cv.time2freq(x)
        out[f"{t}_time"] = x
        out[f"{t}_time2freq"] = hc
        out[f"{t}_freq2time"] = cv.freq2time(hc)
        a, b, c = (rng.standard_normal(N).astype(dt) for _ in range(3))
        out[f"{t}_mix_in"] = np.stack([a, b, c])
        out[f"{t}_mix_scales"] = np.array([0.5, -1.25, 3.0e-3])
        out[f"{t}_mix_INPUT"] = cv.mixnscale([a, b, c], [0.5, -1.25, 3.0e-3], cv.MIXMODE_INPUT)
        out[f"{t}_mix_OUTPUT"] = cv.mixnscale([a, b, c], [0.5, -1.25, 3.0e-3], cv.MIXMODE_OUTPUT)
        out[f"{t}_convolve"] = cv.convolve(a, b)
        out[f"{t}_convolve_add"] = cv.convolve_add(a, b, c.copy())
        out[f"{t}_dirac"] = cv.dirac_convolve(a)
        taps = rng.standard_normal(20).astype(dt)
        out[f"{t}_taps"] = taps
        out[f"{t}_coeffs2cbuf"] = cv.coeffs2cbuf(taps, 0.75)
        if rs == 4:     # the reference's double crossfade branch is broken
            new, old = a.copy(), b.copy()
            cv.crossfade_inplace(new, old)
            out[f"{t}_crossfade"] = new
        # quantiser corner cases (SURVEY.md preamble) + clipping, S24_4LE
        q_in = np.array([-0.5, -1.0, -1.5, -2.5, 0.5, 1.5, 3.8, -3.2, 8388607.4, 8388607.6, -8388608.4, -8388608.6,
                         1e9, -1e9] + [0.0] * 18, dt)
        bf = BufferFormat(parse_sample_format("S24_4LE"), 1, 0)
        raw = np.zeros(32 * 4, np.uint8)
        of = _abi.OverflowC(0, 0, 0.0, float((1 << 23) - 1))
        cv.cbuf2raw(np.concatenate([q_in, np.zeros(32, dt)]), raw, bf, of)
        out[f"{t}_quant_in"] = q_in
        out[f"{t}_quant_raw"] = raw
        out[f"{t}_quant_overflow"] = np.array([of.n_overflows, of.intlargest, of.largest, of.max])
        # (a generator of their own, so that the vectors above keep their values)
        rng2 = np.random.default_rng(4343 + rs)
        # convolver_td_* (fftw_convolver.c:682-782): 31 taps, block length 32, three sliding [previous | current] frames
        td_taps = rng2.standard_normal(31).astype(dt)
        td = cv.td_new(td_taps)
        x = rng2.standard_normal(4 * 32).astype(dt)
        out[f"{t}_td_taps"], out[f"{t}_td_in"] = td_taps, x
        out[f"{t}_td_out"] = np.stack([cv.td_convolve(td, x[k * 32:(k + 2) * 32]) for k in range(3)])
        # convolver_cbuf2raw with dither (HP-TPDF + error feedback), S16_LE, channel 1 of 2, 40 blocks: the table
        # pointer wraps (2001 entries at 100 Hz) and block 7 clips
        cv.dither_init(2, 100)
        bf16 = BufferFormat(parse_sample_format("S16_LE"), 1, 0)
        of = _abi.OverflowC(0, 0, 0.0, 32767.0)
        d_in = rng2.standard_normal((40, 32)) * 3000.0
        d_in[7] *= 20.0
        d_in = d_in.astype(dt)
        raws = []
        for k in range(40):
            raw = np.zeros(32 * 2, np.uint8)
            cv.cbuf2raw_dither(np.concatenate([d_in[k], np.zeros(32, dt)]), raw, bf16, of, 1)
            raws.append(raw)
        out[f"{t}_dither_in"], out[f"{t}_dither_raw"] = d_in, np.stack(raws)
        out[f"{t}_dither_overflow"] = np.array([of.n_overflows, of.intlargest, of.largest, of.max])
    return out


def golden_blocks():
    """Whole-block vectors: small graphs through the reference's block sequence."""
    out = {}
    # (a) two filters, mixed inputs, delay, S24_LE packed in / S16_BE out
    L, P = 16, 4
    inb, nin = interleaved_layout(2, "S24_LE", L)
    outb, nout = interleaved_layout(2, "S16_BE", L)
    g = FilterGraph(L, P, 4, inb, outb, nin, nout,
                    [Filter([0], [0], coeff=0, delayblocks=1), Filter([1, 0], [1, 0], in_scales=[0.5, 0.25],
                                                                      out_scales=[1.0, 0.125], coeff=1)], [P, 2])
    rng = np.random.default_rng(777)
    taps = [rng.standard_normal(L * P).astype(np.float32) / 8, rng.standard_normal(2 * L - 5).astype(np.float32) / 8]
    x = np.round(rng.standard_normal((10, 2, L)) * 0.05 * (1 << 23))
    sig = np.stack([pack_block(x[b], inb, nin) for b in range(10)])
    d = po.BlockDriver("ref", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    out["a_sig"], out["a_taps0"], out["a_taps1"] = sig, taps[0], taps[1]
    out["a_out"] = d.run(sig)
    d.close()
    # (b) xtc topology with crossfade swaps every 5 blocks (float32 only)
    g = configs.config_c5(L=16, P=8)
    g.coeff_n_blocks = [8, 8]
    taps = configs.synthetic_filters(g, 5)
    sig = configs.synthetic_signal(g, 5, 16)
    d = po.BlockDriver("ref", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    outs = []
    for b in range(16):
        if b in (5, 10):
            swap = (b // 5) % 2
            for f in range(4):
                base = f % 2
                d.set_control(f, (base + swap) % 2)
        if b == 13:
            d.set_control(0, -1)
        outs.append(d.process_block(sig[b]))
    out["b_sig"], out["b_taps0"], out["b_taps1"] = sig, taps[0], taps[1]
    out["b_out"] = np.stack(outs)
    d.close()
    # (c) filter -> filter chaining (bench1_config's topology plus a third level and source multipliers)
    g = golden_graph_c()
    rng = np.random.default_rng(778)
    taps = [rng.standard_normal(g.filter_length * n).astype(np.float32) / 6 for n in g.coeff_n_blocks]
    x = np.round(rng.standard_normal((12, 2, g.filter_length)) * 0.05 * (1 << 23))
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(12)])
    d = po.BlockDriver("ref", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    out["c_sig"] = sig
    for c, h in enumerate(taps):
        out[f"c_taps{c}"] = h
    out["c_out"] = d.run(sig)
    d.close()
    # (d) HP-TPDF dither with error feedback on 16-bit outputs, through a table wrap and clipping blocks
    g = golden_graph_d()
    rng = np.random.default_rng(779)
    taps = [rng.standard_normal(g.filter_length * 2).astype(np.float32) / 6 for _ in range(2)]
    x = np.round(rng.standard_normal((36, 2, g.filter_length)) * 0.1 * (1 << 23))
    x[5:8] *= 40
    x = np.clip(x, -(1 << 23), (1 << 23) - 1)
    sig = np.stack([pack_block(x[b], g.in_formats, g.in_bytes) for b in range(36)])
    d = po.BlockDriver("ref", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    out["d_sig"], out["d_taps0"], out["d_taps1"] = sig, taps[0], taps[1]
    out["d_out"] = d.run(sig)
    out["d_overflow"] = np.array([[d.overflow(o).n_overflows, d.overflow(o).intlargest, d.overflow(o).largest]
                                  for o in range(2)])
    d.close()
    return out


def xtc_taps():
    res = {}
    for name in ("directpath", "crosspath"):
        with open(os.path.join(REF, name + ".txt")) as f:
            res[name] = np.array([float(s) for s in f.read().split()], np.float64)
    return res


if __name__ == "__main__":
    po.build()
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **golden_functions())
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **golden_blocks())
    np.savez_compressed(os.path.join(HERE, "xtc_taps.npz"), **xtc_taps())
    print("golden fixtures written to", HERE)
