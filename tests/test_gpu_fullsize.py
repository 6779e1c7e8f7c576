"""Size-independent properties at the FULL BASELINE.json sizes (the oracle would take minutes there):
  C3  64 ch x 1 048 576 taps, 8192 x 128:  shifted unit impulses as filters turn the convolution into an
      exact integer delay -> the output must equal the delayed input exactly; linearity; the MAC stage is
      spot-checked bit-exactly against the oracle's convolve_add on buffers read back from the device.
  C4  32 ch x 262 144 taps, 256 x 1024 (split partition sum): same delay property."""
import numpy as np
import pytest

from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu


def delay_property(g, cid, delays, nb_extra):
    L, P = g.filter_length, g.n_blocks
    n_ch = len(g.filters)
    nb = max(delays) // L + 2 + nb_extra
    sig = configs.synthetic_signal(g, cid, nb, sigma=0.02)
    with Engine(g) as e:
        for c in range(n_ch):
            h = np.zeros(L * P, np.float32)
            h[delays[c]] = 1.0
            e.coeff_from_taps(c, h)
        out = e.run(sig)
        info = e.info()
    x = unpack_run(sig, g.in_formats, L)
    y = unpack_run(out, g.out_formats, L)
    for c in range(n_ch):
        d = delays[c]
        assert np.array_equal(y[c, d:], x[c, : x.shape[1] - d]), c      # exact: integers survive float32 at this level
        assert np.all(y[c, :d] == 0)
    return info


def test_c3_headline_shape_delay_and_linearity(gpu_lib, oracle_libs):
    g = configs.config_c3()
    rng = np.random.default_rng(33)
    delays = [int(v) for v in rng.integers(0, 3 * 8192, 64)]
    delays[0], delays[1], delays[2] = 0, 8191, 8192
    info = delay_property(g, 3, delays, 2)
    assert info.mac_split == 1 and info.mac_bytes_per_block == 1077936128
    # linearity with real filters: y(a) + y(b) == y(a + b) up to float32 rounding, 1 LSB each
    g = configs.config_c3(n_ch=4)
    taps = configs.synthetic_filters(g, 3)
    a = configs.synthetic_signal(g, 3, 6, sigma=0.04)
    b = configs.synthetic_signal(g, 4, 6, sigma=0.04)
    from brutefir_b200.formats import pack_block
    xa, xb = (np.stack([np.stack([unpack_run(s[i:i + 1], g.in_formats, 8192)[c] for c in range(4)]) for i in range(6)])
              for s in (a, b))
    ab = np.stack([pack_block(xa[i] + xb[i], g.in_formats, g.in_bytes) for i in range(6)])
    ys = []
    for s in (a, b, ab):
        with Engine(g) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            ys.append(unpack_run(e.run(s), g.out_formats, 8192))
    assert np.abs(ys[0] + ys[1] - ys[2]).max() <= 2


def test_c3_mac_stage_bit_exact_spot_check(gpu_lib, oracle_libs):
    g = configs.config_c3()
    rng = np.random.default_rng(34)
    sig = configs.synthetic_signal(g, 3, 3)
    o = po.Convolver("oracle", 8192, 4)
    with Engine(g) as e:
        for c in (5, 63):
            e.coeff_from_taps(c, configs.synthetic_filters(configs.config_c3(n_ch=1), 30 + c)[0])
        for b in range(3):
            e.process_block(sig[b])
        for f in (5, 63):
            want = o.convolve(e.debug_read(_abi.DBG_DELAYLINE, f, 2), e.coeff_get_block(f, 0))
            for i in range(1, 128):
                o.convolve_add(e.debug_read(_abi.DBG_DELAYLINE, f, (2 - i) % 128), e.coeff_get_block(f, i), want)
            assert np.array_equal(e.debug_read(_abi.DBG_FILTER_OUTPUT, f), want)


def test_c4_low_latency_shape_delay(gpu_lib):
    g = configs.config_c4()
    rng = np.random.default_rng(44)
    delays = [int(v) for v in rng.integers(0, 6 * 256, 32)]
    info = delay_property(g, 4, delays, 4)
    assert info.mac_split > 1       # the 1024-deep delay line is split to fill the machine


def test_c2_stereo_against_oracle(gpu_lib, oracle_libs):
    g = configs.config_c2()
    taps = configs.synthetic_filters(g, 2)
    sig = configs.synthetic_signal(g, 2, 24)
    with Engine(g) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        y, r = unpack_run(e.run(sig), g.out_formats, 4096), unpack_run(d.run(sig), g.out_formats, 4096)
        d.close()
    assert np.abs(y - r).max() <= 1e-6      # float32 output, full scale = 1.0


def test_c1_bench1_config_graph_on_the_device(gpu_lib, oracle_libs):
    """BASELINE configs[0]: bench1_config as shipped (/root/reference/bench1_config) -- 8192 x 8 partitions, 2 in / 2 out
    S24_4LE, six "dirac pulse" filters, four of them feeding the other two through to_filters.  With unit pulses
    every path is the identity, so out0 = in0 + in1 and out1 = in0 + in1 EXACTLY (integer samples);
    with random coefficients the device follows the oracle's block sequence."""
    g = configs.config_c1_chained()
    L = g.filter_length
    quiet = configs.synthetic_signal(g, 1, 12, sigma=0.004)     # -48 dBFS: two float32 FFT round trips stay < 1/2 LSB
    x = unpack_run(quiet, g.in_formats, L)
    pulse = np.zeros(8 * L, np.float32)
    pulse[0] = 1.0
    with Engine(g) as e:
        for c in range(6):
            e.coeff_from_taps(c, pulse)
        y = unpack_run(e.run(quiet), g.out_formats, L)
    assert np.abs(x).max() > 5e4
    assert np.array_equal(y[0], x[0] + x[1]) and np.array_equal(y[1], x[0] + x[1])
    sig = configs.synthetic_signal(g, 1, 12, sigma=0.05)
    taps = [t * 0.7 for t in configs.synthetic_filters(g, 11)]
    with Engine(g) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref = unpack_run(e.run(sig), g.out_formats, L), unpack_run(d.run(sig), g.out_formats, L)
        d.close()
    assert np.abs(ref).max() > 1e4 and np.abs(got - ref).max() <= 2 and np.mean(np.abs(got - ref) > 1) < 1e-4


def test_c3_batched_pipeline_equals_block_by_block_at_full_size(gpu_lib):
    """The headline mode of bench.py (8 blocks per call, three software-pipelined stage streams) at the full
    64 x 1 048 576-tap size: byte-identical to the block-by-block engine, random filters on a few channels and shifted
    unit impulses elsewhere, 24 blocks through host buffers."""
    g = configs.config_c3()
    L, P = g.filter_length, g.n_blocks
    sig = configs.synthetic_signal(g, 3, 24, sigma=0.02)
    outs = []
    for B in (1, 8):
        with Engine(g, max_batch=B) as e:
            for c in range(64):
                if c % 16 == 5:
                    r = np.random.default_rng(35 + c)       # same taps for both engines
                    h = (r.standard_normal(L * P) * np.exp(-np.arange(L * P) / (L * P / 4.0)) * 3e-3).astype(np.float32)
                else:
                    h = np.zeros(L * P, np.float32)
                    h[(c * 977) % (L * 3)] = 1.0
                e.coeff_from_taps(c, h)
            out = np.zeros((24, g.out_bytes), np.uint8)
            for b0 in range(0, 24, B):
                e.process_blocks_async(sig[b0:b0 + B], out[b0:b0 + B], B)
            e.synchronize()
            outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(unpack_run(outs[0], g.out_formats, L)).max() > 1e4


def test_c3_packed_s24_io_equals_the_4_byte_layout_at_the_headline_partition_size(gpu_lib):
    """The headline shape's channels and partition size (64 x 8192) in massive_config's own sample format, packed S24_LE
    (massive_config:11,17), 8 blocks per call: the same integers as the S24_4LE layout, byte for byte after unpacking
    (the 4-byte layout is the one the full-size parity tests pin against the reference build)."""
    L, P, nb, B = 8192, 4, 16, 8
    outs = []
    for fmt in ("S24_4LE", "S24_LE"):
        g = configs.config_c3(fmt=fmt, P=P)
        taps = configs.synthetic_filters(g, 3)
        sig = configs.synthetic_signal(g, 3, nb, sigma=0.1)
        with Engine(g, max_batch=B) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            for b in range(0, nb, B):
                e.process_blocks_async(sig[b:b + B], out[b:b + B], B)
            e.synchronize()
        outs.append(unpack_run(out, g.out_formats, L))
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() > 2 ** 20
