"""Host-side logic: sample formats, raw block layouts, graph lowering, sharding (single process and a
world_size-2 gloo run of the rank-local logic)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from brutefir_b200 import configs
from brutefir_b200.formats import (interleaved_layout, pack_block, parse_sample_format, planar_layout,
                                   unpack_block)
from brutefir_b200.graph import Filter, FilterGraph
from brutefir_b200.sharding import assign_filters, filter_groups, shard_graph
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fmt", ["S8", "S16_LE", "S16_BE", "S24_LE", "S24_BE", "S24_4LE", "S24_4BE", "S32_LE",
                                 "S32_BE", "FLOAT_LE", "FLOAT_BE", "FLOAT64_LE", "FLOAT64_BE", "s24_3le", "S24_NE"])
def test_sample_format_table_and_round_trip(fmt):
    sf = parse_sample_format(fmt)
    assert sf.scale == (1.0 if sf.isfloat else 2.0 ** -(8 * sf.sbytes - 1))     # bfconf.c:473-477
    assert sf.swap == fmt.upper().endswith("BE") or sf.bytes == 1
    for layout in (interleaved_layout, planar_layout):
        bfs, nb = layout(3, sf, 16)
        assert nb % 32 == 0                                                    # dai.c:571-573
        rng = np.random.default_rng(1)
        v = rng.standard_normal((3, 16)) if sf.isfloat else \
            rng.integers(-(1 << (sf.bits - 1)), 1 << (sf.bits - 1), (3, 16)).astype(np.float64)
        back = unpack_block(pack_block(v, bfs, nb), bfs, 16)
        if sf.isfloat and sf.bytes == 4:
            v = v.astype(np.float32).astype(np.float64)
        assert np.array_equal(back, v)


def test_ambiguous_reference_formats_are_refused():
    for fmt in ("FLOAT_NE", "FLOAT64_NE", "U8"):
        with pytest.raises(ValueError):
            parse_sample_format(fmt)


def test_graph_lowering_matches_c_struct():
    g = configs.config_c5()
    cfg, keep = g.to_config(device=3, flags=1, mac_split=2)
    assert (cfg.filter_length, cfg.n_blocks, cfg.realsize, cfg.n_filters, cfg.n_coeffs) == (64, 64, 4, 4, 2)
    assert cfg.formats[0][1].byte_offset == 3 and cfg.formats[0][1].sample_spacing == 2
    assert cfg.filters[1].channels[0][0] == 1 and cfg.filters[1].channels[1][0] == 0 and cfg.filters[1].crossfade == 1
    assert cfg.device == 3 and cfg.flags == 1 and cfg.mac_split == 2
    with pytest.raises(ValueError):
        FilterGraph(48, 2, 4, g.in_formats, g.out_formats, g.in_bytes, g.out_bytes, [], []).validate()


def test_filter_groups_follow_the_same_output_rule():
    g = configs.config_c5()     # filters 0,1 -> out 0; filters 2,3 -> out 1
    assert filter_groups(g) == [[0, 1], [2, 3]]
    assert assign_filters(g, 2) == [0, 0, 1, 1]
    assert assign_filters(g, 2, split_outputs=True) == [0, 1, 0, 1]
    d = configs.diagonal_graph(8, 64, 2)
    assert len(filter_groups(d)) == 8 and assign_filters(d, 4) == [0, 1, 2, 3, 0, 1, 2, 3]
    chained = configs.config_c1_chained()
    assert filter_groups(chained) == [[0, 3, 4], [1, 2, 5]]


def test_shards_cover_the_graph_once():
    g = configs.config_c3(n_ch=16, L=64, P=4)
    shards = shard_graph(g, 4)
    assert sorted(f for s in shards for f in s.filters) == list(range(16))
    for s in shards:
        assert len(s.filters) == 4 and s.shared_outputs == []
        for lf, gf in zip(s.graph.filters, s.filters):
            assert s.inputs[lf.inputs[0]] == g.filters[gf].inputs[0]
            assert s.graph.in_formats[lf.inputs[0]].byte_offset == g.in_formats[g.filters[gf].inputs[0]].byte_offset
    split = shard_graph(configs.config_c5(), 2, split_outputs=True)
    assert [s.shared_outputs for s in split] == [[0, 1], [0, 1]]


@pytest.mark.parametrize("n_ranks", [2, 3, 4])
def test_split_outputs_every_rank_joins_every_collective(oracle_libs, n_ranks):
    """Outputs whose feeders sit on SOME ranks only: every rank must still list every shared output, in the same
    (global) order, because the engine matches its all-reduces across ranks by position.  The cross-rank sum is
    emulated by adding the ranks' float64 outputs (what ncclAllReduce does to the time-domain rows)."""
    L, P = 32, 3
    inb, nin = interleaved_layout(4, "FLOAT64_LE", L)
    outb, nout = interleaved_layout(3, "FLOAT64_LE", L)
    filters = [Filter([0], [0], coeff=0), Filter([1], [0], coeff=1), Filter([2], [1], coeff=2), Filter([3], [1], coeff=3),
               Filter([0], [2], coeff=1)]
    g = FilterGraph(L, P, 8, inb, outb, nin, nout, filters, [P] * 4)
    taps = configs.synthetic_filters(g, 7)
    sig = configs.synthetic_signal(g, 7, 6)
    full = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        full.coeff_from_taps(c, h)
    want = np.stack([unpack_block(b, g.out_formats, L) for b in full.run(sig)])
    full.close()
    shards = shard_graph(g, n_ranks, split_outputs=True, compact=True)
    n_shared = {len(s.shared_outputs) for s in shards}
    assert len(n_shared) == 1                                       # same number of collectives on every rank
    for s in shards:
        assert [s.outputs[o] for o in s.shared_outputs] == [s0 for s0 in shards[0].outputs if s0 in
                                                            [shards[0].outputs[o] for o in shards[0].shared_outputs]]
    got = np.zeros_like(want)
    for s in shards:
        d = po.BlockDriver("oracle", s.graph)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        y = np.stack([unpack_block(b, s.graph.out_formats, L) for b in d.run(s.slice_input(g, sig))])
        d.close()
        for i, o in enumerate(s.outputs):
            got[:, o] += y[:, i]
    assert np.abs(got - want).max() <= 1e-12 and np.abs(want).max() > 1e-3


def test_sharding_refuses_cross_rank_chains_and_dither():
    chained = configs.config_c1_chained()
    with pytest.raises(ValueError):
        shard_graph(chained, 2, split_outputs=True)
    g = configs.config_c3(n_ch=4, L=64, P=2, fmt="S16_LE")
    g.apply_dither = [True] * 4
    with pytest.raises(ValueError):
        shard_graph(g, 2)
    assert shard_graph(g, 1)[0].graph.apply_dither == [True] * 4


def test_compact_shards_reassemble_to_the_whole_graph(oracle_libs):
    """compact=True: every rank gets an interleaved block of only its channels (one dai device per GPU); slicing
    the input, running each shard and scattering the outputs back must give the unsharded graph's bytes."""
    g = configs.config_c3(n_ch=6, L=32, P=3)
    g.filters[4].outputs, g.filters[4].out_scales = [5], [0.5]      # two filters into output 5: they must stay together
    taps = configs.synthetic_filters(g, 9)
    sig = configs.synthetic_signal(g, 9, 7)
    full = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        full.coeff_from_taps(c, h)
    want = full.run(sig)
    full.close()
    got = np.zeros_like(want)
    shards = shard_graph(g, 3, compact=True)
    assert sum(s.graph.in_bytes for s in shards) <= g.in_bytes + 3 * 32
    for s in shards:
        assert s.graph.in_bytes < g.in_bytes and all(bf.sample_spacing == len(s.inputs) for bf in s.graph.in_formats)
        d = po.BlockDriver("oracle", s.graph)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        s.scatter_output(g, d.run(s.slice_input(g, sig)), got)
        d.close()
    assert np.array_equal(got, want)


def test_world_size_2_gloo(tmp_path):
    """The N > 1 host path on CPU: two ranks over gloo shard a graph, run their shards through the CPU
    oracle (stand-in for the per-rank engines) and rank 0 reassembles the output of the whole graph."""
    script = os.path.join(ROOT, "tests", "dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script, str(tmp_path)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIST_OK" in r.stdout


@pytest.mark.parametrize("count,world,want", [(8, 1, (0,)), (8, 2, (0, 4)), (8, 4, (0, 2, 4, 6)), (8, 8, tuple(range(8))),
                                              (2, 2, (0, 1)), (4, 2, (0, 2)), (6, 4, (0, 1, 2, 3)), (0, 2, (0, 1))])
def test_bench_spreads_ranks_over_the_visible_gpus(monkeypatch, count, world, want):
    """bench.py: a run on fewer ranks than the box shows GPUs takes every (count // world)-th device (the GPUs of an HGX
    box share PCIe switch uplinks in pairs, profiles/r2_copy_skew_n8.txt); otherwise device = LOCAL_RANK."""
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    monkeypatch.setattr(torch.cuda, "device_count", lambda: count)
    monkeypatch.delenv("BENCH_NO_SPREAD", raising=False)
    got = tuple(bench.spread_device(r, world)[0] for r in range(max(world, len(want))))[:len(want)]
    assert got == want
    monkeypatch.setenv("BENCH_NO_SPREAD", "1")
    assert all(bench.spread_device(r, world)[0] == r for r in range(world))
