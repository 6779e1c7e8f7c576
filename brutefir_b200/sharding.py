"""Filter sharding over GPUs: the reference's "one filter process per CPU" rule applied to devices.

``load_balance_filters`` (/root/reference/bfconf.c:2227-2318) groups filters that are connected
(``to_filters``/``from_filters``) or that mix into the same output -- "mixing to an output channel or a
filter input must be done within the same process" (bfconf.c:2893-2931) -- and deals the groups round
robin over the CPUs.  Here a "process" is one rank = one GPU with its own engine: it holds the
coefficient spectra and delay lines of its filters only (HBM is sized per shard), reads just the input
channels those filters use and writes just the output channels they feed.  Diagonal workloads need no
exchange at all.  When the caller insists on splitting the filters of one output over ranks
(``split_outputs=True``, BASELINE config 5), that output is marked shared and its L time-domain samples
are summed over NVLink by the engine (ncclAllReduce) before quantisation.
"""
from __future__ import annotations

import copy
import dataclasses
from typing import Dict, List

import numpy as np

from .formats import interleaved_layout
from .graph import Filter, FilterGraph


def filter_groups(graph: FilterGraph) -> List[List[int]]:
    """Connected components under "shares an output" / "is chained to" (bfconf.c:2234-2298)."""
    n = len(graph.filters)
    group = [-1] * n
    groups: List[List[int]] = []
    for start in range(n):
        if group[start] != -1:
            continue
        gid = len(groups)
        group[start] = gid
        changed = True
        while changed:
            changed = False
            outs = {o for i in range(n) if group[i] == gid for o in graph.filters[i].outputs}
            for i in range(n):
                fi = graph.filters[i]
                if group[i] == gid:
                    for k in fi.from_filters:
                        if group[k] != gid:
                            group[k] = gid
                            changed = True
                    continue
                linked = any(o in outs for o in fi.outputs) or any(group[k] == gid for k in fi.from_filters)
                if linked:
                    group[i] = gid
                    changed = True
        groups.append([i for i in range(n) if group[i] == gid])
    return groups


def assign_filters(graph: FilterGraph, n_ranks: int, split_outputs: bool = False) -> List[int]:
    """rank of every filter.  Default: whole groups round robin (bfconf.c:2300-2316).  With
    ``split_outputs`` filters are dealt round robin individually, ignoring the same-output rule."""
    owner = [0] * len(graph.filters)
    if split_outputs:
        for f in range(len(graph.filters)):
            owner[f] = f % n_ranks
        return owner
    for g, members in enumerate(filter_groups(graph)):
        for f in members:
            owner[f] = g % n_ranks
    return owner


def _compact_layout(bfs, fragsize):
    """Interleave the given channels (all of one sample format) into a block of their own, like
    calc_buffer_format does for one device (dai.c:537-576)."""
    if not bfs:
        return [], 32
    name = bfs[0].sf.name
    if any(bf.sf.name != name for bf in bfs):
        raise ValueError("compact sharding needs one sample format per direction")
    return interleaved_layout(len(bfs), name, fragsize)


@dataclasses.dataclass
class Shard:
    rank: int
    graph: FilterGraph              # this rank's sub-graph (channel numbering is rank-local)
    filters: List[int]              # global filter indices, in local order
    inputs: List[int]               # global input channels, in local order
    outputs: List[int]              # global output channels, in local order
    coeffs: List[int]               # global coefficient sets, in local order
    shared_outputs: List[int]       # LOCAL output indices that other ranks also feed

    def slice_input(self, full_graph: FilterGraph, raw_blocks: np.ndarray) -> np.ndarray:
        """raw_blocks[b, full in_bytes] in the full graph's layout -> this rank's blocks [b, in_bytes]."""
        out = np.zeros((raw_blocks.shape[0], self.graph.in_bytes), np.uint8)
        for i, c in enumerate(self.inputs):
            src = _byte_index(full_graph.in_formats[c], full_graph.filter_length)
            dst = _byte_index(self.graph.in_formats[i], full_graph.filter_length)
            out[:, dst] = raw_blocks[:, src]
        return out

    def scatter_output(self, full_graph: FilterGraph, local_blocks: np.ndarray, full_blocks: np.ndarray) -> None:
        """Write this rank's output blocks [b, out_bytes] into full_blocks[b, full out_bytes] (in place)."""
        for i, c in enumerate(self.outputs):
            src = _byte_index(self.graph.out_formats[i], full_graph.filter_length)
            dst = _byte_index(full_graph.out_formats[c], full_graph.filter_length)
            full_blocks[:, dst] = local_blocks[:, src]


def _byte_index(bf, fragsize):
    """Byte offsets of every sample byte of one channel inside a raw block (dai.c:537-576)."""
    b = bf.sf.bytes
    return (bf.byte_offset + np.arange(fragsize)[:, None] * (bf.sample_spacing * b) + np.arange(b)[None, :]).reshape(-1)


def shard_graph(graph: FilterGraph, n_ranks: int, split_outputs: bool = False, compact: bool = False) -> List[Shard]:
    """One sub-graph per rank.  By default the buffer formats keep their byte offsets and spacings, i.e. every
    rank addresses its channels inside the ORIGINAL raw block layout (the dai buffers, dai.c:537-576): the host
    hands each rank the same block layout and no repacking is needed -- but then every rank moves the whole
    block over its PCIe link.  With ``compact`` each rank gets its own interleaved block holding only its
    channels (the host fans the input slices out and gathers the output slices; in BruteFIR
    terms: one dai device per GPU): ``Shard.slice_input`` / ``Shard.scatter_output`` do that repacking."""
    owner = assign_filters(graph, n_ranks, split_outputs)
    feeders: Dict[int, set] = {}
    for f, flt in enumerate(graph.filters):
        for o in flt.outputs:
            feeders.setdefault(o, set()).add(owner[f])
        for k in flt.from_filters:
            if owner[k] != owner[f]:
                # the evaluated output of a source filter never leaves its rank (bfrun.c:1603-1660 runs inside one
                # filter process; bfconf.c:2234-2298 keeps chained filters together)
                raise ValueError(f"filter {f} is chained to filter {k} on another rank: chained filters cannot be split")
    # Outputs fed from several ranks, in GLOBAL channel order.  The engine issues one all-reduce per entry, matched
    # across ranks by position only (bf_engine.cu), so every rank carries every shared output in the same order -- a
    # rank without a feeder contributes a row of zeros -- and all ranks issue the same sequence of collectives.
    shared_global = sorted(o for o, r in feeders.items() if len(r) > 1)
    dithered = graph.apply_dither is not None and any(graph.apply_dither)
    if dithered and n_ranks > 1:
        # A channel's offset into the dither table is its index among the dithered channels (dither.c:75-96); a shard
        # renumbers its channels, so its sequence would differ from the unsharded (and the reference's) one.  And the
        # error feedback of a shared output would have to run after the cross-rank sum, on one rank.
        raise ValueError("sharding a graph with dither enabled is not supported (per-channel dither table offsets "
                         "follow the global channel numbering)")
    shards = []
    for r in range(n_ranks):
        mine = [f for f in range(len(graph.filters)) if owner[f] == r]
        ins = sorted({c for f in mine for c in graph.filters[f].inputs})
        outs = sorted({c for f in mine for c in graph.filters[f].outputs} | set(shared_global))
        # every coefficient set stays addressable by its global index at run time (cfc), so keep them all
        coeffs = list(range(len(graph.coeff_n_blocks)))
        imap = {c: i for i, c in enumerate(ins)}
        omap = {c: i for i, c in enumerate(outs)}
        fmap = {f: i for i, f in enumerate(mine)}
        local = []
        for f in mine:
            flt = copy.deepcopy(graph.filters[f])
            flt.inputs = [imap[c] for c in flt.inputs]
            flt.outputs = [omap[c] for c in flt.outputs]
            flt.from_filters = [fmap[k] for k in flt.from_filters]
            local.append(flt)
        in_fmts, out_fmts = [graph.in_formats[c] for c in ins], [graph.out_formats[c] for c in outs]
        in_bytes, out_bytes = graph.in_bytes, graph.out_bytes
        if compact:
            in_fmts, in_bytes = _compact_layout(in_fmts, graph.filter_length)
            out_fmts, out_bytes = _compact_layout(out_fmts, graph.filter_length)
        sub = FilterGraph(graph.filter_length, graph.n_blocks, graph.realsize, in_fmts, out_fmts,
                          in_bytes, out_bytes, local, list(graph.coeff_n_blocks),
                          safety_limit=graph.safety_limit, sampling_rate=graph.sampling_rate,
                          apply_dither=None if graph.apply_dither is None else [graph.apply_dither[o] for o in outs],
                          max_dither_table_size=graph.max_dither_table_size,
                          powersave=graph.powersave, analog_powersave=graph.analog_powersave)
        shared = [omap[o] for o in shared_global]
        shards.append(Shard(r, sub, mine, ins, outs, coeffs, shared))
    return shards
