"""Worker of tests/test_host_logic.py::test_world_size_2_gloo (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from brutefir_b200 import configs  # noqa: E402
from brutefir_b200.sharding import shard_graph  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = configs.config_c3(n_ch=6, L=32, P=4)
    taps = configs.synthetic_filters(g, 3)
    sig = configs.synthetic_signal(g, 3, 8)
    shard = shard_graph(g, world, compact=True)[rank]   # each rank moves only its own channels' bytes
    d = po.BlockDriver("oracle", shard.graph)
    for c in shard.coeffs:
        d.coeff_from_taps(c, taps[c])
    out = np.zeros((sig.shape[0], g.out_bytes), np.uint8)
    shard.scatter_output(g, d.run(shard.slice_input(g, sig)), out)     # only this rank's channels are non-zero
    d.close()
    gathered = [torch.zeros(out.shape, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        merged = np.zeros_like(out)
        for part in gathered:
            merged |= part.numpy()
        full = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            full.coeff_from_taps(c, h)
        ref = full.run(sig)
        full.close()
        assert np.array_equal(merged, ref), "sharded result differs from the unsharded graph"
        assert t.item() == float(world)
        print("DIST_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
