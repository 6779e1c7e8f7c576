"""Host-buffer copies alone with every rank of a multi-GPU run copying at once: does the placement of the pinned
buffers (offset inside the allocation), the size of a transfer, or the number of ranks copying at the same time move the
per-rank rate?  (The 8-GPU e2e figure of bench.py sits on this ceiling.)  Launch with torchrun; rank 0 prints a table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from brutefir_b200 import configs
from brutefir_b200.engine import Engine, PinnedBuffer
from brutefir_b200.sharding import shard_graph

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dist.init_process_group("gloo")
B = 8
rows = []


def report(label, ms, h2d, d2h, active=True):
    t = torch.tensor([ms if active else 0.0, h2d if active else 0.0, d2h if active else 0.0, 1.0 if active else 0.0], dtype=torch.float64)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        n = int(sm[3].item())
        rows.append(f"{label:58s} ranks {n}: max {mx[0].item() * 1e3:7.1f} us per round, per-rank mean H2D {sm[1].item() / n:5.1f} "
                    f"D2H {sm[2].item() / n:5.1f} GB/s, aggregate {(sm[1].item() + sm[2].item()):6.1f} GB/s")


for fmt in ("S24_4LE", "S24_LE"):
    g = configs.config_c3(fmt=fmt, P=2)
    sub = shard_graph(g, world, compact=world > 1)[rank].graph
    nin, nout = B * sub.in_bytes, B * sub.out_bytes
    with Engine(sub, device=local, max_batch=B) as e:
        arena_in = PinnedBuffer(nin + (8 << 20), local)
        arena_out = PinnedBuffer(nout + (8 << 20), local)

        def run(label, off_in, off_out, n_blocks=B, active=True, reps=200):
            dist.barrier()
            ms = h = d = 0.0
            if active:
                a = arena_in.array[off_in:off_in + n_blocks * sub.in_bytes]
                b = arena_out.array[off_out:off_out + n_blocks * sub.out_bytes]
                ms, h, d = e.copy_baseline(a, b, n_blocks, reps)
            dist.barrier()
            report(f"{fmt} {label}", ms, h, d, active)

        run("offset 0 (as allocated)", 0, 0)
        run("offset 0 again", 0, 0)
        run("page-aligned offsets, different per rank", 4096 * (1 + 37 * rank), 4096 * (3 + 41 * rank))
        run("offset 1 MiB on every rank", 1 << 20, 1 << 20)
        run("odd offsets (64 B + 256 B x rank)", 64 + 256 * rank, 192 + 256 * rank)
        for nb in (1, 2, 4):
            run(f"{nb} block(s) per copy ({nb * sub.in_bytes >> 10} KiB)", 0, 0, n_blocks=nb)
        if world >= 4:
            run("even ranks only", 0, 0, active=rank % 2 == 0)
            run("first half of the ranks only", 0, 0, active=rank < world // 2)
            run("rank 0 alone", 0, 0, active=rank == 0)
        arena_in.free()
        arena_out.free()
if rank == 0:
    print("\n".join(rows), flush=True)
dist.destroy_process_group()
