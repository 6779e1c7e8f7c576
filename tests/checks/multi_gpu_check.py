"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
  (1) diagonal graph sharded by filter group: no collective, the union of the ranks' outputs == the oracle;
  (2) xtc topology (BASELINE config 5) with the two filters of every output deliberately placed on different
      ranks: the time-domain blocks of the shared outputs are summed over NVLink (ncclAllReduce inside the
      engine) before quantisation, with a crossfaded coefficient swap on the way."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from brutefir_b200.sharding import shard_graph
from oracle import pyoracle as po
from helpers import unpack_run


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def gather_or(out):
        t = torch.from_numpy(out.astype(np.int32)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)      # disjoint channels: sum == or
        return t.cpu().numpy().astype(np.uint8)

    # (1) diagonal, no collective
    g = configs.config_c3(n_ch=8, L=1024, P=8)
    taps = configs.synthetic_filters(g, 3)
    sig = configs.synthetic_signal(g, 3, 14, sigma=0.02)
    sh = shard_graph(g, world)[rank]
    with Engine(sh.graph, device=local) as e:
        for c in sorted({f.coeff for f in sh.graph.filters}):
            e.coeff_from_taps(c, taps[c])
        merged = gather_or(e.run(sig))
    if rank == 0:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        ref = d.run(sig); d.close()
        diff = np.abs(unpack_run(merged, g.out_formats, 1024) - unpack_run(ref, g.out_formats, 1024)).max()
        print(f"[1] diagonal over {world} ranks: max |gpu - oracle| = {diff} LSB", flush=True)
        assert diff <= 1

    # (2) xtc with split outputs -> NCCL sum
    g = configs.config_c5(L=64, P=64)
    taps = configs.synthetic_filters(g, 5)
    sig = configs.synthetic_signal(g, 5, 80, sigma=0.02)
    sh = shard_graph(g, world, split_outputs=True)[rank]
    uid = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    outs = []
    with Engine(sh.graph, device=local) as e:
        e.comm_init(rank, world, uid[0])
        e.comm_shared_outputs(sh.shared_outputs)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        for b in range(80):
            if b in (20, 50):
                for lf, gf in enumerate(sh.filters):
                    e.set_control(lf, (gf % 2 + (1 if b == 20 else 0)) % 2)
            outs.append(e.process_block(sig[b]))
    got = np.stack(outs)        # every rank holds the full sum of the shared outputs
    if rank == 0:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        ref = []
        for b in range(80):
            if b in (20, 50):
                for f in range(4):
                    d.set_control(f, (f % 2 + (1 if b == 20 else 0)) % 2)
            ref.append(d.process_block(sig[b]))
        d.close()
        diff = np.abs(unpack_run(got, g.out_formats, 64) - unpack_run(np.stack(ref), g.out_formats, 64)).max()
        print(f"[2] xtc split over {world} ranks, NCCL output sum + crossfade: max |gpu - oracle| = {diff} LSB", flush=True)
        assert diff <= 1
        print("MULTI_GPU_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
