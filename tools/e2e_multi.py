"""Diagnose the host-buffer (e2e) step time of one rank when several ranks run at once.
Launch with torchrun; env: PIN=1 pins each rank to its own pair of cores, ISOLATE=1 hides the other GPUs from the
process (CUDA_VISIBLE_DEVICES), NODIST=1 skips torch.distributed / NCCL initialisation."""
import os, sys, time
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("PIN") == "1":
    cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // world)
    os.sched_setaffinity(0, set(cores[rank * per:(rank + 1) * per]))
dev = rank
if os.environ.get("ISOLATE") == "1":
    os.environ["CUDA_VISIBLE_DEVICES"] = str(rank); dev = 0
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine, PinnedBuffer
from brutefir_b200.sharding import shard_graph
if os.environ.get("TORCHCUDA") == "1":
    import torch
    torch.cuda.set_device(dev)
    torch.zeros(1).cuda()
if os.environ.get("GLOO") == "1":
    import torch.distributed as dist
    dist.init_process_group("gloo")
elif os.environ.get("NODIST") != "1":
    import torch, torch.distributed as dist
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
B = 8
g = configs.config_c3()
sh = shard_graph(g, world, compact=world > 1)[rank]
sub = sh.graph
rng = np.random.default_rng(0)
with Engine(sub, device=dev, max_batch=B) as e:
    h = (rng.standard_normal(g.taps_per_filter()) * 1e-3).astype(np.float32)
    for c in sorted({f.coeff for f in sub.filters}):
        e.coeff_from_taps(c, h)
    sig = configs.synthetic_signal(g, 3, B)
    if world > 1:
        sig = sh.slice_input(g, sig)
    pin_in = [PinnedBuffer(B * sub.in_bytes) for _ in range(3)]
    pin_out = [PinnedBuffer(B * sub.out_bytes) for _ in range(3)]
    for p in pin_in:
        p.array[:] = sig.reshape(-1)
    e.upload_inputs(sig)
    for mode in ("device", "async"):
        for i in range(30):
            e.process_blocks_device(B) if mode == "device" else e.process_blocks_async(pin_in[i % 3].array, pin_out[i % 3].array, B)
        e.synchronize()
        if os.environ.get("NODIST") != "1":
            dist.barrier()
        K = 300
        e.timer_start(); t0 = time.perf_counter()
        for i in range(K):
            e.process_blocks_device(B) if mode == "device" else e.process_blocks_async(pin_in[i % 3].array, pin_out[i % 3].array, B)
        t1 = time.perf_counter(); ms = e.timer_stop()
        print(f"rank {rank}/{world} PIN={os.environ.get('PIN','0')} ISOLATE={os.environ.get('ISOLATE','0')} NODIST={os.environ.get('NODIST','0')} TORCHCUDA={os.environ.get('TORCHCUDA','0')} GLOO={os.environ.get('GLOO','0')} "
              f"{mode:6s}: host {1e6*(t1-t0)/K:.0f} us/call, device {1e3*ms/K:.0f} us/step", flush=True)
if os.environ.get("NODIST") != "1":
    dist.destroy_process_group()
