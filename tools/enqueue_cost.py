"""Host cost of one enqueue (bfcuda_process_blocks_async / _device) versus the device time of the step, on a small
shard (8 filters of the headline shape = one rank of the 8-GPU run).  Tells whether a rank is launch-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine, PinnedBuffer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = configs.config_c3(n_ch=nch)
rng = np.random.default_rng(0)
with Engine(g, max_batch=B) as e:
    env = np.exp(-np.arange(g.taps_per_filter(), dtype=np.float32) / (g.taps_per_filter() / 4.0))
    h = rng.standard_normal(g.taps_per_filter(), dtype=np.float32) * env * 1e-2
    for c in range(nch):
        e.coeff_from_taps(c, h)
    sig = configs.synthetic_signal(g, 3, B)
    pin_in = [PinnedBuffer(B * g.in_bytes) for _ in range(3)]
    pin_out = [PinnedBuffer(B * g.out_bytes) for _ in range(3)]
    for p in pin_in:
        p.array[:] = sig.reshape(-1)
    e.upload_inputs(sig)
    for mode in ("device", "async"):
        for i in range(50):
            e.process_blocks_device(B) if mode == "device" else e.process_blocks_async(pin_in[i % 3].array, pin_out[i % 3].array, B)
        e.synchronize()
        K = 400
        e.timer_start()
        t0 = time.perf_counter()
        for i in range(K):
            if mode == "device":
                e.process_blocks_device(B)
            else:
                e.process_blocks_async(pin_in[i % 3].array, pin_out[i % 3].array, B)
        t1 = time.perf_counter()
        ms = e.timer_stop()
        print(f"{mode:6s} B={B} filters={nch}: host enqueue {1e6 * (t1 - t0) / K:.1f} us/call, device {1e3 * ms / K:.1f} us/step")
