"""bench1_config's graph (BASELINE configs[0]: 2 in / 2 out, 6 filters of 8192 x 8, four feeding two through to_filters)
on the device: realtime multiple block by block and batched."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine

g = configs.config_c1_chained()
taps = [t * 0.7 for t in configs.synthetic_filters(g, 11)]
for B in (1, 8):
    with Engine(g, max_batch=B) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        e.upload_inputs(configs.synthetic_signal(g, 1, B))
        for _ in range(20):
            e.process_blocks_device(B)
        e.synchronize()
        K = 200
        e.timer_start()
        for _ in range(K):
            e.process_blocks_device(B)
        ms = e.timer_stop()
        per = ms / (K * B)
        print(f"bench1_config graph, {B} block(s) per call: {per * 1000:.1f} us/block, x{(8192 / 44100) / (per * 1e-3):.0f} realtime")
