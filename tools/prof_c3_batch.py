"""Short C3-shaped batched run for ncu.  Usage: python tools/prof_c3_batch.py [B] [calls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = configs.config_c3()
rng = np.random.default_rng(0)
with Engine(g, max_batch=B) as e:
    env = np.exp(-np.arange(g.taps_per_filter(), dtype=np.float32) / (g.taps_per_filter() / 4.0))
    h = [rng.standard_normal(g.taps_per_filter(), dtype=np.float32) * env * 1e-2 for _ in range(4)]
    for c in range(64):
        e.coeff_from_taps(c, h[c % 4])
    sig = configs.synthetic_signal(g, 3, B)
    e.upload_inputs(sig)
    for _ in range(calls):
        e.process_blocks_device(B)
    e.synchronize()
print("done")
