"""Short C3-shaped run for ncu: full headline shape, a handful of blocks.  Usage: python tools/prof_c3.py [blocks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = configs.config_c3()
rng = np.random.default_rng(0)
with Engine(g) as e:
    env = np.exp(-np.arange(g.taps_per_filter(), dtype=np.float32) / (g.taps_per_filter() / 4.0))
    for c in range(64):
        e.coeff_from_taps(c, rng.standard_normal(g.taps_per_filter(), dtype=np.float32) * env * 1e-2)
    sig = configs.synthetic_signal(g, 3, 1)
    e.upload_input(sig[0])
    for _ in range(nb):
        e.process_block_device()
    e.synchronize()
print("done")
