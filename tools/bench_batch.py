"""Device-resident timing of the headline shape for several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import configs, _abi
from brutefir_b200.engine import Engine

F = int(os.environ.get("FILTERS", "64"))       # a rank's share of the 64-filter job on 64 / F GPUs
g = configs.config_c3() if F == 64 else configs.diagonal_graph(F, 8192, 128, 4, "S24_4LE")
rng = np.random.default_rng(0)
taps_n = g.taps_per_filter()
env = np.exp(-np.arange(taps_n, dtype=np.float32) / (taps_n / 4.0))
H = [rng.standard_normal(taps_n, dtype=np.float32) * env * 1e-2 for _ in range(4)]
for B in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    with Engine(g, flags=_abi.FLAG_STAGE_TIMING, max_batch=B) as e:
        for c in range(F):
            e.coeff_from_taps(c, H[c % 4])
        sig = configs.synthetic_signal(g, 3, B)
        e.upload_inputs(sig)
        for _ in range(140 // B + 2):
            e.process_blocks_device(B)
        e.synchronize(); e.stage_times()
        K = 240 // B
        e.timer_start()
        for _ in range(K):
            e.process_blocks_device(B)
        ms = e.timer_stop()
        st, nb, nl = e.stage_times()
        info = e.info()
        per = ms / (K * B)
        print(f"B={B}: {per*1000:.1f} us/block  RT x{(8192/48000)/(per*1e-3):.0f}  stages/block(ms) fwd {st[0]:.4f} mac {st[1]:.4f} inv {st[2]:.4f}"
              f"  mac batch GB/s {info.mac_bytes_per_batch/(st[1]*B*1e-3)/1e9:.0f} (compulsory)  {info.mac_bytes_per_block/(st[1]*1e-3)/1e9:.0f} (per-block formula)", flush=True)
