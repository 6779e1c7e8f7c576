"""GPU vs oracle vs float64 truth: who is how far from what, per config (LSB @24 bit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from oracle import pyoracle as po
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import unpack_run

def truth(x, h):
    n = len(x) + len(h) - 1
    nfft = 1 << int(np.ceil(np.log2(n)))
    return np.fft.irfft(np.fft.rfft(x, nfft) * np.fft.rfft(h, nfft), nfft)[: len(x)]

for (L, P, sigma) in [(8192, 4, 0.1), (8192, 4, 0.01), (1024, 8, 0.1), (8192, 16, 0.1), (256, 64, 0.1), (8192, 128, 0.1), (8192, 128, 0.01)]:
    g = configs.diagonal_graph(2, L, P, 4, "S24_4LE")
    taps = configs.synthetic_filters(g, 11)
    nb = min(P + 6, 40)
    sig = configs.synthetic_signal(g, 11, nb, sigma=sigma)
    with Engine(g, mac_split=1) as e:
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h); d.coeff_from_taps(c, h)
        got = unpack_run(e.run(sig), g.out_formats, L); ref = unpack_run(d.run(sig), g.out_formats, L)
        d.close()
    x = unpack_run(sig, g.in_formats, L)
    tr = np.stack([truth(x[c], taps[c].astype(np.float64)) for c in range(2)])
    dg = np.abs(got - ref)
    print(f"L={L} P={P} sigma={sigma}: |gpu-ref| max {dg.max():.0f} (n>1: {(dg>1).sum()}, n==1: {(dg==1).sum()} of {dg.size}); "
          f"|gpu-truth| max {np.abs(got-tr).max():.3f} rms {np.sqrt(np.mean((got-tr)**2)):.3f}; |ref-truth| max {np.abs(ref-tr).max():.3f} rms {np.sqrt(np.mean((ref-tr)**2)):.3f}; peak {np.abs(tr).max():.3e}", flush=True)
