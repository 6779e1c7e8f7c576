"""Diagnostic (not collected): BASELINE config 4 at -20 dBFS, the engine's automatic partition split against the
reference's left-to-right sum -- how far apart are they, and how far is each from the float64 truth?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import test_gpu_parity_fullsize as t
from brutefir_b200 import configs
from helpers import unpack_run

g = configs.config_c4()
taps = t.fast_unit_energy_filters(g, 2004)
for sigma in (0.1, 0.03):
    sig = configs.synthetic_signal(g, 4, 1032, sigma=sigma)
    ref = t.reference_run(g, taps, sig)
    n_tail = 64 * g.filter_length
    chans = [0, 13, 31]
    truth = t.truth_tail(g, taps, sig, chans, n_tail)
    r = unpack_run(ref, g.out_formats, g.filter_length)
    er = np.concatenate([r[c, -n_tail:] - truth[c] for c in chans])
    print(f"sigma {sigma}: reference vs truth rms {np.sqrt(np.mean(er**2)):.3f} max {np.abs(er).max():.2f} LSB; peak {np.abs(r).max():.0f}")
    for B, split in ((1, 0), (8, 0), (1, 1), (1, 2), (1, 8)):
        got, info = t.engine_run(g, taps, sig, B, mac_split=split)
        y = unpack_run(got, g.out_formats, g.filter_length)
        d = np.abs(y - r)
        eg = np.concatenate([y[c, -n_tail:] - truth[c] for c in chans])
        print(f"  B {B} split {info.mac_split}: |gpu-ref| max {d.max():.0f} frac>1 {np.mean(d > 1):.2e} frac>2 {np.mean(d > 2):.2e}; "
              f"gpu vs truth rms {np.sqrt(np.mean(eg**2)):.3f} max {np.abs(eg).max():.2f}")
