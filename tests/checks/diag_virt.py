import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import unpack_run
L, P = 256, 4
inb, nin = interleaved_layout(3, "S24_4LE", L)
phys, nout = interleaved_layout(2, "S24_4LE", L)
outb = [phys[0], phys[0], phys[0], phys[1]]
filters = [Filter([0], [0], coeff=0), Filter([1], [1], coeff=1), Filter([2], [2], out_scales=[0.5], coeff=2),
           Filter([0, 1], [3], in_scales=[0.4, 0.4], coeff=1)]
g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, P], out_physical=[0, 0, 0, 1])
taps = configs.synthetic_filters(g, 37)
sig = configs.synthetic_signal(g, 37, 3, sigma=0.02)
with Engine(g) as e:
    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        e.coeff_from_taps(c, h); d.coeff_from_taps(c, h)
    for b in range(2):
        out = e.process_block(sig[b]); ref = d.process_block(sig[b])
        rows = [e.debug_read(_abi.DBG_OUTPUT_TIME, o)[:6] for o in range(4)]
        orow = [d.debug_read(_abi.DBG_OUTPUT_TIME, o)[:6] for o in range(4)]
        print("block", b)
        for o in range(4):
            print("  eng row", o, np.round(rows[o], 1), " oracle row", np.round(orow[o], 1))
        y = unpack_run(out[None], g.out_formats, L); r = unpack_run(ref[None], g.out_formats, L)
        print("  out eng", y[0, :6], y[3, :6]); print("  out ref", r[0, :6], r[3, :6])
    d.close()
