import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from brutefir_b200 import configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout
from brutefir_b200.graph import Filter, FilterGraph
from helpers import unpack_run

def run(g, taps, sig, B, split=1, chunk=None):
    with Engine(g, mac_split=split, max_batch=B) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        out = np.zeros((sig.shape[0], g.out_bytes), np.uint8)
        b = 0
        while b < sig.shape[0]:
            n = min(chunk or B, sig.shape[0] - b)
            e.process_blocks_async(sig[b:b+n], out[b:b+n], n); b += n
        e.synchronize()
    return out

def cmp(name, g, nb=13, B=2, split=1, chunk=None):
    taps = configs.synthetic_filters(g, 18); sig = configs.synthetic_signal(g, 18, nb)
    a = run(g, taps, sig, 1, split); b = run(g, taps, sig, B, split, chunk)
    L = g.filter_length
    ya, yb = unpack_run(a, g.out_formats, L), unpack_run(b, g.out_formats, L)
    d = np.abs(ya - yb).reshape(len(g.out_formats), nb, L).max(axis=2)
    print(name, "B", B, "chunk", chunk, "max diff per (out, block):\n", d.astype(int), flush=True)

L, P = 256, 10
cmp("diag", configs.diagonal_graph(2, L, P, 4, "S24_4LE"), B=2)
cmp("diag chunk1 in B2 engine", configs.diagonal_graph(2, L, P, 4, "S24_4LE"), B=2, chunk=1)
cmp("diag", configs.diagonal_graph(2, L, P, 4, "S24_4LE"), B=4)
inb, nin = interleaved_layout(3, "S24_4LE", L); outb, nout = interleaved_layout(3, "S16_LE", L)
cmp("delay", FilterGraph(L, P, 4, inb, outb, nin, nout, [Filter([0],[0],coeff=0, delayblocks=2), Filter([1],[1],coeff=0,delayblocks=9), Filter([2],[2],coeff=-1)], [P]), B=2)
cmp("mix", FilterGraph(L, P, 4, inb, outb, nin, nout, [Filter([1,2],[1],in_scales=[0.7,-0.2],coeff=0), Filter([2],[1,2],out_scales=[0.3,2.0],coeff=-1)], [3]), B=2)

print("---- scripted scenario")
filters = [Filter([0], [0], coeff=0, crossfade=True), Filter([1, 2], [1], in_scales=[0.7, -0.2], coeff=1, delayblocks=2),
           Filter([2], [1, 2], out_scales=[1.0 / 3.0, 2.0], coeff=-1), Filter([0], [2], coeff=2, delayblocks=9)]
g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, 3, 7])
taps = configs.synthetic_filters(g, 18)
nb = 37
script = {5: [(0, dict(coeff=2))], 6: [(1, dict(coeff=1, delayblocks=0, in_scales=[0.1, 0.3]))],
          16: [(0, dict(coeff=-1)), (3, dict(coeff=0))], 29: [(0, dict(coeff=1))]}
for sigma, scale, use_script in ((0.3, 40.0, True), (0.02, 1.0, True), (0.02, 1.0, False), (0.3, 40.0, False)):
    sig = configs.synthetic_signal(g, 18, nb, sigma=sigma)
    def run2(max_batch, split=1):
        with Engine(g, mac_split=split, max_batch=max_batch) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h, scale)
            out = np.zeros((nb, g.out_bytes), np.uint8)
            b = 0
            while b < nb:
                if use_script:
                    for filt, kw in script.get(b, []):
                        e.set_control(filt, **kw)
                n = 1
                while n < max_batch and b + n < nb and (b + n) not in script:
                    n += 1
                e.process_blocks_async(sig[b:b + n], out[b:b + n], n)
                b += n
            e.synchronize()
        return out
    a, b2 = run2(1), run2(2)
    d = np.abs(unpack_run(a, g.out_formats, L) - unpack_run(b2, g.out_formats, L)).reshape(3, nb, L).max(axis=2)
    print("sigma", sigma, "scale", scale, "script", use_script, "\n", d.astype(int), flush=True)
