"""Soak test of the software-pipelined engine: thousands of asynchronous calls with random run-time control changes
(coefficients with crossfade, delays, scales), pipelined stages versus BFCUDA_FLAG_SERIAL_STAGES -- the outputs must be
byte-identical, i.e. no stage ever reads a buffer a neighbouring launch is still writing.
Usage: python tests/checks/soak_pipeline.py [calls] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout
from brutefir_b200.graph import Filter, FilterGraph

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L, P, n = 1024, 12, 6
inb, nin = interleaved_layout(n, "S24_4LE", L)
outb, nout = interleaved_layout(n, "S24_4LE", L)
filters = [Filter([i % 3], [o], out_scales=[0.5], coeff=(o + i) % 4, crossfade=(i == 0)) for o in range(n) for i in range(2)]
filters.append(Filter([1], [0], coeff=-1, from_filters=[0, 3], fscales=[0.5, 0.25]))
g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P, 5, 1])
taps = [t * 0.5 for t in configs.synthetic_filters(g, 51)]
rng = np.random.default_rng(51)
sig = configs.synthetic_signal(g, 51, 64, sigma=0.03)
script = {}
for k in range(20, calls, 37):
    f = int(rng.integers(0, len(filters)))
    kw = dict(coeff=int(rng.integers(-1, 4)), delayblocks=int(rng.integers(0, 4)))
    if rng.random() < 0.5 and f < len(filters) - 1:
        kw["in_scales"] = [float(rng.choice([1.0, 0.5, 0.25]))]
    script[k] = (f, kw)

def run(flags):
    digest = []
    with Engine(g, max_batch=B, flags=flags) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        out = np.zeros((4, B, g.out_bytes), np.uint8)
        for k in range(calls):
            if k in script:
                e.set_control(script[k][0], **script[k][1])
            blocks = np.stack([sig[(k * B + b) % 64] for b in range(B)])
            e.process_blocks_async(blocks, out[k % 4], B)
            if k % 4 == 3 or k == calls - 1:
                if k >= 2:
                    e.lib.bfcuda_wait_previous(e.h, 0)
                digest.append(int(out.astype(np.uint64).sum()))
        e.synchronize()
        digest.append(int(out.astype(np.uint64).sum()))
        rings = e.info().n_streams
    return digest, rings

a, ra = run(0)
b, rb = run(_abi.FLAG_SERIAL_STAGES)
ok = a == b
print(f"soak B={B} calls={calls}: pipelined vs serialised digests {'IDENTICAL' if ok else 'DIFFER'} ({len(a)} checkpoints, rings {ra}/{rb})")
sys.exit(0 if ok else 1)
