"""Randomised differential test: random filter graphs (mixes, chaining, delays, crossfade, sample formats, partition
counts from 1), random run-time control scripts, random batch sizes -- engine against the CPU oracle with the
north_star tolerances.  Scales are irrational on purpose: with a unit pulse ("coeff: -1") and a scale like 0.5 or 0.7
integer samples land exactly on .5 (0.7 x 5), where the last bit of the FFT decides the rounding direction in ANY
implementation.
Usage: python tests/checks/fuzz_parity.py [n_cases] [seed]
Environment: FUZZ_BIG=1 (partitions of 1024..8192 samples, batches up to 16), FUZZ_PMAX=n (up to n partitions),
FUZZ_WIDE=1 (up to 70 channels each way and 80 filters),
FUZZ_LL=1 (BFCUDA_FLAG_LOW_LATENCY),
FUZZ_B=n (force the batch size), FUZZ_ONLY=i,j (run and explain only these cases)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from brutefir_b200 import _abi, configs
from brutefir_b200.engine import Engine
from brutefir_b200.formats import interleaved_layout, planar_layout
from brutefir_b200.graph import Filter, FilterGraph
from oracle import pyoracle as po
from helpers import unpack_run

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
INT_FMTS = ["S16_LE", "S24_LE", "S24_4LE", "S24_4BE", "S32_LE"]
bad = 0
for case in range(n_cases):
    big = os.environ.get("FUZZ_BIG") == "1"
    L = int(rng.choice([1024, 2048, 4096, 8192] if big else [16, 64, 128, 256, 512, 1024, 2048]))
    P = int(rng.integers(1, int(os.environ.get("FUZZ_PMAX", "12" if big else "6")) + 1))
    rs = int(rng.choice([4, 4, 8]))
    wide = os.environ.get("FUZZ_WIDE") == "1"      # more than one 32-channel tile, many filters per delay-line ring
    n_in, n_out = (int(rng.integers(1, 71)), int(rng.integers(1, 71))) if wide else (int(rng.integers(1, 5)), int(rng.integers(1, 5)))
    fin = str(rng.choice(INT_FMTS + ["FLOAT_LE"]))
    fout = str(rng.choice(["S24_4LE", "S24_LE", "FLOAT_LE"] if rs == 4 else INT_FMTS + ["FLOAT_LE", "FLOAT64_LE"]))
    inb, nin = (interleaved_layout if rng.random() < 0.7 else planar_layout)(n_in, fin, L)
    outb, nout = (interleaved_layout if rng.random() < 0.7 else planar_layout)(n_out, fout, L)
    n_coeffs = int(rng.integers(1, 4))
    coeff_blocks = [int(rng.integers(1, P + 1)) for _ in range(n_coeffs)]
    nf = int(rng.integers(1, 81 if wide else 7))
    filters = []
    for f in range(nf):
        srcs = [int(x) for x in rng.choice(f, size=int(rng.integers(1, min(f, 2) + 1)), replace=False)] if f >= 2 and rng.random() < 0.3 else []
        k_in = int(rng.integers(0 if srcs else 1, min(n_in, 2) + 1))
        ins = [int(x) for x in rng.choice(n_in, size=k_in, replace=False)]
        outs = [int(x) for x in rng.choice(n_out, size=int(rng.integers(0, min(n_out, 2) + 1)), replace=False)]
        filters.append(Filter(ins, outs, in_scales=[float(rng.choice([1.0, 0.7071067811865476, -0.3183098861837907])) for _ in ins],
                              out_scales=[float(rng.choice([1.0, 0.6180339887498949])) for _ in outs],
                              coeff=int(rng.integers(-1, n_coeffs)), delayblocks=int(rng.integers(0, P)),
                              crossfade=bool(rng.random() < 0.5 and rs == 4), from_filters=srcs,
                              fscales=[float(rng.choice([1.0, 0.4342944819032518])) for _ in srcs]))
    g = FilterGraph(L, P, rs, inb, outb, nin, nout, filters, coeff_blocks)
    taps = [(rng.standard_normal(L * nb) / (4 * np.sqrt(L * nb / 64))).astype(np.float32 if rs == 4 else np.float64) for nb in coeff_blocks]
    nblk = 3 * P + 8
    # wide graphs sum up to dozens of filters per output: keep the peaks where float32 still resolves 1 LSB @24 bit
    sig = configs.synthetic_signal(g, 100 + case, nblk, sigma=0.001 if wide else 0.01)
    script = {}
    for b in range(2, nblk, int(rng.integers(3, 7))):
        f = int(rng.integers(0, nf))
        kw = dict(coeff=int(rng.integers(-1, n_coeffs)), delayblocks=int(rng.integers(0, P)))
        if rng.random() < 0.4 and filters[f].inputs:
            kw["in_scales"] = [float(rng.choice([1.0, 0.7071067811865476, -0.3183098861837907])) for _ in filters[f].inputs]
        if rng.random() < 0.3 and filters[f].outputs:
            kw["out_scales"] = [float(rng.choice([1.0, 0.6180339887498949])) for _ in filters[f].outputs]
        if rng.random() < 0.3 and filters[f].from_filters:
            kw["fscales"] = [float(rng.choice([1.0, 0.4342944819032518])) for _ in filters[f].from_filters]
        script[b] = (f, kw)
    split = int(rng.choice([1, 1, 0, 3]))
    ll = os.environ.get("FUZZ_LL") == "1"       # the real-time schedule: head + tail partition sums (a two-way split)
    if ll:
        split = 2
    B = int(rng.choice([1, 4, 8, 16] if big and rs == 4 else [1, 1, 2, 4]))
    if os.environ.get("FUZZ_B"):
        B = int(os.environ["FUZZ_B"])
    if os.environ.get("FUZZ_ONLY") and case not in [int(x) for x in os.environ["FUZZ_ONLY"].split(",")]:
        continue
    desc = f"case {case}: L={L} P={P} rs={rs} in={n_in}x{fin} out={n_out}x{fout} filters={nf} B={B} split={split}"
    try:
        d = po.BlockDriver("oracle", g)
        with Engine(g, mac_split=split, max_batch=B, flags=_abi.FLAG_LOW_LATENCY if ll else 0) as e:
            for c, h in enumerate(taps):
                e.coeff_from_taps(c, h)
                d.coeff_from_taps(c, h)
            want, got = [], np.zeros((nblk, g.out_bytes), np.uint8)
            for b in range(nblk):
                if b in script:
                    d.set_control(script[b][0], **script[b][1])
                want.append(d.process_block(sig[b]))
            b = 0
            while b < nblk:
                if b in script:
                    e.set_control(script[b][0], **script[b][1])
                k = 1
                while k < B and b + k < nblk and (b + k) not in script:
                    k += 1
                e.process_blocks_async(sig[b:b + k], got[b:b + k], k)
                b += k
            e.synchronize()
        d.close()
        y, r = unpack_run(got, g.out_formats, L), unpack_run(np.stack(want), g.out_formats, L)
        sf = g.out_formats[0].sf
        diff = np.abs(y - r).max() if y.size else 0.0
        peak = np.abs(r).max() if r.size else 0.0
        if sf.isfloat:
            tol = 1e-6 if rs == 4 or sf.bytes == 4 else 1e-12
        elif rs == 8:
            tol = 0.0 if split == 1 else 1.0     # a split partition sum is a different summation tree
        else:
            # 1 LSB where float32 resolves it, a few ulp above (wide graphs sum dozens of filters per output and split
            # partition sums: twice that)
            tol = max(1.0, (8 if wide else 4) * 2.0 ** -23 * peak)
        ok = diff <= tol
        if not ok and rs == 8 and not sf.isfloat and diff <= 1 and np.mean(np.abs(y - r) > 0) < 0.01:
            # float_bits 64: identical except exact ties -- a unit pulse ("coeff: -1") or power-of-two scales put
            # samples exactly on .5, where the last bit of the FFT decides the rounding direction
            ok = True
        print(f"{'ok  ' if ok else 'FAIL'} {desc}: max diff {diff:g} (tol {tol:g}, peak {peak:g})", flush=True)
        if not ok and os.environ.get("FUZZ_ONLY"):
            bad_blocks = sorted(set(np.nonzero(np.abs(y - r) > tol)[1] // L))
            print("   filters:", [(f.inputs, f.outputs, f.coeff, f.delayblocks, f.crossfade, f.from_filters) for f in filters])
            print("   coeff blocks", coeff_blocks, "script", script, "bad blocks", bad_blocks[:20], "bad channels", sorted(set(np.nonzero(np.abs(y - r) > tol)[0])))
        bad += not ok
    except Exception as exc:
        print(f"ERR  {desc}: {exc!r}", flush=True)
        bad += 1
print(f"{n_cases - bad}/{n_cases} cases within tolerance")
sys.exit(1 if bad else 0)
