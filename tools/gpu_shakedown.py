# first GPU shake-down: small engine vs oracle, MAC variants, then C3 timing
import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from brutefir_b200 import configs, _abi
from brutefir_b200.engine import Engine
from brutefir_b200.formats import unpack_block
from oracle import pyoracle as po

def compare(graph, cid, nb, variant=None, split=0, kind="oracle"):
    if variant is not None: os.environ["BFCUDA_MAC_VARIANT"] = str(variant)
    taps = configs.synthetic_filters(graph, cid)
    sig = configs.synthetic_signal(graph, cid, nb)
    e = Engine(graph, mac_split=split)
    d = po.BlockDriver(kind, graph)
    for c, h in enumerate(taps):
        e.coeff_from_taps(c, h); d.coeff_from_taps(c, h)
    got = e.run(sig); ref = d.run(sig)
    L = graph.filter_length
    g = np.stack([unpack_block(got[b], graph.out_formats, L) for b in range(nb)])
    r = np.stack([unpack_block(ref[b], graph.out_formats, L) for b in range(nb)])
    err = np.abs(g - r).max()
    info = e.info()
    print(f"L={L} P={graph.n_blocks} rs={graph.realsize} variant={variant} split={info.mac_split} max|gpu-oracle|={err:.3e} (ref max {np.abs(r).max():.3e})", flush=True)
    e.close(); d.close()
    return err

print("devices", _abi.load_library().bfcuda_device_count(), flush=True)
for v in (0, 1):
    compare(configs.diagonal_graph(4, 1024, 8, 4, "S24_4LE"), 9, 12, variant=v)
    compare(configs.diagonal_graph(4, 1024, 8, 8, "S24_4LE"), 9, 12, variant=v)
compare(configs.diagonal_graph(3, 64, 16, 4, "FLOAT_LE"), 9, 20, variant=0)
compare(configs.diagonal_graph(3, 64, 16, 4, "FLOAT_LE"), 9, 20, variant=0, split=4)
compare(configs.config_c5(), 5, 70, variant=0)
compare(configs.diagonal_graph(2, 16384, 4, 4, "S16_LE"), 9, 6, variant=0)
compare(configs.diagonal_graph(2, 8192, 4, 8, "S32_LE"), 9, 6, variant=0)

# C3 timing, device resident
for v in (0, 1):
    os.environ["BFCUDA_MAC_VARIANT"] = str(v)
    g = configs.config_c3()
    e = Engine(g, flags=_abi.FLAG_STAGE_TIMING)
    rng = np.random.default_rng(0)
    t0 = time.time()
    for c in range(64):
        h = (rng.standard_normal(g.taps_per_filter()) * 1e-3).astype(np.float32)
        e.coeff_from_taps(c, h)
    print("coeff upload s", time.time() - t0, flush=True)
    sig = configs.synthetic_signal(g, 3, 1)
    e.upload_input(sig[0])
    for _ in range(140): e.process_block_device()
    e.synchronize(); e.stage_times()
    e.timer_start()
    K = 200
    for _ in range(K): e.process_block_device()
    ms = e.timer_stop()
    st, nb, nl = e.stage_times()
    info = e.info()
    print(f"variant {v}: {ms/K*1000:.1f} us/block; stages(ms) {st}; mac GB/s {info.mac_bytes_per_block/ (st[1]*1e-3)/1e9:.1f}; RT x{(8192/48000)/(ms/K*1e-3):.0f}", flush=True)
    e.close()
