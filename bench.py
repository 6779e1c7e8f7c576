#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's headline configuration.

Workload (configs[2], "c3"): 64 inputs -> 64 outputs, 64 filters of 1 048 576 taps, uniformly partitioned
8192 x 128, 48 kHz, float_bits 32, S24_4LE interleaved I/O, dither off, synthetic white noise and random
unit-energy filters.  A "step" is one call of the hot path -- raw2real -> FFT -> delay-line MAC over all
partitions -> IFFT -> real2raw -- over --batch consecutive audio blocks (8192 samples on every channel =
170.67 ms of audio each).  The default is 8 blocks per step AT EVERY N (the offline / file-to-file mode the
reference's own benchmark configs run in: bfio_file, no real-time constraint); --batch 1 is the reference's
block-by-block schedule and is ALWAYS measured too and reported under "streaming" (with the SURVEY.md 8(d)
roofline).

  value   realtime multiple with the raw input block already resident in HBM (device-timed, CUDA events on
          the engine's stream), whole job over all ranks
  e2e     the same through the C ABI with HOST buffers: pinned host -> device copy of every input block and
          device -> host copy of every output block inside the timed region (pipelined streaming), plus the
          fully synchronous per-block latency; `copy_only` beside it is the same copies without any kernel,
          all ranks at once (what the box's host side can carry)
  roofline  the MAC kernel's algorithmic bytes / its measured duration against the measured HBM peak
  cpu_baseline  the reference's own convolver (oracle/_ref) on the host cores, bounded sample
  configs   sub-records for BASELINE configs[1] (c2), [3] (c4), [4] (c5) and the headline shape at float_bits 64
            (N = 1 only): value, e2e, latency, roofline fraction each
  nccl_xtc  N > 1 only: BASELINE configs[4] with the two filters of every output on different ranks -- the
            time-domain output blocks are summed over NVLink by ncclAllReduce inside the engine (bfcuda_comm_*);
            parity against the oracle and the cost of the exchange step per block

N > 1 (torchrun): the 64 filters are sharded by filter over the ranks exactly as the reference deals
filter groups over CPUs (brutefir_b200/sharding.py); diagonal graph => no data-path collective; strong
scaling (the job is fixed, each rank holds 64/N filters); time = max over ranks.

`--impl reference` times the reference's CPU implementation of the same step on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "realtime_multiple"
UNIT = "x realtime (64ch x 1M-tap @48kHz); Gtap-MAC/s and per-block latency alongside"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "m8", "c3f64", "c3s24le"],
                    help="c3 = the headline configuration; c2 / c4 = BASELINE configs[1] / [3]; m8 = 8 x 8 matrix of the "
                         "headline filter shape (64 filters, every input feeds 8 of them: shared delay lines); c3f64 = "
                         "the headline shape at float_bits 64")
    ap.add_argument("--no-sharing", action="store_true", help="give every filter its own delay line (A/B for m8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (configs, nccl_xtc, batch-16 series)")
    ap.add_argument("--quick", action="store_true", help="only the headline measurement (kernel experiments)")
    ap.add_argument("--shard-of", type=int, default=0,
                    help="single process: run rank 0's shard of a K-rank job (profiling a rank's kernels on one GPU)")
    ap.add_argument("--batch", type=int, default=0,
                    help="audio blocks per step; 0 (default) = 8 at every N; 1 = the reference's block-by-block schedule "
                         "(its figures are always measured and reported under 'streaming' too)")
    return ap.parse_args()


def workload_graph(name):
    from brutefir_b200 import configs
    return {"c2": configs.config_c2, "c3": configs.config_c3, "c4": configs.config_c4, "m8": configs.config_matrix,
            "c3f64": lambda: configs.config_c3(realsize=8), "c3s24le": lambda: configs.config_c3(fmt="S24_LE")}[name]()


def workload_config(name, graph, n_gpus):
    return {"workload": f"{name}: {len(graph.filters)} filters x {graph.taps_per_filter()} taps, "
                        f"{graph.filter_length} x {graph.n_blocks} partitions, {graph.sampling_rate} Hz, "
                        f"float_bits {graph.realsize * 8}, {graph.in_formats[0].sf.name} I/O, dither off",
            "n_filters": len(graph.filters), "filter_length": graph.filter_length, "n_blocks": graph.n_blocks,
            "sampling_rate": graph.sampling_rate, "parallelism": f"filters sharded over {n_gpus} GPU(s), no collective; every rank is handed and returns "
                           "only its own channels' interleaved blocks; ranks on every (visible GPUs / ranks)-th device, i.e. one PCIe switch "
                           "uplink each where the box shows more GPUs than the run uses",
            "l2": "per-block working set (coefficients + delay lines) exceeds L2 many times over; nothing is re-read "
                  "from L2 between steps"}


def fast_filters(graph, seed):
    """Random unit-energy decaying filters; float32 generation keeps the set-up short."""
    rng = np.random.default_rng(seed)
    out = []
    for nb in graph.coeff_n_blocks:
        taps = nb * graph.filter_length
        env = np.exp(-np.arange(taps, dtype=np.float32) / (taps / 4.0))
        h = rng.standard_normal(taps, dtype=np.float32) * env
        h /= np.sqrt(np.sum(h.astype(np.float64) ** 2))
        out.append(h.astype(np.float32 if graph.realsize == 4 else np.float64))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.proc = None
        self.lines = []
        self.device_index = device_index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(graph, taps, sig_block, n_warm, n_steps, budget_s=None):
    """Time the reference's own convolver (oracle/_ref, else the oracle port) on all host cores."""
    from oracle import pyoracle as po
    kind, lib_kind = ("reference", "ref") if po.available("ref") else ("port", "oracle")
    if not po.available(lib_kind):
        po.build()
    cores = len(os.sched_getaffinity(0))
    d = po.BlockDriver(lib_kind, graph, n_threads=cores)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    # the reference convolves min(P, blocks seen so far) partitions (bfrun.c:1745-1754): warm up until every
    # partition of the delay line is live, otherwise the timed blocks do a fraction of the work
    d.run_timed(sig_block, max(1, n_warm, graph.n_blocks))
    if budget_s is not None:
        probe = d.run_timed(sig_block, 3) / 3
        n_steps = int(max(5, min(400, budget_s / max(probe, 1e-6))))
    secs = d.run_timed(sig_block, n_steps)
    d.close()
    return secs / n_steps, n_steps, cores, kind


class Dist:
    """Control plane of the benchmark: the timing barrier and max / sum over ranks, over gloo on the host.  The data
    path of the sharded diagonal workload has no collective; an NCCL communicator is created only where the path
    exchanges data (nccl_xtc below)."""

    def __init__(self, world):
        self.world = world
        self.on = world > 1
        if self.on:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            dist.init_process_group("gloo")

    def barrier(self):
        if self.on:
            self.dist.barrier()

    def reduce(self, v, op="max"):
        if not self.on:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def broadcast_bytes(self, payload):
        if not self.on:
            return payload
        box = [payload]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.on:
            self.dist.barrier()
            self.dist.destroy_process_group()


def load_peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "mac_dram_bytes.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    return peak, source, traffic


def measure(ctx, graph, shard, taps, cid, B, steps, warmup, sample_clocks=False, flags=0, tag="c3", light=False):
    """One engine with max_batch = B on this rank's shard; a step = B consecutive audio blocks in one call."""
    from brutefir_b200 import configs
    from brutefir_b200.engine import Engine, PinnedBuffer
    dist, rank, local_rank, world = ctx["dist"], ctx["rank"], ctx["local_rank"], ctx["world"]
    peak, peak_source, traffic = ctx["peak"], ctx["peak_source"], ctx["traffic"]
    sub = shard.graph
    block_s, gtap_unit = graph.block_seconds(), graph.gtap_mac_per_realtime()
    eng = Engine(sub, device=local_rank, flags=flags, max_batch=B)
    for c in sorted({f.coeff for f in sub.filters if f.coeff >= 0}):
        eng.coeff_from_taps(c, taps[c])
    nbuf = int(os.environ.get("BENCH_NBUF", "3"))
    sig_full = configs.synthetic_signal(graph, cid, nbuf * B)
    sliced = ctx["n_shards"] > 1
    sig = shard.slice_input(graph, sig_full) if sliced else sig_full
    pin_in = [PinnedBuffer(B * sub.in_bytes, local_rank) for _ in range(nbuf)]
    pin_out = [PinnedBuffer(B * sub.out_bytes, local_rank) for _ in range(nbuf)]
    for i in range(nbuf):
        pin_in[i].array[:] = sig[i * B:(i + 1) * B].reshape(-1)
    # fill the delay line once so that every partition multiplies real data
    eng.upload_inputs(sig[:B])
    for _ in range(graph.n_blocks // B + 1):
        eng.process_blocks_device(B)
    eng.synchronize()
    info = eng.info()       # after the first block: delay lines that several filters share are merged by now

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(max(3, warmup)):
        eng.process_blocks_device(B)
    eng.synchronize()
    eng.stage_times()
    sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    eng.timer_start()
    for _ in range(steps):
        eng.process_blocks_device(B)
    ms = eng.timer_stop()
    dist.barrier()
    _, _, launches = eng.stage_times()
    ms_step = dist.reduce(ms / steps)

    # the driver's timed region is short (20 steps = 5 ms): the same loop over 1000 steps as the sustained figure
    sustained = None
    if not light:
        n_long = 1000
        dist.barrier()
        eng.timer_start()
        for _ in range(n_long):
            eng.process_blocks_device(B)
        sustained = dist.reduce(eng.timer_stop() / n_long)
        dist.barrier()
        eng.stage_times()

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    for i in range(max(3, warmup)):
        eng.process_blocks_async(pin_in[i % nbuf].array, pin_out[i % nbuf].array, B)
    eng.synchronize()
    dist.barrier()
    eng.timer_start()
    for i in range(steps):
        eng.process_blocks_async(pin_in[i % nbuf].array, pin_out[i % nbuf].array, B)
    e2e_ms = eng.timer_stop()
    eng.synchronize()
    dist.barrier()
    clocks = sampler.stop() if sampler else None        # sampled over both timed regions
    if os.environ.get("BENCH_DEBUG"):
        print(f"[rank {rank}] {tag} B={B} e2e {1e3 * e2e_ms / steps:.1f} us/step, device-resident {1e3 * ms / steps:.1f} us/step",
              file=sys.stderr, flush=True)
    e2e_step = dist.reduce(e2e_ms / steps)
    lat = []
    for i in range(min(40, steps)):
        t0 = time.perf_counter()
        check_rc = eng.lib.bfcuda_process_blocks(eng.h, B, pin_in[i % nbuf].array.ctypes.data,
                                                 pin_out[i % nbuf].array.ctypes.data)
        assert check_rc == 0
        lat.append(time.perf_counter() - t0)
    latency_ms = dist.reduce(float(np.median(lat)) * 1e3)

    # ---- the same copies without any kernel, every rank at the same time -----------------------------
    dist.barrier()
    cp_ms, h2d, d2h = eng.copy_baseline(pin_in[0].array, pin_out[0].array, B, max(10, min(steps, 100)))
    dist.barrier()
    cp_ms_max = dist.reduce(cp_ms)
    agg = dist.reduce((B * sub.in_bytes + B * sub.out_bytes) / (cp_ms * 1e-3) / 1e9, "sum")
    copy_only = {"ms_per_step": cp_ms_max, "value": B * block_s / (cp_ms_max * 1e-3),
                 "h2d_gbs_rank0": h2d, "d2h_gbs_rank0": d2h, "aggregate_gbs_both_directions": agg,
                 "note": "same pinned buffers, same copy streams, both directions at once, all ranks concurrently, no "
                         "kernels: the ceiling the box's host side sets for e2e at this shard size"}
    # host fan-out / gather of a sharded run: every rank is handed only its channels; cutting them out of the full
    # interleaved block (and writing its outputs back into it) is host work outside the timed e2e loop -- its cost:
    fan_ms = None
    if sliced and not light:
        nfr = B * graph.filter_length
        full_i = sig_full[:B].reshape(nfr, -1)
        fr_in = graph.in_formats[0].sample_spacing * graph.in_formats[0].sf.bytes
        w_in = len(sub.in_formats) * graph.in_formats[0].sf.bytes
        if full_i.shape[1] >= fr_in and w_in <= fr_in:
            dst = np.empty((nfr, w_in), np.uint8)
            full_o = np.zeros((nfr, fr_in), np.uint8)
            t0 = time.perf_counter()
            for _ in range(5):
                np.copyto(dst, full_i[:, :w_in])
                np.copyto(full_o[:, :w_in], dst)
            fan_ms = (time.perf_counter() - t0) / 5 * 1e3
    eng.stage_times()

    # ---- per-stage durations for the roofline ------------------------------------------------------
    # The engine overlaps the stages of consecutive launches, so events around a stage in the runs above would
    # include the time it shares the SMs with its neighbours (and recording them costs a few percent, which is
    # why the runs above go without).  Here each stage is timed ALONE: same engine, same data, launches
    # serialised (BFCUDA_FLAG_SERIAL_STAGES), CUDA events on the stage's own stream.
    eng.set_stage_timing(True)
    eng.set_serial_stages(True)
    for _ in range(3):
        eng.process_blocks_device(B)
    eng.synchronize()
    eng.stage_times()
    chunks = []
    for _ in range(1 if light else 3):                  # the roofline uses the best chunk's means
        for _ in range(max(20, steps // 6)):
            eng.process_blocks_device(B)
        eng.synchronize()
        chunks.append(eng.stage_times()[0])             # mean ms per BLOCK of each stage, running alone
    stage_ms = min(chunks, key=lambda c: c[1])
    eng.set_serial_stages(False)
    for _ in range(3):
        eng.process_blocks_device(B)
    eng.synchronize()
    eng.stage_times()
    for _ in range(max(20, steps // 4)):
        eng.process_blocks_device(B)
    eng.synchronize()
    piped_ms, _, _ = eng.stage_times()                 # the same with the stages of neighbouring launches overlapping
    eng.set_stage_timing(False)

    # ---- roofline of the MAC kernel --------------------------------------------------------------
    mac_ms_launch = stage_ms[1] * B                               # one launch covers B blocks
    compulsory = info.mac_bytes_per_batch if B > 1 else info.mac_bytes_per_block
    achieved = compulsory / (mac_ms_launch * 1e-3) / 1e9 if mac_ms_launch > 0 else 0.0
    survey = info.mac_bytes_per_block * B / (mac_ms_launch * 1e-3) / 1e9 if mac_ms_launch > 0 else 0.0
    n_in, n_out = len(sub.in_formats), len(sub.out_formats)
    fwd_bytes = n_in * (graph.filter_length * sub.in_formats[0].sf.bytes + graph.n_fft * graph.realsize)
    inv_bytes = n_out * (graph.filter_length * sub.out_formats[0].sf.bytes + graph.n_fft * graph.realsize)
    flop = 2.5 * graph.n_fft * np.log2(graph.n_fft)
    roof = {"bound": "hbm", "kernel": "k_mac" if B == 1 else f"k_mac_batch2 (B={B})", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get(f"{tag}_n{ctx['n_shards']}_b{B}"),
            "peak_source": peak_source, "algorithmic_bytes_per_launch": compulsory, "kernel_ms": mac_ms_launch,
            "blocks_per_launch": B,
            "timing": "CUDA events around the kernel on its stream, stages serialised (each stage alone); mean over "
                      "the launches of the best chunk",
            "mac_ms_per_block_chunks": [c[1] for c in chunks],
            "stage_ms_per_block": {"forward": stage_ms[0], "mac": stage_ms[1], "inverse": stage_ms[2]},
            "stage_ms_per_block_pipelined": {"forward": piped_ms[0], "mac": piped_ms[1], "inverse": piped_ms[2]},
            "step_over_mac": (ms_step / mac_ms_launch) if mac_ms_launch > 0 else None,
            "fft_stages": {
                "note": "forward = unpack + R2HC into the delay line, inverse = mix + HC2R + pack; byte roofline "
                        "n*(L*bytes + N*rs) per stage and block (SURVEY.md 8(d))",
                "forward_gbs": fwd_bytes / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else None,
                "inverse_gbs": inv_bytes / (stage_ms[2] * 1e-3) / 1e9 if stage_ms[2] > 0 else None,
                # flop side of the same roofline: ~2.5 N log2 N per real transform of N points
                "forward_tflops": n_in * flop / (stage_ms[0] * 1e-3) / 1e12 if stage_ms[0] > 0 else None,
                "inverse_tflops": n_out * flop / (stage_ms[2] * 1e-3) / 1e12 if stage_ms[2] > 0 else None}}
    if B > 1:
        roof["note"] = ("one launch covers B blocks and reads every coefficient / delay-line spectrum ONCE for all "
                        "of them (register reuse): algorithmic bytes = rs*N*(P*F + (P+B-1)*U + B*F).  With the "
                        "per-block formula of SURVEY.md 8(d) times B the same launch rates at "
                        f"{survey:.0f} GB/s-equivalent ({survey / peak:.2f} of peak); at B = 8 the HBM floor and the "
                        "FP32-pipe floor (8 exactly rounded flop per complex MAC) are within 15 % of each other")
        roof["survey_formula_equivalent_gbs"] = survey
    res = {"batch": B, "value": B * block_s / (ms_step * 1e-3), "ms_per_step": ms_step, "ms_per_block": ms_step / B,
           "value_sustained_1000_steps": (B * block_s / (sustained * 1e-3)) if sustained else None,
           "gtap_mac_per_s": B * block_s / (ms_step * 1e-3) * gtap_unit,
           "e2e": {"value": B * block_s / (e2e_step * 1e-3), "unit": UNIT, "ms_per_step": e2e_step,
                   "h2d_bytes_per_step": B * sub.in_bytes, "d2h_bytes_per_step": B * sub.out_bytes,
                   "mode": f"pipelined bfcuda_process_blocks_async({B} block(s) per call), pinned host buffers on the GPU's NUMA node",
                   "sync_call_latency_ms": latency_ms, "copy_only": copy_only,
                   "host_fanout_gather_ms_per_step": fan_ms,
                   "host_fanout_note": None if fan_ms is None else
                   "every rank is handed and returns blocks of ITS channels only (one I/O device or file per GPU, the way bfconf "
                   "configs split channels over devices); if a single interleaved device feeds all GPUs the host must cut the "
                   "channels apart first -- this figure is that repack (in and out) done by one numpy thread, OUTSIDE the timed "
                   "e2e loop"},
           "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
           "engine": {"mac_split": info.mac_split, "kernels_per_step": info.kernels_per_block, "max_batch": B,
                      "device": info.device_name.decode(), "device_bytes": info.device_bytes,
                      "filters_on_rank0": len(sub.filters), "delay_line_rings": info.n_streams,
                      "uses_graph": info.uses_graph}}
    eng.close()
    for b in pin_in + pin_out:
        b.free()
    return res


def measure_low_latency(ctx, graph, shard, taps, cid):
    """Synchronous per-block call latency of the real-time schedule (BFCUDA_FLAG_LOW_LATENCY): the sum over the
    partitions 1 .. P-1 of the next block is made while the host waits for that block, so the timed call only
    multiplies partition 0.  Paced like a real-time host: the engine is idle when the input arrives."""
    from brutefir_b200 import _abi, configs
    from brutefir_b200.engine import Engine, PinnedBuffer
    sub = shard.graph
    eng = Engine(sub, device=ctx["local_rank"], flags=_abi.FLAG_LOW_LATENCY, max_batch=1)
    for c in sorted({f.coeff for f in sub.filters if f.coeff >= 0}):
        eng.coeff_from_taps(c, taps[c])
    sig = configs.synthetic_signal(graph, cid, 2)
    if ctx["n_shards"] > 1:
        sig = shard.slice_input(graph, sig)
    pin_in, pin_out = PinnedBuffer(sub.in_bytes, ctx["local_rank"]), PinnedBuffer(sub.out_bytes, ctx["local_rank"])
    pin_in.array[:] = sig[0].reshape(-1)
    eng.upload_inputs(sig[:1])
    for _ in range(graph.n_blocks + 1):
        eng.process_blocks_device(1)
    eng.synchronize()
    lat = []
    for i in range(40):
        t0 = time.perf_counter()
        rc = eng.lib.bfcuda_process_blocks(eng.h, 1, pin_in.array.ctypes.data, pin_out.array.ctypes.data)
        lat.append(time.perf_counter() - t0)
        assert rc == 0
        eng.synchronize()       # the ahead-of-time sum for the next block finishes in the gap between two blocks
    eng.close()
    pin_in.free()
    pin_out.free()
    return float(np.median(lat[5:])) * 1e3


def summary(rec):
    """The figures a sub-record carries: value, e2e, latency, roofline fraction."""
    return {"batch": rec["batch"], "value": rec["value"], "ms_per_block": rec["ms_per_block"],
            "gtap_mac_per_s": rec["gtap_mac_per_s"], "e2e_value": rec["e2e"]["value"],
            "sync_call_latency_ms": rec["e2e"]["sync_call_latency_ms"], "gpu_launches": rec["gpu_launches"],
            "roofline_frac": rec["roofline"]["frac"], "mac_kernel_ms": rec["roofline"]["kernel_ms"],
            "mac_bytes_per_launch": rec["roofline"]["algorithmic_bytes_per_launch"],
            "step_over_mac": rec["roofline"]["step_over_mac"], "mac_split": rec["engine"]["mac_split"],
            "uses_graph": rec["engine"]["uses_graph"]}


def sub_records(ctx, steps, warmup):
    """BASELINE configs[1], [3], [4] and the headline shape at float_bits 64, on one GPU."""
    from brutefir_b200 import configs
    from brutefir_b200.sharding import shard_graph
    out = {}
    one = dict(ctx, n_shards=1)
    for name, cid in (("c2", 2), ("c4", 4)):
        try:
            g = workload_graph(name)
            taps = fast_filters(g, 2000 + cid)
            sh = shard_graph(g, 1)[0]
            rec = {"workload": workload_config(name, g, 1)["workload"]}
            for B in (1, 8):
                rec["block_by_block" if B == 1 else "batch8"] = summary(
                    measure(one, g, sh, taps, cid, B, steps, warmup, tag=name, light=True))
            rec["latency_ms_per_block_low_latency_schedule"] = measure_low_latency(one, g, sh, taps, cid)
            bound = g.block_seconds() / (rec["block_by_block"]["mac_bytes_per_launch"] / (ctx["peak"] * 1e9))
            rec["hbm_bound_block_by_block"] = bound
            out[name] = rec
        except Exception as exc:
            out[name] = {"error": repr(exc)}
    try:
        out["c5"] = xtc_single_gpu(ctx, steps)
    except Exception as exc:
        out["c5"] = {"error": repr(exc)}
    try:
        g = workload_graph("c3f64")
        taps = fast_filters(g, 2003)
        sh = shard_graph(g, 1)[0]
        rec = {"workload": workload_config("c3f64", g, 1)["workload"],
               "batch8": summary(measure(one, g, sh, taps, 3, 8, max(20, steps // 4), warmup, tag="c3f64", light=True)),
               "block_by_block": summary(measure(one, g, sh, taps, 3, 1, max(20, steps // 4), warmup, tag="c3f64", light=True))}
        if not ctx["args"].no_cpu_baseline:
            sig0 = configs.synthetic_signal(g, 3, 1)[0]
            per_block, n, cores, kind = cpu_reference_run(g, taps, sig0, 2, 0, budget_s=6.0)
            rec["cpu_baseline"] = {"value": g.block_seconds() / per_block, "cores": cores, "kind": kind,
                                   "sample": f"{n} blocks, float_bits 64"}
        out["c3_f64"] = rec
    except Exception as exc:
        out["c3_f64"] = {"error": repr(exc)}
    try:
        # the headline shape in massive_config's OWN sample format (/root/reference/massive_config:11,17: "S24_LE", packed
        # three-byte samples): 12 MiB instead of 16 MiB each way per 8-block step -- what the host-buffer figure, bound by
        # the box's PCIe copies, gains from a quarter fewer bytes
        g = configs.config_c3(fmt="S24_LE")
        taps = fast_filters(g, 2003)
        sh = shard_graph(g, 1)[0]
        rec = {"workload": workload_config("c3 (S24_LE I/O)", g, 1)["workload"],
               "batch8": summary(measure(one, g, sh, taps, 3, 8, max(20, steps // 2), warmup, tag="c3s24le", light=True))}
        rec["batch8"]["h2d_bytes_per_step"] = 8 * g.in_bytes
        rec["batch8"]["d2h_bytes_per_step"] = 8 * g.out_bytes
        out["c3_s24le"] = rec
    except Exception as exc:
        out["c3_s24le"] = {"error": repr(exc)}
    return out


def xtc_graph_and_script(n_blocks, n=2):
    """xtc_config's topology (/root/reference/xtc_config:28-50) as an n x n crosstalk matrix: output o = direct path of
    input o + cross paths of the other inputs, every filter crossfading; coefficient set 0 = direct, 1 = cross, and a
    script that swaps the two every 16 blocks (bench5_config:5-9's cfc mechanism).  n = 2 is config 5 itself."""
    from brutefir_b200 import configs
    from brutefir_b200.formats import interleaved_layout
    from brutefir_b200.graph import Filter, FilterGraph
    if n == 2:
        g = configs.config_c5(L=64, P=64)
    else:
        L, P = 64, 64
        inb, nin = interleaved_layout(n, "S24_LE", L)
        outb, nout = interleaved_layout(n, "S24_LE", L)
        filters = [Filter([i], [o], out_scales=[2.0 / n], coeff=0 if i == o else 1, crossfade=True)
                   for o in range(n) for i in range(n)]
        g = FilterGraph(L, P, 4, inb, outb, nin, nout, filters, [P, P], sampling_rate=44100)
    taps = configs.synthetic_filters(g, 5)
    base = [f.coeff for f in g.filters]
    script, state = {}, 0
    for b in range(10, n_blocks, 16):
        state ^= 1
        script[b] = [(f, base[f] ^ state) for f in range(len(g.filters))]
    return g, taps, script


def xtc_single_gpu(ctx, steps):
    """BASELINE configs[4] on one GPU: xtc topology, L 64 x P 64, all filters crossfading, scripted coefficient swaps;
    block by block (a crossfade block is its own launch), parity against the oracle in the same run."""
    from brutefir_b200 import configs
    from brutefir_b200.engine import Engine
    from brutefir_b200.formats import unpack_block
    from oracle import pyoracle as po
    n = 160
    g, taps, script = xtc_graph_and_script(n)
    sig = configs.synthetic_signal(g, 5, n, sigma=0.05)
    with Engine(g, device=ctx["local_rank"]) as e:
        d = po.BlockDriver("ref" if po.available("ref") else "oracle", g)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
            d.coeff_from_taps(c, h)
        got, ref, lat = [], [], []
        for b in range(n):
            for filt, coeff in script.get(b, ()):
                e.set_control(filt, coeff)
                d.set_control(filt, coeff)
            t0 = time.perf_counter()
            got.append(e.process_block(sig[b]))
            lat.append(time.perf_counter() - t0)
            ref.append(d.process_block(sig[b]))
        d.close()
        # throughput: the same engine, pipelined calls, no swaps
        from brutefir_b200.engine import PinnedBuffer
        pin_in, pin_out = PinnedBuffer(g.in_bytes), PinnedBuffer(g.out_bytes)
        pin_in.array[:] = sig[0]
        e.timer_start()
        for i in range(max(steps, 200)):
            e.process_block_async(pin_in.array, pin_out.array)
        ms = e.timer_stop() / max(steps, 200)
        e.synchronize()
        pin_in.free()
        pin_out.free()
    y = np.stack([unpack_block(b, g.out_formats, 64) for b in got])
    r = np.stack([unpack_block(b, g.out_formats, 64) for b in ref])
    return {"workload": "c5: xtc_config topology, 4 crossfading filters x 4096 taps, 64 x 64 partitions, 44100 Hz, S24_LE",
            "value": g.block_seconds() / (ms * 1e-3), "ms_per_block": ms, "sync_call_latency_ms": float(np.median(lat)) * 1e3,
            "crossfaded_swaps": len(script), "max_abs_diff_lsb_vs_reference": float(np.abs(y - r).max()),
            "parity_ok": bool(np.abs(y - r).max() <= 1 and np.abs(r).max() > 1e4)}


def nccl_xtc(ctx):
    """BASELINE configs[4] with the filters of every output split over the ranks: after the inverse FFT each rank holds a
    partial time-domain block of both outputs; the engine sums them with ncclAllReduce over NVLink (bfcuda_comm_*) and
    then quantises.  Every rank ends up with the full output; rank 0 compares it with the oracle."""
    from brutefir_b200 import configs
    from brutefir_b200.engine import Engine
    from brutefir_b200.formats import unpack_block
    from brutefir_b200.sharding import shard_graph
    from oracle import pyoracle as po
    dist, rank, world, local = ctx["dist"], ctx["rank"], ctx["world"], ctx["local_rank"]
    n = 120
    g, taps, script = xtc_graph_and_script(n, 2 if world <= 4 else 4)
    sig = configs.synthetic_signal(g, 5, n, sigma=0.05)
    sh = shard_graph(g, world, split_outputs=True, compact=False)[rank]
    uid = dist.broadcast_bytes(Engine.comm_unique_id() if rank == 0 else None)
    res = {}
    with Engine(sh.graph, device=local) as e:
        e.comm_init(rank, world, uid)
        e.comm_shared_outputs(sh.shared_outputs)
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        got = []
        for b in range(n):
            for filt, coeff in script.get(b, ()):
                if filt in sh.filters:
                    e.set_control(sh.filters.index(filt), coeff)
            got.append(e.process_block(sig[b]))
        dist.barrier()
        from brutefir_b200.engine import PinnedBuffer
        pin_in, pin_out = PinnedBuffer(sh.graph.in_bytes), PinnedBuffer(sh.graph.out_bytes)
        pin_in.array[:] = sig[0]
        e.timer_start()
        for i in range(300):
            e.process_block_async(pin_in.array, pin_out.array)
        ms_comm = e.timer_stop() / 300
        e.synchronize()
    ms_comm = dist.reduce(ms_comm)
    dist.barrier()
    # the same shard without the exchange step (its outputs are then partial sums: timing only)
    with Engine(sh.graph, device=local) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        for i in range(20):
            e.process_block_async(pin_in.array, pin_out.array)
        e.synchronize()
        e.timer_start()
        for i in range(300):
            e.process_block_async(pin_in.array, pin_out.array)
        ms_local = e.timer_stop() / 300
        e.synchronize()
    ms_local = dist.reduce(ms_local)
    pin_in.free()
    pin_out.free()
    if rank == 0:
        d = po.BlockDriver("ref" if po.available("ref") else "oracle", g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h)
        ref = []
        for b in range(n):
            for filt, coeff in script.get(b, ()):
                d.set_control(filt, coeff)
            ref.append(d.process_block(sig[b]))
        d.close()
        omap = [sh.outputs.index(o) for o in range(len(g.out_formats))]
        y = np.stack([unpack_block(b, sh.graph.out_formats, 64)[omap] for b in got])
        r = np.stack([unpack_block(b, g.out_formats, 64) for b in ref])
        diff = float(np.abs(y - r).max())
        res = {"ranks": world, "workload": f"{len(g.in_formats)} x {len(g.out_formats)} crosstalk matrix, {len(g.filters)} crossfading "
                                           "filters x 4096 taps (64 x 64), filters dealt round robin over the ranks",
               "filters_on_rank0": len(sh.filters), "shared_outputs": len(sh.shared_outputs),
               "collective": "ncclAllReduce(sum) of L samples per shared output and block, on the inverse stream, between the "
                             "inverse FFT and quantisation",
               "crossfaded_swaps": len(script), "max_abs_diff_lsb_vs_reference": diff,
               "parity_ok": bool(diff <= 1 and np.abs(r).max() > 1e4),
               "us_per_block_with_allreduce": ms_comm * 1e3, "us_per_block_without": ms_local * 1e3,
               "allreduce_step_us_per_block": (ms_comm - ms_local) * 1e3}
    return res


def spread_device(local_rank, world):
    """The GPUs of the box hang in PAIRS off one PCIe switch uplink: two ranks copying on neighbouring devices get half the
    host bandwidth each (profiles/r2_copy_skew_n8.txt: ranks 0-3 at once 15-17 GB/s per direction each, ranks 0/2/4/6 at
    once 32-33 GB/s each, one rank alone 39-41 GB/s).  A run on fewer ranks than the box shows GPUs therefore takes every
    (count // world)-th device, so that each rank has an uplink of its own.  BENCH_NO_SPREAD=1 keeps device = LOCAL_RANK."""
    if world < 2 or os.environ.get("BENCH_NO_SPREAD"):
        return local_rank, 1
    try:
        import torch
        count = torch.cuda.device_count()
    except Exception:
        return local_rank, 1
    if count >= 2 * world and count % world == 0:
        return local_rank * (count // world), count // world
    return local_rank, 1


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank, device_stride = spread_device(int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from brutefir_b200 import configs

    graph = workload_graph(args.workload)
    cfg = workload_config(args.workload, graph, world)
    block_s = graph.block_seconds()
    gtap_unit = graph.gtap_mac_per_realtime()
    cid = int(args.workload[1]) if args.workload[0] == "c" else 8

    if args.impl == "reference":
        if rank != 0:
            return
        taps = fast_filters(graph, 2000 + cid)
        sig = configs.synthetic_signal(graph, cid, 1)[0]
        per_block, n, cores, kind = cpu_reference_run(graph, taps, sig, args.warmup, args.steps)
        rt = block_s / per_block
        line = {"impl": "reference", "metric": METRIC, "value": rt, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
                "warmup": args.warmup, "ms_per_step": per_block * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32" if graph.realsize == 4 else "f64", "data": "synthetic",
                "config": cfg, "gtap_mac_per_s": rt * gtap_unit,
                "cpu_baseline": {"value": rt, "unit": UNIT, "cores": cores, "kind": kind,
                                 "sample": f"{n} blocks of the full workload on {cores} host threads"},
                "e2e": {"value": rt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    if world > 1:
        # the one NCCL communicator of this program (nccl_xtc) announces itself: rank / nranks lines on stderr
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    from brutefir_b200 import _abi
    from brutefir_b200.sharding import shard_graph

    dist = Dist(world)
    n_shards = args.shard_of if (world == 1 and args.shard_of > 1) else world
    # each rank owns a contiguous group of filters with their inputs and outputs and moves only ITS channels over
    # its PCIe link: the host fans the interleaved input out into one block per GPU
    shard = shard_graph(graph, n_shards, compact=n_shards > 1)[rank]
    taps = fast_filters(graph, 2000 + cid)
    peak, peak_source, traffic = load_peaks()
    ctx = {"dist": dist, "rank": rank, "local_rank": local_rank, "world": world, "n_shards": n_shards, "peak": peak,
           "peak_source": peak_source, "traffic": traffic, "args": args}
    flags = _abi.FLAG_NO_STREAM_SHARING if args.no_sharing else 0

    B = args.batch if args.batch >= 1 else 8
    if args.quick:
        head = measure(ctx, graph, shard, taps, cid, B, args.steps, args.warmup, False, flags, tag=args.workload, light=True)
        if rank == 0:
            print(json.dumps({"quick": summary(head), "stage_ms_per_block": head["roofline"]["stage_ms_per_block"],
                              "stage_ms_per_block_pipelined": head["roofline"]["stage_ms_per_block_pipelined"]}), flush=True)
        dist.close()
        return
    head = measure(ctx, graph, shard, taps, cid, B, args.steps, args.warmup, True, flags, tag=args.workload)
    stream = head if B == 1 else measure(ctx, graph, shard, taps, cid, 1, max(args.steps, 50), args.warmup, False, flags,
                                         tag=args.workload)
    try:
        low_latency_ms = measure_low_latency(ctx, graph, shard, taps, cid)
    except Exception as exc:        # an extra figure: never takes the headline down with it
        low_latency_ms = float("inf")
        print(f"low-latency measurement failed: {exc!r}", file=sys.stderr)
    low_latency_ms = dist.reduce(low_latency_ms)        # every rank takes part, whatever happened above
    if not np.isfinite(low_latency_ms):
        low_latency_ms = None

    schedule = {"blocks_per_step": B,
                "schedule": ("block by block (the reference's filter_process schedule)" if B == 1 else
                             f"{B} consecutive blocks per call at every N (offline / file-to-file mode, I/O delay +{B - 1} "
                             "blocks; results bit-identical to block by block); the block-by-block figures are under "
                             "'streaming'")}
    if n_shards != world:
        schedule["shard_of"] = n_shards
    if device_stride > 1:
        schedule["device_stride"] = device_stride
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if graph.realsize == 4 else "f64", "data": "synthetic", "config": cfg,
            "schedule": schedule, "gtap_mac_per_s": head["gtap_mac_per_s"], "ms_per_block": head["ms_per_block"],
            "latency_ms_per_block": stream["e2e"]["sync_call_latency_ms"],
            "latency_ms_per_block_low_latency_schedule": low_latency_ms,
            "value_sustained_1000_steps": head["value_sustained_1000_steps"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "clocks": head["clocks"], "engine": head["engine"],
            "streaming": {k: stream[k] for k in ("batch", "value", "ms_per_step", "gtap_mac_per_s", "e2e", "roofline",
                                                 "gpu_launches")}}
    if not args.no_extras and args.workload == "c3":
        sub_steps = max(20, min(args.steps, 200))
        if B != 16 and len(shard.graph.filters) <= 16:
            # second series: small shards amortise their per-step launch cost over more blocks
            try:
                line["batch16_series"] = summary(measure(ctx, graph, shard, taps, cid, 16, sub_steps, args.warmup, False,
                                                         flags, tag=args.workload, light=True))
            except Exception as exc:
                line["batch16_series"] = {"error": repr(exc)}
        if n_shards > 1:
            # the same job in massive_config's own sample format (packed S24_LE: a quarter fewer bytes over PCIe, which is
            # what bounds the host-buffer figure at every N); a second series, the headline stays on S24_4LE
            try:
                g24 = configs.config_c3(fmt="S24_LE")
                sh24 = shard_graph(g24, n_shards, compact=True)[rank]
                r24 = measure(ctx, g24, sh24, taps, cid, B, sub_steps, args.warmup, False, flags, tag="c3s24le", light=True)
                line["s24le_series"] = {"value": r24["value"], "e2e_value": r24["e2e"]["value"],
                                        "h2d_bytes_per_step": r24["e2e"]["h2d_bytes_per_step"],
                                        "d2h_bytes_per_step": r24["e2e"]["d2h_bytes_per_step"],
                                        "copy_only_value": (r24["e2e"].get("copy_only") or {}).get("value"),
                                        "note": "I/O in massive_config's own sample format (packed S24_LE)"}
            except Exception as exc:
                line["s24le_series"] = {"error": repr(exc)}
        if world > 1:
            # A rank that fails before a collective would leave the others spinning in it: if the check has not come
            # back in time, rank 0 prints the line it has and every rank leaves.
            def bail():
                if rank == 0:
                    line["nccl_xtc"] = {"error": "timed out (a rank did not reach the collective)"}
                    print(json.dumps(line), flush=True)
                os._exit(0)
            dog = threading.Timer(float(os.environ.get("BENCH_NCCL_TIMEOUT", "150")), bail)
            dog.daemon = True
            dog.start()
            # NCCL writes its version / INFO lines to stdout: send them to stderr while the communicator lives, so that
            # stdout carries the one JSON line only
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                res = nccl_xtc(ctx)
            except Exception as exc:
                res = {"error": repr(exc)}
                print(f"[rank {rank}] nccl_xtc failed: {exc!r}", file=sys.stderr, flush=True)
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
            line["nccl_xtc"] = res
            dist.barrier()
            dog.cancel()
        elif n_shards == 1:
            line["configs"] = sub_records(ctx, sub_steps, args.warmup)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sig0 = configs.synthetic_signal(graph, cid, 1)[0]
            per_block, n, cores, kind = cpu_reference_run(graph, taps, sig0, 2, 0, budget_s=15.0)
            crt = block_s / per_block
            line["cpu_baseline"] = {"value": crt, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": per_block * 1e3,
                                    "sample": f"{n} blocks of the full workload, {cores} host threads, filters dealt "
                                              "over threads like load_balance_filters"}
        except Exception as exc:     # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"error": repr(exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.close()


if __name__ == "__main__":
    main()
