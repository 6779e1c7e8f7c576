#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's headline configuration.

Workload (configs[2], "c3"): 64 inputs -> 64 outputs, 64 filters of 1 048 576 taps, uniformly partitioned
8192 x 128, 48 kHz, float_bits 32, S24_4LE interleaved I/O, dither off, synthetic white noise and random
unit-energy filters.  A "step" is one call of the hot path -- raw2real -> FFT -> delay-line MAC over all
partitions -> IFFT -> real2raw -- over --batch consecutive audio blocks (8192 samples on every channel =
170.67 ms of audio each).  The default (8 blocks; 16 when a rank holds 16 filters or fewer) is the offline /
file-to-file mode the reference's own benchmark configs run in (bfio_file, no real-time constraint); --batch 1 is
the reference's block-by-block schedule and is ALWAYS measured too and reported under "streaming" (with the
SURVEY.md 8(d) roofline).

  value   realtime multiple with the raw input block already resident in HBM (device-timed, CUDA events on
          the engine's stream), whole job over all ranks
  e2e     the same through the C ABI with HOST buffers: pinned host -> device copy of every input block and
          device -> host copy of every output block inside the timed region (pipelined streaming), plus the
          fully synchronous per-block latency
  roofline  the MAC kernel's algorithmic bytes / its measured duration against the measured HBM peak
  cpu_baseline  the reference's own convolver (oracle/_ref) on the host cores, bounded sample

N > 1 (torchrun): the 64 filters are sharded by filter over the ranks exactly as the reference deals
filter groups over CPUs (brutefir_b200/sharding.py); diagonal graph => no data-path collective; strong
scaling (the job is fixed, each rank holds 64/N filters); time = max over ranks.

`--impl reference` times the reference's CPU implementation of the same step on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
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
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "m8"],
                    help="c3 = the headline configuration; c2 / c4 = BASELINE configs[1] / [3]; m8 = 8 x 8 matrix of the "
                         "headline filter shape (64 filters, every input feeds 8 of them: shared delay lines)")
    ap.add_argument("--no-sharing", action="store_true", help="give every filter its own delay line (A/B for m8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0,
                    help="audio blocks per step.  0 (default) = 8, or 16 when a rank holds 16 filters or fewer (small "
                         "shards amortise their per-step launch cost over more blocks); 1 = the reference's "
                         "block-by-block schedule (its figures are always measured and reported under 'streaming' too)")
    return ap.parse_args()


def workload_graph(name):
    from brutefir_b200 import configs
    return {"c2": configs.config_c2, "c3": configs.config_c3, "c4": configs.config_c4, "m8": configs.config_matrix}[name]()


def workload_config(name, graph, n_gpus):
    return {"workload": f"{name}: {len(graph.filters)} filters x {graph.taps_per_filter()} taps, "
                        f"{graph.filter_length} x {graph.n_blocks} partitions, {graph.sampling_rate} Hz, "
                        f"float_bits {graph.realsize * 8}, {graph.in_formats[0].sf.name} I/O, dither off",
            "n_filters": len(graph.filters), "filter_length": graph.filter_length, "n_blocks": graph.n_blocks,
            "sampling_rate": graph.sampling_rate, "parallelism": f"filters sharded over {n_gpus} GPU(s), no collective; every rank is handed and returns "
                           "only its own channels' interleaved blocks",
            "l2": "per-block working set (coefficients + delay lines) exceeds L2 many times over; nothing is re-read "
                  "from L2 between steps"}


def fast_filters(graph, seed):
    """Random unit-energy decaying filters; float32 generation keeps the set-up short."""
    rng = np.random.default_rng(seed)
    taps = graph.taps_per_filter()
    env = np.exp(-np.arange(taps, dtype=np.float32) / (taps / 4.0))
    out = []
    for _ in range(len(graph.coeff_n_blocks)):
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


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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

    from brutefir_b200 import _abi
    from brutefir_b200.engine import Engine, PinnedBuffer
    from brutefir_b200.sharding import shard_graph

    # The data path has no collective (diagonal graph: every rank owns its filters, inputs and outputs), so the only
    # inter-rank traffic of this program is the timing barrier and the max-over-ranks of the measured times: that
    # control plane runs over gloo on the host.  (Initialising an NCCL communicator here costs the host-buffer path
    # up to 80 us per step at 4-8 ranks -- measured with tools/e2e_multi.py -- for nothing; NCCL is used where the
    # path really exchanges data, bfcuda_comm_* for outputs fed from several ranks, tests/checks/multi_gpu_check.py.)
    distributed = world > 1
    if distributed:
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")

    def barrier():
        if distributed:
            dist.barrier()

    def max_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # each rank owns a contiguous group of filters with their inputs and outputs and moves only ITS channels over
    # its PCIe link: the host fans the interleaved input out into one block per GPU
    shard = shard_graph(graph, world, compact=world > 1)[rank]
    sub = shard.graph
    taps = fast_filters(graph, 2000 + cid)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)"
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "mac_dram_bytes.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass

    def measure(B, steps, warmup, sample_clocks):
        """One engine with max_batch = B; a step = B consecutive audio blocks in one call."""
        eng = Engine(sub, device=local_rank, flags=_abi.FLAG_NO_STREAM_SHARING if args.no_sharing else 0, max_batch=B)
        for c in sorted({f.coeff for f in sub.filters if f.coeff >= 0}):
            eng.coeff_from_taps(c, taps[c])
        nbuf = int(os.environ.get("BENCH_NBUF", "3"))
        sig = configs.synthetic_signal(graph, cid, nbuf * B)
        if world > 1:
            sig = shard.slice_input(graph, sig)
        pin_in = [PinnedBuffer(B * sub.in_bytes) for _ in range(nbuf)]
        pin_out = [PinnedBuffer(B * sub.out_bytes) for _ in range(nbuf)]
        for i in range(nbuf):
            pin_in[i].array[:] = sig[i * B:(i + 1) * B].reshape(-1)
        info = eng.info()
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
        barrier()
        eng.timer_start()
        for _ in range(steps):
            eng.process_blocks_device(B)
        ms = eng.timer_stop()
        barrier()
        _, _, launches = eng.stage_times()
        ms_step = max_over_ranks(ms / steps)

        # ---- end to end through the C ABI with host buffers ------------------------------------------
        for i in range(max(3, warmup)):
            eng.process_blocks_async(pin_in[i % nbuf].array, pin_out[i % nbuf].array, B)
        eng.synchronize()
        barrier()
        eng.timer_start()
        for i in range(steps):
            eng.process_blocks_async(pin_in[i % nbuf].array, pin_out[i % nbuf].array, B)
        e2e_ms = eng.timer_stop()
        eng.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None        # sampled over both timed regions
        if os.environ.get("BENCH_DEBUG"):
            print(f"[rank {rank}] B={B} e2e {1e3 * e2e_ms / steps:.1f} us/step, device-resident {1e3 * ms / steps:.1f} us/step",
                  file=sys.stderr, flush=True)
        e2e_step = max_over_ranks(e2e_ms / steps)
        lat = []
        for i in range(min(40, steps)):
            t0 = time.perf_counter()
            check_rc = eng.lib.bfcuda_process_blocks(eng.h, B, pin_in[i % nbuf].array.ctypes.data,
                                                     pin_out[i % nbuf].array.ctypes.data)
            assert check_rc == 0
            lat.append(time.perf_counter() - t0)
        latency_ms = max_over_ranks(float(np.median(lat)) * 1e3)
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
        for _ in range(3):                                  # three chunks; the roofline uses the best chunk's means
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
        roof = {"bound": "hbm", "kernel": "k_mac" if B == 1 else f"k_mac_batch2 (B={B})", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get(f"{args.workload}_n{world}_b{B}"),
                "peak_source": peak_source, "algorithmic_bytes_per_launch": compulsory, "kernel_ms": mac_ms_launch,
                "blocks_per_launch": B,
                "timing": "CUDA events around the kernel on its stream, stages serialised (each stage alone); mean over "
                          "the launches of the best of three chunks",
                "mac_ms_per_block_chunks": [c[1] for c in chunks],
                "stage_ms_per_block": {"forward": stage_ms[0], "mac": stage_ms[1], "inverse": stage_ms[2]},
                "stage_ms_per_block_pipelined": {"forward": piped_ms[0], "mac": piped_ms[1], "inverse": piped_ms[2]},
                "fft_stages": {
                    "note": "forward = unpack + R2HC into the delay line, inverse = mix + HC2R + pack; byte roofline "
                            "n*(L*bytes + N*rs) per stage and block (SURVEY.md 8(d))",
                    "forward_gbs": (len(sub.in_formats) * (graph.filter_length * sub.in_formats[0].sf.bytes + graph.n_fft * graph.realsize) /
                                    (stage_ms[0] * 1e-3) / 1e9) if stage_ms[0] > 0 else None,
                    "inverse_gbs": (len(sub.out_formats) * (graph.filter_length * sub.out_formats[0].sf.bytes + graph.n_fft * graph.realsize) /
                                    (stage_ms[2] * 1e-3) / 1e9) if stage_ms[2] > 0 else None,
                    # flop side of the same roofline: ~2.5 N log2 N per real transform of N points
                    "forward_tflops": (len(sub.in_formats) * 2.5 * graph.n_fft * np.log2(graph.n_fft) /
                                       (stage_ms[0] * 1e-3) / 1e12) if stage_ms[0] > 0 else None,
                    "inverse_tflops": (len(sub.out_formats) * 2.5 * graph.n_fft * np.log2(graph.n_fft) /
                                       (stage_ms[2] * 1e-3) / 1e12) if stage_ms[2] > 0 else None}}
        if B > 1:
            roof["note"] = ("one launch covers B blocks and reads every coefficient / delay-line spectrum ONCE for all "
                            "of them (register reuse): algorithmic bytes = rs*N*(P*F + (P+B-1)*U + B*F).  With the "
                            "per-block formula of SURVEY.md 8(d) times B the same launch rates at "
                            f"{survey:.0f} GB/s-equivalent ({survey / peak:.2f} of peak); at B = 8 the HBM floor and the "
                            "FP32-pipe floor (8 exactly rounded flop per complex MAC) are within 15 % of each other")
            roof["survey_formula_equivalent_gbs"] = survey
        res = {"batch": B, "value": B * block_s / (ms_step * 1e-3), "ms_per_step": ms_step, "ms_per_block": ms_step / B,
               "gtap_mac_per_s": B * block_s / (ms_step * 1e-3) * gtap_unit,
               "e2e": {"value": B * block_s / (e2e_step * 1e-3), "unit": UNIT, "ms_per_step": e2e_step,
                       "h2d_bytes_per_step": B * sub.in_bytes, "d2h_bytes_per_step": B * sub.out_bytes,
                       "mode": f"pipelined bfcuda_process_blocks_async({B} block(s) per call), pinned host buffers",
                       "sync_call_latency_ms": latency_ms},
               "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
               "engine": {"mac_split": info.mac_split, "kernels_per_step": info.kernels_per_block, "max_batch": B,
                          "device": info.device_name.decode(), "device_bytes": info.device_bytes,
                          "filters_on_rank0": len(sub.filters), "delay_line_rings": info.n_streams}}
        eng.close()
        for b in pin_in + pin_out:
            b.free()
        return res

    def measure_low_latency():
        """Synchronous per-block call latency of the real-time schedule (BFCUDA_FLAG_LOW_LATENCY): the sum over the
        partitions 1 .. P-1 of the next block is made while the host waits for that block, so the timed call only
        multiplies partition 0.  Paced like a real-time host: the engine is idle when the input arrives."""
        eng = Engine(sub, device=local_rank, flags=_abi.FLAG_LOW_LATENCY, max_batch=1)
        for c in sorted({f.coeff for f in sub.filters if f.coeff >= 0}):
            eng.coeff_from_taps(c, taps[c])
        sig = configs.synthetic_signal(graph, cid, 2)
        if world > 1:
            sig = shard.slice_input(graph, sig)
        pin_in, pin_out = PinnedBuffer(sub.in_bytes), PinnedBuffer(sub.out_bytes)
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

    B = args.batch if args.batch >= 1 else (8 if len(sub.filters) > 16 else 16)
    head = measure(B, args.steps, args.warmup, True)
    stream = head if B == 1 else measure(1, max(args.steps, 50), args.warmup, False)
    try:
        low_latency_ms = measure_low_latency()
    except Exception as exc:        # an extra figure: never takes the headline down with it
        low_latency_ms = float("inf")
        print(f"low-latency measurement failed: {exc!r}", file=sys.stderr)
    low_latency_ms = max_over_ranks(low_latency_ms)     # every rank takes part, whatever happened above
    if not np.isfinite(low_latency_ms):
        low_latency_ms = None

    cfg["blocks_per_step"] = B
    cfg["schedule"] = ("block by block (the reference's filter_process schedule)" if B == 1 else
                       f"{B} consecutive blocks per call (offline / file-to-file mode, I/O delay +{B - 1} blocks; results "
                       "bit-identical to block by block); the block-by-block figures are under 'streaming'")
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if graph.realsize == 4 else "f64", "data": "synthetic", "config": cfg,
            "gtap_mac_per_s": head["gtap_mac_per_s"], "ms_per_block": head["ms_per_block"],
            "latency_ms_per_block": stream["e2e"]["sync_call_latency_ms"],
            "latency_ms_per_block_low_latency_schedule": low_latency_ms,
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "clocks": head["clocks"], "engine": head["engine"],
            "streaming": {k: stream[k] for k in ("batch", "value", "ms_per_step", "gtap_mac_per_s", "e2e", "roofline",
                                                 "gpu_launches")}}

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
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
