"""The C host program (host/bfcuda_run.c) -- raw PCM file in, raw PCM file out through the C ABI from plain C,
like the reference's filter process + bfio_file -- against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from brutefir_b200 import configs
from oracle import pyoracle as po
from helpers import unpack_run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_streams_files_like_bfio_file(gpu_lib, oracle_libs, tmp_path):
    exe = os.path.join(ROOT, "host", "bfcuda_run")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L, P, n = 512, 6, 3
    g = configs.diagonal_graph(n, L, P, 4, "S24_4LE")
    taps = configs.synthetic_filters(g, 21)
    sig = configs.synthetic_signal(g, 21, 9, sigma=0.02)
    raw = sig.reshape(-1)[: sig.size - 1000]                  # last block is partial: must be zero-filled
    (tmp_path / "in.raw").write_bytes(raw.tobytes())
    (tmp_path / "taps.f32").write_bytes(np.concatenate(taps).astype("<f4").tobytes())
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-i", "S24_4LE", "-o", "S24_4LE", "-b",
                        "-c", str(tmp_path / "taps.f32"), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "realtime multiple" in r.stderr and "convolve" in r.stderr
    out = np.frombuffer((tmp_path / "out.raw").read_bytes(), np.uint8).reshape(9, g.out_bytes)
    padded = sig.copy().reshape(-1)
    padded[sig.size - 1000:] = 0
    d = po.BlockDriver("oracle", g)
    for c, h in enumerate(taps):
        d.coeff_from_taps(c, h)
    ref = d.run(padded.reshape(9, g.in_bytes))
    d.close()
    assert np.abs(unpack_run(out, g.out_formats, L) - unpack_run(ref, g.out_formats, L)).max() <= 1

    # offline mode: four blocks per call, two calls in flight -- the same bytes
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-i", "S24_4LE", "-o", "S24_4LE", "-B", "4",
                        "-c", str(tmp_path / "taps.f32"), str(tmp_path / "in.raw"), str(tmp_path / "out4.raw")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "out4.raw").read_bytes() == (tmp_path / "out.raw").read_bytes()

    # the real-time schedule (-l): within 1 LSB of the oracle as well
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-i", "S24_4LE", "-o", "S24_4LE", "-l",
                        "-c", str(tmp_path / "taps.f32"), str(tmp_path / "in.raw"), str(tmp_path / "out_l.raw")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out_l = np.frombuffer((tmp_path / "out_l.raw").read_bytes(), np.uint8).reshape(9, g.out_bytes)
    assert np.abs(unpack_run(out_l, g.out_formats, L) - unpack_run(ref, g.out_formats, L)).max() <= 1

    # filter_process()'s per-block pattern (-R: snapshot of every filter, one synchronous block, peak meter of every
    # output) on the real-time schedule: same samples within 1 LSB, and the latency line is printed
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-i", "S24_4LE", "-o", "S24_4LE", "-l", "-R",
                        "-c", str(tmp_path / "taps.f32"), str(tmp_path / "in.raw"), str(tmp_path / "out_r.raw")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "per-block call latency" in r.stderr
    out_r = np.frombuffer((tmp_path / "out_r.raw").read_bytes(), np.uint8).reshape(9, g.out_bytes)
    assert np.abs(unpack_run(out_r, g.out_formats, L) - unpack_run(ref, g.out_formats, L)).max() <= 1

    # text files on both sides (bfio_file `text: true`, bfio_file.c:153-185, 308-420, 509-565): FLOAT64 samples as
    # white-space separated numbers in, "%+.16e" tab separated frames out; unit pulses: out == in within 1e-12
    rng = np.random.default_rng(5)
    frames = rng.uniform(-1, 1, (3 * L + 17, n))                   # the last block is partial
    with open(tmp_path / "in.txt", "w") as f:
        for k, fr in enumerate(frames):
            f.write(("\t" if k % 2 else " ").join(repr(float(v)) for v in fr) + ("\n\n" if k % 7 == 0 else "\n"))
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-r", "64", "-t", str(tmp_path / "in.txt"),
                        str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = (tmp_path / "out.txt").read_text().splitlines()
    assert len(lines) == 4 * L and all(len(ln.split("\t")) == n for ln in lines) and lines[0][0] in "+-"
    got = np.array([[float(v) for v in ln.split("\t")] for ln in lines])
    assert np.abs(got[:len(frames)] - frames).max() <= 1e-12 and np.abs(got[len(frames):]).max() <= 1e-12
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-t", "-o", "S16_LE", str(tmp_path / "in.txt"),
                        str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "No support for text conversion" in r.stderr

    # "dirac pulse" coefficients: the output file equals the input file (bfconf.c:1905-1913)
    r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P), "-r", "64", str(tmp_path / "in.raw"),
                        str(tmp_path / "out2.raw")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out2 = np.frombuffer((tmp_path / "out2.raw").read_bytes(), np.uint8)
    assert np.array_equal(out2, padded)


def test_c_host_coefficient_file_formats(gpu_lib, oracle_libs, tmp_path):
    """load_coeff's file formats (bfconf.c:1725-1821, 1867-2030) in the C host: text (one number per line, blank lines
    skipped), raw integer samples scaled by the format's scale, `skip:` and `attenuation:`; short files are zero
    extended to whole blocks."""
    exe = os.path.join(ROOT, "host", "bfcuda_run")
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True).returncode == 0
    L, P, n = 256, 4, 2
    g = configs.diagonal_graph(n, L, P, 4, "S24_4LE")
    sig = configs.synthetic_signal(g, 23, 7, sigma=0.02)
    (tmp_path / "in.raw").write_bytes(sig.tobytes())
    rng = np.random.default_rng(23)

    def run_host(extra, taps, scale):
        r = subprocess.run([exe, "-n", str(n), "-L", str(L), "-P", str(P)] + extra +
                           [str(tmp_path / "in.raw"), str(tmp_path / "out.raw")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        out = np.frombuffer((tmp_path / "out.raw").read_bytes(), np.uint8).reshape(7, g.out_bytes)
        d = po.BlockDriver("oracle", g)
        for c, h in enumerate(taps):
            d.coeff_from_taps(c, h, scale)
        ref = d.run(sig)
        d.close()
        assert np.abs(unpack_run(ref, g.out_formats, L)).max() > 1e4
        assert np.abs(unpack_run(out, g.out_formats, L) - unpack_run(ref, g.out_formats, L)).max() <= 1

    # text: the second filter is short (zero extended); blank lines and leading blanks are skipped
    t0 = (rng.standard_normal(L * P) / 8).astype(np.float32)
    t1 = (rng.standard_normal(L * P) / 8).astype(np.float32)
    t1[300:] = 0
    lines = ["  %.9e" % v for v in t0] + [""] + ["\t%.9e" % v for v in t1[:300]] + ["0.0"] * (L * P - 300)
    (tmp_path / "taps.txt").write_text("\n".join(lines) + "\n")
    run_host(["-c", str(tmp_path / "taps.txt"), "-f", "text"], [t0, t1], 1.0)
    # raw S16_BE behind a 44-byte header, 6 dB attenuation
    q = rng.integers(-8000, 8000, size=(2, L * P)).astype(">i2")
    (tmp_path / "taps.s16").write_bytes(b"H" * 44 + q.tobytes())
    taps = [(q[c].astype(np.float32) * np.float32(2.0 ** -15)).astype(np.float32) for c in range(2)]
    run_host(["-c", str(tmp_path / "taps.s16"), "-f", "S16_BE", "-k", "44", "-a", "6.0"], taps, 10.0 ** (-6.0 / 20.0))


def test_two_gpus_sharded_and_nccl_output_sum(gpu_lib):
    """tests/checks/multi_gpu_check.py under torchrun on 2 GPUs (skipped on a single-GPU box): the diagonal graph
    sharded by filter with no exchange, and the xtc topology with the filters of each output split over the ranks --
    time-domain blocks summed with ncclAllReduce over NVLink before quantisation, through a crossfaded swap."""
    import sys
    if gpu_lib.bfcuda_device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "checks", "multi_gpu_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29537", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("matrix", [False, True])
@pytest.mark.parametrize("gpus,B", [(2, 1), (3, 4)])
def test_multi_gpu_c_host_equals_the_single_engine_host(gpu_lib, tmp_path, matrix, gpus, B):
    """host/bfcuda_multi.c: load_balance_filters' grouping (bfconf.c:2227-2318) dealt over `gpus` engines from plain C
    (engine k on device k modulo the devices present), per-engine blocks of its own channels, host fan-out and
    gather -- byte-identical to the single-engine host on the same files."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L, P, n = 256, 5, 6
    rng = np.random.default_rng(77)
    n_filters = n * n if matrix else n
    taps = (rng.standard_normal((n_filters, L * P)) * np.exp(-np.arange(L * P) / (L * P / 4.0)) * 0.05).astype("<f4")
    g = configs.diagonal_graph(n, L, P, 4, "S24_4LE")
    sig = configs.synthetic_signal(g, 22, 11, sigma=0.02)
    (tmp_path / "in.raw").write_bytes(sig.tobytes()[: sig.size - 500])
    (tmp_path / "taps.f32").write_bytes(taps.tobytes())
    common = ["-n", str(n), "-L", str(L), "-P", str(P), "-c", str(tmp_path / "taps.f32")] + (["-m"] if matrix else [])
    one = subprocess.run([os.path.join(ROOT, "host", "bfcuda_run")] + common + [str(tmp_path / "in.raw"), str(tmp_path / "one.raw")],
                         capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stderr
    many = subprocess.run([os.path.join(ROOT, "host", "bfcuda_multi"), "-g", str(gpus), "-B", str(B)] + common +
                          [str(tmp_path / "in.raw"), str(tmp_path / "many.raw")], capture_output=True, text=True, timeout=300)
    assert many.returncode == 0, many.stderr
    a, b = (tmp_path / "one.raw").read_bytes(), (tmp_path / "many.raw").read_bytes()
    assert len(a) == len(b) == 11 * g.out_bytes and a == b
    assert np.abs(unpack_run(np.frombuffer(a, np.uint8).reshape(11, -1), g.out_formats, L)).max() > 1e4


def test_c_host_processed_and_shared_memory_coefficients(gpu_lib, tmp_path):
    """`blocks:` (-K), the "processed" coefficient format and System V shared-memory coefficients of load_coeff
    (bfconf.c:823-826, 1924-1957, 1825-1865) in the C host: the same bytes as loading the taps."""
    import ctypes
    from brutefir_b200.engine import Engine
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = os.path.join(ROOT, "host", "bfcuda_run")
    L, P, K, n = 256, 6, 4, 2
    g = configs.diagonal_graph(n, L, P, 4, "S24_4LE", coeff_blocks=K)
    taps = configs.synthetic_filters(g, 27)
    sig = configs.synthetic_signal(g, 27, 10, sigma=0.02)
    (tmp_path / "in.raw").write_bytes(sig.tobytes())
    (tmp_path / "taps.f32").write_bytes(np.concatenate(taps).astype("<f4").tobytes())
    common = [exe, "-n", str(n), "-L", str(L), "-P", str(P), "-K", str(K)]
    r = subprocess.run(common + ["-c", str(tmp_path / "taps.f32"), str(tmp_path / "in.raw"), str(tmp_path / "a.raw")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    # the processed blocks, in the reference's layout, from an engine that loaded the same taps
    with Engine(g) as e:
        for c, h in enumerate(taps):
            e.coeff_from_taps(c, h)
        blocks = np.concatenate([e.coeff_get_block(c, b) for c in range(n) for b in range(K)]).astype("<f4")
    (tmp_path / "proc.bin").write_bytes(blocks.tobytes())
    r = subprocess.run(common + ["-f", "processed", "-c", str(tmp_path / "proc.bin"), str(tmp_path / "in.raw"),
                                 str(tmp_path / "b.raw")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "a.raw").read_bytes() == (tmp_path / "b.raw").read_bytes()
    # ... and the same blocks in two System V shared memory segments (filename: ID/OFFSET/BLOCKS, shared_mem: true)
    libc = ctypes.CDLL(None, use_errno=True)
    libc.shmat.restype = ctypes.c_void_p
    libc.shmat.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    libc.shmdt.argtypes = [ctypes.c_void_p]
    raw = blocks.tobytes()
    half = (n * K // 2) * 2 * L * 4
    ids = []
    try:
        spec = []
        for part, off in ((raw[:half], 64), (raw[half:], 0)):
            shmid = libc.shmget(0, len(part) + off, 0o1000 | 0o600)        # IPC_PRIVATE, IPC_CREAT
            assert shmid >= 0, ctypes.get_errno()
            ids.append(shmid)
            p = libc.shmat(shmid, None, 0)
            ctypes.memmove(p + off, part, len(part))
            libc.shmdt(p)
            spec.append(f"{shmid}/{off}/{len(part) // (2 * L * 4)}")
        r = subprocess.run(common + ["-c", "shm:" + ",".join(spec), str(tmp_path / "in.raw"), str(tmp_path / "c.raw")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
    finally:
        for shmid in ids:
            libc.shmctl(shmid, 0, None)       # IPC_RMID
    assert (tmp_path / "a.raw").read_bytes() == (tmp_path / "c.raw").read_bytes()
