"""The device code of brutefir_b200/csrc/*.cuh compiled for the HOST and run against the oracle:
the shared-memory FFT passes in lock-step thread emulation against a long-double DFT, and the sample
conversion / quantiser bit for bit against oracle/bf_oracle.c.  Catches kernel logic errors without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "host_emul")


def build_and_run(tmp_path, src, extra):
    exe = str(tmp_path / (src + ".bin"))
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-o", exe, os.path.join(EMUL, src)] + extra + ["-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    return r


def test_device_fft_on_host(tmp_path):
    r = build_and_run(tmp_path, "emul_fft.cpp", [])
    assert r.returncode == 0, r.stdout
    assert "FAIL" not in r.stdout and r.stdout.count("ok") >= 25


def test_device_fft2_on_host(tmp_path):
    """The size-specialised FFT core (bf_fft2.cuh): all five sizes, forward / inverse / registers-only last pass,
    and the bank-conflict check of every shared-memory access pattern."""
    r = build_and_run(tmp_path, "emul_fft2.cpp", [])
    assert r.returncode == 0, r.stdout
    assert "FAIL" not in r.stdout and r.stdout.count("ok") == 18


def test_device_sample_conversion_on_host(tmp_path, oracle_libs):
    objs = tmp_path / "orc.o"
    shim = tmp_path / "shim.o"
    for src, out in (("oracle/bf_oracle.c", objs), ("oracle/shim/fft_shim.c", shim)):
        c = subprocess.run(["gcc", "-O2", "-msse2", "-ffp-contract=off", "-c", os.path.join(ROOT, src), "-o", str(out)],
                           capture_output=True, text=True)
        assert c.returncode == 0, c.stderr
    r = build_and_run(tmp_path, "emul_sample.cpp", [str(objs), str(shim)])
    assert r.returncode == 0 and "emul_sample ok" in r.stdout, r.stdout
