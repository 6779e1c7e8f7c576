"""Boundary path A, linked for real: the reference's UNMODIFIED delay.c + firwindow.c + emalloc.c object code against
libbfcuda.so's convolver_* symbols (oracle/_ref/delay_link_gpu, built by oracle/Makefile from the sources under
/root/reference), next to the same objects on the reference's own convolver (oracle/_ref/delay_link_ref).

delay_subsample_init builds 199 windowed-sinc filters through convolver_td_new and delay_subsample_update pushes
fragments through convolver_td_convolve (delay.c:415-510, fftw_convolver.c:682-782): every one of those calls lands in
the CUDA library.  Outputs must agree within FFT rounding."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "delay_link_gpu")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "delay_link_ref")

needs_bins = pytest.mark.skipif(not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN)),
                                reason="oracle/_ref link binaries not built (needs /root/reference at build time)")


@needs_bins
def test_reference_objects_bind_to_the_cuda_library():
    """CPU: the dynamic symbol table of the linked binary takes convolver_* from libbfcuda.so."""
    ldd = subprocess.run(["ldd", GPU_BIN], capture_output=True, text=True).stdout
    assert "libbfcuda.so" in ldd and "libbfref" not in ldd
    nm = subprocess.run(["nm", "-D", "--undefined-only", GPU_BIN], capture_output=True, text=True).stdout
    for sym in ("convolver_init", "convolver_td_new", "convolver_td_convolve", "convolver_td_block_length"):
        assert sym in nm, sym
    lib = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "brutefir_b200", "libbfcuda.so")],
                         capture_output=True, text=True).stdout
    for sym in ("convolver_init", "convolver_td_new", "convolver_td_convolve", "convolver_td_block_length"):
        assert f" T {sym}" in lib, sym
    # and the reference build of the same driver runs here
    out = subprocess.run([REF_BIN, "4", "1024", "15"], capture_output=True)
    assert out.returncode == 0 and len(out.stdout) == 8 * 1024 * 4


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("rs,fragment,half", [(4, 1024, 15), (8, 256, 7), (4, 4096, 31)])
def test_unmodified_delay_c_runs_on_the_cuda_convolver(gpu_lib, rs, fragment, half):
    args = [str(rs), str(fragment), str(half)]
    ref = subprocess.run([REF_BIN] + args, capture_output=True)
    gpu = subprocess.run([GPU_BIN] + args, capture_output=True)
    assert ref.returncode == 0, ref.stderr.decode()
    assert gpu.returncode == 0, gpu.stderr.decode()
    dt = np.float32 if rs == 4 else np.float64
    r, g = np.frombuffer(ref.stdout, dt), np.frombuffer(gpu.stdout, dt)
    assert r.size == g.size == 8 * fragment and np.abs(r).max() > 0.1
    assert np.abs(r - g).max() <= (2e-6 if rs == 4 else 1e-13)
