"""ctypes view of include/bfcuda.h (the C ABI of libbfcuda.so) and the library loader.

The structures are layout-identical to the C ones; the reference types they mirror are named in the
header (dai.h:21-34, bfmod.h:99-133).  There is no fallback: if the shared library is missing the
import of anything that computes fails with an explicit error.
"""
from __future__ import annotations

import ctypes as C
import os

IN, OUT = 0, 1

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFCUDA_LIB", os.path.join(PKG_DIR, "libbfcuda.so"))


class SampleFormatC(C.Structure):
    _fields_ = [("isfloat", C.c_int), ("swap", C.c_int), ("bytes", C.c_int), ("sbytes", C.c_int),
                ("scale", C.c_double), ("format", C.c_int)]


class BufferFormatC(C.Structure):
    _fields_ = [("sf", SampleFormatC), ("sample_spacing", C.c_int), ("byte_offset", C.c_int)]


class OverflowC(C.Structure):
    _fields_ = [("n_overflows", C.c_uint), ("intlargest", C.c_int32), ("largest", C.c_double),
                ("max", C.c_double)]


class FilterC(C.Structure):
    _fields_ = [("crossfade", C.c_int),
                ("n_channels", C.c_int * 2),
                ("channels", C.POINTER(C.c_int) * 2),
                ("scale", C.POINTER(C.c_double) * 2),
                ("n_filters_in", C.c_int),
                ("filters_in", C.POINTER(C.c_int)),
                ("fscale", C.POINTER(C.c_double)),
                ("coeff", C.c_int),
                ("delayblocks", C.c_int)]


class FilterControlC(C.Structure):
    _fields_ = [("coeff", C.c_int), ("delayblocks", C.c_int), ("scale", C.POINTER(C.c_double) * 2),
                ("fscale", C.POINTER(C.c_double))]


class ConfigC(C.Structure):
    _fields_ = [("filter_length", C.c_int), ("n_blocks", C.c_int), ("realsize", C.c_int),
                ("n_channels", C.c_int * 2),
                ("formats", C.POINTER(BufferFormatC) * 2),
                ("n_bytes", C.c_int * 2),
                ("n_filters", C.c_int),
                ("filters", C.POINTER(FilterC)),
                ("n_coeffs", C.c_int),
                ("coeff_n_blocks", C.POINTER(C.c_int)),
                ("safety_limit", C.c_double),
                ("device", C.c_int),
                ("flags", C.c_uint),
                ("mac_split", C.c_int),
                ("apply_dither", C.POINTER(C.c_int)),
                ("sampling_rate", C.c_int),
                ("max_dither_table_size", C.c_int),
                ("max_batch", C.c_int),
                ("powersave", C.c_int),
                ("analog_powersave", C.c_double),
                ("out_physical", C.POINTER(C.c_int))]


class InfoC(C.Structure):
    _fields_ = [("n_fft", C.c_int), ("mac_split", C.c_int), ("n_streams", C.c_int),
                ("kernels_per_block", C.c_int), ("uses_graph", C.c_int), ("sm_count", C.c_int),
                ("mac_bytes_per_block", C.c_size_t), ("device_bytes", C.c_size_t),
                ("device_name", C.c_char * 64),
                ("max_batch", C.c_int), ("mac_bytes_per_batch", C.c_size_t)]


FLAG_STAGE_TIMING = 1
FLAG_NO_GRAPH = 2
FLAG_KEEP_INPUT_SPECTRA = 4
FLAG_SERIAL_STAGES = 8
FLAG_NO_STREAM_SHARING = 16
FLAG_LOW_LATENCY = 32

DBG_INPUT_SPECTRUM, DBG_DELAYLINE, DBG_FILTER_OUTPUT, DBG_OUTPUT_TIME = 1, 2, 3, 4

# every symbol include/bfcuda.h and include/bfcuda_convolver.h declare (checked by the CPU tests)
ENGINE_SYMBOLS = [
    "bfcuda_strerror", "bfcuda_device_count", "bfcuda_create", "bfcuda_destroy",
    "bfcuda_coeff_from_taps", "bfcuda_coeff_set_block", "bfcuda_coeff_get_block",
    "bfcuda_coeff_runtime_block", "bfcuda_set_control", "bfcuda_get_overflow", "bfcuda_reset_overflow",
    "bfcuda_process_block", "bfcuda_process_block_async", "bfcuda_synchronize",
    "bfcuda_process_blocks", "bfcuda_process_blocks_async", "bfcuda_wait_previous", "bfcuda_process_blocks_device",
    "bfcuda_process_block_device", "bfcuda_device_io", "bfcuda_upload_input", "bfcuda_download_output",
    "bfcuda_upload_inputs", "bfcuda_download_outputs",
    "bfcuda_host_alloc", "bfcuda_host_free", "bfcuda_timer_start", "bfcuda_timer_stop",
    "bfcuda_stage_times", "bfcuda_set_serial_stages", "bfcuda_set_stage_timing", "bfcuda_get_info", "bfcuda_debug_read", "bfcuda_comm_unique_id",
    "bfcuda_comm_init", "bfcuda_comm_shared_outputs", "bfcuda_host_alloc_near", "bfcuda_copy_baseline",
    "bfcuda_set_subdelay", "bfcuda_set_mute",
]
class DitherStateC(C.Structure):
    """struct dither_state (dither.h:17-22) == struct bfcuda_dither_state (include/bfcuda_convolver.h)."""
    _fields_ = [("randtab_ptr", C.c_int), ("randtab", C.POINTER(C.c_int8)), ("sf", C.c_float * 2),
                ("sd", C.c_double * 2)]


CONVOLVER_SYMBOLS = [
    "convolver_init", "convolver_cbufsize", "convolver_raw2cbuf", "convolver_time2freq",
    "convolver_mixnscale", "convolver_convolve_inplace", "convolver_convolve",
    "convolver_crossfade_inplace", "convolver_convolve_add", "convolver_dirac_convolve",
    "convolver_dirac_convolve_inplace", "convolver_freq2time", "convolver_convolve_eval",
    "convolver_cbuf2raw", "convolver_coeffs2cbuf", "convolver_runtime_coeffs2cbuf",
    "convolver_verify_cbuf", "convolver_debug_dump_cbuf", "convolver_fftplan",
    "convolver_td_block_length", "convolver_td_new", "convolver_td_convolve",
    "bfcuda_convolver_td_delete", "bfcuda_convolver_set_dither_table", "bfcuda_convolver_set_host",
    "bfcuda_convolver_last_error",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen brutefir_b200/libbfcuda.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C brutefir_b200/csrc`). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    lib.bfcuda_strerror.restype = C.c_char_p
    lib.bfcuda_host_alloc.restype = C.c_void_p
    lib.bfcuda_host_alloc.argtypes = [C.c_size_t]
    lib.bfcuda_host_free.argtypes = [C.c_void_p]
    lib.bfcuda_host_alloc_near.restype = C.c_void_p
    lib.bfcuda_host_alloc_near.argtypes = [C.c_int, C.c_size_t]
    lib.bfcuda_copy_baseline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.bfcuda_create.argtypes = [C.POINTER(ConfigC), C.POINTER(C.c_void_p)]
    lib.bfcuda_destroy.argtypes = [C.c_void_p]
    lib.bfcuda_destroy.restype = None
    lib.bfcuda_coeff_from_taps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double]
    lib.bfcuda_coeff_set_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.bfcuda_coeff_get_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.bfcuda_coeff_runtime_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.bfcuda_set_control.argtypes = [C.c_void_p, C.c_int, C.POINTER(FilterControlC)]
    lib.bfcuda_set_subdelay.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.bfcuda_set_mute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.bfcuda_get_overflow.argtypes = [C.c_void_p, C.c_int, C.POINTER(OverflowC)]
    lib.bfcuda_reset_overflow.argtypes = [C.c_void_p]
    lib.bfcuda_process_block.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bfcuda_process_block_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bfcuda_synchronize.argtypes = [C.c_void_p]
    lib.bfcuda_wait_previous.argtypes = [C.c_void_p, C.c_int]
    lib.bfcuda_process_block_device.argtypes = [C.c_void_p]
    lib.bfcuda_process_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bfcuda_process_blocks_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bfcuda_process_blocks_device.argtypes = [C.c_void_p, C.c_int]
    lib.bfcuda_upload_inputs.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bfcuda_download_outputs.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bfcuda_device_io.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.bfcuda_upload_input.argtypes = [C.c_void_p, C.c_void_p]
    lib.bfcuda_download_output.argtypes = [C.c_void_p, C.c_void_p]
    lib.bfcuda_timer_start.argtypes = [C.c_void_p]
    lib.bfcuda_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.bfcuda_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long),
                                       C.POINTER(C.c_long)]
    lib.bfcuda_set_serial_stages.argtypes = [C.c_void_p, C.c_int]
    lib.bfcuda_set_stage_timing.argtypes = [C.c_void_p, C.c_int]
    lib.bfcuda_get_info.argtypes = [C.c_void_p, C.POINTER(InfoC)]
    lib.bfcuda_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.bfcuda_comm_unique_id.argtypes = [C.c_void_p]
    lib.bfcuda_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.bfcuda_comm_shared_outputs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    _lib = lib
    return lib


class BfcudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"bfcuda error {code}: {message}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().bfcuda_strerror()
        raise BfcudaError(rc, msg.decode() if msg else "")
