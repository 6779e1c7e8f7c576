"""brutefir_b200 -- B200-native implementation of BruteFIR's partitioned-convolution hot path.

The product is the CUDA library ``libbfcuda.so`` (brutefir_b200/csrc, C ABI in include/); this package is
the thin host-side mirror of the reference's interfaces for that path:

* :mod:`brutefir_b200.engine`     block-level engine (include/bfcuda.h)
* :mod:`brutefir_b200.convolver`  per-call convolver.h surface (include/bfcuda_convolver.h)
* :mod:`brutefir_b200.graph`      filter graph (struct bfconf's hot-path part)
* :mod:`brutefir_b200.formats`    sample formats and raw block layouts (bfconf.c:358-533, dai.c:537-576)
* :mod:`brutefir_b200.configs`    the BASELINE.json workloads
* :mod:`brutefir_b200.sharding`   filter-group sharding over GPUs (bfconf.c:2227-2318)
"""
from .formats import BufferFormat, SampleFormat, interleaved_layout, parse_sample_format, planar_layout  # noqa: F401
from .graph import Filter, FilterGraph  # noqa: F401

__all__ = ["BufferFormat", "SampleFormat", "Filter", "FilterGraph", "interleaved_layout", "planar_layout",
           "parse_sample_format"]
