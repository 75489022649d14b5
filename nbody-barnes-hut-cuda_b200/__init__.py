"""B200-native Barnes-Hut step engine — Python host mirror of the C ABI (include/bh.h).

The product is ``libbh.so`` (hand-written sm_100a CUDA behind ``extern "C"``); this package only
loads it with ctypes and mirrors the reference's step interface
(``simulationStep()``, /root/reference/nbody_v5_bench.cu:255-283) for tests and bench.py.
There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every
compute entry point raises if the library or the device is missing.
"""
from .engine import (  # noqa: F401
    BHEngine,
    BHError,
    BHParams,
    MultiGpu,
    mg_unique_id,
    DBG,
    PHASE,
    STAT,
    build_library,
    ic_lib,
    ic_plummer,
    ic_refdisk,
    ic_two_disks,
    ic_two_disks_range,
    ic_uniform_cube,
    lib,
    library_path,
    probe_fp32_tflops,
    probe_fp32x2_tflops,
    probe_hbm_gbs,
    sort_pairs_u32,
)
