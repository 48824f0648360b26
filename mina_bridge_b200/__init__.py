"""B200-native batch verifier for Mina bridge proofs (proof-of-state / proof-of-account).

The package is a thin Python mirror of the C ABI in include/*.h; all arithmetic runs in the native
library (CUDA kernels for sm_100a + a C++ host driver).  There is no CPU fallback.
"""
from .ffi import *  # noqa: F401,F403
from .ffi import (  # noqa: F401
    MinaB200Error,
    field_op,
    host_blake2b512,
    host_field_op,
    host_srs_derive,
    init,
    launch_count,
    library_path,
    load,
    msm,
    msm_srs,
    msm_srs_device,
    msm_configure,
    point_add,
    shutdown,
    srs_points,
)

CURVE_PALLAS, CURVE_VESTA = 0, 1
FIELD_FP, FIELD_FQ = 0, 1
