"""avir_b200 -- B200-native execution of AVIR's resize hot path.

The product is native: ``libavirb200.so`` (hand-written sm_100a kernels behind the C ABI in
``include/avirb200.h``) and the header-only C++ front-ends ``include/avir_b200.h`` /
``include/lancir_b200.h`` that keep upstream's ``avir::CImageResizer<>`` / ``avir::CLancIR``
API.  This Python package is only a thin ctypes mirror of that API for the tests, the
benchmark and torch-side buffer management; it performs no image arithmetic and has no CPU
fallback -- every call goes to the CUDA library and raises if that fails.
"""
from .api import (  # noqa: F401
    FP_DEF, FP_FLOAT4, FP_FLOAT8_DIL, FP_DEF_ERRD, FP_FLOAT4_ERRD, FP_FLOAT8_DIL_ERRD, CImageResizer, CImageResizerVars, CLancIR,
    CLancIRParams, AvirB200Error, lib, host_lib, device_count, set_option,
    OPT_KERNEL_FAMILY, OPT_STREAM_VARIANT_H, OPT_STREAM_VARIANT_V, OPT_HOST_BANDS, OPT_ALL_STREAM_CHAINS, OPT_OVERLAP_HALO,
)
