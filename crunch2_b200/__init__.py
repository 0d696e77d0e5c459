"""crunch2_b200 -- B200-native (sm_100a) implementation of crnlib's data-parallel hot path.

Only what the path needs lives here: `csrc/` (hand-written CUDA kernels + the C-ABI library
libcrn_b200.so declared in include/crn_b200.h) and the thin host-side mirror of the reference's
interface for this path (`api.py`).  There is no CPU fallback: importing works anywhere, but every
compute call raises unless the nvcc-built library is present and a CUDA device is visible.
"""
from .api import (  # noqa: F401
    CrnGpuError, Context, PackParams, FMT_DXT1, FMT_DXT1A, FMT_DXT3, FMT_DXT5, FMT_DXT5A, FMT_DXN_XY, FMT_DXN_YX,
    bytes_per_block, library_path, load_library, Texture, texture_info, crn_params, crn_hc_params, crn_write,
)

__all__ = ["CrnGpuError", "Context", "PackParams", "Texture", "texture_info", "bytes_per_block", "library_path", "load_library", "crn_params", "crn_hc_params", "crn_write",
           "FMT_DXT1", "FMT_DXT1A", "FMT_DXT3", "FMT_DXT5", "FMT_DXT5A", "FMT_DXN_XY", "FMT_DXN_YX"]
