"""B200-native per-timestep PIC hot path for WumingPIC2D (push, Esirkepov deposit, sort,
implicit field solve) behind a C ABI.  This package is only the Python-side binding of
include/wumingpic2d.h; the product is wumingpic2d_b200/libwumingpic2d.so (CUDA, sm_100a).
There is no CPU fallback: loading fails loudly if the library has not been built, and every
call fails if no B200-class device is present.
"""
from .api import (Context, LoopbackGroup, WmConfig, WmError, load_library, library_path,  # noqa: F401
                  WM_BC_PERIODIC, WM_BC_RECONNECTION, WM_BC_SHOCK, WM_FLAG_EXACT_PUSH)
