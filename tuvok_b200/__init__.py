"""tuvok_b200 -- B200 (sm_100a) brick-pool volume raycaster behind Tuvok's renderer interfaces.

The product is tuvok_b200/libtvkcuda.so (C ABI in include/tvk.h, kernels in tuvok_b200/csrc/).
This package is the thin Python host mirror of the reference renderer interface used by the
tests and bench.py.  It never imports the CPU oracle (oracle/) and has no CPU fallback.
"""
from ._lib import (BI_CHILD_EMPTY, BI_EMPTY, BI_FLAG_COUNT, BI_MISSING, BS_ONLY_NEEDED, BS_REQUEST_ALL,
                   BS_SKIP_ONE_LEVEL, BS_SKIP_TWO_LEVELS, F32, LIB_PATH, RM_1DTRANS, RM_2DTRANS, RM_ISOSURFACE, RGBA8, U8,
                   U16, TvkError, lib)
from .renderer import CudaGridLeaper, mip_ortho_projection, mip_rotation, rotation_x, rotation_y, translation
from .tf import TransferFunction1D, TransferFunction2D

__all__ = ["CudaGridLeaper", "TransferFunction1D", "TransferFunction2D", "TvkError", "lib", "LIB_PATH",
           "rotation_x", "rotation_y", "translation", "mip_rotation", "mip_ortho_projection", "U8", "U16", "F32", "RGBA8", "RM_1DTRANS", "RM_2DTRANS",
           "RM_ISOSURFACE", "BI_MISSING", "BI_CHILD_EMPTY", "BI_EMPTY", "BI_FLAG_COUNT", "BS_ONLY_NEEDED",
           "BS_REQUEST_ALL", "BS_SKIP_ONE_LEVEL", "BS_SKIP_TWO_LEVELS"]
