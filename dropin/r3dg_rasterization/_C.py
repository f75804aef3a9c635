"""`r3dg_rasterization._C`: the three functions the reference binds in ext.cpp:15-19."""
from relightable3dgaussian_b200._C_raster import (  # noqa: F401
    mark_visible, rasterize_gaussians, rasterize_gaussians_backward)
