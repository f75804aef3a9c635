"""ctypes binding of libr3dg_b200.so (C ABI: include/r3dg_b200.h).

There is NO fallback: if the CUDA library is missing or does not export a declared symbol this
module raises at import time, so a GPU run can never silently use anything but the sm_100a
kernels.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# R3DG_LIB_PATH: developer knob to A/B kernel builds (tools/variant_sweep.sh); unset = the in-tree library
LIB_PATH = os.environ.get("R3DG_LIB_PATH") or os.path.join(_HERE, "libr3dg_b200.so")

c_void_p, c_int, c_float, c_size_t, c_ll = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_size_t, ctypes.c_longlong)


class RasterFwdArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("P", "S", "D", "M", "W", "H")] +
        [(n, c_void_p) for n in ("background", "means3D", "shs", "colors_precomp", "features",
                                 "opacities", "scales", "rotations", "cov3D_precomp", "viewmatrix",
                                 "projmatrix", "campos")] +
        [(n, c_float) for n in ("scale_modifier", "tan_fovx", "tan_fovy", "cx", "cy")] +
        [(n, c_int) for n in ("prefiltered", "computer_pseudo_normal", "debug")] +
        [(n, c_void_p) for n in ("out_color", "out_opacity", "out_depth", "out_feature",
                                 "out_normal", "out_surface_xyz", "out_weights", "radii",
                                 "n_contrib")] +
        [("geom", c_void_p), ("geom_bytes", c_size_t), ("img", c_void_p), ("img_bytes", c_size_t),
         ("binning", c_void_p), ("binning_bytes", c_size_t), ("num_rendered_host", c_void_p), ("count_ready_event", c_void_p)])


class RasterBwdArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("P", "S", "D", "M", "W", "H")] +
        [(n, c_void_p) for n in ("background", "means3D", "shs", "colors_precomp", "features",
                                 "scales", "rotations", "cov3D_precomp", "viewmatrix",
                                 "projmatrix", "campos")] +
        [(n, c_float) for n in ("scale_modifier", "tan_fovx", "tan_fovy")] +
        [(n, c_int) for n in ("backward_geometry", "debug")] +
        [(n, c_void_p) for n in ("dL_dout_color", "dL_dout_opacity", "dL_dout_depth",
                                 "dL_dout_feature", "dL_dmeans2D", "dL_dcolors", "dL_dopacity",
                                 "dL_dmeans3D", "dL_dfeatures", "dL_dcov3D", "dL_dsh",
                                 "dL_dscales", "dL_drotations", "dL_dsh_factor")] +
        [("geom", c_void_p), ("geom_bytes", c_size_t), ("img", c_void_p), ("img_bytes", c_size_t),
         ("binning", c_void_p), ("binning_bytes", c_size_t)])


class ShadeArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("P", "N", "sh_coeffs", "env_h", "env_w")] +
        [(n, c_void_p) for n in ("base_color", "roughness", "normals", "viewdirs", "incidents", "env",
                                 "env_transform", "visibility", "incident_dirs", "incident_areas",
                                 "pbr", "diffuse_light", "specular", "mean_incident_lights",
                                 "mean_local_lights", "mean_global_lights", "mean_visibility",
                                 "incident_lights", "local_incident_lights", "global_incident_lights",
                                 "dL_dpbr", "dL_ddiffuse_light", "dL_dspecular", "dL_dbase_color",
                                 "dL_droughness", "dL_dviewdirs", "dL_dincidents", "dL_denv")])


class ExchangeArgs(ctypes.Structure):
    _fields_ = ([(n, c_int) for n in ("P", "D", "M", "world", "rank")] +
                [("means3D", c_void_p), ("campos", c_void_p), ("factors", c_void_p * 64), ("dL_dsh", c_void_p),
                 ("n_dense", c_ll), ("dense", c_void_p * 64), ("dense_multicast", c_void_p)])


class PackSrc(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("width", c_int)]


class CompactTensor(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("row_bytes", c_ll)]


class AdamTensor(ctypes.Structure):
    _fields_ = ([(n, c_void_p) for n in ("param", "grad", "exp_avg", "exp_avg_sq")] +
                [("n", c_ll), ("step", c_ll)] +
                [(n, ctypes.c_double) for n in ("lr", "beta1", "beta2", "eps")])


# every symbol include/r3dg_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("r3dg_version", ctypes.c_char_p, []),
    ("r3dg_raster_geom_bytes", c_size_t, [c_int, c_int]),
    ("r3dg_raster_img_bytes", c_size_t, [c_int, c_int]),
    ("r3dg_raster_binning_bytes", c_size_t, [c_ll]),
    ("r3dg_raster_img_n_contrib_offset", c_size_t, [c_int, c_int]),
    ("r3dg_raster_forward", c_int, [ctypes.POINTER(RasterFwdArgs), c_void_p]),
    ("r3dg_raster_backward", c_int, [ctypes.POINTER(RasterBwdArgs), c_void_p]),
    ("r3dg_sh_grad_from_factors", c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    ("r3dg_exchange_p2p", c_int, [ctypes.POINTER(ExchangeArgs), c_void_p]),
    ("r3dg_mark_visible", c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("r3dg_bvh_build_tmp_bytes", c_size_t, [c_int]),
    ("r3dg_bvh_trace_tmp_bytes", c_size_t, [c_int]),
    ("r3dg_bvh_leaf_aabbs", c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("r3dg_bvh_build", c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    ("r3dg_bvh_trace_opacity", c_int, [c_int, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    ("r3dg_sample_incident_dirs", c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("r3dg_bvh_bake_visibility", c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    ("r3dg_unpremultiply_forward", c_int, [c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("r3dg_unpremultiply_backward", c_int, [c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("r3dg_render_equation_forward", c_int, [ctypes.POINTER(ShadeArgs), c_void_p]),
    ("r3dg_render_equation_backward", c_int, [ctypes.POINTER(ShadeArgs), c_void_p]),
    ("r3dg_pack_features_forward", c_int, [c_int, c_int, c_void_p, c_void_p, c_int, ctypes.POINTER(PackSrc), c_void_p, c_void_p]),
    ("r3dg_pack_features_backward", c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(PackSrc), c_void_p, c_void_p]),
    ("r3dg_compact_tmp_bytes", c_size_t, [c_int]),
    ("r3dg_compact_scan", c_int, [c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    ("r3dg_compact_rows", c_int, [c_int, c_int, ctypes.POINTER(CompactTensor), c_void_p, c_void_p, c_void_p]),
    ("r3dg_knn_tmp_bytes", c_size_t, [c_int]),
    ("r3dg_knn_dist2", c_int, [c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    ("r3dg_adam_step", c_int, [c_int, ctypes.POINTER(AdamTensor), c_void_p]),
    ("r3dg_launch_count", ctypes.c_ulonglong, []),
    ("r3dg_tune", c_int, [ctypes.c_char_p, c_int, ctypes.POINTER(c_int)]),
    ("r3dg_prof_begin", c_int, [c_int]),
    ("r3dg_prof_end", c_int, [ctypes.POINTER(c_float), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    ("r3dg_raster_debug_copy", c_ll, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p, c_ll, c_void_p]),
]

_lib = None


def load():
    """Load the library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found — build it with `python -m relightable3dgaussian_b200.build` "
            "(there is no CPU/PyTorch fallback for the rasterizer hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def tune(key, value):
    """Set a tuning knob (include/r3dg_b200.h: r3dg_tune); returns the previous value."""
    prev = c_int(0)
    check(load().r3dg_tune(key.encode(), int(value), ctypes.byref(prev)), f"r3dg_tune({key})")
    return prev.value


def check(rc, what):
    if rc != 0:
        if rc in (-10001, -10002):
            raise RuntimeError(f"{what}: {'bad argument' if rc == -10001 else 'unsupported configuration'} ({rc})")
        raise RuntimeError(f"{what}: CUDA error {-rc}")
