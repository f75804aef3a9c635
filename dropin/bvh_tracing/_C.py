"""`bvh_tracing._C`: the three functions the reference binds in bvh/src/bindings.cpp:8-13."""
from relightable3dgaussian_b200._C_bvh import create_bvh, trace_bvh, trace_bvh_opacity  # noqa: F401
