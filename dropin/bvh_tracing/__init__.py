"""Drop-in `bvh_tracing` package: the reference's bvh/__init__.py:9 does `from bvh_tracing import _C`."""
from relightable3dgaussian_b200 import _C_bvh as _C  # noqa: F401
from relightable3dgaussian_b200.raytracer import RayTracer  # noqa: F401
