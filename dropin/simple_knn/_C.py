"""`simple_knn._C`: the one function the reference binds in submodules/simple-knn/ext.cpp:15-17."""
from relightable3dgaussian_b200._C_knn import distCUDA2  # noqa: F401
