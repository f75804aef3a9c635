"""Drop-in `simple_knn` package (scene/gaussian_model.py:13 does `from simple_knn._C import distCUDA2`)."""
