"""B200-native drop-in for the reference package `simple_knn` (submodules/simple-knn):
`from simple_knn._C import distCUDA2` keeps working (scene/gaussian_model.py:20)."""
