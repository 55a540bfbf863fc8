"""Drop-in replacement for the reference's `simple_knn` plugin (README.md:39-44): `from simple_knn._C import distCUDA2`."""
