"""Launch counter shared by the kernel wrappers (``bench.py`` reports it as ``gpu_launches``)."""
LAUNCHES = {"count": 0}      # kernels of this package launched so far
