"""B200-native YOLOv2 (Darknet) detection forward pass.

The product is `libyolo2_b200.so` (sm_100a CUDA kernels + a C host runtime exporting the
reference's own C API, see include/); this package is the Python mirror of that API used by
the tests and the benchmark.  There is no CPU or PyTorch fallback anywhere in it.
"""
__version__ = "0.1.0"
