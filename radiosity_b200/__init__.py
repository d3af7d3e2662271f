"""radiosity_b200 — B200-native progressive-radiosity shooting loop.

The product is two in-tree native libraries:

* ``librad_cuda.so``      hand-written sm_100a CUDA behind the C ABI of ``include/rad_cuda.h``
* ``libradiosity_host.so`` C++17 host library with the reference's host API
  (Config, Patch, Model, PrimitiveModel, WaveFrontModel, ModelContainer, FormFactors, Camera, the
  headless driver RadiositySolver) plus the ``radiosity`` command-line driver

This Python package is only the ctypes view used by tests and ``bench.py``.  There is no Python or CPU
compute path: importing ``radiosity_b200.api`` raises if the native libraries have not been built.
"""
from .api import (  # noqa: F401
    RadConfig, RadStats, Context, Scene, RadError, cuda_lib, host_lib,
    SELECT_REFERENCE, SELECT_TOPK, FLAG_KEEP_ITEMBUFFER, FACE_NAMES,
)
