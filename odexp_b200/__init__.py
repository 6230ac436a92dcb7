"""odexp_b200 — B200-native population integrator behind odexp's model files and C stepper entry points.

Layout: popfile (model-file parser), cudagen (CUDA emitter), build (nvcc driver), capi (ctypes binding of include/odx.h),
csrc/ (kernels, runtime, C host driver).  The product path needs the compiled CUDA library; there is no CPU fallback.
"""
from .popfile import parse, parse_file, PopModel, PopError   # noqa: F401
