"""lisflood_code_b200 -- B200-native (sm_100a) implementation of LISFLOOD's raster time-step hot path.

Host side stays Python and mirrors the reference's module paths and operator API
(ec-jrc/lisflood-code: src/lisflood/hydrological_modules/...); the arithmetic runs in hand-written CUDA
kernels behind the C ABI declared in include/lisflood_b200.h (lisflood_code_b200/csrc/liblisf_b200.so).
There is no CPU fallback: importing works anywhere, but every compute entry point raises when the
CUDA library or an sm_100 device is missing.
"""
__version__ = "0.1.0"
