"""parpe_b200 -- B200-native objective/gradient hot path of parPE.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels and the C
ABI), ``codegen/`` + ``models/`` (model functions as device code), ``capi``
(ctypes binding of ``include/parpe_b200.h``) and ``problem`` (host-side mirror
of the reference's ``AmiciSummedGradientFunction`` interface).
"""
__version__ = "0.1.0"
