"""inference_tools_b200 -- B200-native Gaussian-process regression behind the inference-tools
``GpRegressor`` API.

    from inference_tools_b200.gp import GpRegressor, SquaredExponential, RationalQuadratic, WhiteNoise

Every array operation of the reference's GP path (covariance assembly, Cholesky, triangular solves,
inverse, gradient traces, batched prediction, expected improvement) runs in hand-written sm_100a CUDA
kernels inside ``libgpb200.so`` (C ABI in ``include/gpb200.h``), bound here with ctypes.  There is no
CPU fallback: importing :mod:`inference_tools_b200.gp` works anywhere, but creating an engine without
the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
