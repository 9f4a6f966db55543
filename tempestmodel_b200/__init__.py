"""tempestmodel_b200: B200-native dynamical-core timestep hot path of the
Tempest atmospheric model behind the reference's plugin interfaces.

The compute path is hand-written FP64 CUDA (sm_100a) in
``libtempest_b200.so``, reached through the C ABI of
``include/tempest_b200.h``.  There is no CPU fallback.
"""
from ._lib import (DATA_ALL, DATA_STATE, DATA_TRACERS, EQN_PRIMITIVE_NONHYDRO,
                   EQN_SHALLOW_WATER, PRODUCT_LIBRARY, LibraryMissing, load)
from .device import DeviceContext, TempestError

__all__ = ["DeviceContext", "TempestError", "LibraryMissing", "load",
           "PRODUCT_LIBRARY", "DATA_ALL", "DATA_STATE", "DATA_TRACERS",
           "EQN_SHALLOW_WATER", "EQN_PRIMITIVE_NONHYDRO"]
