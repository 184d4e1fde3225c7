"""ars_b200 -- B200-native (sm_100a) render hot path of Audio Raytracing Studio.

`ars_b200.raytracer_studio` mirrors the reference module's hot-path functions; the array
work runs in libars_b200.so (hand-written CUDA behind a C ABI, include/ars_b200.h).
"""
from . import _capi
from ._capi import ArsError
from . import raytracer_studio

__all__ = ["raytracer_studio", "ArsError", "_capi"]
__version__ = "0.1.0"
