"""vct_b200 -- B200-native (sm_100a CUDA) implementation of the three data-parallel passes of
AlerianEmperor/Voxel-Cone-Tracing behind the reference's own `Voxel_Cone_Tracing` interface.

Layout: csrc/ (CUDA kernels + C ABI, built into lib/libvct_b200.so), host/ (C++ facade mirroring
the reference struct), capi.py (ctypes binding of include/vct_c_api.h), renderer.py (Python mirror
of the reference struct), scenes.py / uniforms.py / glmath.py (synthetic inputs and the reference's
host-side matrix set-up)."""
from . import capi, glmath, images, objloader, parallel, renderer, scenes, uniforms  # noqa: F401
from .capi import Context, VctError, load_library  # noqa: F401
from .renderer import Camera, Voxel_Cone_Tracing  # noqa: F401

__all__ = ["capi", "glmath", "images", "objloader", "parallel", "renderer", "scenes", "uniforms", "Context", "VctError", "load_library", "Camera",
           "Voxel_Cone_Tracing"]
