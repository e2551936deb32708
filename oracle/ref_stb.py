"""The reference's own image decoder, compiled here (oracle/_ref/libstb_image_ref.so).  TEST INFRASTRUCTURE ONLY.

The reference decodes every material texture with `stbi_load(path, &w, &h, &n, 0)` (Model.h:152) from the stb_image
v2.26 it vendors (stb_image.h + stb_image.cpp, plain C, no dependencies).  Unlike the rest of the reference that part
compiles with g++ in this image, so it is built FROM THE SOURCES WHERE THEY LIE under /root/reference (nothing is copied
into the repository; the .so is git-ignored but travels to the GPU box) and used by tests/test_images_vs_stb.py as the
reference for vct_b200/images.py, byte for byte."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/Voxel_Cone_Tracing_Final/stb_image.cpp"
LIB = os.path.join(_HERE, "_ref", "libstb_image_ref.so")


def source_available():
    return os.path.exists(SRC)


def build(force=False):
    """g++ on the reference's stb_image.cpp, output only into oracle/_ref/.  Returns the path, or None when neither the
    source nor a previously built library is present."""
    if not source_available():
        return LIB if os.path.exists(LIB) else None
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # stb_image.cpp starts with `#pragma once` (harmless in a main file) and defines STB_IMAGE_IMPLEMENTATION
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-w", "-x", "c++", SRC, "-o", LIB], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libstb_image_ref.so is not built and /root/reference is absent")
        L = C.CDLL(path)
        L.stbi_load_from_memory.restype = C.POINTER(C.c_ubyte)
        L.stbi_load_from_memory.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        L.stbi_load.restype = C.POINTER(C.c_ubyte)
        L.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        L.stbi_image_free.argtypes = [C.c_void_p]
        L.stbi_failure_reason.restype = C.c_char_p
        _lib = L
    return _lib


def _take(L, p, w, h, n):
    if not p:
        raise ValueError((L.stbi_failure_reason() or b"?").decode())
    out = np.ctypeslib.as_array(p, shape=(h.value, w.value, n.value)).copy()
    L.stbi_image_free(p)
    return out


def load_from_memory(data: bytes) -> np.ndarray:
    """stbi_load_from_memory(data, ..., req_comp = 0) -> uint8 (h, w, n), as Model.h:152 calls it"""
    L = lib()
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    return _take(L, L.stbi_load_from_memory(data, len(data), C.byref(w), C.byref(h), C.byref(n), 0), w, h, n)


def load(path: str) -> np.ndarray:
    L = lib()
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    return _take(L, L.stbi_load(path.encode(), C.byref(w), C.byref(h), C.byref(n), 0), w, h, n)
