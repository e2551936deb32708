"""ctypes binding of the CPU oracle (oracle/libvct_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvct_oracle.so")
MAX_CONES = 16


class OrcParams(C.Structure):
    _fields_ = [
        ("VoxelDimensions", C.c_int32), ("VoxelGridWorldSize", C.c_float), ("ShadowMapSize", C.c_int32),
        ("screen_width", C.c_int32), ("screen_height", C.c_int32),
        ("ModelMatrix", C.c_float * 16), ("ModelViewMatrix", C.c_float * 16), ("ProjectionMatrix", C.c_float * 16),
        ("DepthModelViewProjectionMatrix", C.c_float * 16),
        ("ProjX", C.c_float * 16), ("ProjY", C.c_float * 16), ("ProjZ", C.c_float * 16),
        ("CameraPosition", C.c_float * 3), ("LightDirection", C.c_float * 3), ("ambientFactor", C.c_float),
        ("NumDiffuseCones", C.c_int32), ("ConeDirections", C.c_float * (MAX_CONES * 3)),
        ("ConeWeights", C.c_float * MAX_CONES),
        ("DiffuseTanHalfAngle", C.c_float), ("SpecularTanHalfAngle", C.c_float), ("StepMultiplier", C.c_float),
        ("MaxDistance", C.c_float), ("MaxAlpha", C.c_float), ("PcfRadius", C.c_int32), ("ShadowBias", C.c_float),
        ("CoveragePolicy", C.c_int32), ("VoxelStoreMode", C.c_int32), ("Bounces", C.c_int32),
        ("FilterMode", C.c_int32), ("GridFormat", C.c_int32), ("RasterOrigin", C.c_int32),
    ]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("vct_oracle.cpp", "vct_oracle.h", "Makefile")]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={k: v for k, v in os.environ.items() if k != "CXX"})
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_cone_samples.restype = C.c_uint64
        L.orc_fragment_count.restype = C.c_uint64
        L.orc_pcf.restype = C.c_float
        for name in ("orc_destroy", "orc_set_params", "orc_upload_texture", "orc_set_material", "orc_upload_mesh",
                     "orc_draw_depth", "orc_draw_voxels", "orc_draw_voxels_range", "orc_resolve_and_mip", "orc_render",
                     "orc_get_depth", "orc_get_counts", "orc_get_sums", "orc_set_accum", "orc_get_grid",
                     "orc_set_grid_level0", "orc_build_mips", "orc_get_visibility", "orc_get_frame", "orc_cone_samples",
                     "orc_fragment_count", "orc_sample_voxels", "orc_cone", "orc_sample_texture", "orc_pcf"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def set_num_threads(n=0):
    """OpenMP threads of the oracle (n > 0 sets them); returns the number in effect."""
    return int(lib().orc_set_num_threads(int(n)))


class Oracle:
    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.p = OrcParams()
        self.L.orc_default_params(C.byref(self.p))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters -----------------------------------------------------------------------------
    def set_uniforms(self, u: dict):
        p = self.p
        for k, v in u.items():
            if k == "ConeDirections":
                d = np.asarray(v, dtype=np.float32).reshape(-1, 3)
                p.NumDiffuseCones = d.shape[0]
                for i, x in enumerate(d.reshape(-1)):
                    p.ConeDirections[i] = float(x)
            elif k == "ConeWeights":
                for i, x in enumerate(np.asarray(v, dtype=np.float32).reshape(-1)):
                    p.ConeWeights[i] = float(x)
            else:
                cur = getattr(p, k)
                if isinstance(cur, (int, float)):
                    setattr(p, k, type(cur)(v))
                else:
                    a = np.asarray(v, dtype=np.float32).reshape(-1)
                    assert len(a) == len(cur), k
                    for i, x in enumerate(a):
                        cur[i] = float(x)
        rc = self.L.orc_set_params(self.h, C.byref(p))
        if rc:
            raise ValueError("orc_set_params rejected the parameters")

    def load_scene(self, scene):
        for i, t in enumerate(scene.textures):
            self.upload_texture(i, t)
        for i, m in enumerate(scene.materials):
            self.L.orc_set_material(self.h, i, int(m[0]), int(m[1]), int(m[2]), C.c_float(m[3]))
        self.upload_mesh(scene.verts, scene.idx, scene.tri_material)

    def upload_texture(self, i, t):
        t = np.ascontiguousarray(t, dtype=np.uint8)
        h, w = t.shape[:2]
        c = 1 if t.ndim == 2 else t.shape[2]
        assert self.L.orc_upload_texture(self.h, i, w, h, c, _p(t, C.c_uint8)) == 0

    def upload_mesh(self, verts, idx, tri_material=None):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 14)
        ix = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        tm = None if tri_material is None else np.ascontiguousarray(tri_material, dtype=np.uint16)
        rc = self.L.orc_upload_mesh(self.h, _p(v, C.c_float), C.c_size_t(v.shape[0]), _p(ix, C.c_uint32),
                                    C.c_size_t(ix.shape[0]), None if tm is None else _p(tm, C.c_uint16))
        assert rc == 0

    # ---- passes ---------------------------------------------------------------------------------
    def draw_depth(self):
        assert self.L.orc_draw_depth(self.h) == 0

    def draw_voxels(self):
        assert self.L.orc_draw_voxels(self.h) == 0

    def draw_voxels_range(self, tb, te, clear_first=True):
        assert self.L.orc_draw_voxels_range(self.h, C.c_size_t(tb), C.c_size_t(te), int(clear_first)) == 0

    def resolve_and_mip(self):
        assert self.L.orc_resolve_and_mip(self.h) == 0

    def render(self):
        assert self.L.orc_render(self.h) == 0

    def render_rows(self, y0, y1):
        assert self.L.orc_render_rows(self.h, int(y0), int(y1)) == 0

    # ---- read-back ------------------------------------------------------------------------------
    @property
    def V(self):
        return self.p.VoxelDimensions

    def depth(self):
        S = self.p.ShadowMapSize
        a = np.empty((S, S), dtype=np.uint32)
        assert self.L.orc_get_depth(self.h, _p(a, C.c_uint32)) == 0
        return a

    def counts(self):
        V = self.V
        a = np.empty((V, V, V), dtype=np.uint32)
        self.L.orc_get_counts(self.h, _p(a, C.c_uint32))
        return a

    def sums(self):
        V = self.V
        a = np.empty((V, V, V, 3), dtype=np.uint32)
        self.L.orc_get_sums(self.h, _p(a, C.c_uint32))
        return a

    def set_accum(self, counts, sums):
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        s = np.ascontiguousarray(sums, dtype=np.uint32)
        self.L.orc_set_accum(self.h, _p(c, C.c_uint32), _p(s, C.c_uint32))

    def grid(self, level=0):
        """(n,n,n,4) uint8 for RGBA8 grids, float16 for RGBA16F grids"""
        n = self.V >> level
        a = np.empty((n, n, n, 4), dtype=np.float16 if self.p.GridFormat == 1 else np.uint8)
        assert self.L.orc_get_grid(self.h, level, a.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        return a

    def set_grid_level0(self, rgba, build_mips=True):
        a = np.ascontiguousarray(rgba, dtype=np.float16 if self.p.GridFormat == 1 else np.uint8)
        assert a.size == self.V ** 3 * 4
        self.L.orc_set_grid_level0(self.h, a.ctypes.data_as(C.POINTER(C.c_uint8)))
        if build_mips:
            self.L.orc_build_mips(self.h)

    def visibility(self):
        a = np.empty((self.p.screen_height, self.p.screen_width), dtype=np.uint32)
        assert self.L.orc_get_visibility(self.h, _p(a, C.c_uint32)) == 0
        return a

    def frame(self):
        a = np.empty((self.p.screen_height, self.p.screen_width, 4), dtype=np.uint8)
        assert self.L.orc_get_frame(self.h, _p(a, C.c_uint8)) == 0
        return a

    def cone_samples(self):
        return int(self.L.orc_cone_samples(self.h))

    def fragment_count(self):
        return int(self.L.orc_fragment_count(self.h))

    def debug_pixel(self, i, j):
        out = (C.c_float * 110)()
        rgba = (C.c_uint8 * 4)()
        rc = self.L.orc_debug_pixel(self.h, int(i), int(j), out, rgba)
        a = np.array(out[:], dtype=np.float32)
        return rc, a[:88].reshape(8, 11), a[88:], np.array(rgba[:])

    # ---- probes ---------------------------------------------------------------------------------
    def sample_voxels(self, pos, lod):
        p = (C.c_float * 3)(*map(float, pos))
        o = (C.c_float * 4)()
        self.L.orc_sample_voxels(self.h, p, C.c_float(lod), o)
        return np.array(o[:], dtype=np.float32)

    def cone(self, start, direction, tan_half):
        s = (C.c_float * 3)(*map(float, start))
        d = (C.c_float * 3)(*map(float, direction))
        o = (C.c_float * 4)()
        n = C.c_int(0)
        self.L.orc_cone(self.h, s, d, C.c_float(tan_half), o, C.byref(n))
        return np.array(o[:], dtype=np.float32), n.value

    def select_axis(self, w0, w1, w2):
        f = lambda v: (C.c_float * 3)(*map(float, v))
        return int(self.L.orc_select_axis(f(w0), f(w1), f(w2)))

    def sample_texture(self, tex, u, v, lod):
        o = (C.c_float * 4)()
        self.L.orc_sample_texture(self.h, int(tex), C.c_float(u), C.c_float(v), C.c_float(lod), o)
        return np.array(o[:], dtype=np.float32)

    def pcf(self, dc):
        d = (C.c_float * 4)(*map(float, dc))
        return float(self.L.orc_pcf(self.h, d))
