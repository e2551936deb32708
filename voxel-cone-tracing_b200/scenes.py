"""Seeded procedural scenes for the BASELINE.json configs.

The reference loads `sponza.obj` through assimp from a hard-coded Windows path
(/root/reference/Voxel_Cone_Tracing_Final/Voxel_Cone_Tracing.h:77, Model.h:39-139); neither the asset
nor assimp exists here, so the scenes are synthetic.  What is kept from the reference is the OUTPUT
layout of its loader: interleaved 14-float vertices `Position, Normal, TexCoords, Tangents, Bi_Tangents`
(Mesh.h:12-19), u32 triangle indices, and per-mesh diffuse/specular/height textures (Mesh.h:89-111).
Geometry is authored in MODEL units = world units * 20 so that the reference's
`ModelMatrix = scale(0.05)` (Voxel_Cone_Tracing.h:183) and its un-normalised-normal quirk apply.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F = np.float32
MODEL_SCALE = 0.05
INV_SCALE = 20.0


@dataclass
class Scene:
    name: str
    verts: np.ndarray            # (nv, 14) float32, model units
    idx: np.ndarray              # (nt, 3) uint32
    tri_material: np.ndarray     # (nt,) uint16
    textures: list               # list of (h, w, c) uint8 arrays, c in {1, 3, 4}
    materials: list              # list of (diffuse, specular, height, shininess)
    camera_pos: tuple = (0.0, 4.0, 0.0)   # Voxel_Cone_Tracing.h:8
    yaw: float = -90.0                    # Camera.h:21
    pitch: float = 0.0
    fov_deg: float = 45.0                 # Camera.h:25 (Zoom)
    meta: dict = field(default_factory=dict)

    @property
    def n_tris(self):
        return int(self.idx.shape[0])


class _Builder:
    def __init__(self):
        self.v, self.i, self.m = [], [], []
        self.nv = 0

    def patch(self, P, uv, mat, flip=False):
        """P: (rv+1, ru+1, 3) WORLD positions on a grid; uv: (rv+1, ru+1, 2)."""
        P = np.asarray(P, dtype=np.float64)
        rv, ru = P.shape[0] - 1, P.shape[1] - 1
        T = np.gradient(P, axis=1) if ru > 0 else np.zeros_like(P)
        B = np.gradient(P, axis=0) if rv > 0 else np.zeros_like(P)

        def _n(a):
            l = np.linalg.norm(a, axis=-1, keepdims=True)
            return a / np.maximum(l, 1e-20)

        T, B = _n(T), _n(B)
        N = _n(np.cross(T, B))
        if flip:
            N = -N
        V = np.concatenate([P * INV_SCALE, N, uv, T, B], axis=-1).reshape(-1, 14)
        a = (np.arange(rv)[:, None] * (ru + 1) + np.arange(ru)[None, :]).reshape(-1)
        b, c, d = a + 1, a + ru + 2, a + ru + 1
        if flip:
            tri = np.stack([np.stack([a, c, b], 1), np.stack([a, d, c], 1)], 1).reshape(-1, 3)
        else:
            tri = np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3)
        self.v.append(V.astype(F))
        self.i.append((tri + self.nv).astype(np.uint32))
        self.m.append(np.full(tri.shape[0], mat, dtype=np.uint16))
        self.nv += V.shape[0]

    def quad(self, p00, p10, p11, p01, mat, ru=1, rv=1, uv_scale=(1.0, 1.0), flip=False):
        """Bilinear quad; u runs p00->p10, v runs p00->p01.  Normal = du x dv."""
        p00, p10, p11, p01 = (np.asarray(p, dtype=np.float64) for p in (p00, p10, p11, p01))
        u = np.linspace(0, 1, ru + 1)[None, :, None]
        v = np.linspace(0, 1, rv + 1)[:, None, None]
        P = (1 - u) * (1 - v) * p00 + u * (1 - v) * p10 + u * v * p11 + (1 - u) * v * p01
        uv = np.concatenate([np.broadcast_to(u * uv_scale[0], P.shape[:2] + (1,)),
                             np.broadcast_to(v * uv_scale[1], P.shape[:2] + (1,))], -1)
        self.patch(P, uv, mat, flip)

    def box(self, lo, hi, mat, res=1, bottom=True, uv_scale=(1.0, 1.0)):
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        q = lambda a, b, c, d: self.quad(a, b, c, d, mat, res, res, uv_scale)
        q((x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0))      # top    +y
        if bottom:
            q((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1))  # bottom -y
        q((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1))      # front  +z
        q((x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0))      # back   -z
        q((x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1))      # right  +x
        q((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0))      # left   -x

    def build(self):
        return (np.concatenate(self.v, 0), np.concatenate(self.i, 0), np.concatenate(self.m, 0))


# ------------------------------------------------------------------------------------------ textures

def value_noise(size, cells, rng, octaves=3):
    """Tileable value noise in [0, 1], shape (size, size)."""
    out = np.zeros((size, size), dtype=np.float64)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        c = cells * (2 ** o)
        lat = rng.random((c, c))
        t = np.arange(size) * (c / size)
        i0 = np.floor(t).astype(int)
        f = t - i0
        f = f * f * (3 - 2 * f)
        i1 = (i0 + 1) % c
        a = lat[i0][:, i0] * (1 - f)[None, :] + lat[i0][:, i1] * f[None, :]
        b = lat[i1][:, i0] * (1 - f)[None, :] + lat[i1][:, i1] * f[None, :]
        out += amp * (a * (1 - f)[:, None] + b * f[:, None])
        tot += amp
        amp *= 0.5
    return out / tot


def _u8(a):
    return np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)


def solid_texture(rgb, channels=3):
    t = np.zeros((1, 1, channels), dtype=np.uint8)
    t[0, 0, :len(rgb)] = rgb
    if channels == 4 and len(rgb) == 3:
        t[0, 0, 3] = 255
    return t


# ------------------------------------------------------------------------------------------ config 1

def cornell(name="cornell"):
    """BASELINE config 1: Cornell-style room, world extent [-60, 60]^3, skylight in the ceiling because
    the reference's only light is directional from above (Voxel_Cone_Tracing.h:14)."""
    b = _Builder()
    W, RED, GREEN = 0, 1, 2
    s = 60.0
    b.quad((-s, -s, s), (s, -s, s), (s, -s, -s), (-s, -s, -s), W)            # floor, normal +y
    b.quad((-s, -s, -s), (s, -s, -s), (s, s, -s), (-s, s, -s), W)            # back wall, normal +z
    b.quad((-s, -s, s), (-s, -s, -s), (-s, s, -s), (-s, s, s), RED)          # left wall, normal +x
    b.quad((s, -s, -s), (s, -s, s), (s, s, s), (s, s, -s), GREEN)            # right wall, normal -x
    h = 30.0  # skylight half-size
    for x0, x1, z0, z1 in ((-s, s, -s, -h), (-s, s, h, s), (-s, -h, -h, h), (h, s, -h, h)):
        b.quad((x0, s, z0), (x1, s, z0), (x1, s, z1), (x0, s, z1), W)        # ceiling ring, normal -y
        b.quad((x0, s, z1), (x1, s, z1), (x1, s, z0), (x0, s, z0), W)        # outer side so it shadows
    b.box((-40.0, -s, -35.0), (-5.0, 15.0, 0.0), W, bottom=False)            # tall box
    b.box((8.0, -s, 5.0), (40.0, -25.0, 37.0), W, bottom=False)              # short box
    v, i, m = b.build()
    textures = [solid_texture((200, 200, 200)), solid_texture((200, 30, 30)), solid_texture((30, 200, 30)),
                solid_texture((128,), 1), solid_texture((128, 128, 128))]
    materials = [(0, 3, 4, 20.0), (1, 3, 4, 20.0), (2, 3, 4, 20.0)]
    return Scene(name, v, i, m, textures, materials, camera_pos=(0.0, 0.0, 205.0), yaw=-90.0, pitch=0.0,
                 meta={"config": 1})


# ------------------------------------------------------------------------------------------ config 2

def _cylinder(b, cx, cz, y0, y1, r, seg, rings, mat, uv_scale=(2.0, 4.0)):
    th = np.linspace(0, 2 * np.pi, seg + 1)[None, :]
    y = np.linspace(y0, y1, rings + 1)[:, None]
    # slight entasis so ring normals are not all identical
    rr = r * (1.0 - 0.12 * ((y - y0) / max(y1 - y0, 1e-6)) ** 2)
    P = np.stack([cx + rr * np.cos(-th), np.broadcast_to(y, (rings + 1, seg + 1)), cz + rr * np.sin(-th)], -1)
    uv = np.stack([np.broadcast_to(th / (2 * np.pi) * uv_scale[0], P.shape[:2]),
                   np.broadcast_to((y - y0) / max(y1 - y0, 1e-6) * uv_scale[1], P.shape[:2])], -1)
    b.patch(P, uv, mat)


def _arch(b, x0, x1, y0, z, depth, thick, seg, rad, mat):
    """Half-ring arch between two columns in the x direction at height y0, extruded `depth` in z."""
    cx, R = 0.5 * (x0 + x1), 0.5 * (x1 - x0)
    a = np.linspace(0, np.pi, seg + 1)[None, :]
    for (r, flip) in ((R - thick, True), (R, False)):
        zz = np.linspace(z - depth / 2, z + depth / 2, rad + 1)[:, None]
        P = np.stack([cx - r * np.cos(a) + 0 * zz, y0 + r * np.sin(a) + 0 * zz, zz + 0 * a], -1)
        uv = np.stack([np.broadcast_to(a / np.pi * 3.0, P.shape[:2]), np.broadcast_to((zz - z) / depth, P.shape[:2])], -1)
        b.patch(P, uv, mat, flip=flip)
    for (zf, flip) in ((z + depth / 2, False), (z - depth / 2, True)):
        rr = np.linspace(R - thick, R, 3)[:, None]
        P = np.stack([cx - rr * np.cos(a), y0 + rr * np.sin(a), zf + 0 * a + 0 * rr], -1)
        uv = np.stack([np.broadcast_to(a / np.pi * 3.0, P.shape[:2]), np.broadcast_to((rr - R) / thick, P.shape[:2])], -1)
        b.patch(P, uv, mat, flip=flip)


def atrium_textures(seed=1234, size=512, n_solid=18, n_cutout=4):
    """~25 materials: value-noise albedo in [0.2, 0.9], height = noise, greyscale or RGB specular,
    cutout materials carry a binary alpha mask with ~50 % coverage."""
    rng = np.random.default_rng(seed)
    textures, materials = [], []

    def add(t):
        textures.append(np.ascontiguousarray(t))
        return len(textures) - 1

    for k in range(n_solid + n_cutout):
        base = rng.random(3) * 0.5 + 0.3
        n = value_noise(size, 4 + (k % 5) * 2, rng, octaves=4)
        alb = 0.2 + 0.7 * np.clip(base[None, None, :] * (0.55 + 0.9 * n[..., None]), 0, 1)
        alb = np.clip(alb, 0.2, 0.9)
        if k >= n_solid:
            m = value_noise(size, 6, rng, octaves=2)
            mask = (m > np.median(m)).astype(np.float64)
            d = add(_u8(np.concatenate([alb, mask[..., None]], -1)))
        else:
            d = add(_u8(alb))
        hmap = add(_u8(value_noise(size // 2, 8, rng, octaves=3))[..., None])
        if k % 3 == 0:
            spec = add(_u8(0.15 + 0.5 * value_noise(size // 2, 5, rng, octaves=2))[..., None])       # .rrra path
        else:
            spec = add(_u8(np.clip(0.1 + 0.5 * value_noise(size // 2, 5, rng, 2)[..., None] * (0.6 + 0.4 * rng.random(3)), 0, 1)))
        materials.append((d, spec, hmap, 20.0))   # Shininess = 20 for every mesh, Mesh.h:86
    return textures, materials


def atrium(seed=1234, detail=1.0, tex_size=512, name="atrium"):
    """BASELINE config 2: Sponza-scale two-storey atrium, ~260 K triangles at detail=1.
    Bounds (world): x in [-95, 90], y in [-6, 72], z in [-58, 59] -- wider than the +-75 voxel grid in x,
    as the real asset is."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    d = lambda n: max(1, int(round(n * detail)))
    x0, x1, y0, y1, z0, z1 = -95.0, 90.0, -6.0, 72.0, -58.0, 59.0
    M_FLOOR, M_WALL, M_WALL2, M_COL, M_COL2, M_ARCH, M_SLAB, M_TRIM = 0, 1, 2, 3, 4, 5, 6, 7
    N_SOLID, N_CUT = 18, 4
    # floor and outer walls (inward facing), open roof
    b.quad((x0, y0, z1), (x1, y0, z1), (x1, y0, z0), (x0, y0, z0), M_FLOOR, d(180), d(110), (12, 8))
    b.quad((x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), M_WALL, d(120), d(48), (8, 4))
    b.quad((x1, y0, z1), (x0, y0, z1), (x0, y1, z1), (x1, y1, z1), M_WALL, d(120), d(48), (8, 4))
    b.quad((x0, y0, z1), (x0, y0, z0), (x0, y1, z0), (x0, y1, z1), M_WALL2, d(80), d(48), (6, 4))
    b.quad((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0), M_WALL2, d(80), d(48), (6, 4))
    # wall backs so the shadow pass (back-face culled) sees a closed shell from the light
    b.quad((x0, y0, z0 - 1), (x0, y1, z0 - 1), (x1, y1, z0 - 1), (x1, y0, z0 - 1), M_WALL, d(8), d(4))
    b.quad((x0, y0, z1 + 1), (x1, y0, z1 + 1), (x1, y1, z1 + 1), (x0, y1, z1 + 1), M_WALL, d(8), d(4))
    # two arcades (z = -30, +30), two storeys
    ncol = 12
    xs = np.linspace(x0 + 12, x1 - 12, ncol)
    storeys = ((y0, 26.0), (30.0, 58.0))
    for zi, zc in enumerate((-30.0, 30.0)):
        for si, (ya, yb) in enumerate(storeys):
            for ci, cx in enumerate(xs):
                mat = M_COL if (ci + si) % 2 == 0 else M_COL2
                r = 2.6 if si == 0 else 2.0
                _cylinder(b, cx, zc, ya, yb - 7.0, r, d(28), d(22), mat)
                b.box((cx - r * 1.4, ya, zc - r * 1.4), (cx + r * 1.4, ya + 1.5, zc + r * 1.4), M_TRIM, d(3))
                b.box((cx - r * 1.3, yb - 7.0, zc - r * 1.3), (cx + r * 1.3, yb - 6.0, zc + r * 1.3), M_TRIM, d(3))
            for ci in range(ncol - 1):
                _arch(b, xs[ci], xs[ci + 1], yb - 6.0 - 0.0, zc, 5.0, 1.8, d(28), d(6), M_ARCH)
            # entablature / gallery slab above the arches, reaching the outer wall
            zo = z0 if zc < 0 else z1
            lo_z, hi_z = (zo, zc + 3.0) if zc < 0 else (zc - 3.0, zo)
            b.box((x0, yb + 2.2, lo_z), (x1, yb + 4.0, hi_z), M_SLAB, d(28), uv_scale=(10, 2))
    # cutout cards: hanging banners and foliage clusters (~8 % of triangles)
    for k in range(10):
        cx = xs[1 + k % (ncol - 2)] + 7.0
        zc = -24.0 if k % 2 == 0 else 24.0
        mat = N_SOLID + (k % N_CUT)
        top = 24.0 if k < 6 else 56.0
        b.quad((cx - 4, top - 16, zc), (cx + 4, top - 16, zc), (cx + 4, top, zc), (cx - 4, top, zc), mat, d(10), d(20), (1, 2))
        b.quad((cx + 4, top - 16, zc), (cx - 4, top - 16, zc), (cx - 4, top, zc), (cx + 4, top, zc), mat, d(10), d(20), (1, 2))
    nfol = d(3300)
    fx = rng.uniform(x0 + 8, x1 - 8 - 30.0, nfol)
    fx = np.where(fx > -15.0, fx + 30.0, fx)          # keep |x| < 15 around the default camera clear
    pos = np.stack([fx, rng.uniform(y0 + 2.3, y0 + 9.0, nfol), rng.uniform(-16.0, 16.0, nfol)], -1)
    ang = rng.uniform(0, np.pi, nfol)
    sz = rng.uniform(0.8, 2.2, nfol)
    for k in range(nfol):
        c, a, s = pos[k], ang[k], sz[k]
        dx, dz = np.cos(a) * s, np.sin(a) * s
        mat = N_SOLID + (k % N_CUT)
        b.quad((c[0] - dx, c[1] - s, c[2] - dz), (c[0] + dx, c[1] - s, c[2] + dz),
               (c[0] + dx, c[1] + s, c[2] + dz), (c[0] - dx, c[1] + s, c[2] - dz), mat)
        b.quad((c[0] + dx, c[1] - s, c[2] + dz), (c[0] - dx, c[1] - s, c[2] - dz),
               (c[0] - dx, c[1] + s, c[2] - dz), (c[0] + dx, c[1] + s, c[2] + dz), mat)
    # a few free-standing props on the floor (boxes + spheres) using the remaining solid materials
    for k in range(14):
        cx, cz = rng.uniform(x0 + 15, x1 - 15), rng.uniform(-18, 18)
        mat = 8 + (k % (N_SOLID - 8))
        if k % 2 == 0:
            s = rng.uniform(2.0, 4.5)
            b.box((cx - s, y0, cz - s), (cx + s, y0 + 2 * s, cz + s), mat, d(6), bottom=False)
        else:
            r = rng.uniform(2.0, 4.0)
            th = np.linspace(0, 2 * np.pi, d(40) + 1)[None, :]
            ph = np.linspace(0.02, np.pi - 0.02, d(20) + 1)[:, None]
            P = np.stack([cx + r * np.sin(ph) * np.cos(-th), y0 + r + r * np.cos(np.pi - ph) + 0 * th,
                          cz + r * np.sin(ph) * np.sin(-th)], -1)
            uv = np.stack([np.broadcast_to(th / (2 * np.pi) * 2, P.shape[:2]), np.broadcast_to(ph / np.pi, P.shape[:2])], -1)
            b.patch(P, uv, mat)
    v, i, m = b.build()
    textures, materials = atrium_textures(seed, tex_size, N_SOLID, N_CUT)
    return Scene(name, v, i, m, textures, materials, camera_pos=(0.0, 4.0, 0.0), yaw=-90.0, pitch=0.0,
                 meta={"config": 2, "seed": seed, "detail": detail})


# ------------------------------------------------------------------------------------------ config 4

def torus_knot_positions(nu=1024, nv=512, t=0.0, p=2, q=3, R=34.0, r=9.0):
    """WORLD positions (nv+1, nu+1, 3) of a displaced (p,q) torus-knot tube; `t` animates the
    sine displacement (BASELINE config 4: re-voxelised every frame)."""
    u = np.linspace(0, 2 * np.pi, nu + 1)[None, :]
    v = np.linspace(0, 2 * np.pi, nv + 1)[:, None]
    cu = np.stack([(R + 12.0 * np.cos(q * u)) * np.cos(p * u), 14.0 * np.sin(q * u) + 20.0,
                   (R + 12.0 * np.cos(q * u)) * np.sin(p * u)], -1)
    du = np.gradient(cu, axis=1)
    T = du / np.linalg.norm(du, axis=-1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    Nn = np.cross(T, up)
    Nn /= np.linalg.norm(Nn, axis=-1, keepdims=True)
    Bn = np.cross(T, Nn)
    rr = r * (1.0 + 0.25 * np.sin(7 * u + 3.0 * t) * np.cos(5 * v + 2.0 * t))
    return cu + rr[..., None] * (np.cos(v)[..., None] * Nn + np.sin(v)[..., None] * Bn)


def dynamic_knot(nu=1024, nv=512, t=0.0, name="knot"):
    """BASELINE config 4: nu*nv*2 triangles (1 048 576 at the defaults), single material."""
    b = _Builder()
    P = torus_knot_positions(nu, nv, t)
    uv = np.stack(np.broadcast_arrays(np.linspace(0, 16, nu + 1)[None, :], np.linspace(0, 4, nv + 1)[:, None]), -1)
    b.patch(P, uv, 0)
    v, i, m = b.build()
    rng = np.random.default_rng(7)
    alb = _u8(0.25 + 0.6 * np.stack([value_noise(256, 6, rng)] * 3, -1) * np.array([0.9, 0.6, 0.4]))
    textures = [alb, solid_texture((96,), 1), _u8(value_noise(128, 8, rng))[..., None]]
    materials = [(0, 1, 2, 20.0)]
    return Scene(name, v, i, m, textures, materials, camera_pos=(0.0, 24.0, 110.0), yaw=-90.0, pitch=-3.0,
                 meta={"config": 4, "nu": nu, "nv": nv})


def probe_cameras(n=64, seed=5):
    """BASELINE config 5: n cameras on a 4x4x4 lattice inside the atrium with seeded jitter."""
    rng = np.random.default_rng(seed)
    k = int(round(n ** (1 / 3)))
    xs, ys, zs = np.linspace(-60, 60, k), np.linspace(2, 50, k), np.linspace(-18, 18, k)
    cams = []
    for x in xs:
        for y in ys:
            for z in zs:
                j = rng.uniform(-1.5, 1.5, 3)
                cams.append(((x + j[0], y + j[1], z + j[2]), float(rng.uniform(0, 360)), float(rng.uniform(-20, 20))))
    return cams[:n]
