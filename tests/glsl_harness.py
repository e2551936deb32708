"""Fixed-function glue around tests/glsl_run.py: runs the reference's shader FILES over a scene (test infrastructure).

What comes from the reference: every line of Shader/Shadow.vs, Voxelization.{vs,gs,fs} and VoxelConeTracing.{vs,fs},
executed from the files by the interpreter.  What comes from the GL 4.3 specification, restated here in float64: the
viewport transform (13.6.1), point sampling at pixel centres (14.6.1), barycentric / perspective-correct attribute
interpolation (14.6.1 eq. 14.9-14.10), texture filtering and level-of-detail selection (8.14), unorm conversion (2.3.5).
The GL state those rules depend on is the reference's: glViewport(0,0,V,V) for the voxel pass (Voxel_Cone_Tracing.h:218),
depth texture GL_LINEAR + CLAMP_TO_EDGE, material textures GL_REPEAT + LINEAR_MIPMAP_LINEAR (Model.h:171-176), voxel
texture GL_REPEAT x3 + LINEAR_MIPMAP_LINEAR (Voxel_Cone_Tracing.h:105-113).

Used by tests/golden/make_reference_shader_vectors.py (writes the fixture) and tests/test_reference_glsl.py.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

import glsl_run
from vct_b200 import scenes, uniforms

SHADER_DIR = "/root/reference/Voxel_Cone_Tracing_Final/Shader"
SHADERS = ["Shadow.vs", "Voxelization.vs", "Voxelization.gs", "Voxelization.fs", "VoxelConeTracing.vs", "VoxelConeTracing.fs"]
FRAME = dict(V=32, width=48, height=40, shadow_map_size=256)
VOXEL = dict(V=16, width=16, height=16, shadow_map_size=128)
JITTER = [(0.0, 0.0), (1 / 256, 1 / 256), (-1 / 256, 1 / 256), (1 / 256, -1 / 256), (-1 / 256, -1 / 256)]
EDGE_PX = 0.01          # pixel centres closer than this to a triangle edge are left to the fill rule (sub-pixel snap: 1/256 px)


def reference_available():
    return all(os.path.exists(os.path.join(SHADER_DIR, s)) for s in SHADERS)


def shader_hashes():
    return {s: hashlib.sha256(open(os.path.join(SHADER_DIR, s), "rb").read()).hexdigest() for s in SHADERS}


def load(name, dtype):
    return glsl_run.Program(open(os.path.join(SHADER_DIR, name)).read(), dtype)


def fixture_scene():
    """Cornell box with multi-texel albedo / specular / height maps on every surface (seeded): exercises bilinear REPEAT
    fetches, CalcBumpNormal, RGB specular maps, magnified and minified texture LODs."""
    sc = scenes.cornell()
    rng = np.random.default_rng(8)
    tex = [rng.integers(40, 256, (8, 8, 3), dtype=np.uint8) for _ in range(3)]
    spec = rng.integers(0, 256, (4, 8, 3), dtype=np.uint8)
    height = rng.integers(0, 256, (8, 4, 1), dtype=np.uint8)
    sc.textures = tex + [spec, height]
    sc.materials = [(0, 3, 4, 20.0), (1, 3, 4, 20.0), (2, 3, 4, 20.0)]
    return sc


def atrium_scene():
    """6126 triangles, 22 materials: 32^2 RGB albedo maps (minified at this frame size), single-channel specular and
    height maps, alpha cut-out cards -- the scene of tests/golden/atrium_v32_conservative.npz"""
    return scenes.atrium(detail=0.1, tex_size=32)


def config4_scene(step=3):
    """The 1 048 576-triangle knot of BASELINE config 4 at one time step of bench.py's animation (positions displaced
    along the base normals by 12 sin(phase + 0.21 step), float32; normals, tangents and uv stay those of the base mesh)."""
    sc = scenes.dynamic_knot()
    base, nrm = sc.verts[:, :3].astype(np.float32), sc.verts[:, 3:6].astype(np.float32)
    phase = (base[:, 0] * np.float32(0.004) + base[:, 2] * np.float32(0.003)).astype(np.float32)
    amp = (np.float32(12.0) * np.sin(phase + np.float32(0.21 * step)).astype(np.float32)).astype(np.float32)
    sc.verts[:, :3] = (base + nrm * amp[:, None]).astype(np.float32)
    return sc


def scene_uniforms(sc, kind):
    if kind.endswith("_msaa4"):
        return uniforms.scene_uniforms(sc, coverage="msaa4", **{"voxel_msaa4": VOXEL, "shards_msaa4": SHARDS}[kind])
    kw = dict({"frame": FRAME, "voxel": VOXEL, "card": CARD, "shards": SHARDS, "config1": CONFIG1, "atrium": ATRIUM, "config2": CONFIG2, "config4": CONFIG4}[kind])
    kw.setdefault("coverage", "center")
    return uniforms.scene_uniforms(sc, **kw)


# ------------------------------------------------------------------------------- fixed function: textures
def _expand(c):
    """channel rules of the reference's uploads (Model.h:159-169): RED -> (r, 0, 0, 1), RGB -> alpha 1"""
    c = np.atleast_1d(c)
    if len(c) == 1:
        return np.array([c[0], 0.0, 0.0, 1.0])
    return np.append(c[:3], 1.0) if len(c) == 3 else c


def _bilinear(level, s, t, wrap):
    h, w = level.shape[:2]
    x, y = s * w - 0.5, t * h - 0.5
    i, j = int(np.floor(x)), int(np.floor(y))
    a, b = x - i, y - j
    if wrap == "repeat":
        at = lambda ii, jj: level[jj % h, ii % w]
    else:
        at = lambda ii, jj: level[min(max(jj, 0), h - 1), min(max(ii, 0), w - 1)]
    return (1 - b) * ((1 - a) * at(i, j) + a * at(i + 1, j)) + b * ((1 - a) * at(i, j + 1) + a * at(i + 1, j + 1))


def box_mips_2d(tex):
    """glGenerateMipmap on an RGBA8 / RGB8 / R8 texture, as the oracle defines it: 2x2 box, (sum + 2) >> 2"""
    levels = [tex.astype(np.int64)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        p = levels[-1]
        h, w = max(p.shape[0] // 2, 1), max(p.shape[1] // 2, 1)
        if p.shape[0] == 1:
            p = np.concatenate([p, p], 0)
        if p.shape[1] == 1:
            p = np.concatenate([p, p], 1)
        levels.append((p.reshape(h, 2, w, 2, -1).sum((1, 3)) + 2) >> 2)
    return [l.astype(np.float64) / 255.0 for l in levels]


def box_mips_3d(g0):
    levels = [g0.astype(np.int64)]
    while levels[-1].shape[0] > 1:
        p = levels[-1]
        n = p.shape[0] // 2
        levels.append((p.reshape(n, 2, n, 2, n, 2, 4).sum((1, 3, 5)) + 4) >> 3)
    return [l.astype(np.float64) / 255.0 for l in levels]


class Sampler2D:
    """A bound 2D texture.  `lod` is the fragment's level of detail for THIS texture, set by the rasteriser glue from
    the screen-space derivatives of the interpolated coordinate (an offset added in the shader does not change them)."""

    def __init__(self, levels, wrap):
        self.levels, self.wrap, self.lod = levels, wrap, 0.0

    def size(self):
        return np.array([self.levels[0].shape[1], self.levels[0].shape[0]], dtype=np.float64)

    def set_lod_from(self, duv_dx, duv_dy):
        rho = max(np.linalg.norm(duv_dx * self.size()), np.linalg.norm(duv_dy * self.size()))     # GL 4.3 eq. 8.7-8.8
        self.lod = float(np.log2(max(rho, 1e-30)))

    def sample(self, s, t):
        lod = min(max(self.lod, 0.0), len(self.levels) - 1.0)
        l0 = int(np.floor(lod)); f = lod - l0
        c = _bilinear(self.levels[l0], s, t, self.wrap)
        if f > 0 and l0 + 1 < len(self.levels):
            c = (1 - f) * c + f * _bilinear(self.levels[l0 + 1], s, t, self.wrap)
        return _expand(c)


class Sampler3D:
    """`gain` scales every fetch: the generator reruns the fragment stage with gain 1 +- 4e-6 (a few float32 ulps of
    filter arithmetic, which GL leaves open) and drops pixels whose colour depends on it -- in practice those whose
    cone loop ends with alpha within ~1e-6 of MAX_ALPHA."""

    def __init__(self, levels, gain=1.0):
        self.levels, self.gain = levels, gain

    def _level(self, level, uvw):
        n = level.shape[0]
        t = np.asarray(uvw, dtype=np.float64) * n - 0.5
        i0 = np.floor(t).astype(int)
        f = t - i0
        out = np.zeros(4)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    w = (f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])
                    out += w * level[(i0[2] + dz) % n, (i0[1] + dy) % n, (i0[0] + dx) % n]
        return out

    def sample(self, uvw, lod):
        lod = min(max(float(lod), 0.0), len(self.levels) - 1.0)
        l0 = int(np.floor(lod)); f = lod - l0
        a = self._level(self.levels[l0], uvw)
        if f > 0 and l0 + 1 < len(self.levels):
            a = (1 - f) * a + f * self._level(self.levels[l0 + 1], uvw)
        return a * self.gain


def install_texture_hooks(prog):
    F = prog.F
    prog.hooks["texture"] = lambda smp, uv: smp.sample(float(uv[0]), float(uv[1])).astype(F)
    prog.hooks["textureLod"] = lambda smp, uvw, lod: smp.sample(uvw.astype(np.float64), lod).astype(F)


def scene_samplers(sc):
    return [Sampler2D(box_mips_2d(t if t.ndim == 3 else t[..., None]), "repeat") for t in sc.textures]


def shadow_sampler(d24):
    return Sampler2D([(d24.astype(np.float64) / 16777215.0)[..., None]], "clamp")


def col_major(u, name):
    return np.asarray(u[name], dtype=np.float64).reshape(4, 4).T


# ------------------------------------------------------------------------------------------ Shadow.vs
def shadow_reference_depths(sc, u, dtype=np.float32):
    """Runs Shadow.vs on every vertex, then (fixed function, float64) finds for every shadow-map texel whose centre is
    well inside exactly one nearest front-facing triangle the window-space depth.  Returns (ij [n,2], z_window [n],
    slack [n]) -- slack is the depth change a 1/256 px vertex snap can cause on that triangle."""
    S = int(u["ShadowMapSize"])
    vs = load("Shadow.vs", dtype)
    vs.set_uniform("DepthModelViewProjectionMatrix", u["DepthModelViewProjectionMatrix"])
    vs.globals["gl_Position"] = np.zeros(4, dtype=dtype)
    clip = np.zeros((len(sc.verts), 4))
    for k, v in enumerate(sc.verts):
        vs.globals["position"] = v[:3].astype(dtype)
        vs.run()
        clip[k] = vs.globals["gl_Position"]
    win = np.empty((len(clip), 3))
    win[:, 0] = (clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * S
    win[:, 1] = (clip[:, 1] / clip[:, 3] * 0.5 + 0.5) * S
    win[:, 2] = clip[:, 2] / clip[:, 3] * 0.5 + 0.5
    tri = win[sc.idx.astype(np.int64)]                         # [nt][3][xyz]
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    area = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    ok = area > 1e-9                                           # back faces are culled (Voxel_Cone_Tracing.h:194)
    safe = np.where(ok, area, 1.0)
    # window-space depth gradient of each triangle: GL snaps vertices to a sub-pixel grid (>= 4 bits, 8 on the oracle
    # and on NVIDIA hardware), which moves the interpolated depth by up to |grad z| / 256 px
    dzdx = ((b[:, 2] - a[:, 2]) * (c[:, 1] - a[:, 1]) - (c[:, 2] - a[:, 2]) * (b[:, 1] - a[:, 1])) / safe
    dzdy = ((c[:, 2] - a[:, 2]) * (b[:, 0] - a[:, 0]) - (b[:, 2] - a[:, 2]) * (c[:, 0] - a[:, 0])) / safe
    slack = (np.abs(dzdx) + np.abs(dzdy)) / 256.0
    out_ij, out_z, out_tol = [], [], []
    for j in range(S):
        for i in range(S):
            px, py = i + 0.5, j + 0.5
            w0 = ((b[:, 0] - px) * (c[:, 1] - py) - (b[:, 1] - py) * (c[:, 0] - px)) / safe
            w1 = ((c[:, 0] - px) * (a[:, 1] - py) - (c[:, 1] - py) * (a[:, 0] - px)) / safe
            w2 = 1.0 - w0 - w1
            m = np.minimum(w0, np.minimum(w1, w2))
            inside = ok & (m > 0.01)
            if (ok & (m > -0.01) & ~inside).any() or not inside.any():
                continue                                        # near an edge of some triangle, or background
            z = np.where(inside, w0 * a[:, 2] + w1 * b[:, 2] + w2 * c[:, 2], np.inf)
            k = int(np.argmin(z))
            zs = np.sort(z)
            if len(zs) > 1 and zs[1] - zs[0] < 1e-4:
                continue
            if not 0.0 <= z[k] <= 1.0:
                continue
            out_ij.append((i, j)); out_z.append(z[k]); out_tol.append(slack[k])
    return np.array(out_ij, dtype=np.int32), np.array(out_z), np.array(out_tol)


# ------------------------------------------------------------------------ Voxelization.vs / .gs / .fs
MSAA4 = [(0.375, 0.125), (0.875, 0.375), (0.125, 0.625), (0.625, 0.875)]      # the standard 4x pattern (GL 4.3 14.3.1 leaves it
                                                                              # to the implementation; D3D / NVIDIA / the oracle use this)


def voxel_reference_fragments(sc, u, shadow_d24, dtype, tri_ids=None, jitter=False, coverage="center"):
    """Runs the voxelisation program over triangles of the scene.  Returns a list of fragments
    ([(voxel index or -1 when the store is out of bounds, rgba8 bytes stored)], certain, triangle) -- one result for the
    pixel centre and, with `jitter`, four more with the attributes evaluated 1/256 px away (the freedom GL's sub-pixel
    vertex snap leaves); `certain` is False when the pixel centre lies within EDGE_PX of a triangle edge, where the
    fill rule, not the shaders, decides.  coverage = "msaa4": the reference's default framebuffer has four samples
    (main.cpp:30), a fragment is generated when any of them is inside and -- no `centroid` qualifier in the shaders --
    its inputs and gl_FragCoord are evaluated at the pixel centre, inside the triangle or not."""
    V = int(u["VoxelDimensions"])
    vs, gs, fs = load("Voxelization.vs", dtype), load("Voxelization.gs", dtype), load("Voxelization.fs", dtype)
    for p in (vs, gs, fs):
        install_texture_hooks(p)
    vs.set_uniform("DepthModelViewProjectionMatrix", u["DepthModelViewProjectionMatrix"])
    vs.set_uniform("ModelMatrix", u["ModelMatrix"])
    for n in ("ProjX", "ProjY", "ProjZ"):
        gs.set_uniform(n, u[n])
    fs.set_uniform("ShadowMapSize", int(u["ShadowMapSize"]))
    fs.set_uniform("VoxelDimensions", V)
    fs.globals["ShadowMap"] = shadow_sampler(shadow_d24)
    fs.globals["VoxelTexture"] = "image"
    samplers = scene_samplers(sc)
    white = Sampler2D([np.ones((1, 1, 4))], "repeat")
    stores = []
    fs.hooks["imageStore"] = lambda img, p, v: stores.append((p.copy(), v.copy()))
    emitted = []
    gs.hooks["EmitVertex"] = lambda: emitted.append(dict(pos=gs.globals["gl_Position"].astype(np.float64),
                                                         uv=gs.globals["TexCoord"].astype(np.float64),
                                                         dc=gs.globals["DepthCoord"].astype(np.float64),
                                                         axis=int(gs.globals["axis"])))
    gs.hooks["EndPrimitive"] = lambda: None
    vs.globals["gl_Position"] = np.zeros(4, dtype=dtype)
    gs.globals["gl_Position"] = np.zeros(4, dtype=dtype)
    frags = []
    for ti in (range(len(sc.idx)) if tri_ids is None else tri_ids):
        gl_in, vertices = [], []
        for vi in sc.idx[ti]:
            v = sc.verts[int(vi)]
            vs.globals["position"] = v[:3].astype(dtype)
            vs.globals["tex"] = v[6:8].astype(dtype)
            vs.run()
            gl_in.append(glsl_run.Block(gl_Position=vs.globals["gl_Position"].copy()))
            vertices.append(glsl_run.Block(TexCoord=vs.globals["TexCoord"].copy(), DepthCoord=vs.globals["DepthCoord"].copy()))
        gs.globals["gl_in"], gs.globals["vertices"] = gl_in, vertices
        del emitted[:]
        gs.run()
        assert len(emitted) == 3
        # viewport transform, glViewport(0, 0, V, V), depth range [0, 1]
        win = np.array([[(e["pos"][0] / e["pos"][3] * 0.5 + 0.5) * V, (e["pos"][1] / e["pos"][3] * 0.5 + 0.5) * V,
                         e["pos"][2] / e["pos"][3] * 0.5 + 0.5] for e in emitted])
        a, b, c = win
        area = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
        if abs(area) < 1e-12:
            continue
        mat = sc.materials[int(sc.tri_material[ti])] if getattr(sc, "tri_material", None) is not None else sc.materials[0]
        diffuse = samplers[mat[0]] if mat[0] is not None and mat[0] >= 0 else white
        uv = np.array([e["uv"] for e in emitted]); dc = np.array([e["dc"] for e in emitted])

        def bary(px, py):
            w0 = ((b[0] - px) * (c[1] - py) - (b[1] - py) * (c[0] - px)) / area
            w1 = ((c[0] - px) * (a[1] - py) - (c[1] - py) * (a[0] - px)) / area
            return np.array([w0, w1, 1.0 - w0 - w1])
        # smallest altitude of the triangle in pixels: barycentric w = distance to the opposite edge / altitude
        edges = [np.linalg.norm((c - b)[:2]), np.linalg.norm((a - c)[:2]), np.linalg.norm((b - a)[:2])]
        px_per_bary = abs(area) / max(edges)
        z_slack = (abs(bary(1.5, 0.5) @ win[:, 2] - bary(0.5, 0.5) @ win[:, 2]) +
                   abs(bary(0.5, 1.5) @ win[:, 2] - bary(0.5, 0.5) @ win[:, 2])) / 256.0
        # the projections are orthographic (w = 1): attributes are affine in window space, so are their derivatives
        diffuse.set_lod_from(bary(1.5, 0.5) @ uv - bary(0.5, 0.5) @ uv, bary(0.5, 1.5) @ uv - bary(0.5, 0.5) @ uv)
        fs.globals["DiffuseTexture"] = diffuse
        i0, i1 = int(np.floor(win[:, 0].min())), int(np.ceil(win[:, 0].max()))
        j0, j1 = int(np.floor(win[:, 1].min())), int(np.ceil(win[:, 1].max()))
        for j in range(max(j0, 0), min(j1, V - 1) + 1):
            for i in range(max(i0, 0), min(i1, V - 1) + 1):
                w = bary(i + 0.5, j + 0.5)
                if coverage == "msaa4":
                    d = [bary(i + sx, j + sy).min() * px_per_bary for sx, sy in MSAA4]
                    inside = any(x > EDGE_PX for x in d)
                    if not inside and not any(abs(x) <= EDGE_PX for x in d):
                        continue
                else:
                    if w.min() * px_per_bary < -EDGE_PX:
                        continue
                    inside = bool(w.min() * px_per_bary > EDGE_PX)
                z = float(w @ win[:, 2])
                # the fill rule decides within EDGE_PX of an edge; a sub-pixel vertex snap can move z across a slice
                certain = inside and abs(z * V - np.rint(z * V)) > z_slack * V + 1e-5
                if not 0.0 <= z <= 1.0:
                    continue                                    # depth clipping (13.5), glDepthRange default
                results = []
                for ox, oy in (JITTER if jitter else JITTER[:1]):
                    wj = bary(i + 0.5 + ox, j + 0.5 + oy)
                    fs.globals["gs"] = glsl_run.Block(TexCoord=(wj @ uv).astype(dtype), axis=emitted[0]["axis"],
                                                      DepthCoord=(wj @ dc).astype(dtype))
                    fs.globals["gl_FragCoord"] = np.array([i + 0.5, j + 0.5, float(wj @ win[:, 2]), 1.0], dtype=dtype)
                    del stores[:]
                    fs.run()                                    # Voxelization.fs has no discard
                    assert len(stores) == 1
                    p, val = stores[0]
                    inb = bool((p >= 0).all() and (p < V).all())  # out-of-bounds image stores are dropped (8.26)
                    rgba = tuple(int(x) for x in np.rint(np.clip(val.astype(np.float64), 0, 1) * 255.0))
                    results.append((int((p[2] * V + p[1]) * V + p[0]) if inb else -1, rgba))
                frags.append((results, certain, ti))
    return frags


def voxel_reference_accumulator(sc, u, shadow_d24, tri_ids=None, coverage="center"):
    """Per-voxel fragment counts and byte sums the reference's voxelisation program produces.  A voxel is `certain`
    when every fragment landing in it (a) is interior to its triangle, (b) targets the same voxel in float32 and float64
    execution and under the sub-pixel jitter, and (c) stores bytes that move by at most 1 under either.  Returns (voxel
    index, count, rgb sums, uncertain voxel indices, fragments run)."""
    f32 = voxel_reference_fragments(sc, u, shadow_d24, np.float32, tri_ids, jitter=True, coverage=coverage)
    f64 = voxel_reference_fragments(sc, u, shadow_d24, np.float64, tri_ids, coverage=coverage)
    assert len(f32) == len(f64)
    acc, bad = {}, set()
    for (r32, certain, _), (r64, _, _) in zip(f32, f64):
        v0, c0 = r32[0]
        others = r32[1:] + r64
        stable = certain and all(v == v0 and max(abs(x - y) for x, y in zip(c, c0)) <= 1 for v, c in others)
        if not stable:
            bad.update(v for v, _ in r32 + r64 if v >= 0)
            continue
        if v0 < 0:
            continue
        e = acc.setdefault(v0, [0, 0, 0, 0])
        e[0] += 1; e[1] += c0[0]; e[2] += c0[1]; e[3] += c0[2]
    keys = sorted(k for k in acc if k not in bad)
    idx = np.array(keys, dtype=np.int64)
    cnt = np.array([acc[k][0] for k in keys], dtype=np.uint32)
    sums = np.array([acc[k][1:] for k in keys], dtype=np.uint32).reshape(-1, 3)
    return idx, cnt, sums, np.array(sorted(bad), dtype=np.int64), len(f32)


# ----------------------------------------------------------------------- VoxelConeTracing.vs / .fs
class FrameStage:
    """VoxelConeTracing.vs + .fs bound to a scene, its uniforms and the fixed-function inputs of the fragment stage."""

    NAMES = ["tex", "Position_world", "Normal_world", "Tangent_world", "BiTangent_world", "CameraDirection_world", "Position_depth"]

    def __init__(self, sc, u, shadow_d24, grid0, dtype, voxel_gain=1.0):
        self.sc, self.dtype = sc, dtype
        self.W, self.H = int(u["screen_width"]), int(u["screen_height"])
        self.vs, self.fs = load("VoxelConeTracing.vs", dtype), load("VoxelConeTracing.fs", dtype)
        vs, fs = self.vs, self.fs
        install_texture_hooks(fs)
        for n in ("ModelMatrix", "ModelViewMatrix", "ProjectionMatrix", "DepthModelViewProjectionMatrix"):
            vs.set_uniform(n, u[n])
        vs.set_uniform("CameraPosition", u["CameraPosition"])
        vs.globals["gl_Position"] = np.zeros(4, dtype=dtype)
        fs.set_uniform("LightDirection", u["LightDirection"])
        fs.set_uniform("ambientFactor", u.get("ambientFactor", 0.1))
        fs.set_uniform("ShadowMapSize", int(u["ShadowMapSize"]))
        fs.set_uniform("VoxelGridWorldSize", u["VoxelGridWorldSize"])
        fs.set_uniform("VoxelDimensions", int(u["VoxelDimensions"]))
        fs.globals["ShadowMap"] = shadow_sampler(shadow_d24)
        fs.globals["VoxelTexture"] = Sampler3D(box_mips_3d(grid0), voxel_gain)
        self.samplers = scene_samplers(sc)
        self.cache = {}

    def triangle(self, ti):
        """vertex stage for the three vertices: clip positions, window xy + depth, varyings"""
        if ti not in self.cache:
            vs, rows = self.vs, []
            for vi in self.sc.idx[ti]:
                v = self.sc.verts[int(vi)]
                for name, lo, hi in (("Position", 0, 3), ("Normal", 3, 6), ("TexCoord", 6, 8), ("Tangent", 8, 11), ("BiTangent", 11, 14)):
                    vs.globals[name] = v[lo:hi].astype(self.dtype)
                vs.run()
                rows.append((vs.globals["gl_Position"].astype(np.float64), [vs.globals[k].astype(np.float64) for k in self.NAMES]))
            clip = np.array([r[0] for r in rows])
            win = np.stack([(clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * self.W, (clip[:, 1] / clip[:, 3] * 0.5 + 0.5) * self.H,
                            clip[:, 2] / clip[:, 3] * 0.5 + 0.5], 1)
            a, b, c = win
            area = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
            self.cache[ti] = (rows, clip, win, area)
        return self.cache[ti]

    def linear_bary(self, ti, px, py):
        _, _, (a, b, c), area = self.triangle(ti)
        w0 = ((b[0] - px) * (c[1] - py) - (b[1] - py) * (c[0] - px)) / area
        w1 = ((c[0] - px) * (a[1] - py) - (c[1] - py) * (a[0] - px)) / area
        return np.array([w0, w1, 1.0 - w0 - w1])

    def persp(self, ti, px, py):
        q = self.linear_bary(ti, px, py) / self.triangle(ti)[1][:, 3]        # GL 4.3 eq. 14.9
        return q / q.sum()

    def shade(self, ti, i, j, ox=0.0, oy=0.0):
        """fragment stage at the centre of pixel (i, j) of triangle ti (moved by (ox, oy) px: what a sub-pixel vertex snap
        does to the interpolated inputs); None when the shader discards"""
        rows = self.triangle(ti)[0]
        fs = self.fs
        w = self.persp(ti, i + 0.5 + ox, j + 0.5 + oy)
        for k, name in enumerate(self.NAMES):
            fs.globals[name] = sum(w[m] * rows[m][1][k] for m in range(3)).astype(self.dtype)
        uv = fs.globals["tex"].astype(np.float64)
        uvx = sum(self.persp(ti, i + 1.5, j + 0.5)[m] * rows[m][1][0] for m in range(3))
        uvy = sum(self.persp(ti, i + 0.5, j + 1.5)[m] * rows[m][1][0] for m in range(3))
        mat = self.sc.materials[int(self.sc.tri_material[ti])]
        d, s, h = self.samplers[mat[0]], self.samplers[mat[1]], self.samplers[mat[2]]
        for smp in (d, s, h):
            smp.set_lod_from(uvx - uv, uvy - uv)
        fs.globals["DiffuseTexture"], fs.globals["SpecularTexture"], fs.globals["HeightTexture"] = d, s, h
        fs.set_uniform("HeightTextureSize", h.size())
        fs.set_uniform("Shininess", mat[3])
        fs.globals["color"] = np.zeros(4, dtype=self.dtype)
        try:
            fs.run()
        except glsl_run.Discard:
            return None
        return fs.globals["color"].astype(np.float64)

    def candidates(self, i, j):
        """Front-facing triangles (back faces are culled, main.cpp:57-58) whose interior holds the centre of pixel
        (i, j), nearest first; None when the centre is within EDGE_PX of an edge of any of them (fill rule territory)
        or two of them are closer than 1e-6 in window depth."""
        found = []
        for ti in range(len(self.sc.idx)):
            _, clip, win, area = self.triangle(ti)
            if area <= 1e-12 or (clip[:, 3] <= 0).any():
                continue
            lin = self.linear_bary(ti, i + 0.5, j + 0.5)
            edges = [np.linalg.norm(win[2, :2] - win[1, :2]), np.linalg.norm(win[0, :2] - win[2, :2]), np.linalg.norm(win[1, :2] - win[0, :2])]
            dist = lin.min() * area / max(edges)
            if dist < -EDGE_PX:
                continue
            if dist < EDGE_PX:
                return None
            z = float(lin @ win[:, 2])
            if 0.0 <= z <= 1.0:
                found.append((z, ti))
        found.sort()
        if any(b[0] - a[0] < 1e-6 for a, b in zip(found, found[1:])):
            return None
        return [ti for _, ti in found]


def frame_reference_pixels(sc, u, shadow_d24, grid0, visibility, pixels, dtype, voxel_gain=1.0, offset=(0.0, 0.0)):
    """Runs VoxelConeTracing.vs on the three vertices of the triangle the visibility pass found at each pixel,
    interpolates the seven varyings perspective-correctly at the pixel centre (fixed function, float64) and runs
    VoxelConeTracing.fs.  Returns float colours [n, 4] (NaN rows for background pixels and discarded fragments)."""
    st = FrameStage(sc, u, shadow_d24, grid0, dtype, voxel_gain)
    out = np.full((len(pixels), 4), np.nan)
    for n, (i, j) in enumerate(pixels):
        ti = int(visibility[j, i])
        if ti != 0xFFFFFFFF:
            c = st.shade(ti, i, j, *offset)
            if c is not None:
                out[n] = c
    return out


def frame_reference_depth_ordered(sc, u, shadow_d24, grid0, pixels, dtype, voxel_gain=1.0):
    """The same without a visibility map: per pixel the fragment shader runs on the covering front-facing triangles from
    the nearest on, and the first fragment that is not discarded is what depth test LESS leaves in the framebuffer
    (a discarded fragment writes neither colour nor depth).  Returns (triangle id [n] -- 0xFFFFFFFF background, -1 pixel
    left to the fill rule --, colours [n, 4], fragments discarded)."""
    st = FrameStage(sc, u, shadow_d24, grid0, dtype, voxel_gain)
    tri = np.full(len(pixels), -1, dtype=np.int64)
    out = np.full((len(pixels), 4), np.nan)
    discarded = 0
    for n, (i, j) in enumerate(pixels):
        cands = st.candidates(i, j)
        if cands is None:
            continue
        tri[n] = 0xFFFFFFFF
        for ti in cands:
            c = st.shade(ti, i, j)
            if c is None:
                discarded += 1
                continue
            tri[n], out[n] = ti, c
            break
    return tri, out, discarded


def _quad(p00, p10, p11, p01, scale=20.0):
    """two triangles, model units = world * 20 (ModelMatrix = scale(0.05)), uv (0,0)-(1,1)"""
    P = np.array([p00, p10, p11, p01], dtype=np.float64) * scale
    v = np.zeros((4, 14), dtype=np.float32)
    v[:, :3] = P
    n = np.cross(P[1] - P[0], P[3] - P[0]); n /= np.linalg.norm(n)
    t = (P[1] - P[0]) / np.linalg.norm(P[1] - P[0]); b = (P[3] - P[0]) / np.linalg.norm(P[3] - P[0])
    v[:, 3:6], v[:, 8:11], v[:, 11:14] = n, t, b
    v[:, 6:8] = [[0, 0], [1, 0], [1, 1], [0, 1]]
    return v, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)


CARD = dict(V=32, width=40, height=30, shadow_map_size=256)
SHARDS = dict(V=32, width=16, height=16, shadow_map_size=128)
CONFIG1 = dict(V=64, width=256, height=256, shadow_map_size=1024, coverage="msaa4")     # BASELINE.json configs[0]
ATRIUM = dict(V=32, width=96, height=54, shadow_map_size=512, coverage="conservative")  # the golden atrium case
ATRIUM_STRIDE = 3
CONFIG2 = dict(V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")   # BASELINE.json configs[1], the headline
CONFIG4 = dict(V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")   # BASELINE.json configs[3]
CONFIG4_STRIDE = 307
CONFIG4_STEP = 3                 # = the default of config4_scene()
CONFIG2_STRIDE = 691                                                                           # ~3000 of the 2 M pixels
CONFIG1_STRIDE = 13                                                                   # every 13th covered pixel


def shards_scene():
    """Forty randomly oriented triangles in general position (seeded), half of them textured: every dominant axis and
    both windings of Voxelization.gs:25-41, every voxelPos permutation of Voxelization.fs:70-86, on tilted geometry."""
    rng = np.random.default_rng(77)
    n = 40
    centre = rng.uniform(-45, 45, (n, 1, 3))
    tri = centre + rng.normal(0, 20, (n, 3, 3))                      # world units; the grid spans +-75
    v = np.zeros((n * 3, 14), dtype=np.float32)
    v[:, :3] = (tri * 20.0).reshape(-1, 3)                           # model units (ModelMatrix = scale(0.05))
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    v[:, 3:6] = np.repeat(nrm, 3, axis=0)
    v[:, 6:8] = rng.uniform(0, 2, (n * 3, 2))
    v[:, 8:11], v[:, 11:14] = (1, 0, 0), (0, 1, 0)
    tex = rng.integers(30, 256, (8, 8, 3), dtype=np.uint8)
    flat = np.array([[[200, 160, 90]]], dtype=np.uint8)
    grey = np.array([[[128]]], dtype=np.uint8)
    return scenes.Scene("shards", v, np.arange(n * 3, dtype=np.uint32).reshape(n, 3), (np.arange(n) % 2).astype(np.uint16),
                        [tex, flat, grey], [(0, 2, 2, 20.0), (1, 2, 2, 20.0)])


def card_scene():
    """A card with a 4x4 alpha pattern (cut-out) in front of a wall, seen slightly from the side: exercises
    `if (alpha < 0.5f) discard;` (VoxelConeTracing.fs:169-170) and the single-channel specular map (`specColor.rrra`)."""
    rng = np.random.default_rng(31)
    card_tex = np.zeros((4, 4, 4), dtype=np.uint8)
    card_tex[..., :3] = rng.integers(60, 256, (4, 4, 3))
    card_tex[..., 3] = np.where(rng.random((4, 4)) < 0.5, 0, 255)
    wall_tex = np.array([[[90, 120, 200]]], dtype=np.uint8)
    grey = rng.integers(0, 256, (2, 2, 1), dtype=np.uint8)
    cv, ci = _quad((-30, -25, 20), (30, -25, 20), (30, 25, 20), (-30, 25, 20))
    wv, wi = _quad((-70, -60, -40), (70, -60, -40), (70, 60, -40), (-70, 60, -40))
    return scenes.Scene("card", np.concatenate([cv, wv]), np.concatenate([ci, wi + 4]).astype(np.uint32),
                        np.array([0, 0, 1, 1], dtype=np.uint16), [card_tex, wall_tex, grey],
                        [(0, 2, 2, 20.0), (1, 2, 2, 8.0)], camera_pos=(6.0, 3.0, 140.0), yaw=-92.0, pitch=-1.0)


GAINS = (1.0 + 4e-6, 1.0 - 4e-6)


def colours_agree(a, b):
    """rows of two float colour arrays equal to a quarter of an 8-bit step (NaN = no fragment, must match too)"""
    na, nb = np.isnan(a[:, 0]), np.isnan(b[:, 0])
    close = np.abs(np.clip(np.nan_to_num(a), 0, 1) - np.clip(np.nan_to_num(b), 0, 1)).max(1) * 255.0 < 0.25
    return (na == nb) & (na | close)


def to_unorm8(c):
    return np.rint(np.clip(c, 0.0, 1.0) * 255.0).astype(np.int32)
