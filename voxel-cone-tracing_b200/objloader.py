"""Wavefront OBJ / MTL -> scenes.Scene: the real-asset path of the reference.

The reference loads `sponza.obj` through assimp with `aiProcess_Triangulate | aiProcess_GenSmoothNormals |
aiProcess_FlipUVs | aiProcess_CalcTangentSpace` (/root/reference/Voxel_Cone_Tracing_Final/Model.h:43) and flattens it
into `struct Vertex` (Mesh.h:12-19) + indices + per-mesh textures (Model.h:75-139).  assimp is a binary-only Windows
dependency of the reference (Lib/assimp-vc140-mt.dll), so this module restates those four post-processing steps:

  * Triangulate: polygons are fanned around their first vertex.
  * GenSmoothNormals: only when the file has no `vn`; area-weighted face normals accumulated per position.
  * FlipUVs: v -> 1 - v.
  * CalcTangentSpace: per-triangle tangent / bitangent from the uv deltas, accumulated per vertex, normalised.

Texture slots follow the reference's (quirky) mapping, Model.h:126-136 and Mesh.h:95-109: map_Kd -> DiffuseTexture,
map_Ks -> SpecularTexture, map_Ka (aiTextureType_AMBIENT) -> HeightTexture; bump maps (aiTextureType_HEIGHT) are loaded
by the reference but never bound, so they are ignored here.  A material without one of the three gets a 1x1 texture
(Kd / Ks colour, flat height): the reference would otherwise inherit the previous mesh's binding (undefined, A.6 #9).
Images are decoded by images.py (PNG / JPEG / TGA / BMP / PNM on numpy + zlib, byte-identical to the stb_image the
reference calls at Model.h:152 -- tests/test_images_vs_stb.py), with Pillow as the fallback for the rest (GIF, PSD, HDR)."""
from __future__ import annotations

import os

import numpy as np

from .scenes import Scene


def _load_image(path):
    from . import images
    try:                                  # built-in decoders (PNG / JPEG / TGA / BMP / PNM, stb_image's results)
        return images.load_image(path)
    except images.UnsupportedImage:
        pass
    from PIL import Image                 # anything else (GIF, PSD, ...): Pillow, if it is installed
    im = Image.open(path)
    if im.mode not in ("L", "RGB", "RGBA"):
        im = im.convert("RGBA" if "A" in im.getbands() else "RGB")
    a = np.asarray(im, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    # stb_image loads top row first and the reference passes aiProcess_FlipUVs instead of flipping the image,
    # so row 0 of the array stays the top row: nothing to do here
    return np.ascontiguousarray(a)


def parse_mtl(path):
    mats, cur = {}, None
    if not os.path.exists(path):
        return mats
    base = os.path.dirname(path)
    for line in open(path, errors="ignore"):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        k = t[0].lower()
        if k == "newmtl":
            cur = {"Kd": (0.8, 0.8, 0.8), "Ks": (0.0, 0.0, 0.0)}
            mats[" ".join(t[1:])] = cur
        elif cur is None:
            continue
        elif k in ("kd", "ks"):
            cur[t[0][0].upper() + "d" if k == "kd" else "Ks"] = tuple(float(x) for x in t[1:4])
        elif k in ("map_kd", "map_ks", "map_ka"):
            cur[k] = os.path.join(base, t[-1].replace("\\", "/"))
    return mats


def load_obj(path, name=None):
    pos, uvs, nrm = [], [], []
    corners = {}          # (vi, ti, ni) -> output vertex
    out_v = []            # (vi, ti, ni)
    tris, tri_mat = [], []
    mtl, mat_names, cur_mat = {}, [], 0
    base = os.path.dirname(os.path.abspath(path))

    def mat_index(n):
        if n not in mat_names:
            mat_names.append(n)
        return mat_names.index(n)

    for line in open(path, errors="ignore"):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        k = t[0]
        if k == "v":
            pos.append([float(x) for x in t[1:4]])
        elif k == "vt":
            uvs.append([float(t[1]), float(t[2]) if len(t) > 2 else 0.0])
        elif k == "vn":
            nrm.append([float(x) for x in t[1:4]])
        elif k == "mtllib":
            mtl.update(parse_mtl(os.path.join(base, " ".join(t[1:]))))
        elif k == "usemtl":
            cur_mat = mat_index(" ".join(t[1:]))
        elif k == "f":
            if not mat_names:
                cur_mat = mat_index("__default__")
            idx = []
            for c in t[1:]:
                p = (c.split("/") + ["", ""])[:3]
                vi = int(p[0]); ti = int(p[1]) if p[1] else 0; ni = int(p[2]) if p[2] else 0
                vi = vi - 1 if vi > 0 else len(pos) + vi
                ti = (ti - 1 if ti > 0 else len(uvs) + ti) if p[1] else -1
                ni = (ni - 1 if ni > 0 else len(nrm) + ni) if p[2] else -1
                key = (vi, ti, ni)
                if key not in corners:
                    corners[key] = len(out_v)
                    out_v.append(key)
                idx.append(corners[key])
            for j in range(1, len(idx) - 1):          # aiProcess_Triangulate: fan
                tris.append((idx[0], idx[j], idx[j + 1]))
                tri_mat.append(cur_mat)
    if not tris:
        raise ValueError(f"{path}: no faces")
    P = np.asarray(pos, dtype=np.float64)
    keys = np.asarray(out_v, dtype=np.int64)
    tri = np.asarray(tris, dtype=np.int64)
    vp = P[keys[:, 0]]
    vt = np.zeros((len(keys), 2))
    if uvs:
        T = np.asarray(uvs, dtype=np.float64)
        has = keys[:, 1] >= 0
        vt[has] = T[keys[has, 1]]
    vt[:, 1] = 1.0 - vt[:, 1]                          # aiProcess_FlipUVs
    e1, e2 = vp[tri[:, 1]] - vp[tri[:, 0]], vp[tri[:, 2]] - vp[tri[:, 0]]
    fn = np.cross(e1, e2)                              # area-weighted face normals
    if nrm and np.all(keys[:, 2] >= 0):
        vn = np.asarray(nrm, dtype=np.float64)[keys[:, 2]]
    else:                                              # aiProcess_GenSmoothNormals: smooth over shared positions
        acc = np.zeros_like(P)
        for c in range(3):
            np.add.at(acc, keys[tri[:, c], 0], fn)
        vn = acc[keys[:, 0]]
    vn = vn / np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    # aiProcess_CalcTangentSpace
    d1, d2 = vt[tri[:, 1]] - vt[tri[:, 0]], vt[tri[:, 2]] - vt[tri[:, 0]]
    det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    r = np.where(np.abs(det) > 1e-20, 1.0 / np.where(det == 0, 1, det), 0.0)[:, None]
    tan = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) * r
    bit = (e2 * d1[:, 0:1] - e1 * d2[:, 0:1]) * r
    vtan, vbit = np.zeros_like(vp), np.zeros_like(vp)
    for c in range(3):
        np.add.at(vtan, tri[:, c], tan)
        np.add.at(vbit, tri[:, c], bit)

    def _fix(t, n, fallback_axis):
        t = t - n * np.sum(t * n, axis=1, keepdims=True)          # Gram-Schmidt against the normal, as assimp does
        l = np.linalg.norm(t, axis=1, keepdims=True)
        bad = l[:, 0] < 1e-12
        if bad.any():                                            # no usable uv: any vector perpendicular to n
            a = np.where(np.abs(n[bad, fallback_axis:fallback_axis + 1]) < 0.9, np.eye(3)[fallback_axis], np.eye(3)[(fallback_axis + 1) % 3])
            tb = np.cross(n[bad], a)
            t[bad] = tb
            l[bad] = np.linalg.norm(tb, axis=1, keepdims=True)
        return t / np.maximum(l, 1e-20)

    vtan = _fix(vtan, vn, 0)
    vbit = _fix(vbit, vn, 1)
    verts = np.concatenate([vp, vn, vt, vtan, vbit], axis=1).astype(np.float32)

    textures, materials, cache = [], [], {}

    def tex_of(path_or_none, colour, channels=3):
        if path_or_none and os.path.exists(path_or_none):
            if path_or_none not in cache:                        # texture cache by path, Model.h:200-207
                textures.append(_load_image(path_or_none)); cache[path_or_none] = len(textures) - 1
            return cache[path_or_none]
        key = ("solid", tuple(colour), channels)
        if key not in cache:
            t = np.zeros((1, 1, channels), dtype=np.uint8)
            t[0, 0, :] = np.clip(np.rint(np.asarray(colour[:channels]) * 255), 0, 255)
            textures.append(t); cache[key] = len(textures) - 1
        return cache[key]

    for n in mat_names:
        m = mtl.get(n, {"Kd": (0.8, 0.8, 0.8), "Ks": (0.0, 0.0, 0.0)})
        materials.append((tex_of(m.get("map_kd"), m.get("Kd", (0.8, 0.8, 0.8))), tex_of(m.get("map_ks"), m.get("Ks", (0, 0, 0))),
                          tex_of(m.get("map_ka"), (0.5, 0.5, 0.5)), 20.0))     # Shininess = 20, Mesh.h:86
    return Scene(name or os.path.basename(path), verts, tri.astype(np.uint32), np.asarray(tri_mat, dtype=np.uint16),
                 textures, materials, meta={"source": os.path.abspath(path), "materials": list(mat_names)})


def save_frame_png(frame_rgba, path):
    """Frame as returned by vct_render (row 0 = bottom row, GL window order) -> PNG, top row first."""
    from . import images
    images.save_png(np.ascontiguousarray(frame_rgba[::-1, :, :3]), path)
