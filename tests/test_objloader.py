"""OBJ/MTL loader = the reference's assimp post-processing restated (Model.h:43): triangulate, smooth normals, flipped
UVs, tangent space; texture slots as Model.h:126-136 / Mesh.h:95-109 bind them."""
import os

import numpy as np

from vct_b200 import objloader


def _write_asset(d):
    from PIL import Image
    Image.fromarray(np.full((4, 4, 3), (200, 40, 40), np.uint8)).save(os.path.join(d, "red.png"))
    Image.fromarray(np.full((2, 2), 128, np.uint8)).save(os.path.join(d, "grey.png"))
    open(os.path.join(d, "box.mtl"), "w").write(
        "newmtl wall\nKd 0.5 0.6 0.7\nKs 0.1 0.1 0.1\nmap_Kd red.png\nmap_Ka grey.png\n\nnewmtl plain\nKd 0.2 0.4 0.6\nKs 0.3 0.3 0.3\n")
    # a unit cube from quads, no normals; one face uses the second material
    v = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    quads = [(4, 3, 2, 1), (6, 7, 8, 5), (2, 6, 5, 1), (3, 7, 6, 2), (4, 8, 7, 3), (8, 4, 1, 5)]   # CCW seen from outside
    with open(os.path.join(d, "box.obj"), "w") as f:
        f.write("mtllib box.mtl\n")
        for p in v:
            f.write("v %g %g %g\n" % p)
        f.write("vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl wall\n")
        for k, q in enumerate(quads):
            if k == 5:
                f.write("usemtl plain\n")
            f.write("f " + " ".join(f"{a}/{j + 1}" for j, a in enumerate(q)) + "\n")
    return os.path.join(d, "box.obj")


def test_obj_loader_matches_the_reference_loader_semantics(tmp_path):
    sc = objloader.load_obj(_write_asset(str(tmp_path)))
    assert sc.n_tris == 12 and sc.verts.shape[1] == 14                      # quads fanned into triangles
    assert sc.idx.max() < len(sc.verts) and len(sc.tri_material) == 12
    assert list(sc.tri_material[:10]) == [0] * 10 and list(sc.tri_material[10:]) == [1, 1]
    n, t, b = sc.verts[:, 3:6], sc.verts[:, 8:11], sc.verts[:, 11:14]
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1, atol=1e-5)      # GenSmoothNormals: unit, pointing outwards
    centre = sc.verts[:, :3] - 0.5
    assert np.all(np.sum(n * centre, axis=1) > 0)
    np.testing.assert_allclose(np.linalg.norm(t, axis=1), 1, atol=1e-5)      # CalcTangentSpace
    np.testing.assert_allclose(np.sum(t * n, axis=1), 0, atol=1e-5)
    np.testing.assert_allclose(np.sum(b * n, axis=1), 0, atol=1e-5)
    uv = sc.verts[:, 6:8]
    assert uv.min() >= 0 and uv.max() <= 1 and {0.0, 1.0} == set(np.unique(uv[:, 1]))   # FlipUVs keeps the range
    d, s, h, shin = sc.materials[0]
    assert sc.textures[d].shape == (4, 4, 3) and tuple(sc.textures[d][0, 0]) == (200, 40, 40)   # map_Kd -> DiffuseTexture
    assert sc.textures[h].shape == (2, 2, 1) and shin == 20.0                                   # map_Ka -> HeightTexture
    assert sc.textures[s].shape == (1, 1, 3) and tuple(sc.textures[s][0, 0]) == (26, 26, 26)     # Ks colour, 1x1
    d2, s2, h2, _ = sc.materials[1]
    assert tuple(sc.textures[d2][0, 0]) == (51, 102, 153) and sc.textures[h2].shape == (1, 1, 3)  # every material has all three


def test_obj_scene_runs_through_the_oracle(tmp_path, oracle):
    from vct_b200 import uniforms
    sc = objloader.load_obj(_write_asset(str(tmp_path)))
    sc.verts[:, :3] = (sc.verts[:, :3] - 0.5) * 60.0 * 20.0                 # a 60-unit cube in model units
    u = uniforms.reference_uniforms(V=32, width=64, height=64, shadow_map_size=256, camera_pos=(0, 20, 120), pitch=-8.0)
    oracle.set_uniforms(u); oracle.load_scene(sc)
    oracle.draw_depth(); oracle.draw_voxels(); oracle.render()
    assert oracle.counts().sum() > 500
    f = oracle.frame()
    assert (f[..., :3] != 128).any()                                          # the cube is visible
    objloader.save_frame_png(f, str(tmp_path / "frame.png"))
    assert os.path.getsize(tmp_path / "frame.png") > 100


def test_builtin_decoders_against_pillow_and_round_trip(tmp_path):
    """images.py (the stb_image stand-in): PNG with every filter type / colour type, PPM, TGA (raw and RLE, both origins)."""
    import zlib, struct
    from PIL import Image
    from vct_b200 import images
    rng = np.random.default_rng(3)
    smooth = (np.add.outer(np.arange(37), np.arange(53))[..., None] * np.array([3, 5, 7, 2]) % 256).astype(np.uint8)
    noisy = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    for arr in (smooth, noisy):
        for mode, a in (("L", arr[..., 0]), ("RGB", arr[..., :3]), ("RGBA", arr), ("LA", arr[..., :2])):
            p = str(tmp_path / f"t_{mode}.png")
            Image.fromarray(a, mode).save(p, optimize=True)          # Pillow picks filters per row (all five occur)
            got = images.load_image(p)
            want = np.asarray(Image.open(p).convert("RGBA" if mode == "LA" else mode))
            assert np.array_equal(got.reshape(want.shape), want), mode
    pal = Image.fromarray(noisy[..., :3], "RGB").quantize(16)
    pal.save(str(tmp_path / "pal.png"))
    assert np.array_equal(images.load_image(str(tmp_path / "pal.png")), np.asarray(pal.convert("RGB")))
    # every filter type explicitly, written by hand
    img = noisy[:5, :11, :3]
    def filt(ft, row, prev):
        row, prev = row.astype(int).reshape(-1), prev.astype(int).reshape(-1)
        out = np.zeros_like(row)
        for x in range(len(row)):
            a = row[x - 3] if x >= 3 else 0; b = prev[x]; c = prev[x - 3] if x >= 3 else 0
            pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
            paeth = a if (pa <= pb and pa <= pc) else b if pb <= pc else c
            out[x] = (row[x] - [0, a, b, (a + b) // 2, paeth][ft]) % 256
        return bytes([ft]) + bytes(out.astype(np.uint8))
    raw = b"".join(filt(y % 5, img[y], img[y - 1] if y else np.zeros_like(img[0])) for y in range(5))
    ck = lambda k, b: struct.pack(">I", len(b)) + k + b + struct.pack(">I", zlib.crc32(k + b))
    png = b"\x89PNG\r\n\x1a\n" + ck(b"IHDR", struct.pack(">IIBBBBB", 11, 5, 8, 2, 0, 0, 0)) + ck(b"IDAT", zlib.compress(raw)) + ck(b"IEND", b"")
    assert np.array_equal(images.decode_png(png), img)
    # own encoder <-> own decoder <-> Pillow
    for a in (noisy[..., 0], noisy[..., :3], noisy):
        b = images.encode_png(a)
        assert np.array_equal(images.decode_png(b).reshape(a.shape), a)
        open(tmp_path / "e.png", "wb").write(b)
        assert np.array_equal(np.asarray(Image.open(tmp_path / "e.png")).reshape(a.shape), a)
    # PPM / PGM
    Image.fromarray(noisy[..., :3], "RGB").save(str(tmp_path / "t.ppm"))
    assert np.array_equal(images.load_image(str(tmp_path / "t.ppm")), noisy[..., :3])
    # TGA: raw + RLE, bottom-left and top-left origin
    for rle in (False, True):
        for mode, a in (("RGB", smooth[..., :3]), ("RGBA", smooth), ("L", smooth[..., 0])):
            p = str(tmp_path / f"t_{mode}_{int(rle)}.tga")
            Image.fromarray(a, mode).save(p, compression="tga_rle" if rle else None, orientation=-1 if rle else 1)
            got = images.load_image(p)
            assert np.array_equal(got.reshape(a.shape), a), (mode, rle)
    with np.testing.assert_raises(images.UnsupportedImage):
        images.load_image(__file__)
