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
