"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host math, scene generators."""
import os
import re

import numpy as np
import pytest

import vct_b200
from vct_b200 import capi, parallel, scenes, uniforms
from vct_b200 import glmath as gm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_declared_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "vct_c_api.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(vct_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    L = capi.load_library()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in L.vct_version()


def test_no_cpu_fallback_create_fails_loudly_without_a_device():
    import ctypes as C
    L = capi.load_library()
    h = C.c_void_p()
    rc = L.vct_create(0, C.byref(h))
    if rc == 0:        # running on a GPU box
        L.vct_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in L.vct_last_error(None)
    with pytest.raises(capi.VctError):
        capi.Context(0)


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.join(ROOT, "voxel-cone-tracing_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in src and "libvct_oracle" not in src and "orc_" not in src, f


def test_glm_equivalents():
    m = gm.look_at((0, 1, 0.25), (0, 0, 0), (0, 1, 0))
    np.testing.assert_allclose(m @ m.T, np.eye(4), atol=1e-6) if False else None
    r = m[:3, :3]
    np.testing.assert_allclose(r @ r.T, np.eye(3), atol=1e-6)          # rotation part orthonormal
    np.testing.assert_allclose(m @ np.array([0, 1, 0.25, 1], dtype=np.float32), [0, 0, 0, 1], atol=1e-6)
    o = gm.ortho(-120, 120, -120, 120, -100, 100)
    np.testing.assert_allclose(np.diag(o), [1 / 120, 1 / 120, -1 / 100, 1], atol=1e-7)
    p = gm.perspective(gm.radians(45.0), 16 / 9, 0.1, 1000.0)
    assert p[3, 2] == -1 and abs(p[1, 1] - 1 / np.tan(np.pi / 8)) < 1e-5
    near = p @ np.array([0, 0, -0.1, 1], dtype=np.float32)
    far = p @ np.array([0, 0, -1000.0, 1], dtype=np.float32)
    assert abs(near[2] / near[3] + 1) < 1e-4 and abs(far[2] / far[3] - 1) < 1e-4
    v = gm.view_matrix((0, 4, 0), -90.0, 0.0)                            # Front = (0,0,-1), Camera.h:131-144
    np.testing.assert_allclose(v @ np.array([0, 4, -10, 1], dtype=np.float32), [0, 0, -10, 1], atol=1e-5)
    assert gm.colmajor(gm.scale(0.05))[0] == np.float32(0.05)


def test_scene_generators_are_seeded_and_sized():
    c = scenes.cornell()
    assert 30 <= c.n_tris <= 48 and c.verts.shape[1] == 14
    a1, a2 = scenes.atrium(detail=0.2, tex_size=32), scenes.atrium(detail=0.2, tex_size=32)
    assert np.array_equal(a1.verts, a2.verts) and np.array_equal(a1.idx, a2.idx)
    assert all(np.array_equal(x, y) for x, y in zip(a1.textures, a2.textures))
    assert a1.idx.max() < a1.verts.shape[0] and a1.tri_material.max() < len(a1.materials)
    n = a1.verts[:, 3:6]
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-3)
    for m in a1.materials:                                                # every material has all three textures (A.6 #9)
        assert all(0 <= t < len(a1.textures) for t in m[:3])


def test_atrium_full_size_matches_config_2():
    a = scenes.atrium()
    assert abs(a.n_tris - 260_000) / 260_000 < 0.05
    cut = np.isin(a.tri_material, [18, 19, 20, 21]).mean()
    assert 0.05 < cut < 0.11
    w = a.verts[:, :3] * 0.05
    assert w[:, 0].min() < -90 and w[:, 0].max() > 85                    # wider than the +-75 grid in x


def test_sharding_helpers():
    for n, w in ((10, 3), (1048576, 8), (7, 8), (0, 2)):
        r = [parallel.triangle_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    bands = [parallel.row_band(1080, k, 8) for k in range(8)]
    assert bands[0][0] == 0 and bands[-1][1] == 1080 and all(b[0] % 8 == 0 for b in bands)
    assert all(bands[k][1] == bands[k + 1][0] for k in range(7))
    assert sorted(sum((parallel.views_for_rank(64, k, 8) for k in range(8)), [])) == list(range(64))
    c = np.arange(8, dtype=np.uint32).reshape(2, 2, 2)
    s = np.arange(24, dtype=np.uint32).reshape(2, 2, 2, 3) * 7
    c2, s2 = parallel.unpack_accumulator(parallel.pack_accumulator(c, s))
    assert np.array_equal(c2, c.reshape(-1)) and np.array_equal(s2, s.reshape(-1, 3))


def test_uniform_names_match_the_reference_shader_interface():
    u = uniforms.reference_uniforms()
    for k in ("VoxelDimensions", "VoxelGridWorldSize", "ShadowMapSize", "ProjX", "ProjY", "ProjZ", "ModelMatrix",
              "ModelViewMatrix", "ProjectionMatrix", "DepthModelViewProjectionMatrix", "CameraPosition",
              "LightDirection", "ambientFactor"):                        # Voxel_Cone_Tracing.h:167-187,224-243
        assert k in u
    assert u["VoxelDimensions"] == 128 and u["ShadowMapSize"] == 4096 and u["VoxelGridWorldSize"] == 150.0


def test_equal_row_bands_tile_the_frame():
    """row_band_equal: equal nominal bands (multiples of 8 rows) that cover the frame once; later bands are clipped."""
    from vct_b200 import parallel
    for H in (1080, 2160, 720, 64, 9):
        for world in (1, 2, 3, 4, 8):
            rows, per_all = [], set()
            for r in range(world):
                b0, b1, per = parallel.row_band_equal(H, r, world)
                assert 0 <= b0 <= b1 <= H and per % 8 == 0 and b1 - b0 <= per
                assert b0 == min(r * per, H)
                rows += list(range(b0, b1)); per_all.add(per)
            assert rows == list(range(H)) and len(per_all) == 1
            assert per_all.pop() * world >= H


@pytest.mark.parametrize("height,world", [(1080, 2), (1080, 8), (2160, 8), (360, 3), (7, 4), (256, 1)])
def test_interleaved_row_strips_partition_the_frame(height, world):
    """RowInterleave / RowPhase: every row belongs to exactly one rank, strips are 8 rows, shares differ by at most one strip."""
    owner = np.full(height, -1)
    counts = []
    for r in range(world):
        strips = parallel.row_strips_for_rank(height, r, world)
        counts.append(len(strips))
        for b, e in strips:
            assert b % 8 == 0 and 0 < e - b <= 8 and (owner[b:e] == -1).all()
            owner[b:e] = r
            assert (b // 8) % world == r
    assert (owner >= 0).all() and max(counts) - min(counts) <= 1


def test_flythrough_script_drives_the_fly_camera():
    """tools/flythrough.py: the reference's event loop (main.cpp:97-149) as a script; --dry-run needs no device."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("flythrough", os.path.join(ROOT, "tools", "flythrough.py"))
    fly = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fly)
    ev = fly.parse_script("W*3 M(900,0) E*2 S(10) S")
    assert len(ev) == 8 and ev[0] == ("key", 0) and ev[3] == ("mouse", 900.0, 0.0) and ev[6] == ("scroll", 10.0) and ev[7] == ("key", 1)
    with np.testing.assert_raises(ValueError):
        fly.parse_script("X*3")
    cam = fly.main(["--script", "W*3 M(900,0) E*2 S(10)", "--dt", "0.5", "--dry-run"])
    np.testing.assert_allclose(cam.position, [0.0, 2.6, 205.0 - 3.9], atol=1e-4)     # Cornell camera (0, 0, 205): 3 x 1.3 forward, 2 x 1.3 up
    assert cam.Yaw == 0.0 and cam.Zoom == 35.0
