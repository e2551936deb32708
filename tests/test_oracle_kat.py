"""Known-answer tests that pin the CPU oracle.  The reference has no tests of its own, so
these hand-derivable vectors -- each derived from a cited line of the reference -- are the pins."""
import numpy as np
import pytest

from vct_b200 import scenes, uniforms
from vct_b200 import glmath as gm

F = np.float32


def quad_mesh(p00, p10, p11, p01, scale=20.0):
    """two triangles, model units = world * 20 (ModelMatrix = scale(0.05))"""
    P = np.array([p00, p10, p11, p01], dtype=np.float64) * scale
    v = np.zeros((4, 14), dtype=F)
    v[:, :3] = P
    n = np.cross(P[1] - P[0], P[3] - P[0]); n /= np.linalg.norm(n)
    t = (P[1] - P[0]) / np.linalg.norm(P[1] - P[0]); b = (P[3] - P[0]) / np.linalg.norm(P[3] - P[0])
    v[:, 3:6], v[:, 8:11], v[:, 11:14] = n, t, b
    v[:, 6:8] = [[0, 0], [1, 0], [1, 1], [0, 1]]
    return v, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)


def setup(o, V=128, coverage="center", **kw):
    u = uniforms.reference_uniforms(V=V, coverage=coverage, shadow_map_size=kw.pop("shadow_map_size", 256), **kw)
    o.set_uniforms(u)
    o.upload_texture(0, scenes.solid_texture((255, 255, 255)))
    o.L.orc_set_material(o.h, 0, 0, 0, 0, __import__("ctypes").c_float(20.0))
    return u


# --- Voxelization.gs:34-39 ------------------------------------------------------------------------
@pytest.mark.parametrize("n,axis", [((0, 0, 1), 3), ((0, 1, 0), 2), ((1, 0, 0), 1), ((0.3, 0.5, 0.81), 3),
                                    ((0.6, -0.7, 0.2), 2), ((-0.9, 0.1, 0.3), 1)])
def test_axis_selection(oracle, n, axis):
    n = np.array(n, dtype=np.float64); n /= np.linalg.norm(n)
    a = np.cross(n, [0.3, -0.7, 0.2]); a /= np.linalg.norm(a)
    b = np.cross(n, a)   # a x b is parallel to n
    w0 = np.array([1.0, 2.0, 3.0])
    # Voxelization.gs:25-28: n = cross(p0 - p1, p2 - p0)
    assert oracle.select_axis(w0, w0 + a, w0 + b) == axis
    assert oracle.select_axis(w0, w0 + b, w0 + a) == axis      # sign of the normal is irrelevant (abs)


@pytest.mark.parametrize("e1,e2,axis", [
    ((-1, 1, 0), (0, 0, 1), 1),     # cross = (1,1,0):  nx == ny  -> X wins the tie (>=), Voxelization.gs:34
    ((1, 0, 0), (0, -1, 1), 2),     # cross = (0,-1,-1): ny == nz -> Y (:36)
    ((0, 1, 0), (1, 0, -1), 1),     # cross = (-1,0,-1): nx == nz -> X
    ((1, 1, 0), (-1, 0, 1), 1),     # cross = (1,-1,1):  all equal -> X
])
def test_axis_selection_exact_ties(oracle, e1, e2, axis):
    # integer coordinates make the cross product exact, so the ties are real ties
    w0 = np.array([3.0, -2.0, 5.0])
    w1 = w0 - np.array(e1, dtype=float)      # e1 = p0 - p1
    w2 = w0 + np.array(e2, dtype=float)      # e2 = p2 - p0
    assert oracle.select_axis(w0, w1, w2) == axis


def test_axis_degenerate_is_z(oracle):
    p = [1.0, 1.0, 1.0]
    assert oracle.select_axis(p, p, p) == 3                    # normalize(0) = NaN: every comparison false
    assert oracle.select_axis([0, 0, 0], [1, 1, 1], [2, 2, 2]) == 3


# --- Voxelization.fs:58-86: world -> voxel map under the three axis paths ----------------------------
@pytest.mark.parametrize("p,vox", [((0.3, 0.3, 0.3), (64, 64, 64)), ((-74.9, 10.2, 33.3), (0, 72, 92)),
                                   ((74.9, -74.9, 0.1), (127, 0, 64)), ((10.0, 20.0, -30.0), (72, 81, 38))])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_world_to_voxel_map(oracle, p, vox, axis):
    setup(oracle, V=128)
    p = np.array(p)
    e = np.eye(3)
    a, b = e[(axis + 1) % 3], e[(axis + 2) % 3]
    s = 1.4   # a little more than one voxel (150/128 = 1.17) so that a pixel centre is always covered
    v, i = quad_mesh(p - s * a - s * b, p + s * a - s * b, p + s * a + s * b, p - s * a + s * b)
    oracle.upload_mesh(v, i)
    oracle.draw_depth()
    oracle.draw_voxels()
    c = oracle.counts()
    zz, yy, xx = np.nonzero(c)
    occ = np.stack([xx, yy, zz], 1)
    assert len(occ) > 0
    # every fragment lies in the slab floor((p/150 + 0.5) * 128) along the projection axis ...
    assert np.all(occ[:, axis] == vox[axis])
    # ... and the voxel containing p is among them
    assert any((o == np.array(vox)).all() for o in occ)
    expect = np.floor((p / 150.0 + 0.5) * 128).astype(int)
    assert tuple(expect) == vox


# --- A.1 matrices -----------------------------------------------------------------------------------
def test_reference_matrices_numeric_form():
    u = uniforms.reference_uniforms(V=128, model_scale=1.0)
    m = lambda k: np.asarray(u[k]).reshape(4, 4).T     # back to math layout
    # ProjX: x_ndc = -z/75, y_ndc = y/75, z_ndc = -x/75 ; ProjY: x/75, -z/75, -y/75 ; ProjZ: x/75, y/75, -z/75
    w = np.array([10.0, 20.0, -30.0, 1.0], dtype=F)
    np.testing.assert_allclose(m("ProjX") @ w, [30 / 75, 20 / 75, -10 / 75, 1], atol=1e-6)
    np.testing.assert_allclose(m("ProjY") @ w, [10 / 75, 30 / 75, -20 / 75, 1], atol=1e-6)
    np.testing.assert_allclose(m("ProjZ") @ w, [10 / 75, 20 / 75, 30 / 75, 1], atol=1e-6)
    d = m("DepthModelViewProjectionMatrix")
    expect = np.array([[0.008333, 0, 0, 0], [0, 0.002021, -0.008085, 0], [0, -0.009701, -0.002425, 0.010308], [0, 0, 0, 1]])
    np.testing.assert_allclose(d, expect, atol=2e-6)
    L = gm.normalize((0, 1, 0.25))
    np.testing.assert_allclose(L, [0, 0.970143, 0.242536], atol=1e-6)


def test_cone_table_is_the_references():
    d, w = uniforms.cone_set("6+1")
    assert abs(float(w.sum()) - 1.0) < 1e-6                       # VoxelConeTracing.fs:48
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)   # :49-57
    for k in ("5+1", "9+1"):
        d, w = uniforms.cone_set(k)
        assert abs(float(w.sum()) - 1.0) < 1e-6
        np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)


# --- A.5 cone stepping in free space (VoxelConeTracing.fs:94-104) -----------------------------------
@pytest.mark.parametrize("V,nd,ns", [(64, 5, 18), (128, 6, 23), (256, 7, 29)])
def test_cone_step_counts_empty_grid(oracle, V, nd, ns):
    setup(oracle, V=V)
    col, n = oracle.cone((0, 0, 0), (0, 0, 1), 0.577)
    assert n == nd and np.all(col == 0)
    col, n = oracle.cone((0, 0, 0), (0, 0, 1), 0.07)
    assert n == ns and np.all(col == 0)


def test_diffuse_cone_distance_sequence_v128():
    # dist_{k+1} = dist_k + max(vws, 2*0.577*dist_k), dist_0 = vws = 150/128
    vws = 150.0 / 128
    d, seq, lods = vws, [], []
    while d < 75.0:
        dia = max(vws, 2 * 0.577 * d)
        seq.append(d); lods.append(np.log2(dia / vws)); d += dia
    np.testing.assert_allclose(seq, [1.1719, 2.5242, 5.4372, 11.7117, 25.2269, 54.3388], atol=2e-3)
    np.testing.assert_allclose(lods, [0.207, 1.314, 2.421, 3.528, 4.635, 5.742], atol=2e-3)


@pytest.mark.parametrize("V", [64, 128])
def test_uniform_grid_cone_exits_after_one_sample(oracle, V):
    setup(oracle, V=V)
    c = 153
    g = np.empty((V, V, V, 4), dtype=np.uint8); g[..., :3] = c; g[..., 3] = 255
    oracle.set_grid_level0(g)
    vws = 150.0 / V
    # specular aperture: first diameter = max(vws, 0.14*vws) = vws  => occ = 1/(1+0.03*vws)  (0.966038 at V=128)
    col, n = oracle.cone((1.0, 2.0, 3.0), (0.0, 1.0, 0.0), 0.07)
    assert n == 1                                                   # alpha = 1 >= 0.95 after the first sample
    np.testing.assert_allclose(col[:3], c / 255.0, atol=1e-6)
    np.testing.assert_allclose(col[3], 1.0 / (1.0 + 0.03 * vws), atol=1e-5)
    if V == 128:
        np.testing.assert_allclose(col[3], 0.966038, atol=1e-5)
    # diffuse aperture: first diameter = 2*0.577*vws
    col, n = oracle.cone((1.0, 2.0, 3.0), (0.0, 1.0, 0.0), 0.577)
    assert n == 1
    np.testing.assert_allclose(col[3], 1.0 / (1.0 + 0.03 * 2 * 0.577 * vws), atol=1e-5)


def test_mip_rounding_single_voxel(oracle):
    V = 16
    setup(oracle, V=V)
    g = np.zeros((V, V, V, 4), dtype=np.uint8)
    g[5, 9, 3] = 255
    oracle.set_grid_level0(g)
    assert oracle.grid(1)[2, 4, 1, 0] == 32        # round(255/8); truncation would give 31
    assert oracle.grid(2)[1, 2, 0, 0] == 4
    assert oracle.grid(3)[0, 1, 0, 0] == 1         # round(4/8) = (4+4)>>3 = 1; truncation would give 0
    assert oracle.grid(4)[0, 0, 0, 0] == 0
    assert int(oracle.grid(1).astype(int).sum()) == 32 * 4


def test_sample_voxels_wraps_like_gl_repeat(oracle):
    # Voxel_Cone_Tracing.h:110-113 never sets a wrap mode => GL_REPEAT: +75 and -75 are the same texel edge
    V = 16
    setup(oracle, V=V)
    rng = np.random.default_rng(0)
    g = rng.integers(0, 256, (V, V, V, 4), dtype=np.uint8)
    oracle.set_grid_level0(g)
    a = oracle.sample_voxels((74.0 + 150.0, 3.0, -20.0), 0.0)
    b = oracle.sample_voxels((74.0, 3.0, -20.0), 0.0)
    np.testing.assert_allclose(a, b, atol=2e-6)
    # texel centre of voxel (x=3,y=4,z=5): exact value, no filtering
    vws = 150.0 / V
    p = (np.array([3, 4, 5]) + 0.5) * vws - 75.0
    np.testing.assert_allclose(oracle.sample_voxels(p, 0.0), g[5, 4, 3] / 255.0, atol=1e-6)
    # lod = 1 exactly: pure level-1 trilinear
    p1 = (np.array([2, 1, 3]) + 0.5) * 2 * vws - 75.0
    np.testing.assert_allclose(oracle.sample_voxels(p1, 1.0), oracle.grid(1)[3, 1, 2] / 255.0, atol=1e-6)


# --- fill rule / depth slice (A.7 last bullet) --------------------------------------------------------
@pytest.mark.parametrize("V", [32, 128])
def test_axis_aligned_quad_fills_exactly_one_slab(oracle, V):
    setup(oracle, V=V, coverage="center")
    v, i = quad_mesh((-75, 10.2, 75), (75, 10.2, 75), (75, 10.2, -75), (-75, 10.2, -75))
    oracle.upload_mesh(v, i)
    oracle.draw_depth()
    oracle.draw_voxels()
    c = oracle.counts()
    vy = int(np.floor((10.2 / 150 + 0.5) * V))
    assert np.all(c[:, vy, :] == 1)            # every (vx, vz) exactly once: shared diagonal owned by one triangle
    assert c.sum() == V * V
    g = oracle.grid(0)
    assert np.all(g[:, vy, :, 3] == 255) and g[..., 3].astype(int).sum() == 255 * V * V


@pytest.mark.parametrize("coverage,expect_min", [("center", 1), ("msaa4", 1), ("conservative", 1)])
def test_shared_edge_no_double_count_center(oracle, coverage, expect_min):
    V = 64
    setup(oracle, V=V, coverage=coverage)
    v, i = quad_mesh((-40.3, -3.1, 33.7), (51.9, -3.1, 20.2), (44.4, -3.1, -47.8), (-36.6, -3.1, -29.9))
    oracle.upload_mesh(v, i)
    oracle.draw_depth()
    oracle.draw_voxels()
    c = oracle.counts()
    if coverage == "center":
        assert c.max() == 1                    # top-left rule: no pixel of the shared diagonal is hit twice
    else:
        assert c.max() <= 2                    # any-sample / conservative coverage may touch both triangles
    assert (c > 0).sum() > 1000


def test_coverage_policies_are_nested(oracle):
    V = 64
    occ = {}
    for cov in ("center", "msaa4", "conservative"):
        setup(oracle, V=V, coverage=cov)
        v, i = quad_mesh((-40.3, -3.1, 33.7), (51.9, 9.1, 20.2), (44.4, 12.0, -47.8), (-36.6, 2.0, -29.9))
        oracle.upload_mesh(v, i)
        oracle.draw_depth()
        oracle.draw_voxels()
        occ[cov] = oracle.counts().sum()
    assert occ["center"] <= occ["msaa4"] <= occ["conservative"]


# --- shading known answers ------------------------------------------------------------------------
def test_empty_grid_pixel_colour_matches_closed_form(oracle):
    """Grid empty => every cone returns 0 => colour = albedo*(shadow*max(N.L,0) + ambient) + spec*shadow*specColor,
    shadow = 25*0.111 = 2.775 when fully lit (VoxelConeTracing.fs:158,205,223-227)."""
    albedo = np.array([40, 60, 80]) / 255.0
    # light along +z so that the +z-facing quad is hit head on (no shadow acne at the reference's 0.002 bias)
    u = uniforms.reference_uniforms(V=32, width=64, height=64, shadow_map_size=256, camera_pos=(0, 0, 100.0),
                                    light_direction=(0.0, 0.0, 1.0))
    oracle.set_uniforms(u)
    oracle.upload_texture(0, scenes.solid_texture((40, 60, 80)))
    oracle.upload_texture(1, scenes.solid_texture((0, 0, 0)))          # specular colour 0: no specular term
    oracle.upload_texture(2, scenes.solid_texture((128, 128, 128)))    # flat height map
    oracle.L.orc_set_material(oracle.h, 0, 0, 1, 2, __import__("ctypes").c_float(20.0))
    v, i = quad_mesh((-30, -30, 0), (30, -30, 0), (30, 30, 0), (-30, 30, 0))   # faces +z, towards camera and light
    oracle.upload_mesh(v, i)
    oracle.draw_depth()
    oracle.render()                                                    # no draw_voxels: grid stays empty
    f = oracle.frame()
    expect = albedo * (2.775 * 1.0 + 0.1)
    got = f[32, 32, :3] / 255.0
    np.testing.assert_allclose(got, np.clip(expect, 0, 1), atol=1.0 / 255 + 1e-6)
    assert f[32, 32, 3] == 255
    assert tuple(f[0, 0]) == (128, 128, 128, 255)                      # clear colour 0.5 grey, Voxel_Cone_Tracing.h:156-159
    assert oracle.cone_samples() > 0


def test_pcf_normalisations(oracle):
    """Voxelization.fs:46 divides by 25; with nothing in the shadow map every tap is lit -> 1.0."""
    u = uniforms.reference_uniforms(V=32, shadow_map_size=64)
    oracle.set_uniforms(u)
    v, i = quad_mesh((-1, -1, -90), (1, -1, -90), (1, 1, -90), (-1, 1, -90))
    oracle.upload_mesh(v, i)
    oracle.draw_depth()
    assert oracle.pcf((0.5, 0.5, 0.3, 1.0)) == 1.0
    assert oracle.pcf((0.5, 0.5, 1.5, 1.0)) == 0.0    # beyond the cleared depth (1.0) + bias


def test_texture_channel_rules(oracle):
    """Model.h:159-169: 1 channel -> GL_RED (r,0,0,1); 3 -> (r,g,b,1); mip chain is a 2x2 box filter."""
    t = np.zeros((2, 2, 1), dtype=np.uint8); t[..., 0] = [[0, 255], [255, 255]]
    oracle.upload_texture(0, t)
    np.testing.assert_allclose(oracle.sample_texture(0, 0.25, 0.25, 0.0), [0, 0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(oracle.sample_texture(0, 0.75, 0.25, 0.0), [1, 0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(oracle.sample_texture(0, 0.5, 0.5, 1.0), [191 / 255, 0, 0, 1], atol=1e-6)  # (765+2)>>2
    np.testing.assert_allclose(oracle.sample_texture(0, 0.5, 0.5, 0.0), [0.75, 0, 0, 1], atol=1e-6)       # bilinear centre


# --- accumulator resolve (DESIGN.md "Defined semantics"; Voxelization.fs:88 writes alpha 1) ---------------------
def test_resolve_rule_from_a_hand_set_accumulator(oracle):
    V = 8
    setup(oracle, V=V)
    counts = np.zeros((V, V, V), dtype=np.uint32); sums = np.zeros((V, V, V, 3), dtype=np.uint32)
    counts[1, 2, 3], sums[1, 2, 3] = 3, (10, 11, 765)        # 10/3 = 3.33 -> 3, 11/3 = 3.67 -> 4, 765/3 = 255
    counts[0, 0, 0], sums[0, 0, 0] = 2, (1, 3, 255)          # ties round up: (1+1)/2 = 1, (3+1)/2 = 2, (255+1)/2 = 128
    oracle.set_accum(counts, sums); oracle.resolve_and_mip()
    g = oracle.grid(0)
    assert tuple(g[1, 2, 3]) == (3, 4, 255, 255)
    assert tuple(g[0, 0, 0]) == (1, 2, 128, 255)
    assert int(g[..., 3].astype(int).sum()) == 2 * 255       # every other voxel stays (0, 0, 0, 0)
    # level 1 = (sum of 8 + 4) >> 3 per channel, alpha included
    assert tuple(oracle.grid(1)[0, 0, 0]) == ((1 + 4) >> 3, (2 + 4) >> 3, (128 + 4) >> 3, (255 + 4) >> 3)


def test_fp16_resolve_and_mip_order(oracle):
    """RGBA16F grid (BASELINE config 3): level 0 = half(sum / (count * 255)), alpha = 1; a mip texel is the fp32 sum of the
    eight parents in the order ((a00 + a10) + a01) + a11 over (y, z) of the x-pair sums, times 0.125, rounded to half."""
    V = 4
    setup(oracle, V=V, grid_format=1)
    counts = np.zeros((V, V, V), dtype=np.uint32); sums = np.zeros((V, V, V, 3), dtype=np.uint32)
    counts[0, 0, 0], sums[0, 0, 0] = 3, (100, 200, 765)
    counts[1, 1, 1], sums[1, 1, 1] = 1, (255, 1, 0)
    oracle.set_accum(counts, sums); oracle.resolve_and_mip()
    g0 = oracle.grid(0)
    assert g0.dtype == np.float16
    exp = np.array([np.float32(100) / np.float32(765), np.float32(200) / np.float32(765), 1.0, 1.0], dtype=np.float32).astype(np.float16)
    assert np.array_equal(g0[0, 0, 0], exp)
    assert np.array_equal(g0[1, 1, 1], np.array([1.0, np.float32(1) / np.float32(255), 0.0, 1.0], dtype=np.float32).astype(np.float16))
    p = g0[:2, :2, :2].astype(np.float32)                    # [z][y][x][c]
    pair = p[:, :, 0] + p[:, :, 1]                           # x pairs, fp32
    s = ((pair[0, 0] + pair[0, 1]) + pair[1, 0]) + pair[1, 1]
    assert np.array_equal(oracle.grid(1)[0, 0, 0], (s * np.float32(0.125)).astype(np.float16))
    assert not oracle.grid(1)[1:, 1:, 1:].any()


# --- sharding hooks of the oracle are exact restatements of the unsharded passes ---------------------------------
def test_row_bands_and_triangle_ranges_reproduce_the_whole(oracle):
    sc = scenes.cornell()
    u = uniforms.scene_uniforms(sc, V=32, width=64, height=48, shadow_map_size=256)
    oracle.set_uniforms(u); oracle.load_scene(sc); oracle.draw_depth(); oracle.draw_voxels(); oracle.render()
    frame, counts, sums, g1 = oracle.frame().copy(), oracle.counts().copy(), oracle.sums().copy(), oracle.grid(1).copy()
    n = sc.n_tris
    oracle.draw_voxels_range(0, n // 3, clear_first=True)
    oracle.draw_voxels_range(n // 3, n, clear_first=False)
    oracle.resolve_and_mip()
    assert np.array_equal(oracle.counts(), counts) and np.array_equal(oracle.sums(), sums)
    assert np.array_equal(oracle.grid(1), g1)
    for y0, y1 in ((0, 16), (16, 40), (40, 48)):      # a band call renders only its rows (the bench extrapolates from one)
        oracle.render_rows(y0, y1)
        assert np.array_equal(oracle.frame()[y0:y1], frame[y0:y1])


def test_extra_bounce_only_adds_light_and_keeps_occupancy(oracle):
    """Bounces = 3 (extension, README.md:14 claims it, no reference code): re-injection adds gathered radiance to
    occupied voxels only; alpha (occupancy) is untouched."""
    sc = scenes.cornell()
    u2 = uniforms.scene_uniforms(sc, V=32, width=32, height=32, shadow_map_size=256, bounces=2)
    oracle.set_uniforms(u2); oracle.load_scene(sc); oracle.draw_depth(); oracle.draw_voxels()
    g2 = oracle.grid(0).astype(int)
    u3 = dict(u2); u3["Bounces"] = 3
    oracle.set_uniforms(u3); oracle.draw_voxels()
    g3 = oracle.grid(0).astype(int)
    assert np.array_equal(g2[..., 3], g3[..., 3])
    assert np.all(g3[..., :3] >= g2[..., :3]) and g3[..., :3].sum() > g2[..., :3].sum()
    assert not g3[g3[..., 3] == 0].any()
