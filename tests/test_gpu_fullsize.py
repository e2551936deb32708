"""GPU parity at BASELINE size for configs 3, 4 and 5 (config 2 at size lives in test_gpu_parity.py), and the
config-2 frame against the oracle's OTHER filter model (fp32 trilinear weights).  Run with -m gpu on a B200.

Bars (BASELINE.json north_star): shadow map, voxel occupancy and fragment counts bit-exact; every mip level
bit-exact given level 0; radiance and frames PSNR >= 40 dB and |diff| <= 2/255 on >= 99.9 % of pixels.  The
measured fractions are printed (pytest -s / the captured output of a failure) so that a pass says by how much.
"""
import numpy as np
import pytest

from conftest import frac_within, psnr
from vct_b200 import scenes, uniforms

pytestmark = pytest.mark.gpu

PSNR_MIN, LSB_TOL, FRAC_MIN = 40.0, 2, 0.999


@pytest.fixture(scope="module")
def atrium_full():
    return scenes.atrium()


def report(what, fg, fo):
    p, f = psnr(fg[..., :3], fo[..., :3]), frac_within(fg, fo, LSB_TOL)
    print(f"[parity] {what}: psnr {p:.2f} dB, {100 * f:.4f} % of pixels within {LSB_TOL}/255")
    return p, f


def assert_frame(what, fg, fo):
    p, f = report(what, fg, fo)
    assert p >= PSNR_MIN and f >= FRAC_MIN, f"{what}: psnr {p:.2f}, frac {f:.5f}"


def knot_positions(sc, step):
    """The config-4 animation of bench.py (base + normal * 12 sin(phase + 0.21 step)), in float32 on the host so that
    the oracle and the device see the same bits."""
    base, nrm = sc.verts[:, :3].astype(np.float32), sc.verts[:, 3:6].astype(np.float32)
    phase = (base[:, 0] * np.float32(0.004) + base[:, 2] * np.float32(0.003)).astype(np.float32)
    amp = (np.float32(12.0) * np.sin(phase + np.float32(0.21 * step)).astype(np.float32)).astype(np.float32)
    return (base + nrm * amp[:, None]).astype(np.float32)


def test_config3_full_size_vs_oracle(gpu_ctx, oracle, atrium_full):
    """BASELINE config 3 itself: 259 608 triangles, 512^3 RGBA16F grid, 3840x2160, 9 diffuse cones + specular."""
    sc = atrium_full
    u = uniforms.scene_uniforms(sc, V=512, width=3840, height=2160, shadow_map_size=4096, coverage="conservative",
                                cones="9+1", grid_format=1)
    c = gpu_ctx
    c.set_uniforms(u); c.load_scene(sc); c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
    oracle.set_uniforms(u); oracle.load_scene(sc); oracle.draw_depth(); oracle.draw_voxels(); oracle.render()
    assert np.array_equal(c.depth(), oracle.depth())
    cg = c.counts()
    assert np.array_equal(cg, oracle.counts()) and cg.sum() > 4_000_000
    occ = cg > 0
    del cg
    # level 0: same occupancy; rgb = half(sum / (255 count)) with sums that differ by <= 1 per fragment (hardware-
    # filtered albedo), i.e. within the radiance bar
    g0 = c.grid(0)
    o0 = oracle.grid(0)
    assert np.array_equal(g0[..., 3] > 0, occ) and np.array_equal(o0[..., 3] > 0, occ)
    d = np.abs(g0[occ].astype(np.float32) - o0[occ].astype(np.float32)).max(-1)
    print(f"[parity] config 3 level 0: {occ.sum()} voxels, max |diff| {d.max() * 255:.3f}/255, "
          f"{100 * (d <= LSB_TOL / 255).mean():.4f} % within {LSB_TOL}/255")
    assert (d <= LSB_TOL / 255).mean() >= FRAC_MIN
    del o0, d, occ
    vis_g, vis_o = c.visibility(), oracle.visibility()
    print(f"[parity] config 3 visibility: {(vis_g != vis_o).sum()} of {vis_g.size} pixels differ")
    assert (vis_g != vis_o).mean() <= 1e-4
    assert_frame("config 3 frame (512^3 RGBA16F, 4K, 9+1 cones)", c.read_frame(), oracle.frame())
    assert abs(c.cone_samples() - oracle.cone_samples()) <= 2e-4 * oracle.cone_samples()
    # every mip level bit for bit, given the device's level 0 (the mip rule at full size)
    oracle.set_grid_level0(g0)
    del g0
    for l in range(1, 10):
        assert np.array_equal(c.grid(l).view(np.uint16), oracle.grid(l).view(np.uint16)), f"level {l}"


def test_config4_full_size_vs_oracle(gpu_ctx, oracle):
    """BASELINE config 4 itself: the 1 048 576-triangle knot, V = 256, re-voxelised at two time steps of the
    animation bench.py runs.  Shadow map and fragment counts bit-exact at both, frame within the bar."""
    sc = scenes.dynamic_knot()
    assert sc.n_tris == 1_048_576
    u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
    c = gpu_ctx
    c.set_uniforms(u); c.load_scene(sc)
    oracle.set_uniforms(u); oracle.load_scene(sc)
    prev = None
    for step in (3, 11):
        P = knot_positions(sc, step)
        c.update_positions(P)
        c.draw_depth(); c.frame(); c.sync()
        v = sc.verts.copy(); v[:, :3] = P
        oracle.upload_mesh(v, sc.idx, sc.tri_material)
        oracle.draw_depth(); oracle.draw_voxels(); oracle.render()
        assert np.array_equal(c.depth(), oracle.depth()), f"step {step}: shadow map"
        cg = c.counts()
        assert np.array_equal(cg, oracle.counts()) and cg.sum() > 2_000_000, f"step {step}: counts"
        for l in range(9):          # noise-textured albedo, hardware filtered: radiance within the bar, occupancy equal
            g, o = c.grid(l), oracle.grid(l)
            assert np.array_equal(g[..., 3], o[..., 3]), f"step {step}: level {l} alpha"
            assert np.abs(g.astype(int) - o.astype(int)).max() <= LSB_TOL, f"step {step}: level {l}"
        assert (c.visibility() != oracle.visibility()).mean() <= 1e-4
        assert_frame(f"config 4 frame, step {step}", c.read_frame(), oracle.frame())
        assert prev is None or not np.array_equal(prev, cg > 0)      # the mesh moved
        prev = cg > 0


def test_config5_full_size_vs_oracle(gpu_ctx, oracle, atrium_full):
    """BASELINE config 5 itself: V = 256, Bounces = 3 (re-injection, extension), 1024^2 views of the probe lattice."""
    sc = atrium_full
    cams = scenes.probe_cameras(64)
    base = dict(V=256, width=1024, height=1024, shadow_map_size=4096, coverage="conservative", bounces=3)
    u0 = uniforms.scene_uniforms(sc, **base)
    c = gpu_ctx
    c.set_uniforms(u0); c.load_scene(sc); c.draw_depth(); c.draw_voxels(); c.sync()
    oracle.set_uniforms(u0); oracle.load_scene(sc); oracle.draw_depth(); oracle.draw_voxels()
    assert np.array_equal(c.counts(), oracle.counts())
    g, o = c.grid(0), oracle.grid(0)
    assert np.array_equal(g[..., 3], o[..., 3])
    occ = g[..., 3] > 0
    d = np.abs(g[occ].astype(int) - o[occ].astype(int)).max(-1)
    print(f"[parity] config 5 level 0 after re-injection: max |diff| {d.max()}/255, {100 * (d <= LSB_TOL).mean():.4f} % within {LSB_TOL}")
    assert (d <= LSB_TOL).mean() >= FRAC_MIN
    # The reference's cone loop exits on a hard threshold (alpha < 0.95, VoxelConeTracing.fs:94): a sub-LSB filtering
    # difference on a sample that lands on it adds or drops a whole step.  How often that happens depends on the view:
    # probe 0 (a corner camera 3 units above the floor) is the worst of the lattice -- the oracle's two legitimate
    # filter models (FilterMode 0 vs 1) agree on only 99.885 % of ITS pixels, 99.98-99.9995 % for the others.  So the
    # north_star bar is asserted on the pooled pixels of the views, and every single view must stay above 99.5 %.
    within = total = 0
    for k in (0, 21, 42, 63):
        pos, yaw, pitch = cams[k]
        u = uniforms.scene_uniforms(sc, camera_pos=pos, yaw=yaw, pitch=pitch, **base)
        c.set_uniforms(u); c.render(); c.sync()
        oracle.set_uniforms(u); oracle.render()
        fg, fo = c.read_frame(), oracle.frame()
        p, f = report(f"config 5 probe {k}", fg, fo)
        assert p >= PSNR_MIN and f >= 0.995, f"probe {k}: psnr {p:.2f}, frac {f:.5f}"
        within += f * fg.shape[0] * fg.shape[1]; total += fg.shape[0] * fg.shape[1]
    print(f"[parity] config 5, four probes pooled: {100 * within / total:.4f} % of pixels within {LSB_TOL}/255")
    assert within / total >= FRAC_MIN
    assert np.array_equal(c.grid(0), g)


def test_config2_vs_oracle_with_fp32_filter_weights(gpu_ctx, oracle, atrium_full):
    """The oracle's default FilterMode = 1 models the 8-bit trilinear weights measured on the B200's texture units.
    GL leaves the weight precision open; this test holds the device to the SAME bar against the oracle's fp32-weight
    model (FilterMode = 0), so the parity claim does not rest on an oracle that was tuned to the device."""
    sc = atrium_full
    u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
    c = gpu_ctx
    c.set_uniforms(u); c.load_scene(sc); c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
    uo = dict(u); uo["FilterMode"] = 0
    oracle.set_uniforms(uo); oracle.load_scene(sc); oracle.draw_depth(); oracle.draw_voxels(); oracle.render()
    assert_frame("config 2 frame vs fp32-weight oracle (FilterMode 0)", c.read_frame(), oracle.frame())
    n_g, n_o = c.cone_samples(), oracle.cone_samples()
    print(f"[parity] cone samples: device {n_g}, fp32-weight oracle {n_o} ({abs(n_g - n_o) / n_o:.2e} relative)")
    assert abs(n_g - n_o) <= 1e-3 * n_o
