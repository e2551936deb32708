"""Writes tests/golden/reference_shader_vectors.npz:  python tests/golden/make_reference_shader_vectors.py

Vectors produced by EXECUTING THE REFERENCE'S SHADER FILES (/root/reference/Voxel_Cone_Tracing_Final/Shader/*.vs|.gs|.fs,
read in place, never copied) with the GLSL interpreter in tests/glsl_run.py and the fixed-function glue in
tests/glsl_harness.py.  This is the nearest thing to "outputs of the reference itself run here" this image allows: there
is no GL stack and no GPU in the build container, so the C++ host cannot run, but its programmable stages can.

Contents (scene: tests/glsl_harness.fixture_scene(), centre-sample coverage):
  shadow_*   Shadow.vs over every vertex -> window depth at every shadow-map texel well inside one front-facing triangle
  vox_*      Voxelization.vs -> .gs -> .fs over every triangle at V = 16: per voxel, the number of fragments stored and
             the sum of the RGB bytes imageStore would write, for voxels whose fragments are all `certain`
  shards_*   the same program over forty randomly oriented triangles (tests/glsl_harness.shards_scene())
  voxm_*, shardsm_*  both again with 4x MSAA coverage (the reference's default framebuffer): a fragment wherever any of the
             four samples is inside, shaded at the pixel centre
  frame_*    VoxelConeTracing.vs -> .fs for every covered pixel of a 48 x 40 frame at V = 32, given the shadow map, the
             voxel grid and the triangle-per-pixel map stored next to them (the fixed-function inputs of that stage)
  config1_*  the same stage at BASELINE config 1 (Cornell box, 64^3, 256 x 256, 1024^2 shadow map, 4x MSAA voxel coverage)
             on every 13th covered pixel; the inputs are pinned by CRC-32 instead of being stored
  atrium_*   the same stage on the 6126-triangle atrium (22 materials, minified textures, cut-out cards, conservative voxel
             coverage; the scene of atrium_v32_conservative.npz) on every 3rd covered pixel; inputs pinned by CRC-32
  config2_*  the same stage at BASELINE config 2, the headline benchmark configuration (259 608 triangles, 256^3, 1920 x 1080,
             4096^2 shadow map, 512^2 textures), on every 691st covered pixel (3000 pixels); inputs pinned by CRC-32
  config4_*  the same stage at BASELINE config 4 (the 1 048 576-triangle knot at time step 3, 256^3, 1920 x 1080) on every
             307th covered pixel; inputs pinned by CRC-32
  card_*     the same stage on an alpha cut-out card in front of a wall WITHOUT a triangle-per-pixel map: the fragment
             shader runs on the covering triangles nearest first and its `discard` decides which one is seen
The float32 run is the vector; a float64 run, a +-4e-6 gain on the voxel-texture fetches (frame stages) and a 1/256 px
jitter (voxel stages) mark the entries that are numerically stable, i.e. whose value does not hinge on a rounding the GL
specification leaves open (in practice: a cone loop ending with alpha within ~1e-6 of MAX_ALPHA).  The sampled full-size
frames also carry *_slack: per pixel, how far the colour moves when the fragment's inputs are interpolated 1/256 px off
centre (GL's sub-pixel vertex snap) -- a shadow-map tap or texel boundary that close decides by itself.
"""
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, ROOT)
sys.path.insert(0, TESTS)
import vct_b200  # noqa: E402,F401
import glsl_harness as gh  # noqa: E402
from vct_b200 import scenes  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

OUT = os.path.join(HERE, "reference_shader_vectors.npz")


def fixed_function_inputs(kind):
    """Shadow map, voxel grid and visibility for the fixture scene.  They are INPUTS of the stages under test (the
    reference's own frame would get them from the GL pipeline); the oracle supplies them here and the tests check that
    whatever they compare was given the same ones."""
    sc = {"card": gh.card_scene, "shards": gh.shards_scene, "shards_msaa4": gh.shards_scene, "config1": scenes.cornell,
          "atrium": gh.atrium_scene, "config2": scenes.atrium, "config4": gh.config4_scene}.get(kind, gh.fixture_scene)()
    u = gh.scene_uniforms(sc, kind)
    u["FilterMode"] = 0
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth(); o.draw_voxels(); o.render()
    out = dict(depth=o.depth().copy(), grid0=o.grid(0).copy(), visibility=o.visibility().copy())
    o.close()
    return sc, u, out


def stable_frame(sc, u, ff, pixels):
    """float32 colours of the executed fragment stage + which of them survive float64 execution and a +-4e-6 gain on the
    voxel fetches unchanged (to a quarter of an 8-bit step)"""
    run = lambda dtype, gain=1.0: gh.frame_reference_pixels(sc, u, ff["depth"], ff["grid0"], ff["visibility"], pixels, dtype, gain)
    c32 = run(np.float32)
    stable = gh.colours_agree(c32, run(np.float64))
    for gain in gh.GAINS:
        stable &= gh.colours_agree(c32, run(np.float32, gain))
    return c32, stable


def snap_slack(sc, u, ff, pixels, c32):
    """How far (in 8-bit steps, per pixel) the executed fragment stage moves when its inputs are interpolated 1/256 px
    off the pixel centre in the four diagonal directions -- the freedom GL's sub-pixel vertex snap leaves to a
    conforming rasteriser (a PCF tap or a texel boundary crossed by that much)."""
    slack = np.zeros(len(pixels))
    for off in gh.JITTER[1:]:
        c = gh.frame_reference_pixels(sc, u, ff["depth"], ff["grid0"], ff["visibility"], pixels, np.float32, offset=off)
        slack = np.maximum(slack, np.abs(np.clip(np.nan_to_num(c), 0, 1) - np.clip(np.nan_to_num(c32), 0, 1)).max(1) * 255.0)
    return slack.astype(np.float32)


def generate(frame_stride=1, voxel_tris=None, card_stride=1, shard_tris=None, config1_stride=1, atrium_stride=1, config2_stride=1, config4_stride=1, log=print):
    t0 = time.time()
    out = {}
    sc, u, ff = fixed_function_inputs("voxel")
    ij, z, tol = gh.shadow_reference_depths(sc, u)
    out.update(shadow_ij=ij, shadow_z=z, shadow_tol=tol)
    log(f"shadow: {len(ij)} texels  [{time.time() - t0:.1f} s]")
    idx, cnt, sums, bad, nf = gh.voxel_reference_accumulator(sc, u, ff["depth"], voxel_tris)
    out.update(vox_index=idx, vox_count=cnt, vox_sums=sums, vox_uncertain=bad, vox_fragments=np.int64(nf),
               vox_depth_in=ff["depth"])
    log(f"voxel: {nf} fragments, {len(idx)} certain voxels, {len(bad)} uncertain  [{time.time() - t0:.1f} s]")
    sc, u, ff = fixed_function_inputs("shards")
    idx, cnt, sums, bad, nf = gh.voxel_reference_accumulator(sc, u, ff["depth"], shard_tris)
    out.update(shards_index=idx, shards_count=cnt, shards_sums=sums, shards_uncertain=bad, shards_fragments=np.int64(nf))
    log(f"shards: {nf} fragments, {len(idx)} certain voxels, {len(bad)} uncertain  [{time.time() - t0:.1f} s]")
    # the same two scenes under the reference's own default: a 4-sample framebuffer (main.cpp:30)
    for kind, key, tris in (("voxel_msaa4", "voxm", voxel_tris), ("shards_msaa4", "shardsm", shard_tris)):
        sc, u, ff = fixed_function_inputs(kind)
        idx, cnt, sums, bad, nf = gh.voxel_reference_accumulator(sc, u, ff["depth"], tris, coverage="msaa4")
        out.update({f"{key}_index": idx, f"{key}_count": cnt, f"{key}_sums": sums, f"{key}_uncertain": bad, f"{key}_fragments": np.int64(nf)})
        log(f"{kind}: {nf} fragments, {len(idx)} certain voxels, {len(bad)} uncertain  [{time.time() - t0:.1f} s]")
    sc, u, ff = fixed_function_inputs("frame")
    W, H = int(u["screen_width"]), int(u["screen_height"])
    pixels = [(i, j) for j in range(H) for i in range(W) if ff["visibility"][j, i] != 0xFFFFFFFF][::frame_stride]
    c32, stable = stable_frame(sc, u, ff, pixels)
    out.update(frame_px=np.array(pixels, dtype=np.int32), frame_rgba=c32.astype(np.float32), frame_stable=stable,
               frame_depth_in=ff["depth"], frame_grid0_in=ff["grid0"], frame_visibility_in=ff["visibility"])
    log(f"frame: {len(pixels)} pixels, {int(stable.sum())} stable  [{time.time() - t0:.1f} s]")
    # alpha cut-out: no visibility map is given, the discard in the fragment shader decides what is seen
    sc, u, ff = fixed_function_inputs("card")
    W, H = int(u["screen_width"]), int(u["screen_height"])
    pixels = [(i, j) for j in range(H) for i in range(W)][::card_stride]
    t32, c32, nd = gh.frame_reference_depth_ordered(sc, u, ff["depth"], ff["grid0"], pixels, np.float32)
    t64, c64, _ = gh.frame_reference_depth_ordered(sc, u, ff["depth"], ff["grid0"], pixels, np.float64)
    stable = (t32 >= 0) & (t32 == t64) & gh.colours_agree(c32, c64)
    for gain in gh.GAINS:
        tg, cg, _ = gh.frame_reference_depth_ordered(sc, u, ff["depth"], ff["grid0"], pixels, np.float32, gain)
        stable &= (tg == t32) & gh.colours_agree(c32, cg)
    out.update(card_px=np.array(pixels, dtype=np.int32), card_tri=t32, card_rgba=c32.astype(np.float32), card_stable=stable,
               card_discarded=np.int64(nd), card_depth_in=ff["depth"], card_grid0_in=ff["grid0"])
    log(f"card: {len(pixels)} pixels, {int(stable.sum())} stable, {nd} fragments discarded  [{time.time() - t0:.1f} s]")
    # BASELINE config 1 (Cornell box, 64^3, 256 x 256, 1024^2 shadow map, 4x MSAA voxel coverage) at its full size
    sc, u, ff = fixed_function_inputs("config1")
    W, H = int(u["screen_width"]), int(u["screen_height"])
    pixels = [(i, j) for j in range(H) for i in range(W) if ff["visibility"][j, i] != 0xFFFFFFFF][::gh.CONFIG1_STRIDE * config1_stride]
    c32, stable = stable_frame(sc, u, ff, pixels)
    px = np.array(pixels, dtype=np.int32)
    out["config1_slack"] = snap_slack(sc, u, ff, pixels, c32)
    out.update(config1_px=px, config1_rgba=c32.astype(np.float32), config1_stable=stable,
               config1_tri=ff["visibility"][px[:, 1], px[:, 0]].astype(np.int64),
               config1_depth_crc=np.uint32(zlib.crc32(ff["depth"].tobytes())), config1_grid0_crc=np.uint32(zlib.crc32(ff["grid0"].tobytes())))
    log(f"config 1: {len(pixels)} pixels, {int(stable.sum())} stable  [{time.time() - t0:.1f} s]")
    # the atrium: many materials, minified textures (implicit LOD), single-channel maps, cut-out cards, conservative coverage
    sc, u, ff = fixed_function_inputs("atrium")
    W, H = int(u["screen_width"]), int(u["screen_height"])
    pixels = [(i, j) for j in range(H) for i in range(W) if ff["visibility"][j, i] != 0xFFFFFFFF][::gh.ATRIUM_STRIDE * atrium_stride]
    c32, stable = stable_frame(sc, u, ff, pixels)
    px = np.array(pixels, dtype=np.int32)
    out["atrium_slack"] = snap_slack(sc, u, ff, pixels, c32)
    out.update(atrium_px=px, atrium_rgba=c32.astype(np.float32), atrium_stable=stable,
               atrium_tri=ff["visibility"][px[:, 1], px[:, 0]].astype(np.int64),
               atrium_depth_crc=np.uint32(zlib.crc32(ff["depth"].tobytes())), atrium_grid0_crc=np.uint32(zlib.crc32(ff["grid0"].tobytes())))
    log(f"atrium: {len(pixels)} pixels, {int(stable.sum())} stable  [{time.time() - t0:.1f} s]")
    if config2_stride:
        # BASELINE config 2, the headline: 259 608 triangles, 256^3, 1920 x 1080, 4096^2 shadow map, 512^2 textures
        sc, u, ff = fixed_function_inputs("config2")
        W, H = int(u["screen_width"]), int(u["screen_height"])
        pixels = [(i, j) for j in range(H) for i in range(W) if ff["visibility"][j, i] != 0xFFFFFFFF][::gh.CONFIG2_STRIDE * config2_stride]
        c32, stable = stable_frame(sc, u, ff, pixels)
        px = np.array(pixels, dtype=np.int32)
        out["config2_slack"] = snap_slack(sc, u, ff, pixels, c32)
        out.update(config2_px=px, config2_rgba=c32.astype(np.float32), config2_stable=stable,
                   config2_tri=ff["visibility"][px[:, 1], px[:, 0]].astype(np.int64),
                   config2_depth_crc=np.uint32(zlib.crc32(ff["depth"].tobytes())), config2_grid0_crc=np.uint32(zlib.crc32(ff["grid0"].tobytes())))
        log(f"config 2: {len(pixels)} pixels, {int(stable.sum())} stable  [{time.time() - t0:.1f} s]")
    if config4_stride:
        # BASELINE config 4: the 1 048 576-triangle knot (sub-pixel triangles), one time step of the animation
        sc, u, ff = fixed_function_inputs("config4")
        W, H = int(u["screen_width"]), int(u["screen_height"])
        pixels = [(i, j) for j in range(H) for i in range(W) if ff["visibility"][j, i] != 0xFFFFFFFF][::gh.CONFIG4_STRIDE * config4_stride]
        c32, stable = stable_frame(sc, u, ff, pixels)
        px = np.array(pixels, dtype=np.int32)
        out["config4_slack"] = snap_slack(sc, u, ff, pixels, c32)
        out.update(config4_px=px, config4_rgba=c32.astype(np.float32), config4_stable=stable,
                   config4_tri=ff["visibility"][px[:, 1], px[:, 0]].astype(np.int64),
                   config4_depth_crc=np.uint32(zlib.crc32(ff["depth"].tobytes())), config4_grid0_crc=np.uint32(zlib.crc32(ff["grid0"].tobytes())))
        log(f"config 4: {len(pixels)} pixels, {int(stable.sum())} stable  [{time.time() - t0:.1f} s]")
    return out


if __name__ == "__main__":
    if not gh.reference_available():
        sys.exit("the reference's shader files are not at " + gh.SHADER_DIR)
    vectors = generate()
    meta = dict(shader_sha256=gh.shader_hashes(), frame=gh.FRAME, voxel=gh.VOXEL, card=gh.CARD, shards=gh.SHARDS, config1=gh.CONFIG1, config1_stride=gh.CONFIG1_STRIDE, atrium=gh.ATRIUM, atrium_stride=gh.ATRIUM_STRIDE, config2=gh.CONFIG2, config2_stride=gh.CONFIG2_STRIDE, config4=gh.CONFIG4, config4_stride=gh.CONFIG4_STRIDE,
                config4_step=gh.CONFIG4_STEP,
                edge_px=gh.EDGE_PX,
                note="float32 execution of the reference's GLSL text by tests/glsl_run.py")
    np.savez_compressed(OUT, meta=np.array(json.dumps(meta)), **vectors)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
