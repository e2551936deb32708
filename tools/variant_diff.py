"""Where do cone_trace variants disagree?  usage: variant_diff.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.atrium(detail=0.3, tex_size=64)
u = uniforms.scene_uniforms(sc, V=128, width=640, height=360, shadow_map_size=2048, coverage="conservative", cones="9+1")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
ref, n_ref = c.read_frame().astype(int), c.cone_samples()
print("default samples", n_ref)
for v in (1, 3, 4):
    c.set_i("DebugConeVariant", v); c.render(); c.sync()
    f = c.read_frame().astype(int)
    d = np.abs(f - ref).max(-1)
    bad = np.argwhere(d > 0)
    print(f"variant {v}: samples {c.cone_samples()} ({c.cone_samples() - n_ref:+d}), {len(bad)} pixels differ, max diff {d.max()}",
          f"rows {bad[:,0].min()}..{bad[:,0].max()} cols {bad[:,1].min()}..{bad[:,1].max()}" if len(bad) else "")
    if len(bad):
        for (y, x) in bad[:6]:
            print("   ", y, x, "lane", (x % 8) + 8 * (y % 4), ref[y, x], f[y, x])
        lanes = (bad[:, 1] % 8) + 8 * (bad[:, 0] % 4)
        print("    lane histogram", np.bincount(lanes, minlength=32))
for su in (1, 2, 4):
    c.set_i("DebugConeVariant", 0); c.set_i("DebugSpecAhead", su); c.render(); c.sync()
    f = c.read_frame().astype(int)
    print(f"default template, spec-ahead {su}: samples {c.cone_samples() - n_ref:+d}, {(np.abs(f - ref).max(-1) > 0).sum()} pixels differ")
