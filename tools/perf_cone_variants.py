import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.atrium()
u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.sync()
ref = None
for lm in (0, 1, 0, 1):
    c.set_i("DebugLaneMap", lm)
    t = []
    for i in range(12):
        c.render(); c.sync(); t.append(c.pass_time_us("cone"))
    f = c.read_frame()
    if ref is None: ref = f
    print(f"lane_map {lm}: cone {np.median(t[2:]):.1f} us  identical frame: {np.array_equal(f, ref)}")
