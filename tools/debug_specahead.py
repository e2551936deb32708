import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.atrium(detail=0.1, tex_size=32)
u = uniforms.scene_uniforms(sc, V=32, width=96, height=54, shadow_map_size=512, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.sync()
fr = {}
for su in (1, 2, 4, 1):
    c.set_i("DebugSpecAhead", su); c.render(); c.sync(); fr[su] = c.read_frame().copy()
    print(su, "samples", c.cone_samples())
for su in (2, 4):
    d = np.abs(fr[su].astype(int) - fr[1].astype(int)).max(-1)
    print("su", su, "pixels differing", (d > 0).sum(), "max", d.max())
    ys, xs = np.nonzero(d > 0)
    print(list(zip(ys[:20], xs[:20])))
vis = c.visibility()
print("bg pixels", (vis == 0xFFFFFFFF).sum())
