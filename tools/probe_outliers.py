import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
from oracle.oracle_py import Oracle
np.set_printoptions(precision=5, suppress=True, linewidth=220)
sc = scenes.cornell()
u = uniforms.scene_uniforms(sc, V=32, width=96, height=96, shadow_map_size=512, coverage="center")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
o = Oracle(); o.set_uniforms(u); o.load_scene(sc); o.draw_depth(); o.draw_voxels(); o.render()
fg, fo = c.read_frame(), o.frame()
d = np.abs(fg.astype(int) - fo.astype(int)).max(-1)
ys, xs = np.nonzero(d > 2)
print("outliers", len(ys), "grids equal", all(np.array_equal(c.grid(l), o.grid(l)) for l in range(6)))
for (j, i) in list(zip(ys, xs))[:6] + [(40, 40)]:
    rc, cones, scal, rgba = o.debug_pixel(i, j)
    use = [0, 1, 2, 3, 4, 5, 7]
    out, steps = c.trace_cones(cones[use, 0:3], cones[use, 3:6], cones[use, 6])
    osteps = [o.cone(cones[k, 0:3], cones[k, 3:6], float(cones[k, 6]))[1] for k in use]
    print(f"pixel ({j},{i}) gpu {fg[j, i, :3]} oracle {fo[j, i, :3]} shadow {scal[0]:.3f}")
    print(np.c_[cones[use, 7:11], out, steps, osteps])
