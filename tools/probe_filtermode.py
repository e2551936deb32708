import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
from oracle.oracle_py import Oracle
def psnr(a, b):
    mse = ((a.astype(float) - b.astype(float)) ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
cases = [("cornell V32 96^2", scenes.cornell(), dict(V=32, width=96, height=96, shadow_map_size=512, coverage="center")),
         ("cornell V64 256^2", scenes.cornell(), dict(V=64, width=256, height=256, shadow_map_size=1024, coverage="msaa4")),
         ("atrium.3 V128 640x360", scenes.atrium(detail=0.3, tex_size=128), dict(V=128, width=640, height=360, shadow_map_size=2048, coverage="conservative"))]
for name, sc, kw in cases:
    u = uniforms.scene_uniforms(sc, **kw)
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
    c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
    fg = c.read_frame()
    frames = {}
    for mode in (0, 1):
        o = Oracle(); uu = dict(u); uu["FilterMode"] = mode
        o.set_uniforms(uu); o.load_scene(sc); o.draw_depth(); o.draw_voxels(); o.render()
        fo = o.frame(); frames[mode] = fo
        d = np.abs(fg.astype(int) - fo.astype(int)).max(-1)
        print(f"{name}: oracle FilterMode={mode}: psnr {psnr(fg[...,:3], fo[...,:3]):.2f} frac<=2 {(d<=2).mean():.5f} frac<=1 {(d<=1).mean():.5f} max {d.max()} samples gpu {c.cone_samples()} oracle {o.cone_samples()}")
        o.close()
    d = np.abs(frames[0].astype(int) - frames[1].astype(int)).max(-1)
    print(f"{name}: oracle fp32-weights vs 8-bit-weights: psnr {psnr(frames[0][...,:3], frames[1][...,:3]):.2f} frac<=2 {(d<=2).mean():.5f} max {d.max()}")
    c.close()
