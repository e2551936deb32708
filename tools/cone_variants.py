"""cone_trace tuning variants (DebugConeVariant) on config 2: median cone-pass time, Gsamples/s and frame CRC.
usage: cone_variants.py [config]"""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = scenes.atrium()
if cfg == 3:
    u = uniforms.scene_uniforms(sc, V=512, width=3840, height=2160, shadow_map_size=4096, coverage="conservative", cones="9+1", grid_format=1)
else:
    u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.sync()
names = {0: "SU4 64thr minb10 (default)", 1: "SU4 64thr minb8 (round 1)", 2: "SU2 64thr minb10", 3: "SU4 32thr minb16", 4: "SU4 32thr minb20",
         5: "SU2 32thr minb20", 8: "SU4 128thr minb4"}
if cfg == 3:
    names = {0: "NC9 SU4 64thr minb8 (default)", 1: "NC16 SU2 64thr minb8 (round 1)", 3: "NC9 SU2 64thr minb10", 4: "NC9 SU4 32thr minb16"}
for v in sorted(names):
    c.set_i("DebugConeVariant", v)
    t = []
    for i in range(25):
        c.render(); c.sync()
        if i >= 5:
            t.append(c.pass_time_us("cone"))
    n = c.cone_samples()
    print(f"variant {v} {names[v]:28s}: cone {np.median(t):7.1f} us  {n / np.median(t) * 1e-3:6.1f} Gsamples/s  crc {zlib.crc32(c.read_frame().tobytes()):08x}", flush=True)
