import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.atrium()
u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.sync()
c.set_i("Profile", 0)
N = 150
for pad in (0, 24000, 28000, 33000, 41000, 52000, 0):
    c.set_i("ConeSmemPad", pad)
    res = []
    for pipe in (1, 0):
        c.set_i("PipelineFrames", pipe)
        for i in range(5): c.frame()
        c.sync(); t0 = time.perf_counter()
        for i in range(N): c.frame()
        c.sync(); res.append((time.perf_counter() - t0) / N * 1e6)
    c.set_i("Profile", 1); c.frame(); c.sync(); cone = c.pass_time_us("cone"); c.set_i("Profile", 0)
    print(f"pad {pad:6d}: pipelined {res[0]:.1f} us  unpipelined {res[1]:.1f} us  cone alone {cone:.1f} us")
