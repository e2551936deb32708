"""Small frames for compute-sanitizer (memcheck / racecheck / initcheck): every pass incl. the medium raster path, fp16
fused mip, sparse clear / resolve / mip on both slots, Bounces = 3, a two-handle sharded frame on one device."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import capi, scenes, uniforms
for fmt, cones, bounces in ((0, "6+1", 3), (1, "9+1", 2)):
    sc = scenes.atrium(detail=0.12, tex_size=32)
    u = uniforms.scene_uniforms(sc, V=64, width=192, height=112, shadow_map_size=512, coverage="conservative", cones=cones, grid_format=fmt, bounces=bounces)
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc); c.draw_depth()
    for i in range(3):
        c.frame()
    c.sync()
    ref = c.read_frame()
    c.close()
# (the two-handle sharded frame needs the two barrier kernels to run CONCURRENTLY; compute-sanitizer serialises kernels,
#  so under it the device barrier times out -- cleanly, VCT_ERR_STATE after 10 s -- instead of completing: pass --multi
#  only when running without the sanitizer)
if "--multi" not in sys.argv:
    k = scenes.dynamic_knot(nu=96, nv=48)
    u = uniforms.scene_uniforms(k, V=64, width=192, height=112, shadow_map_size=512, coverage="msaa4")
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(k); c.draw_depth(); c.frame(); c.frame(); c.sync(); c.close()
    # a world-of-one sharded frame still runs push / resolve through the library's segment
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(k); c.draw_depth()
    c.comm_init(0, 1, "sanitize")
    for i in range(3):
        c.frame_sharded(None)
    c.frame_sharded_wait(); c.close()
    print("sanitize workload done")
    sys.exit(0)
m = capi.MultiContext([0, 0])
u = uniforms.scene_uniforms(sc, V=64, width=192, height=112, shadow_map_size=512, coverage="conservative")
for r in m.ctx:
    r.set_uniforms(u); r.load_scene(sc); r.draw_depth()
m.comm_init()
host = np.zeros((112, 192, 4), np.uint8)
for i in range(3):
    m.frame_sharded(host)
m.wait(); m.close()
k = scenes.dynamic_knot(nu=96, nv=48)
u = uniforms.scene_uniforms(k, V=64, width=192, height=112, shadow_map_size=512, coverage="msaa4")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(k); c.draw_depth(); c.frame(); c.frame(); c.sync(); c.close()
print("sanitize workload done")
