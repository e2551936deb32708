"""Timeline of the last four pipelined frames (all streams on one axis).  usage: timeline.py [config] [sharded]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vct_b200
from vct_b200 import parallel
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
sharded = "sharded" in sys.argv
a = bench.parse(["--config", cfg])
sc, u = bench.make_scene_and_uniforms(a)
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc); c.draw_depth()
shared = parallel.SharedAccumulator(c, rank=0, world=1, session="tl") if sharded else None
def frame(i):
    bench.set_camera(c, sc, i, 0)
    shared.frame(None) if sharded else c.frame(None)
for i in range(6): frame(i)
c.sync()
c.set_i("PipelineFrames", 1); c.set_i("Profile", 1)
for i in range(12): frame(100 + i)
shared.wait() if sharded else c.sync()
names = ["vox_clear", "vox_cover", "vox_shade"] + (["exchange_push", "exchange_merge"] if sharded else []) + ["resolve", "mip", "visibility", "cone"]
t0 = None
for back in (3, 2, 1, 0):
    row = []
    for n in names:
        try:
            b, e = c.pass_timeline(back, n)
        except Exception:
            continue
        t0 = b if t0 is None else t0
        row.append(f"{n} {b - t0:7.0f}-{e - t0:7.0f}")
    print(f"frame -{back}: " + " | ".join(row))
