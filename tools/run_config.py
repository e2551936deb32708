"""Renders a few frames of a BASELINE config on one GPU (for ncu captures).  usage: run_config.py <config> [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
a = bench.parse(["--config", sys.argv[1]])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import vct_b200
sc, u = bench.make_scene_and_uniforms(a)
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.set_i("PipelineFrames", 0); c.set_i("OverlapVisibility", 0)      # one stream: honest serial per-pass times
c.draw_depth()
for i in range(n):
    c.frame(); c.sync()
names = ["depth", "vox_clear", "vox_cover", "vox_shade", "resolve", "mip", "visibility", "cone"]
print("  ".join(f"{k} {c.pass_time_us(k):.1f}" for k in names), "frags", c.fragment_count(), "items", c.debug_counter(0))
