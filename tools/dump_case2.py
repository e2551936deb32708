import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.cornell()
u = uniforms.scene_uniforms(sc, V=32, width=96, height=96, shadow_map_size=512, coverage="center")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
rng = np.random.default_rng(0)
P = rng.uniform(-110, 110, (60000, 3)).astype(np.float32)
lod = rng.choice(np.arange(0, 5.01, 0.25), 60000).astype(np.float32)
out = c.sample_voxels(P, lod)
np.savez_compressed("gpurun_out/cornell_v32_samples.npz", P=P, lod=lod, out=out)
