import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.cornell()
u = uniforms.scene_uniforms(sc, V=32, width=96, height=96, shadow_map_size=512, coverage="center")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.draw_voxels(); c.render(); c.sync()
np.savez_compressed("gpurun_out/cornell_v32_gpu.npz", frame=c.read_frame(), vis=c.visibility())
